#!/bin/bash
# gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_experimental_n2.sh'
# The experimental kernels on row strips over NCCL: bitwise strip == single check (default, then with the fused non-advection
# kernel on the strips), then the N=2 bench line with and without them.
set -u
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== strip check (default kernels)"; run 29561 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|Error" | head -5
echo "== strip check (fused non-advection on the strips)"; FS2D_FUSED_NONADV=1 run 29562 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|Error" | head -5
echo "== strip check (all experimental kernels)"; FS2D_STRIP_TUNING="1=8,4=1,5=1,limitskip=1" FS2D_FUSED_NONADV=1 run 29564 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|Error" | head -5
for t in "" "1=8,4=1,5=1,nonadv=1,limitskip=1"; do
  echo "== bench N=2, FS2D_TUNING='$t'"
  FS2D_TUNING="$t" run 29563 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tuning', d.get('tuning'), 'ms/step', round(d['ms_per_step'],3), 'G cell-updates/s', round(d['value']/1e9,2))"
done | tee gpurun_out/bench_experimental_n2.txt
