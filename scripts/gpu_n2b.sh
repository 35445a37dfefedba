#!/bin/bash
# gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_n2b.sh'   (strips: quick bitwise check + bench line)
set -u
mkdir -p gpurun_out
n=${NGPU:-2}
echo "== strip check x$n (NCCL)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|Error|error|FAIL" | tail -5
echo "== bench N=$n"
for extra in "" "--graph-strips"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config --state quiescent $extra > gpurun_out/bench_n${n}b.json 2> gpurun_out/bench_n${n}b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n${n}b.json')); print('$extra', {k:d[k] for k in ('n_gpus','value','ms_per_step','stepping')}, d['roofline']['ms_per_sweep'])" || tail -5 gpurun_out/bench_n${n}b.err
done
