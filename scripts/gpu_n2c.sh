#!/bin/bash
# gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_n2c.sh'   (strips: how many SMs the persistent kernels should leave to NCCL)
set -u
mkdir -p gpurun_out
n=${NGPU:-2}
for r in 0 2 4 8 16; do
FS2D_RESERVE_SMS=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$((r % 10)) bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config --state quiescent > gpurun_out/bench_n${n}_r$r.json 2> gpurun_out/bench_n${n}_r$r.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n${n}_r$r.json')); print('reserve $r', {k:d[k] for k in ('n_gpus','value','ms_per_step')}, d['roofline']['ms_per_sweep'])" || tail -5 gpurun_out/bench_n${n}_r$r.err
done
echo "== strip check x$n (NCCL) with the default reservation"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|Error|error|FAIL" | tail -3
