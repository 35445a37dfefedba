#!/bin/bash
# gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_n2.sh'   (row strips on 2 GPUs: bitwise strip-vs-single check, then the bench line)
set -u
mkdir -p gpurun_out
echo "== split test"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_into" 2>&1 | tail -3
echo "== strip check x2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|Error|error" | head -20
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['ms_per_sweep'], d['e2e']['value'])" || tail -5 gpurun_out/bench_n2.err
