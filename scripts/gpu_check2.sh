#!/bin/bash
# gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_check2.sh'
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench N=1"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['ms_per_sweep'], d['e2e']['value'], d['cpu_baseline']['value'])"; tail -3 gpurun_out/bench_n1.err
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
