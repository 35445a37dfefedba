#!/bin/bash
# gpurun --gpus 8 --timeout 600 -- 'bash scripts/gpu_scale.sh'   (weak scaling point at 8 GPUs: 8192^2 cells per GPU)
set -u
mkdir -p gpurun_out
n=${NGPU:-8}
echo "== bench N=$n"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['ms_per_sweep'])" || tail -5 gpurun_out/bench_n$n.err
