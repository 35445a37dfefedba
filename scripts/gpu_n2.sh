#!/bin/bash
# gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_n2.sh'
# 2 x B200: the -m gpu test that needs two GPUs (strips == single domain over NCCL, tests/test_distributed.py) and the N=2
# bench line (= BASELINE config 4, res=8192) with its end-to-end leg
set -u
mkdir -p gpurun_out
n=${NGPU:-2}
echo "== pytest tests/test_distributed.py -m gpu"
timeout 500 python -m pytest tests/test_distributed.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_n$n.txt
echo "== bench N=$n"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches','setup_s')}, d['roofline']['ms_per_sweep'], 'e2e', d['e2e'] and d['e2e']['value'])" || tail -5 gpurun_out/bench_n$n.err
