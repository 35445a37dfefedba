#!/bin/bash
# gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_n2.sh'   (row strips on 2 GPUs: bitwise strip-vs-single check over NCCL, bench lines)
set -u
mkdir -p gpurun_out
n=${NGPU:-2}
echo "== strip check x$n (NCCL)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|case ok|Error|error" | tee gpurun_out/mp_strip_check_n$n.txt | tail -20
echo "== bench N=$n (eager launches on strips)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-extra-config > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches','setup_s')}, d['roofline']['frac'], d['roofline']['ms_per_sweep'], 'e2e', d['e2e']['value'], d['numa'])" || tail -5 gpurun_out/bench_n$n.err
echo "== bench N=$n (CUDA graphs on strips)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config --graph-strips --state quiescent > gpurun_out/bench_n${n}_graph.json 2> gpurun_out/bench_n${n}_graph.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n${n}_graph.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step','stepping')}, d['roofline']['ms_per_sweep'])" || tail -8 gpurun_out/bench_n${n}_graph.err
