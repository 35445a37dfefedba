#!/bin/bash
# gpurun --gpus 4 --timeout 600 -- 'bash scripts/gpu_n4.sh'   (4 x B200: NCCL strip check, config 5 with equal vs work-balanced strips, weak-scaling point)
set -u
mkdir -p gpurun_out
n=${NGPU:-4}
echo "== strip check x$n (NCCL)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|case ok|Error|error" | tee gpurun_out/mp_strip_check_n$n.txt | tail -6
for mode in "--equal-strips" ""; do
echo "== BASELINE config 5 on $n GPUs $mode"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29535 bench.py --config 5 --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --state quiescent $mode > gpurun_out/bench_config5_n${n}${mode}.json 2> gpurun_out/bench_config5_n$n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config5_n${n}${mode}.json')); r=d['roofline']; print({k:d[k] for k in ('n_gpus','value','ms_per_step','setup_s','strip_bounds')}, 'ms/sweep', r['ms_per_sweep'])" || tail -5 gpurun_out/bench_config5_n$n.err
done
echo "== weak scaling point N=$n (8192^2 cells per GPU)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches','setup_s')}, d['roofline']['ms_per_sweep'], 'developed', d['value_developed_state']['ms_per_step'])" || tail -5 gpurun_out/bench_n$n.err
