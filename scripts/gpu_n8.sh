#!/bin/bash
# gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_n8.sh'
# 8 x B200: (1) BASELINE config 5 bench line (bc5, res=16384, 200 sweeps), strips of equal work and equal strips; (2) the
# weak-scaling point (8192^2 cells per GPU) with the end-to-end leg; (3) strips == single domain bitwise over NCCL, incl. the
# config-5 grid itself
set -u
mkdir -p gpurun_out
n=${NGPU:-8}
for mode in "" "--equal-strips"; do
echo "== BASELINE config 5 on $n GPUs $mode"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29535 bench.py --config 5 --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $mode > gpurun_out/bench_config5_n$n$mode.json 2> gpurun_out/bench_config5_n$n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config5_n$n$mode.json')); r=d['roofline']; print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches','setup_s','strip_bounds')}, 'ms/sweep', r['ms_per_sweep'], 'sched', r.get('update_schedule'), 'dev', d['value_developed_state'] and d['value_developed_state']['ms_per_step'])" || tail -5 gpurun_out/bench_config5_n$n.err
done
echo "== weak scaling point N=$n (8192^2 cells per GPU)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches','setup_s')}, d['roofline']['ms_per_sweep'], 'e2e', d['e2e'] and d['e2e']['value'], d['numa'])" || tail -5 gpurun_out/bench_n$n.err
echo "== strip check x$n (NCCL), with the config-5 grid"
FS2D_STRIP_BIG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|case ok|Error|error" | tee gpurun_out/mp_strip_check_n$n.txt | tail -22
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo_n$n.txt
