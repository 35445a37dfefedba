#!/bin/bash
# gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scale.sh'
set -u
mkdir -p gpurun_out
for n in 8 4; do
  echo "== bench N=$n"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])" || tail -5 gpurun_out/bench_n$n.err
done
echo "== strip check x4"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 tests/mp_strip_check.py 2>&1 | grep -E "MP_CHECK|Error" | head -3
