#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_configs.sh'   (BASELINE configs 2 and 3 on one B200, device-resident stepping)
set -u
mkdir -p gpurun_out
for c in 2 3; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_cfg$c.json')); print('config $c', {k:d[k] for k in ('value','ms_per_step','ms_per_step_eager')}, 'steps/s', 1e3/d['ms_per_step'], d['roofline']['frac'], d['roofline']['poisson_share_of_step'])" || tail -3 gpurun_out/bench_cfg$c.err
done
