#!/bin/bash
# gpurun --timeout 1200 -- 'bash scripts/gpu_full.sh'   (smoke, all GPU parity tests, the bench line)
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','ms_per_step_eager','gpu_launches','clocks')}); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_sweep','poisson_share_of_step')}); print(d['e2e']['value'], d['cpu_baseline']['value'])"; tail -3 gpurun_out/bench_n1.err
