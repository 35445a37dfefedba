#!/bin/bash
# Shortest useful GPU check (~3 min):  gpurun --timeout 300 -- 'bash scripts/gpu_sanity_short.sh'
# smoke, the small parity tests (fixtures, trajectories, fused passes), then a 3-step device-only bench line.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest (small parity cases)"
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "reference_fixture or fused_pass_equals_literal or random_mask or division_special or split_into" 2>&1 | tail -6
echo "== bench (device only)"
timeout 120 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config > gpurun_out/sanity_bench.json 2> gpurun_out/sanity_bench.err
tail -c 1500 gpurun_out/sanity_bench.json; tail -3 gpurun_out/sanity_bench.err
