#!/bin/bash
# gpurun --timeout 1200 -- 'bash scripts/gpu_check3.sh'   (fused-kernel variant 3 bring-up: parity first, then timings)
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest fused"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused" 2>&1 | tail -30
echo "== sweep bench v3"; FUSED_VARIANT=3 timeout 300 python scripts/sweep_bench.py 2>&1 | tail -20
echo "== sweep bench v1"; FUSED_VARIANT=1 timeout 300 python scripts/sweep_bench.py 2>&1 | grep -E "T= ?(4|8|12):"
echo "== kernel bench"; timeout 300 python scripts/kernel_bench.py 2>&1 | tail -12
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
