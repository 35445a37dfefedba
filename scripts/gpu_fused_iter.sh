#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_fused_iter.sh'   (fused-kernel iteration loop: parity, then timings per variant)
set -u
mkdir -p gpurun_out
echo "== pytest fused"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused" 2>&1 | tail -8
for v in ${VARIANTS:-5 3}; do echo "== sweep bench v$v"; FUSED_VARIANT=$v timeout 300 python scripts/sweep_bench.py 2>&1 | grep -E "^T=|pass_cost"; done
