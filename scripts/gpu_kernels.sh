#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_kernels.sh'   (non-Poisson step kernels: parity, then timings)
set -u
mkdir -p gpurun_out
echo "== pytest (kernels + trajectories)"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not fused" 2>&1 | tail -8
echo "== kernel bench"; for c in 1 0; do STREAM_CFG=$c FIELDS=random timeout 300 python scripts/kernel_bench.py 2>&1 | grep -E "fields|advect|vort|sum"; done
