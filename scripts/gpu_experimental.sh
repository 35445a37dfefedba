#!/bin/bash
# Round-2 first call:  gpurun --timeout 600 -- 'bash scripts/gpu_experimental.sh'
# The experimental kernels (off by default, logic-verified on the CPU emulation only): parity on the GPU, then timings.
set -u
mkdir -p gpurun_out
echo "== experimental parity tests (GPU)"
FS2D_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nonadv_fused or fused_non_advection" 2>&1 | tail -8
echo "== kernel bench, random fields"
timeout 240 python scripts/kernel_bench.py 2>&1 | tee gpurun_out/kernel_bench_experimental.txt | tail -14
echo "== kernel bench, uniform fields (quiescent-like)"
FIELDS=uniform timeout 240 python scripts/kernel_bench.py 2>&1 | tee -a gpurun_out/kernel_bench_experimental.txt | tail -14
echo "== fused Jacobi: variant 6 (pair barriers, experimental) vs variant 5 (default)"
FS2D_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair_barrier" 2>&1 | tail -4
for v in 5 6; do FUSED_VARIANT=$v timeout 200 python scripts/sweep_bench.py 2>&1 | tee -a gpurun_out/sweep_bench_variants.txt | grep -E "variant|T= ?(4|8|12):"; done
