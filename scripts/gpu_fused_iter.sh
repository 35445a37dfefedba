#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_fused_iter.sh'   (fused kernel iteration: parity, pass timings, step)
set -u
mkdir -p gpurun_out
echo "== fused parity tests"
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or tile_list or two_literal or jacobi_row_range or bench_workload" 2>&1 | tail -4
echo "== sweep bench"
timeout 300 python scripts/sweep_bench.py 2>&1 | tee gpurun_out/sweep_bench.txt | grep -E "^T=|pass_cost"
echo "== step (device only)"
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config 2>gpurun_out/bench_err.txt | tee gpurun_out/bench_line.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'ms/sweep', round(d['roofline']['ms_per_sweep'],5), d['roofline'].get('update_schedule'), 'developed', d['value_developed_state']['ms_per_step']); print(d['roofline'].get('kernel_alone'))" || tail -5 gpurun_out/bench_err.txt
