#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_fused2.sh'   (fused kernel iteration: parity, pass timings, ncu source capture, step)
set -u
mkdir -p gpurun_out
echo "== fused parity tests"
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or tile_list or two_literal or jacobi_row_range" 2>&1 | tail -4
echo "== sweep bench"
timeout 300 python scripts/sweep_bench.py 2>&1 | tee gpurun_out/sweep_bench.txt | grep -E "tile list|^T=|pass_cost|us/sweep"
echo "== ncu full, fused T=8 (3rd launch)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_fused -s 2 -c 1 -f -o gpurun_out/fused_T8 \
    python scripts/fused_prof.py 8 8 8 > gpurun_out/ncu_fused.log 2>&1
tail -1 gpurun_out/ncu_fused.log
echo "== racecheck"
timeout 280 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|=========.*(Error|hazard|Invalid)" | head -8
echo "== step (device only)"
timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config 2>gpurun_out/bench_err.txt | tee gpurun_out/bench_line.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'ms/sweep', round(d['roofline']['ms_per_sweep'],5), d['roofline'].get('update_schedule'), 'developed', d.get('value_developed_state')); print(d['roofline'].get('kernel_alone'))" || tail -5 gpurun_out/bench_err.txt
