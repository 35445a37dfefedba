#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_fused.sh'   (fused Jacobi kernel: parity, timings per pass size, sanitizer, step time)
set -u
mkdir -p gpurun_out
echo "== fused parity tests"
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or tile_list or two_literal or jacobi_row_range or config5_grid" 2>&1 | tail -6
echo "== sweep bench"
timeout 300 python scripts/sweep_bench.py 2>&1 | tee gpurun_out/sweep_bench.txt | grep -E "tile list|^T=|pass_cost|us/sweep"
echo "== compute-sanitizer"
for tool in memcheck racecheck; do
  echo "-- $tool"; timeout 280 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|=========.*(Error|hazard|Invalid)" | head -8
done
echo "== step (device only)"
for t in "" "4=0" "limitskip=1"; do
  FS2D_TUNING="$t" timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config 2>gpurun_out/bench_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tuning', d.get('tuning'), 'ms/step', round(d['ms_per_step'],3), 'ms/sweep', round(d['roofline']['ms_per_sweep'],5), d['roofline'].get('update_schedule'))" || tail -5 gpurun_out/bench_err.txt
done | tee gpurun_out/bench_variants.txt
