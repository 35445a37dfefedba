#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_sanitize.sh'
set -u
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 280 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|=========.*(Error|hazard|Invalid)" | head -12
done
echo "== fused tests + sweep (swizzled source reads)"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused" 2>&1 | tail -2
timeout 200 python scripts/sweep_bench.py 2>&1 | grep -E "T= ?(4|6|8|12):|pass_cost"
