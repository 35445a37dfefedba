#!/bin/bash
# gpurun --timeout 1100 -- 'bash scripts/gpu_verify.sh'
# One B200: the -m gpu suite, the default main.py path, sanitizers on the small grids, and the PCIe probe of the e2e leg.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== default main.py path"
timeout 300 python scripts/default_path_bench.py 2 2048 4096 2>&1 | tee gpurun_out/default_path_bench.txt | grep -E "==|graph replay|sum"
echo "== PCIe probe"
timeout 200 python scripts/probes/h2d_probe.py 2>&1 | tee gpurun_out/h2d_probe.txt
echo "== compute-sanitizer"
for tool in memcheck racecheck; do
  echo "-- $tool"; timeout 280 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|=========.*(Error|hazard|Invalid)" | head -8
done | tee gpurun_out/sanitizer.txt
