#!/bin/bash
# gpurun --timeout 1200 -- 'bash scripts/gpu_r02b.sh'  (one B200: new parity tests, the default main.py path, then the ncu captures of a bench step)
set -u
mkdir -p gpurun_out
echo "== parity: rbsor fused colours, graphs, dye"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rbsor or cuda_graph or dye or facade or trajectory_matches" 2>&1 | tail -4
echo "== default main.py path"
timeout 400 python scripts/default_path_bench.py 2 2048 4096 2>&1 | tee gpurun_out/default_path_bench.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config --state quiescent > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
echo "== ncu full (each kernel once)"
timeout 800 ncu --set full --clock-control none --import-source on --kernel-id ::regex:k_:3 -f -o gpurun_out/step_full \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config --state quiescent > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | grep -E "launches|step_full"
