#!/bin/bash
# Run on the GPU box via:  gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
# smoke + GPU parity tests + a short bench + ncu launch list + one full ncu capture of the Jacobi sweep.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
echo "== ncu full (jacobi)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_fused -s 4 -c 2 -f -o gpurun_out/jacobi_full \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
