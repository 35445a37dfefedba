#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_ncu_step.sh'
# (1) launch list of one bench run (durations only), (2) one --set full capture of each distinct kernel of a step
#     (its 3rd invocation, i.e. after warm-up), from the same bench command.
set -u
mkdir -p gpurun_out
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
echo "== ncu full (each kernel once)"
timeout 800 ncu --set full --clock-control none --import-source on --kernel-id ::regex:k_:3 -f -o gpurun_out/step_full \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | head -30
