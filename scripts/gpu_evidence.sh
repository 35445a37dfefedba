#!/bin/bash
# gpurun --timeout 1800 -- 'bash scripts/gpu_evidence.sh'
# One B200, everything the round's single-GPU evidence consists of: the -m gpu suite, smoke, the bench line and the reference
# arm, BASELINE configs 2 / 3 / 5 on their own grids, the default main.py path, sanitizers, and the ncu captures.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default command line)"
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']; print({k:d[k] for k in ('value','ms_per_step','ms_per_step_eager','gpu_launches','setup_s')}); print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'cfg2', d['baseline_config_2'] and d['baseline_config_2'].get('steps_per_s')); print('roofline', {k: r.get(k) for k in ('bound','achieved','frac','frac_dram','step_dram_gbs','ms_per_sweep','traffic','kernel_alone','update_schedule')}); print('developed', d['value_developed_state'] and d['value_developed_state']['ms_per_step'], 'clocks', d['clocks'])" || tail -5 gpurun_out/bench_n1.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json; python -c "
import json; d=json.load(open('gpurun_out/bench_reference.json')); print(d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['kind'], d['cpu_baseline']['repeat_values'])"
echo "== BASELINE configs on their own grids"
for c in 2 3 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_cfg$c.json')); r=d['roofline']; print('config $c', {k:d[k] for k in ('value','ms_per_step','ms_per_step_eager','setup_s')}, 'steps/s', 1e3/d['ms_per_step'], 'ms/sweep', r['ms_per_sweep'], 'tiles', r.get('tile_lists'), 'sched', r.get('update_schedule'))" || tail -3 gpurun_out/bench_cfg$c.err
done
echo "== default main.py path"
timeout 400 python scripts/default_path_bench.py 2 2048 4096 2>&1 | tee gpurun_out/default_path_bench.txt | grep -E "==|graph replay|sum"
echo "== compute-sanitizer"
for tool in memcheck racecheck; do
  echo "-- $tool"; timeout 280 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|=========.*(Error|hazard|Invalid)" | head -8
done | tee gpurun_out/sanitizer.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config --state quiescent > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full (each kernel once)"
timeout 800 ncu --set full --clock-control none --import-source on --kernel-id ::regex:k_:3 -f -o gpurun_out/step_full \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config --state quiescent > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
echo "== ncu full + source, fused T=8 (3rd launch)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_fused -s 2 -c 1 -f -o gpurun_out/fused_T8 \
    python scripts/fused_prof.py 8 8 8 > gpurun_out/ncu_fused.log 2>&1
tail -1 gpurun_out/ncu_fused.log
ls -la gpurun_out | grep -E "ncu-rep|launches"
