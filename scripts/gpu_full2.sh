#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_full2.sh'   (one B200: the whole -m gpu suite, the bench line, the default main.py path)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default command line)"
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']; print({k:d[k] for k in ('value','ms_per_step','ms_per_step_eager','gpu_launches','setup_s')}); print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'cfg2', d['baseline_config_2'] and d['baseline_config_2'].get('steps_per_s')); print('roofline', {k: r.get(k) for k in ('bound','achieved','frac','frac_dram','ms_per_sweep','traffic','kernel_alone','tile_lists','update_schedule')}); print('developed', d['value_developed_state'] and d['value_developed_state']['ms_per_step'], 'clocks', d['clocks'])" || tail -5 gpurun_out/bench_n1.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['kind'], d['cpu_baseline']['repeat_values'], d['config']['reference_sample'])"
echo "== default main.py path"
timeout 400 python scripts/default_path_bench.py 2 2048 4096 2>&1 | tee gpurun_out/default_path_bench.txt
echo "== kernel bench"
timeout 240 python scripts/kernel_bench.py 2>&1 | tee gpurun_out/kernel_bench.txt | tail -12
