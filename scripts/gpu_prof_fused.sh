#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_prof_fused.sh'
# (1) iteration-body probe (scalar vs packed f32x2), (2) ncu --set full + source of one T=8 fused pass at 8192^2, (3) step time
set -u
mkdir -p gpurun_out
echo "== jacobi body probe"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -DMAIN -o /tmp/jbp scripts/probes/jacobi_body_probe.cu && /tmp/jbp | tee gpurun_out/jacobi_body_probe.txt
echo "== ncu full, fused T=8 (3rd launch)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_fused -s 2 -c 1 -f -o gpurun_out/fused_T8 \
    python scripts/fused_prof.py 8 8 8 > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
echo "== step (device only)"
for t in "" "limitskip=1"; do
  FS2D_TUNING="$t" timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config 2>gpurun_out/bench_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tuning', d.get('tuning'), 'ms/step', round(d['ms_per_step'],3), 'ms/sweep', round(d['roofline']['ms_per_sweep'],5), d['roofline'].get('update_schedule'))" || tail -5 gpurun_out/bench_err.txt
done | tee gpurun_out/bench_variants.txt
