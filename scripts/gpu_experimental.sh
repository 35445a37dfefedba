#!/bin/bash
# Round-2 first call:  gpurun --timeout 600 -- 'bash scripts/gpu_experimental.sh'
# The experimental kernels (off by default, logic-verified on the CPU emulation only): parity on the GPU, then timings.
set -u
mkdir -p gpurun_out
echo "== experimental parity tests (GPU)"
FS2D_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nonadv_fused or fused_non_advection or marching_vorticity or limit_skip" 2>&1 | tail -8
echo "== kernel bench, random fields"
timeout 240 python scripts/kernel_bench.py 2>&1 | tee gpurun_out/kernel_bench_experimental.txt | tail -14
echo "== kernel bench, uniform fields (quiescent-like)"
FIELDS=uniform timeout 240 python scripts/kernel_bench.py 2>&1 | tee -a gpurun_out/kernel_bench_experimental.txt | tail -14
echo "== fused Jacobi: variants 6 (pair barriers), 7 (resolved slow-cell table), 8 (both) vs variant 5 (default)"
FS2D_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair_barrier" 2>&1 | tail -4
for v in 5 6 7 8; do FUSED_VARIANT=$v timeout 200 python scripts/sweep_bench.py 2>&1 | tee -a gpurun_out/sweep_bench_variants.txt | grep -E "variant|T= ?(4|8|12):"; done
echo "== emitting tail pass (fs2d_set_tuning(4, 1)): parity, then the update with and without it"
FS2D_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "emitting_tail" 2>&1 | tail -4
timeout 200 python - <<'PY' 2>&1 | tee gpurun_out/tail_emit_timing.txt
import sys; sys.path.insert(0, "2d-fluid-simulator_b200")
import torch
from fs import _lib
from fs.boundary_condition import BoundaryCondition, build_scene
from fs.double_buffer import DoubleBuffer, Field
from fs.pressure_updater import JacobiPressureUpdater
X = Y = 8192
const, mask = build_scene(2, X, Y)
bc = BoundaryCondition(const, mask)
jac = JacobiPressureUpdater(bc, 0.05 / Y, 1.0 / Y, 80)
p, v = DoubleBuffer((X, Y), 1), Field((X, Y), 2)
for tail in (0, 1, 0, 1):
    _lib.load().fs2d_set_tuning(4, tail)
    for _ in range(2):
        jac.update(p, v)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        jac.update(p, v)
    e1.record(); torch.cuda.synchronize()
    print(f"tail={tail}: {e0.elapsed_time(e1) / 5:.3f} ms per 80-iteration update", flush=True)
PY
echo "== compute-sanitizer on the experimental kernels"
for tool in memcheck racecheck synccheck; do
  echo "-- $tool"; FS2D_EXPERIMENTAL=1 timeout 280 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|=========.*(Error|hazard|Invalid)" | head -8
done
echo "== whole step with the experimental kernels (device-only bench line), default first"
for t in "" "1=7" "1=8" "1=8,4=1" "1=8,4=1,5=1" "1=8,4=1,5=1,nonadv=1,limitskip=1"; do
  FS2D_TUNING="$t" timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-config 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tuning', d.get('tuning'), 'ms/step', round(d['ms_per_step'],3), 'ms/sweep', round(d['roofline']['ms_per_sweep'],5))"
done | tee gpurun_out/bench_experimental.txt
