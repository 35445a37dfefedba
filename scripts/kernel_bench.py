"""CUDA-event timing of the non-Poisson step kernels at 8192x8192 (inputs >> L2)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "2d-fluid-simulator_b200"))
import torch
from fs.boundary_condition import BoundaryCondition, build_scene
from fs.fluid_simulator import make_solver
from fs.solver import VELOCITY_LIMIT, limit_field

X = Y = 8192
const, mask = build_scene(2, X, Y)
bc = BoundaryCondition(const, mask)
s = make_solver(bc, 0.05 / Y, 1.0 / Y, 1e5, 5.0, "cip", pressure="jacobi", n_iter=80)
import os
MODE = os.environ.get("FIELDS", "random")   # random | uniform (v = const: Laplacians / vorticity exactly 0, like a quiescent start)
for f in (s.v, s.vx, s.vy, s.p):
    if MODE == "random":
        f.current.tensor.uniform_(-1, 1); f.next.tensor.uniform_(-1, 1)
    else:
        f.current.tensor.fill_(0.25 if f is s.v else 0.0); f.next.tensor.zero_()
from fs import _lib
_lib.load().fs2d_set_tuning(3, int(os.environ.get("STREAM_CFG", "0")))
_lib.load().fs2d_set_tuning(6, int(os.environ.get("NONADV_VEC", "1")))   # 0: the one-cell-per-thread cip_nonadv kernel
print("fields:", MODE, "stream cfg", os.environ.get("STREAM_CFG", "0"))
vc = s.vorticity_confinement
cases = {
    "cip_nonadv (21 B)": (lambda: s._non_advection_phase(s.v.next, s.v.current, s.p.current), 21),
    "cip_nonadv_grad (49 B)": (lambda: s._non_advection_phase_grad(s.vx.next, s.vy.next, s.vx.current, s.vy.current, s.v.current, s.v.next), 49),
    "cip_advect (49 B)": (lambda: s._advection_phase(s.v.next, s.vx.next, s.vy.next, s.v.current, s.vx.current, s.vy.current, s.v.current), 49),
    "vort_calc (17 B)": (lambda: vc._calc_vorticity(s.v.current), 17),
    "vort_add (25 B)": (lambda: vc._add_vorticity(s.v.next, s.v.current), 25),
    "vort_apply fused (25 B)": (lambda: vc._apply_fused(s.v.next, s.v.current), 25),
    "p_source (16 B)": (lambda: s.pressure_updater._source(s.v.current), 16),
    "limit (8 B)": (lambda: limit_field(s.v.current, VELOCITY_LIMIT, bc=bc), 8),
}
tot = 0.0
for name, (fn, b) in cases.items():
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    tot += us
    print(f"{name:26s} {us:8.1f} us   {b * X * Y / us / 1e6:7.2f} TB/s algorithmic", flush=True)
print(f"sum {tot:.0f} us")
