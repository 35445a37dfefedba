"""Micro-benchmark of the single Jacobi sweep variants at 8192x8192 (CUDA events, L2-exceeding inputs)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "2d-fluid-simulator_b200"))
import torch
from fs import _lib
from fs.boundary_condition import BoundaryCondition, build_scene
from fs.double_buffer import Field
from fs.pressure_updater import JacobiPressureUpdater

X = Y = 8192
const, mask = build_scene(2, X, Y)
bc = BoundaryCondition(const, mask)
jac = JacobiPressureUpdater(bc, 0.05 / Y, 1.0 / Y, 1)
a, b, v = Field((X, Y), 1), Field((X, Y), 1), Field((X, Y), 2)
a.tensor.uniform_(-1, 1); v.tensor.uniform_(-1, 1)
src = jac._source(v)
lib = _lib.load()
for inline in (False,):
    for rows in (4,):
        lib.fs2d_set_tuning(0, rows)
        for _ in range(3):
            jac._sweep(b, a, src, inline); jac._sweep(a, b, src, inline)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            jac._sweep(b, a, src, inline); jac._sweep(a, b, src, inline)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"inline={inline} rows/warp={rows:2d}: {ms*1e3:7.1f} us/sweep  algorithmic {12*X*Y/ms/1e6:7.1f} GB/s  actual~{17*X*Y/ms/1e6:7.1f} GB/s", flush=True)

print("fused passes (T iterations per pass), us per iteration:")
for T in (1, 4, 8, 12):
    if bc.fused_ok(T):
        bc.fused_order(T)
        print("  tile list T =", T, "(entries, slow, dropped) =", [v[1:] for k, v in bc._fused_orders.items() if k[0] == T][0])
costs = {}
for T in range(1, 13):
    if not bc.fused_ok(T):
        print("T", T, "not valid for this mask"); continue
    for _ in range(2):
        _lib.call("fs2d_jacobi_fused", b.ptr(), a.ptr(), src.ptr(), _lib.ptr(bc._pcode), bc.dom, T, _lib.ptr(bc.fused_order(T)[0]), bc.fused_order(T)[1], _lib.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.call("fs2d_jacobi_fused", b.ptr(), a.ptr(), src.ptr(), _lib.ptr(bc._pcode), bc.dom, T, _lib.ptr(bc.fused_order(T)[0]), bc.fused_order(T)[1], _lib.stream())
        _lib.call("fs2d_jacobi_fused", a.ptr(), b.ptr(), src.ptr(), _lib.ptr(bc._pcode), bc.dom, T, _lib.ptr(bc.fused_order(T)[0]), bc.fused_order(T)[1], _lib.stream())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    costs[T] = ms * 1e3
    print(f"T={T:2d}: {ms*1e3:8.1f} us/pass  {ms*1e3/T:7.1f} us/iteration  algorithmic {12*X*Y*T/ms/1e6:8.1f} GB/s", flush=True)
print("pass_cost table:", "{0, " + ", ".join(f"{costs.get(T, 1e9):.0f}" for T in range(1, 13)) + "}")
