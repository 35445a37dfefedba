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
jac = JacobiPressureUpdater(bc, 0.05 / Y, 1.0 / Y, 1, fuse=0)
a, b, v = Field((X, Y), 1), Field((X, Y), 1), Field((X, Y), 2)
a.tensor.uniform_(-1, 1); v.tensor.uniform_(-1, 1); b.tensor.copy_(a.tensor)
src = jac._source(v)
for T in [int(x) for x in sys.argv[1:]]:
    _lib.call("fs2d_jacobi_fused", b.ptr(), a.ptr(), src.ptr(), _lib.ptr(bc._pcode), bc.dom, T, _lib.ptr(bc.fused_order(T)[0]), bc.fused_order(T)[1], _lib.stream())
torch.cuda.synchronize()
