"""Small launches of every shared-memory / TMA kernel for compute-sanitizer (memcheck, racecheck, synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "2d-fluid-simulator_b200"))
import numpy as np
import torch
from fs import _lib
from fs.boundary_condition import BoundaryCondition, build_scene
from fs.fluid_simulator import make_solver

lib = _lib.load()
# (1, 300, 480) has open-fluid tiles: the autonomous-warp path of the fused Jacobi kernel (per-warp TMA refills, progress counters)
for num, X, Y in ((2, 256, 128), (3, 200, 176), (1, 300, 480)):
    const, mask = build_scene(num, X, Y)
    bc = BoundaryCondition(const, mask)
    s = make_solver(bc, 0.05 / Y, 1.0 / 128, 1e3, 5.0, "cip", pressure="jacobi", n_iter=12)
    rng = np.random.default_rng(1)
    for k in ("v", "vx", "vy", "p"):
        getattr(s, k).current.from_numpy(rng.uniform(-1, 1, getattr(s, k).current.tensor.shape).astype(np.float32))
    for tail in (1, 0):                 # the two tails of the Jacobi update (fs2d_set_tuning(4, .))
        lib.fs2d_set_tuning(4, tail)
        for stream in (2, 0, 1):        # streaming-kernel selection; ends on the default
            lib.fs2d_set_tuning(2, stream)
            s.update()
    lib.fs2d_set_tuning(4, 1)
    torch.cuda.synchronize()
    print("ok", num, X, Y, float(s.p.current.tensor.abs().sum()))

# the path main.py runs by default: RB-SOR (both colours in one kernel), dye, clamp
from fs.fluid_simulator import DyeFluidSimulator
sim = DyeFluidSimulator.create(2, 96, 0.05 / 96, 1.0 / 96, 1e4, 5.0, "cip")
for _ in range(3):
    sim.step()
torch.cuda.synchronize()
print("ok default path", float(sim.solver.p.current.tensor.abs().sum()))
