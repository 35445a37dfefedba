"""The path main.py runs by default: DyeFluidSimulator.create(bc, res, dt, dx, re, 5.0, "cip") = CIP + vorticity confinement +
RedBlackSorPressureUpdater(omega 1.3, 2 iterations) + 3-channel dye carried with CIP (/root/reference/main.py:80-82,
fs/fluid_simulator.py:76-78,144-146).  Steps/s with CUDA-graph replay, and one eager step timed per library call.

    python scripts/default_path_bench.py [bc] [res ...]
"""
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "2d-fluid-simulator_b200"))
import torch  # noqa: E402

from fs import _lib  # noqa: E402
from fs.fluid_simulator import DyeFluidSimulator  # noqa: E402

bc_num = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for res in [int(x) for x in sys.argv[2:]] or [2048, 4096]:
    dt, dx = 0.05 / res, 1.0 / res
    sim = DyeFluidSimulator.create(bc_num, res, dt, dx, 1e6, 5.0, "cip")
    for _ in range(3):
        sim.step()
    # per-call timing of one eager step: wrap _lib.call
    calls, real_call = [], _lib.call

    def timed(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); real_call(name, *args); e1.record()
        calls.append((name, e0, e1))

    _lib.call = timed
    sim.step()
    torch.cuda.synchronize()
    _lib.call = real_call
    cells = 2 * res * res
    agg = {}
    for name, e0, e1 in calls:
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1) * 1e3
    print(f"== bc{bc_num} res={res} ({2 * res}x{res} = {cells / 1e6:.1f} M cells), default path (cip + vc + rbsor x2 + dye): one eager step by library call")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:28s} x{n:2d} {us:9.1f} us   {us * 1e3 / cells / n:7.3f} ns/cell/call")
    print(f"  sum {sum(v[1] for v in agg.values()):9.1f} us")
    sim.enable_cuda_graph()
    for _ in range(5):
        sim.step()
    torch.cuda.synchronize()
    n = 50
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        sim.step()
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / n
    print(f"  graph replay: {ms:.3f} ms/step = {1e3 / ms:.1f} steps/s = {cells / ms / 1e6:.2f} G cell-updates/s", flush=True)
    del sim
    torch.cuda.empty_cache()
