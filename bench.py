#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (one solver time step) on N B200s of one box.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle)

Workload (`config.workload`): BASELINE config 4's per-GPU shard -- scene bc=2, CIP advection,
Re=1e5, vorticity confinement 5.0 (main.py default), dt = 0.05/8192, dx = 1/8192, 80 Jacobi sweeps
per step, fp32, quiescent start -- with 8192 x 8192 cells PER GPU (weak scaling): the global grid is
(8192*N) x 8192 rows x columns, split into N row strips.  N=2 is exactly BASELINE config 4
(res=8192, 16384 x 8192).  metric = cell-updates/s = cells * steps / time (all cells).

One JSON line on rank 0.  `value`: device-resident stepping (state stays in HBM).  `e2e`: the same
step through the public API with HOST buffers: every step uploads the state (v, vx, vy, p) from
pinned host memory and downloads (v, p) (the `field_to_numpy()` payload) inside the timed region.
`roofline`: the Jacobi sweep (12 algorithmic B/cell/sweep) timed with CUDA events around
the pressure update inside the timed steps.  `cpu_baseline`: the CPU oracle (a C/OpenMP restatement
of the reference's Taichi kernels -- real Taichi cannot be installed offline) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
for _p in (str(REPO), str(REPO / "2d-fluid-simulator_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CELLS_PER_GPU_ROWS = 8192
Y_COLS = 8192
SCENE, SCHEME, RE, VC, N_JACOBI = 2, "cip", 1e5, 5.0, 80
ALGO_BYTES_PER_CELL_SWEEP = 12.0       # p read + source read + p write; mask excluded
ALGO_BYTES_PER_CELL_STEP = 139.0 + 12.0 * N_JACOBI


def parse() -> argparse.Namespace:
    global SCENE, RE, VC, N_JACOBI, Y_COLS, ALGO_BYTES_PER_CELL_STEP
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rows-per-gpu", type=int, default=CELLS_PER_GPU_ROWS, help="override for quick checks")
    ap.add_argument("--cols", type=int, default=Y_COLS)
    ap.add_argument("--jacobi", type=int, default=N_JACOBI)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--balance-strips", action="store_true", help="N > 1: strips of equal work (walls cost less) instead of equal rows; default for --config 5")
    ap.add_argument("--equal-strips", action="store_true", help="--config 5: equal rows per strip")
    ap.add_argument("--graph-strips", action="store_true", help="N > 1: capture the strips' steps (kernels + NCCL SendRecvs) into CUDA graphs too")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-config", action="store_true", help="skip the BASELINE configs[1] side measurement (N=1)")
    ap.add_argument("--cpu-sample-rows", type=int, default=1024, help="rows of the CPU-baseline sample grid")
    ap.add_argument("--state", choices=["quiescent", "developed", "both"], default="both",
                    help="initial state of the device-resident timing: the reference's all-zero start (the headline `value`), "
                         "seeded non-uniform fields (`developed`), or both (default; the second one as value_developed_state)")
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 5],
                    help="run a BASELINE.json config on its own grid instead of the headline workload: 2 = bc2 res=2048 Re=1e4 "
                         "80 sweeps, 3 = bc3 res=4096 Re=1e8 vc=10 100 sweeps, 5 = bc5 res=16384 Re=1e6 200 sweeps (rows split over --gpus)")
    a = ap.parse_args()
    if a.config:
        SCENE, res, RE, VC, N_JACOBI = {2: (2, 2048, 1e4, 5.0, 80), 3: (3, 4096, 1e8, 10.0, 100), 5: (5, 16384, 1e6, 5.0, 200)}[a.config]
        Y_COLS = a.cols = res
        a.rows_per_gpu, a.jacobi = 2 * res // a.gpus, N_JACOBI
        ALGO_BYTES_PER_CELL_STEP = 139.0 + 12.0 * N_JACOBI
    return a


def peaks() -> tuple[float, str]:
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha256() -> str:
    """identity of the dominant kernel's source: an ncu capture only describes the code it was taken from"""
    import hashlib

    h = hashlib.sha256()
    for name in ("fs2d_fused.cu", "fs2d_common.cuh"):
        h.update((REPO / "2d-fluid-simulator_b200" / "csrc" / name).read_bytes())
    return h.hexdigest()


def traffic_from_profile():
    """dram__bytes_read+write per launch of the dominant kernel (and of a whole step) from the committed ncu capture
    (profiles/dominant_kernel_traffic.json, written by scripts/summarize_profiles.py).  The capture records the hash of the
    kernel sources it profiled; a capture of other code is refused (None) instead of silently going stale."""
    f = REPO / "profiles" / "dominant_kernel_traffic.json"
    try:
        t = json.loads(f.read_text())
    except Exception:  # noqa: BLE001
        return None
    if t.get("source_sha256") != kernel_source_sha256():
        return {"stale": True, "note": f"{f.name} was captured from other kernel sources ({str(t.get('source_sha256'))[:12]}..., "
                                        f"now {kernel_source_sha256()[:12]}...): re-run scripts/gpu_evidence.sh and scripts/summarize_profiles.py"}
    return t


def pin_to_gpu_numa(local_rank: int) -> dict:
    """CPU affinity of this rank = the NUMA node its GPU hangs off, set BEFORE pinned host buffers are allocated (first touch
    puts them on that node): at N = 8 the host->device streams of the e2e leg otherwise share one root complex."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        node = int(Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node").read_text())
        if node < 0:
            return {"numa_node": None}
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "error": f"{type(e).__name__}: {e}"[:120]}


def developed_like_state(solver, seed: int = 0) -> None:
    """Non-uniform fields in every buffer (smooth structures of a few dozen cells plus cell-scale noise; both signs of both
    velocity components): the instruction paths a developed flow takes -- non-zero Laplacians, vorticity gradients and source
    terms, all four CIP upwind directions -- instead of the exactly uniform regions of a quiescent start (fdiv_z's zero
    dividends, the confinement NaN rule).  A real developed flow at 8192^2 is out of reach (CFL 0.05: 1e5+ steps)."""
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + seed)
    for name, amp in (("v", 1.0), ("vx", 8.0), ("vy", 8.0), ("p", 0.5)):
        buf = getattr(solver, name)
        for f in (buf.current, buf.next):
            t = f.tensor
            shape = t.shape if t.dim() == 3 else t.shape + (1,)
            coarse = torch.rand((1, shape[2], shape[0] // 32 + 2, shape[1] // 32 + 2), device=t.device, generator=g) * 2 - 1
            smooth = torch.nn.functional.interpolate(coarse, size=(shape[0], shape[1]), mode="bilinear", align_corners=False)[0]
            smooth = smooth.permute(1, 2, 0).reshape(t.shape)
            t.copy_(amp * (smooth + 0.05 * (torch.rand(t.shape, device=t.device, generator=g) * 2 - 1)))
            del coarse, smooth
            f.dirty = True
    buf = solver.p                      # the never-written wall cells of the two pressure buffers must agree for fused passes
    buf.next.tensor.copy_(buf.current.tensor)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms DURING the timed region (one long-lived process)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.samples, self._proc, self._t = index, [], None, None

    def _read(self) -> None:
        for line in self._proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) == 6:
                self.samples.append(parts)

    def __enter__(self):
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                           "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._read, daemon=True)
            self._t.start()
            time.sleep(0.25)     # first sample lands before the timed region starts; the rest fall inside it
        except Exception:  # noqa: BLE001
            self._proc = None
        return self

    def __exit__(self, *a):
        if self._proc is not None:
            time.sleep(0.12)
            self._proc.terminate()
            try:
                self._proc.wait(timeout=3)
            except Exception:  # noqa: BLE001
                self._proc.kill()
            self._t.join(timeout=3)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the oracle; the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def taichi_reference(rows: int, cols: int, n_jacobi: int, steps: int, warmup: int, repeats: int):
    """The reference ITSELF under a real Taichi runtime (`ti.init(arch=ti.cpu)`, /root/reference/main.py:65-66), when both can be
    imported on this box: taichi, and the reference's `fs` package from baseline/_ref (or $FS2D_REFERENCE_DIR).  Neither exists
    in the build container nor on the pod's GPU boxes (taichi==1.7.4 has no wheel in /opt/wheelhouse, requires-python >= 3.13;
    `pip install --target baseline/_ref /root/reference` fails at metadata generation -- DESIGN.md), so this returns None and the
    oracle port below is timed instead.  When it fires, the line says kind "taichi" and the first step is also compared with the
    oracle (the shim's lowering choices: fast_math, NaN min/max -- SURVEY T2/T5)."""
    ref_dir = os.environ.get("FS2D_REFERENCE_DIR") or str(REPO / "baseline" / "_ref")
    if not (Path(ref_dir) / "fs" / "solver.py").exists():
        return None
    try:
        import taichi as ti
    except Exception:  # noqa: BLE001
        return None
    import numpy as np

    from fs.boundary_condition import build_scene as our_scene
    from oracle import oracle as orc

    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "fs" or k.startswith("fs.")}
    sys.path.insert(0, ref_dir)
    try:
        ti.init(arch=ti.cpu)
        from fs.boundary_condition import BoundaryCondition as RefBC      # the reference's own classes from here on
        from fs.pressure_updater import JacobiPressureUpdater as RefJacobi
        from fs.solver import CipMacSolver as RefCip
        from fs.vorticity_confinement import VorticityConfinement as RefVC

        const, mask = our_scene(SCENE, rows, cols)
        dt, dx = 0.05 / Y_COLS, 1.0 / Y_COLS
        bc = RefBC(const, mask)
        solver = RefCip(bc, RefJacobi(bc, dt, dx, n_jacobi), dt, dx, RE, RefVC(bc, dt, dx, VC))
        solver.update()
        ti.sync()
        check = orc.OracleSolver(mask, const, dt, dx, RE, SCHEME, VC, ("jacobi", n_jacobi))
        check.update()
        v_ref, v_orc = solver.v.current.to_numpy(), check.v.current
        agree = {"max_abs_diff_v_step1": float(np.nanmax(np.abs(v_ref - v_orc))), "bit_identical_v_step1": bool(np.array_equal(v_ref, v_orc, equal_nan=True))}
        for _ in range(max(warmup, 4)):          # every ti.template() kernel is specialised per field tuple: A/B swaps compile twice
            solver.update()
        ti.sync()
        rates = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            for _ in range(steps):
                solver.update()
            ti.sync()
            rates.append(rows * cols * steps / (time.perf_counter() - t0))
        v = statistics.median(rates)
        return {"value": v, "unit": "cell-updates/s", "cores": os.cpu_count() or 1, "kind": "taichi",
                "sample": f"the reference's CipMacSolver + JacobiPressureUpdater({n_jacobi}) under taichi {ti.__version__} arch=cpu on a "
                          f"{rows}x{cols} grid, median of {repeats} x {steps} steps; vs oracle after step 1: {agree}",
                "ms_per_step": rows * cols / v * 1e3, "repeat_values": rates}
    except Exception as e:  # noqa: BLE001
        print(f"[bench] taichi reference unavailable: {type(e).__name__}: {e}", file=sys.stderr)
        return None
    finally:
        sys.path.remove(ref_dir)
        for k in [k for k in sys.modules if k == "fs" or k.startswith("fs.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def cpu_reference(rows: int, cols: int, n_jacobi: int, steps: int, warmup: int, repeats: int = 3) -> dict:
    """The reference's CPU path on `rows x cols` cells of the workload: real Taichi if this box has it (see above), else the
    C/OpenMP oracle port on ALL host cores.  The team size is set through the OpenMP runtime: torch.distributed.run exports
    OMP_NUM_THREADS=1 to every rank, which starved this arm at N > 1 in round 1."""
    r = taichi_reference(rows, cols, n_jacobi, steps, warmup, repeats)
    if r is not None:
        return r
    from fs.boundary_condition import build_scene
    from oracle import oracle as orc

    cores = orc.set_threads(os.cpu_count() or 1)
    const, mask = build_scene(SCENE, rows, cols)
    dt, dx = 0.05 / Y_COLS, 1.0 / Y_COLS
    s = orc.OracleSolver(mask, const, dt, dx, RE, SCHEME, VC, ("jacobi", n_jacobi))
    for _ in range(warmup):
        s.update()
    rates = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        for _ in range(steps):
            s.update()
        rates.append(rows * cols * steps / (time.perf_counter() - t0))
    v = statistics.median(rates)
    return {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port",
            "sample": f"CPU oracle (C/OpenMP restatement of the reference's Taichi kernels; Taichi itself is not installable "
                      f"offline), {cores} OpenMP threads, same scene/scheme/{n_jacobi} sweeps on a {rows}x{cols} grid, median of "
                      f"{repeats} x {steps} steps after {warmup} warm-up, {rows * cols / v * 1e3:.0f} ms/step on that grid",
            "ms_per_step": rows * cols / v * 1e3, "repeat_values": rates}


def run_reference(a: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = min(a.cpu_sample_rows, a.rows_per_gpu * a.gpus)
    r = cpu_reference(rows, a.cols, a.jacobi, a.steps, a.warmup)
    cfg = workload_config(a, a.gpus)
    # what was actually timed: a bounded sample of the workload; the value is a per-cell rate, so the full grid's step time is
    # sample ms/step x (global cells / sample cells)
    cfg["reference_sample"] = {"rows": rows, "cols": a.cols, "cells": rows * a.cols, "extrapolation": "per cell (cell-updates/s)",
                               "ms_per_step_on_sample": r["ms_per_step"],
                               "ms_per_step_extrapolated_to_global_grid": r["ms_per_step"] * (a.rows_per_gpu * a.gpus) / rows}
    line = {"impl": "reference", "metric": "cell-updates/s", "value": r["value"], "unit": "cell-updates/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "gpu_launches": 0,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "repeat_values")},
            "e2e": {"value": r["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(a: argparse.Namespace, n: int) -> dict:
    what = f"BASELINE config {a.config}" if a.config else "BASELINE config 4 shard"
    return {"workload": f"{what}: bc={SCENE} {SCHEME} Re={RE:g} vc={VC} dt=0.05/{Y_COLS} dx=1/{Y_COLS} "
                        f"jacobi={a.jacobi}/step, grid {a.rows_per_gpu * n}x{a.cols} ({a.rows_per_gpu}x{a.cols} cells/GPU"
                        + ("" if a.config else "; N=2 == res=8192") + "), quiescent start, -no_dye",
            "global_grid": [a.rows_per_gpu * n, a.cols], "cells_per_gpu": a.rows_per_gpu * a.cols,
            "parallelism": f"row-strips x{n}", "l2_policy": f"working set {a.rows_per_gpu * a.cols * 82 / 1e9:.1f} GB/GPU (82 B/cell) >> 126 MB L2, no flush needed"}


def time_baseline_config(scene: int, res: int, re: float, vc: float, n_jacobi: int, steps: int, warmup: int) -> dict:
    """One BASELINE.json config on its own (2*res x res) grid on the current GPU: device-resident CUDA-graph stepping, CUDA
    events.  The p / source working set of res=2048 (100 MB) is close to the 126 MB L2, so this is NOT an HBM-roofline number."""
    import torch

    from fs.fluid_simulator import FluidSimulator

    dt, dx = 0.05 / res, 1.0 / res
    sim = FluidSimulator.create(scene, res, dt, dx, re, vc, "cip", pressure="jacobi", n_iter=n_jacobi)
    sim.enable_cuda_graph()
    for _ in range(warmup):
        sim.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sim.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cells = 2 * res * res
    return {"workload": f"BASELINE configs[1]: bc={scene} res={res} cip Re={re:g} vc={vc} dt=auto jacobi={n_jacobi}/step, grid {2 * res}x{res}, "
                        "quiescent start, device-resident, cuda-graph replay",
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "cell_updates_per_s": cells * 1e3 / ms,
            "note": "working set near the L2 size: not an HBM-roofline figure"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a: argparse.Namespace) -> None:
    import torch
    import torch.distributed as dist

    from fs import _lib
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.distributed import Partition
    from fs.fluid_simulator import make_solver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa(local)
    # stdout carries exactly ONE JSON line: whatever libraries print there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    assert lib.fs2d_device_ok(), lib.fs2d_last_error().decode()
    # experimental kernel selection for A/B runs (default: none): FS2D_TUNING="4=0,2=2" -> fs2d_set_tuning(key, value) pairs;
    # recorded in the JSON line as "tuning"
    tuning = {}
    for item in filter(None, os.environ.get("FS2D_TUNING", "").split(",")):
        k, v = item.split("=")
        tuning[k] = int(v)
        _lib.call("fs2d_set_tuning", int(k), int(v))

    X, Y = a.rows_per_gpu * world, a.cols
    dt, dx = 0.05 / Y_COLS, 1.0 / Y_COLS
    t_setup = time.perf_counter()
    bounds = None
    if world > 1 and (a.balance_strips or (a.config == 5 and not a.equal_strips)):
        # strips of equal WORK instead of equal rows: the fused Jacobi passes skip tiles inside walls and pay double for tiles
        # with BC cells, the other kernels stream every cell (fs.boundary_condition.scene_row_cost)
        from fs.boundary_condition import scene_row_cost
        from fs.distributed import balanced_bounds

        bounds = balanced_bounds(scene_row_cost(SCENE, X, Y, a.jacobi), world, min_rows=128)
    part = Partition(X, rank, world, 0 if world == 1 else 33, bounds)  # halo 33: four fused Jacobi passes of up to 8 iterations per exchange
    if world == 1:
        const, mask = build_scene(SCENE, X, Y)
        bc = BoundaryCondition(const, mask, partition=part)
    else:   # every rank builds (and uploads) only its strip of the scene: O(strip) host and device memory, not O(global grid)
        a0, a1 = BoundaryCondition.strip_rows(part)
        const, mask = build_scene(SCENE, X, Y, rows=(a0, a1))
        bc = BoundaryCondition(const, mask, partition=part, row_offset=a0)
    del const, mask
    solver = make_solver(bc, dt, dx, RE, VC, SCHEME, pressure="jacobi", n_iter=a.jacobi)
    solver.pressure_updater.plan(solver.p)      # setup work (fused-pass validity analysis, tile lists) happens here, not in the warm-up
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    cells_total = X * Y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing -------------------------------------------------------------
    # (1) eager pass with CUDA events around pressure_updater.update() -> the roofline numbers;
    # (2) the headline `value`: the same K steps replayed as CUDA graphs on one rank (FluidSimulator.enable_cuda_graph,
    #     one graph launch per step), eagerly on strips (the NCCL exchanges are issued from the host).
    from fs.fluid_simulator import FluidSimulator

    use_graph = (world == 1 or a.graph_strips) and not a.no_graph

    def time_device(solver, sample_clocks: bool):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        solver.timing_events = None
        for _ in range(a.warmup):
            solver.update()
        barrier()
        e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record()
        for k in range(a.steps):
            solver.timing_events = ev[k]      # the solver records these around pressure_updater.update()
            solver.update()
        e_stop.record()
        barrier()
        solver.timing_events = None
        ms_eager = max_over_ranks(e_start.elapsed_time(e_stop)) / a.steps
        sim = FluidSimulator(solver)
        if use_graph:
            sim.enable_cuda_graph(strips=world > 1)
        for _ in range(a.warmup):
            sim.step()
        barrier()
        n0 = lib.fs2d_launch_count()
        t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = ClockSampler(local) if sample_clocks else None
        if clk:
            clk.__enter__()
        t_start.record()
        for k in range(a.steps):
            sim.step()
        t_stop.record()
        barrier()
        if clk:
            clk.__exit__()
        launches = lib.fs2d_launch_count() - n0
        if use_graph:   # replayed kernels do not pass through the library's counter: count what one eager step launches
            sim._graphs = None
            c0 = lib.fs2d_launch_count()
            solver.update()
            torch.cuda.synchronize()
            launches = (lib.fs2d_launch_count() - c0) * a.steps
        ms_total = max_over_ranks(t_start.elapsed_time(t_stop))
        ms_poisson = max_over_ranks(sum(s.elapsed_time(e) for s, e in ev) / a.steps)
        return {"ms_total": ms_total, "ms_step": ms_total / a.steps, "ms_step_eager": ms_eager, "ms_poisson": ms_poisson,
                "launches": int(launches), "clocks": clk.summary() if clk else None}

    developed = None
    if a.state in ("developed", "both"):
        developed_like_state(solver)
        r = time_device(solver, sample_clocks=a.state == "developed")
        developed = {"value": cells_total * a.steps / (r["ms_total"] * 1e-3), "unit": "cell-updates/s", "ms_per_step": r["ms_step"],
                     "ms_per_sweep": r["ms_poisson"] / a.jacobi, "ms_non_poisson": r["ms_step_eager"] - r["ms_poisson"],
                     "finite": bool(torch.isfinite(solver.v.current.tensor).all()),
                     "what": "same workload started from seeded non-uniform fields (smooth structures + cell-scale noise in v, vx, vy, p): "
                             "the instruction paths of a developed flow instead of the exactly uniform regions of a quiescent start"}
        if a.state == "both":      # back to the reference's initial state (all fields zero, fs/double_buffer.py:14-18) for the headline
            for name in ("v", "vx", "vy", "p"):
                getattr(solver, name).reset()
            vcf = solver.vorticity_confinement
            if vcf is not None:
                vcf.vorticity.fill(0)
                vcf.vorticity_abs.fill(0)
    if a.state != "developed":
        r = time_device(solver, sample_clocks=True)
    clocks, launches = r["clocks"], r["launches"]
    ms_total, ms_step, ms_step_eager, ms_poisson = r["ms_total"], r["ms_step"], r["ms_step_eager"], r["ms_poisson"]
    value = cells_total * a.steps / (ms_total * 1e-3)
    ms_sweep = ms_poisson / a.jacobi
    try:    # the update's schedule (host-side query): fused pass sizes, 0 = one literal {BC, sweep} iteration
        schedule = solver.pressure_updater.plan(solver.p)
    except Exception:  # noqa: BLE001
        schedule = None

    # the dominant kernel on its own: CUDA events around back-to-back T = 8 passes on this rank's strip (the pass size the
    # schedule uses most), ping-pong between the two pressure buffers
    kernel_alone = None
    jac = solver.pressure_updater
    t_dom = max(set(t for t in (schedule or []) if t > 0), key=(schedule or []).count, default=0)
    if t_dom and world == 1:
        src = jac._src
        pa, pb = solver.p.current, solver.p.next
        save = (pa.tensor.clone(), pb.tensor.clone())
        for _ in range(2):
            jac._fused(pb, pa, src, t_dom)
            jac._fused(pa, pb, src, t_dom)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        k0.record()
        for _ in range(reps):
            jac._fused(pb, pa, src, t_dom)
            jac._fused(pa, pb, src, t_dom)
        k1.record()
        torch.cuda.synchronize()
        pa.tensor.copy_(save[0]); pb.tensor.copy_(save[1])
        del save
        kernel_alone = {"T": t_dom, "us_per_launch": k0.elapsed_time(k1) * 1e3 / (2 * reps), "launches_timed": 2 * reps}

    peak, peak_src = peaks()
    cells_gpu = a.rows_per_gpu * Y
    achieved = ALGO_BYTES_PER_CELL_SWEEP * cells_gpu / (ms_sweep * 1e-3) / 1e9
    traffic = traffic_from_profile()
    usable = bool(traffic) and not traffic.get("stale") and not a.config and a.rows_per_gpu == CELLS_PER_GPU_ROWS and Y == 8192
    roofline = {"bound": "issue/latency" , "kernel": "k_jacobi_fused (T Jacobi iterations per pass on a 96x128 register tile, autonomous warps; update = "
                                                     "source pre-pass + fused passes + 1 literal sweep)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "frac_algorithmic": achieved / peak,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL_SWEEP * cells_gpu,
                "ms_per_sweep": ms_sweep, "poisson_share_of_step": ms_poisson / ms_step_eager,
                "traffic": traffic.get("dram_bytes_per_launch") if usable else None, "traffic_detail": traffic,
                "update_schedule": schedule,
                "timed": "CUDA events around pressure_updater.update() inside every timed step (all its launches: source pre-pass, "
                         "fused passes, the literal tail with its sparse BC kernels) / sweeps per update",
                "bound_note": "the 12 B/cell/sweep algorithmic figure stops being a floor under temporal blocking (T iterations per pass over "
                              "HBM): frac_algorithmic > 1 says how far the kernel is beyond what a one-sweep-per-pass kernel could reach; "
                              "frac_dram = ACTUAL dram bytes (ncu, traffic) / launch time / peak says how far it is below the HBM roof. "
                              "It is bounded by instruction issue / latency (12 warps per SM, 156 registers: the register tile holds "
                              "p, t2, t3 of 12288 cells), see issue_active_pct",
                "whole_step_algorithmic_gbs": ALGO_BYTES_PER_CELL_STEP * cells_gpu / (ms_step * 1e-3) / 1e9}
    try:    # tile classes of the fused passes on this rank: {T: [tiles listed, of which slow (BC cells / edges), dropped (nothing to store)]}
        roofline["tile_lists"] = {str(k[0]): list(v[1:]) for k, v in sorted(bc._fused_orders.items()) if k[5:] == (0, 0) and k[1:3] == (bc.dom.r0, bc.dom.r1)}
    except Exception:  # noqa: BLE001
        pass
    if kernel_alone:
        kernel_alone["algorithmic_gbs"] = ALGO_BYTES_PER_CELL_SWEEP * cells_gpu * kernel_alone["T"] / (kernel_alone["us_per_launch"] * 1e-6) / 1e9
        kernel_alone["frac_algorithmic"] = kernel_alone["algorithmic_gbs"] / peak
        if usable and traffic.get("iterations_per_launch") == kernel_alone["T"]:
            kernel_alone["dram_gbs"] = traffic["dram_bytes_per_launch"] / (kernel_alone["us_per_launch"] * 1e-6) / 1e9
            roofline["frac_dram"] = kernel_alone["dram_gbs"] / peak
            roofline["issue_active_pct"] = traffic.get("issue_active_pct")
        roofline["kernel_alone"] = kernel_alone
    if usable and traffic.get("step_dram_bytes"):
        roofline["step_dram_gbs"] = traffic["step_dram_bytes"] / (ms_step * 1e-3) / 1e9
        roofline["step_frac_dram"] = roofline["step_dram_gbs"] / peak

    # ---- end-to-end through the public API with host buffers ----------------------------------
    # fs.pipeline.HostPipelinedStepper: every step uploads the state (v, vx, vy, p) from pinned host memory,
    # runs solver.update() and downloads (v, p); 2 solver instances + 3 streams overlap PCIe with the kernels.
    e2e = None
    if not a.no_e2e:
        from fs.pipeline import HostPipelinedStepper

        host_in = {k: torch.empty_like(getattr(solver, k).current.owned(), device="cpu").pin_memory()
                   for k in ("v", "vx", "vy", "p")}
        for k, t in host_in.items():
            t.copy_(getattr(solver, k).current.owned())
        del solver
        torch.cuda.empty_cache()
        pipe = HostPipelinedStepper(bc, dt, dx, RE, VC, SCHEME, pressure="jacobi", n_iter=a.jacobi)
        e2e_steps = max(4, min(a.steps, 8))
        for _ in range(2):
            pipe.submit(host_in)
        pipe.synchronize()
        barrier()
        t0 = time.perf_counter()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        last = 0
        for _ in range(e2e_steps):
            last = pipe.submit(host_in)
        res = pipe.results(last)                      # host-visible result of the last step
        pipe.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3    # events on 3 streams: time the host-visible completion
        barrier()
        ms = max_over_ranks(wall_ms)
        checksum = float(res["p"][::997, ::991].double().sum())
        e2e = {"value": cells_total * e2e_steps / (ms * 1e-3), "unit": "cell-updates/s",
               "h2d_bytes_per_step": pipe.h2d_bytes * world, "d2h_bytes_per_step": pipe.d2h_bytes * world,
               "steps": e2e_steps, "ms_per_step": ms / e2e_steps, "result_checksum": checksum,
               "what": "fs.pipeline.HostPipelinedStepper: per step pinned-host state (v,vx,vy,p) -> device, "
                       "solver.update(), (v,p) -> pinned host; uploads/downloads overlap the kernels of the "
                       "neighbouring steps (2 solver instances, 3 streams); host wall clock to the last result"}

    # ---- BASELINE configs[1] (bc=2 res=2048 CIP Re=1e4, 80 sweeps) on its own grid, for the record (N=1 only) -------------
    extra = None
    if world == 1 and not a.config and not a.no_extra_config:
        try:
            extra = time_baseline_config(2, 2048, 1e4, 5.0, 80, steps=max(a.steps, 20), warmup=max(a.warmup, 3))
        except Exception as e:  # noqa: BLE001  (never lose the headline line over the side measurement)
            extra = {"error": f"{type(e).__name__}: {e}"[:200]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference(min(a.cpu_sample_rows, a.rows_per_gpu), a.cols, a.jacobi, 2, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
                "stepping": "cuda-graph replay (1 launch/step)" if use_graph else "eager launches",
                "ms_per_step_eager": ms_step_eager, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "baseline_config_2": extra, "value_developed_state": developed, "numa": numa, "setup_s": t_setup, "strip_bounds": list(bounds) if bounds else None,
                "state": "quiescent start (all fields zero, the reference's initial state)" if a.state != "developed" else "developed-like seeded fields",
                "tuning": tuning or None}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
