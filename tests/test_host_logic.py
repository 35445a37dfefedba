"""Host-layer logic on CPU: the Python mirror of the reference API (fs/solver.py, fs/pressure_updater.py,
fs/vorticity_confinement.py, fs/boundary_condition.py, fs/halo.py) driven through tests/fake_fs2d.py, a stand-in for
libfs2d.so that executes every fs2d_* call with the CPU oracle while honouring the fs2d_dom row-window contract.

Checked against `OracleSolver` (the reference's orchestration restated, pinned by tests/test_oracle_golden.py), every
physical buffer, bit for bit:
  * single domain: kernel sequence, write targets and swap counts of every solver x pressure updater x VC x dye combination;
    the sparse BC tables; the Jacobi schedule (fused passes + literal iterations == n literal iterations);
  * the N > 1 path under gloo, world_size 2 and 3: fs/halo.py's distributed operator bodies -- which halo rows are
    exchanged when, the split-phase overlap windows, fused passes across strip edges, rank-consistent schedules.
No statement about the CUDA kernels is made here (see tests/test_gpu_parity.py, tests/mp_strip_check.py for those).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import REPO, assert_bitexact

from oracle import oracle as orc


def _buffers(s) -> dict:
    d = {"v_cur": s.v.current, "v_nxt": s.v.next, "p_cur": s.p.current, "p_nxt": s.p.next}
    if hasattr(s, "vx"):
        d.update(vx_cur=s.vx.current, vx_nxt=s.vx.next, vy_cur=s.vy.current, vy_nxt=s.vy.next)
    if s.vorticity_confinement is not None:
        d.update(vort=s.vorticity_confinement.vorticity, vort_abs=s.vorticity_confinement.vorticity_abs)
    if hasattr(s, "dye"):
        d.update(dye_cur=s.dye.current, dye_nxt=s.dye.next)
        if hasattr(s, "dyex"):
            d.update(dyex_cur=s.dyex.current, dyex_nxt=s.dyex.next, dyey_cur=s.dyey.current, dyey_nxt=s.dyey.next)
    return d


def _seed_state(X, Y, dx, keys, seed, same_p=False):
    """the same seeded global state for the host layer and the oracle (all physical buffers, incl. `.next`)"""
    rng = np.random.default_rng(seed)
    out = {}
    for k in keys:
        c = 3 if k.startswith("dye") else (2 if k[0] == "v" and not k.startswith("vort") else 1)
        shape = (X, Y, c) if c > 1 else (X, Y)
        scale = 0.05 / dx if k[:2] in ("vx", "vy") or k[:4] in ("dyex", "dyey") else (0.5 if c == 2 else 1.0)
        a = (rng.uniform(-1, 1, shape) * scale).astype(np.float32)
        if k == "vort_abs" or k.startswith("dye_"):
            a = np.abs(a)
        out[k] = a
    if same_p:      # equal pressure buffers: their never-written wall cells agree, so fused Jacobi passes are allowed
        out["p_nxt"] = out["p_cur"].copy()
    return out


def _random_scene(seed, X, Y):
    """thin walls, ragged inflow / outflow, stray inflow / outflow cells inside the domain (as in tests/test_gpu_parity.py)"""
    rng = np.random.default_rng(100 + seed)
    mask = np.zeros((X, Y), dtype=np.uint8)
    mask[:, :2] = 1; mask[:, -2:] = 1
    for _ in range(int(rng.integers(10, 20))):
        i, j = int(rng.integers(4, X - 8)), int(rng.integers(2, Y - 6))
        mask[i:i + int(rng.integers(1, 7)), j:j + int(rng.integers(1, 7))] = 1
    mask[:2, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.8, 2, mask[:2, 2:-2])
    mask[-2:, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.7, 3, mask[-2:, 2:-2])
    for _ in range(4):
        mask[int(rng.integers(3, X - 3)), int(rng.integers(3, Y - 3))] = int(rng.integers(2, 4))
    const = np.zeros((X, Y, 2), dtype=np.float32)
    const[mask == 2] = (1.0, 0.0)
    return const, mask


def _oracle_pressure(pkw):
    return ("jacobi", pkw["n_iter"]) if pkw["pressure"] == "jacobi" else ("rbsor", pkw.get("relaxation_factor", 1.3), pkw["n_iter"])


def _build(num, X, Y, scheme, vc, pkw, dye, partition=None):
    from fs.boundary_condition import BoundaryCondition, DyeBoundaryCondition, build_scene
    from fs.fluid_simulator import make_solver

    res = Y
    dt, dx, re = 0.05 / res, 1.0 / res, 1e4
    if isinstance(num, str):            # "rand<seed>": adversarial random mask (walls anywhere, also across strip edges)
        const, mask = _random_scene(int(num[4:]), X, Y)
        bc_dye = None
        bc = BoundaryCondition(const, mask, device="cpu", partition=partition)
    elif dye:
        const, mask, bc_dye = build_scene(num, X, Y, with_dye=True)
        bc = DyeBoundaryCondition(const, bc_dye, mask, device="cpu", partition=partition)
    else:
        const, mask = build_scene(num, X, Y)
        bc_dye = None
        bc = BoundaryCondition(const, mask, device="cpu", partition=partition)
    solver = make_solver(bc, dt, dx, re, vc, scheme, dye=dye, **pkw)
    return solver, (mask, const, bc_dye, dt, dx, re)


SINGLE_CASES = [
    # bc, X, Y, scheme, vc, pressure kwargs, dye, steps
    (2, 64, 32, "cip", 5.0, dict(pressure="jacobi", n_iter=4), False, 3),
    (1, 48, 32, "cip", None, dict(pressure="jacobi", n_iter=3), False, 3),        # 2 swaps/step: v.current stays buffer A (T1)
    (3, 80, 48, "cip", 10.0, dict(pressure="rbsor", n_iter=2), False, 3),
    (5, 64, 32, "kk", 5.0, dict(pressure="rbsor", n_iter=2), False, 3),
    (4, 48, 32, "upwind", None, dict(pressure="jacobi", n_iter=5), False, 3),
    (2, 256, 128, "cip", 5.0, dict(pressure="jacobi", n_iter=21), False, 2),      # Y % 16 == 0: fused passes in the schedule
    (1, 48, 32, "cip", 5.0, dict(pressure="rbsor", n_iter=2), True, 3),           # main.py's default object graph (dye on)
    (2, 64, 32, "upwind", 5.0, dict(pressure="jacobi", n_iter=2), True, 3),
    (5, 64, 32, "kk", None, dict(pressure="rbsor", n_iter=3), True, 2),
]


@pytest.mark.parametrize("case", SINGLE_CASES, ids=lambda c: f"bc{c[0]}_{c[3]}_{c[5]['pressure']}{c[5]['n_iter']}_vc{c[4]}_dye{int(c[6])}")
def test_single_domain_host_layer_equals_reference_orchestration(case, monkeypatch):
    from fake_fs2d import FakeFs2d
    from fs import _lib

    num, X, Y, scheme, vc, pkw, dye, steps = case
    fake = FakeFs2d(_lib.load()).install(monkeypatch)
    solver, (mask, const, bc_dye, dt, dx, re) = _build(num, X, Y, scheme, vc, pkw, dye)
    ref = orc.OracleSolver(mask, const, dt, dx, re, scheme, vc, _oracle_pressure(pkw), bc_dye=bc_dye)
    state = _seed_state(X, Y, dx, list(_buffers(solver)), seed=100 + (num if isinstance(num, int) else 50 + int(num[4:])))
    for k, f in _buffers(solver).items():
        f.from_numpy(state[k])
    ref.load_state({k: v for k, v in state.items()})
    for _ in range(steps):
        solver.update()
        ref.update()
    want = ref.state()
    for k, f in _buffers(solver).items():
        assert_bitexact(f"{k} after {steps} steps", f.to_numpy(), want[k])
    # the reference's kernel sequence of one step, pure host helpers (plan / tile queries) left out
    kern = [c for c in fake.trace if c not in ("fs2d_jacobi_plan", "fs2d_fused_tile")]
    if scheme == "cip":
        assert kern[0] == "fs2d_set_grad"                           # CipMacSolver.__init__ (fs/solver.py:190)
        kern = [c for c in kern if c not in ("fs2d_set_grad", "fs2d_dye_set_grad")]
    assert len(kern) % steps == 0
    step = kern[:len(kern) // steps]
    assert kern == step * steps
    head = (["fs2d_vel_bc", "fs2d_cip_nonadv", "fs2d_cip_nonadv_grad", "fs2d_cip_advect"] if scheme == "cip"
            else ["fs2d_vel_bc", "fs2d_mac_update"]) + (["fs2d_vort_apply"] if vc is not None else [])
    assert step[:len(head)] == head and step[len(head)] == "fs2d_pressure_source"
    v_part = step[:step.index("fs2d_limit") + 1]
    if pkw["pressure"] == "jacobi":
        assert v_part[len(head) + 1:] == ["fs2d_jacobi_update", "fs2d_limit"]
    else:
        assert v_part[len(head) + 1:] == ["fs2d_pressure_bc", "fs2d_rbsor_iteration"] * pkw["n_iter"] + ["fs2d_limit"]
    dye_part = step[len(v_part):]
    if not dye:
        assert dye_part == []
    elif scheme == "cip":
        assert dye_part == ["fs2d_dye_bc", "fs2d_dye_nonadv", "fs2d_dye_nonadv_grad", "fs2d_dye_cip_advect", "fs2d_clamp"]
    else:
        assert dye_part == ["fs2d_dye_bc", "fs2d_dye_mac", "fs2d_clamp"]
    if pkw["pressure"] == "jacobi" and X >= 256:
        # the two pressure buffers were seeded with DIFFERENT random values, so their never-written wall cells disagree
        # and the host layer must refuse fused passes for them (DESIGN.md "stale cells"; include/fs2d.h fs2d_jacobi_fused)
        assert solver._bc._exposed_stale.numel() > 0 and solver.pressure_updater.fuse_mask(solver.p) == 0


def test_swap_counts_follow_the_reference(monkeypatch):
    """CIP without VC swaps v twice per step (v.current is always physical buffer A); with VC three times
    (A/B alternate); p swaps n_iter times per update."""
    from fake_fs2d import FakeFs2d
    from fs import _lib

    FakeFs2d(_lib.load()).install(monkeypatch)
    for vc, n_iter in ((None, 3), (5.0, 3), (5.0, 4)):
        solver, _ = _build(2, 32, 16, "cip", vc, dict(pressure="jacobi", n_iter=n_iter), False)
        v_a, p_a = solver.v.current, solver.p.current
        solver.update()
        assert (solver.v.current is v_a) == (vc is None)
        assert (solver.p.current is p_a) == (n_iter % 2 == 0)
        solver.update()
        assert solver.v.current is v_a and solver.p.current is p_a


@pytest.mark.parametrize("tail", [1, 0])
def test_jacobi_schedule_with_fused_passes_equals_literal_iterations(monkeypatch, tail):
    """jacobi_update_distributed's building blocks on one domain: running the plan entry by entry (fused passes through
    fs2d_jacobi_fused, split into interior + edge launches, literal iterations in between) == fs2d_jacobi_update.
    tail 1 (default): the plan ends {emitting fused pass, one literal iteration}; 0: two literal iterations."""
    from fake_fs2d import FakeFs2d
    from fs import _lib
    from fs.halo import split_windows

    FakeFs2d(_lib.load()).install(monkeypatch)
    _lib.load().fs2d_set_tuning(4, tail)
    monkeypatch.setattr(_lib, "_tail_restore", None, raising=False)
    X, Y, n_iter = 512, 64, 19
    solver, _ = _build(2, X, Y, "cip", None, dict(pressure="jacobi", n_iter=n_iter), False)
    jac, bc = solver.pressure_updater, solver._bc
    rng = np.random.default_rng(3)
    solver.v.current.from_numpy((rng.uniform(-1, 1, (X, Y, 2)) * 0.5).astype(np.float32))
    p0 = rng.uniform(-1, 1, (X, Y)).astype(np.float32)
    for f in (solver.p.current, solver.p.next):
        f.from_numpy(p0)
    plan = jac.plan(solver.p)
    _lib.load().fs2d_set_tuning(4, 1)     # (the schedule is in hand; leave the library in its default state)
    assert sum(t if t else 1 for t in plan) == n_iter and any(t > 0 for t in plan)
    assert plan[-1] == 0 and (plan[-2] > 0) == bool(tail)
    tail_at = len(plan) - 2 if tail else -1
    jac.update(solver.p, solver.v.current)
    want = {k: _buffers(solver)[k].to_numpy().copy() for k in ("p_cur", "p_nxt")}
    for f in (solver.p.current, solver.p.next):
        f.from_numpy(p0)
    # replay by hand, the way a strip does it
    import ctypes

    p = solver.p
    a0 = p.current
    src = jac._source(solver.v.current)
    for k, t in enumerate(plan):
        if t > 0:
            rows, hr = ctypes.c_int(), ctypes.c_int()
            _lib.load().fs2d_fused_tile(t, ctypes.byref(rows), None, ctypes.byref(hr), None, None)
            mid, m = split_windows(bc.dom, rows.value - 2 * hr.value, t)
            assert mid is not None
            jac._fused(p.next, p.current, src, t, dom=mid, emit=k == tail_at)
            jac._fused(p.next, p.current, src, t, skip=(1, m - 1), emit=k == tail_at)
        else:
            bc.set_pressure_boundary_condition(p.current)
            jac._sweep(p.next, p.current, src, inline_bc=False)
        p.swap()
    assert (p.current is a0) == (len(plan) % 2 == 0)
    assert_bitexact("p_cur", p.current.to_numpy(), want["p_cur"])
    assert_bitexact("p_nxt", p.next.to_numpy(), want["p_nxt"])


# ---------------------------------------------------------------------------------------------------------------------
# N > 1: row strips under gloo
# ---------------------------------------------------------------------------------------------------------------------
STRIP_CASES = [
    # bc, X, Y, scheme, vc, pressure kwargs, dye, steps, halo
    (2, 96, 32, "cip", 5.0, dict(pressure="jacobi", n_iter=4), False, 3, 2),
    (3, 120, 48, "cip", 10.0, dict(pressure="rbsor", n_iter=2), False, 2, 3),
    (5, 96, 32, "kk", 5.0, dict(pressure="rbsor", n_iter=2), False, 3, 2),
    (1, 96, 32, "upwind", None, dict(pressure="jacobi", n_iter=3), False, 3, 4),
    (2, 640, 64, "cip", 5.0, dict(pressure="jacobi", n_iter=13), False, 2, 9),     # fused passes across the strip edges, overlap windows
    (2, 640, 64, "cip", 5.0, dict(pressure="jacobi", n_iter=30), False, 2, 20),    # deep halo: several fused passes share one exchange
    (1, 96, 32, "cip", 5.0, dict(pressure="rbsor", n_iter=2), True, 2, 2),         # dye (CIP) on strips
    (4, 96, 32, "upwind", 5.0, dict(pressure="jacobi", n_iter=2), True, 2, 2),     # dye (upwind) on strips
] + [(f"rand{seed}", 60, 48, scheme, vc, pkw, False, 2, halo) for seed in range(4) for scheme, vc, pkw, halo in (
    ("cip", 5.0, dict(pressure="jacobi", n_iter=5), 2), ("kk", None, dict(pressure="rbsor", n_iter=2), 3),
    ("cip", 5.0, dict(pressure="jacobi", n_iter=11), 5))]     # walls, inflow and outflow cells on and next to the strip edges


def _strip_worker(rank: int, world: int, port: int, q, tuning=()) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for p in (str(REPO), str(REPO / "2d-fluid-simulator_b200"), str(REPO / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from fake_fs2d import FakeFs2d
    from fs import _lib
    from fs.distributed import Partition
    from fs.halo import exchanger_for, gather_owned

    orc.set_threads(1)
    fake = FakeFs2d(_lib.load()).install_plain()
    for key, value in tuning:
        _lib.call("fs2d_set_tuning", key, value)
    failures, info = [], []
    for num, X, Y, scheme, vc, pkw, dye, steps, halo in STRIP_CASES:
        part = Partition(X, rank, world, halo)
        solver, (mask, const, bc_dye, dt, dx, re) = _build(num, X, Y, scheme, vc, pkw, dye, partition=part)
        state = _seed_state(X, Y, dx, list(_buffers(solver)), seed=200 + (num if isinstance(num, int) else 50 + int(num[4:])), same_p=pkw["n_iter"] > 8)
        g0, g1 = part.owned()
        for k, f in _buffers(solver).items():
            f.from_numpy(state[k][g0:g1])
        n0 = len(fake.trace)
        for _ in range(steps):
            solver.update()
        hx = exchanger_for(solver._bc)
        info.append((num, scheme, hx.n_exchanges, sum(1 for c in fake.trace[n0:] if c.startswith("fs2d_jacobi_fused")),
                     sum(1 for c in fake.trace[n0:] if c == "fs2d_jacobi_fused_tail")))
        ref = None
        if rank == 0:
            ref = orc.OracleSolver(mask, const, dt, dx, re, scheme, vc, _oracle_pressure(pkw), bc_dye=bc_dye)
            ref.load_state(state)
            for _ in range(steps):
                ref.update()
        for k, f in _buffers(solver).items():
            got = gather_owned(f, part)
            if rank == 0:
                try:
                    assert_bitexact(f"bc{num} {scheme} {pkw} dye={dye} world={world}: {k}", got.numpy(), ref.state()[k])
                except AssertionError as e:
                    failures.append(str(e))
        dist.barrier()
    q.put((rank, failures, info))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_strips_equal_single_domain_under_gloo(world):
    from test_distributed import free_port

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_strip_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for rank, failures, info in res:
        assert not failures, "\n".join(failures[:5])
    info = res[0][2]
    assert all(i[2] > 0 for i in info)                             # every case really exchanged halos
    assert info[4][3] > 0 and info[4][4] > 0                        # the 13-iteration case ran fused passes on the strips, the last one emitting


def test_strips_with_the_two_literal_tail_under_gloo():
    """fs2d_set_tuning(4, 0) on strips: the schedule ends with two literal iterations instead of the default {emitting fused
    pass, one literal iteration}; every physical buffer -- the wall cells of both pressure buffers included -- still equals
    the reference's orchestration."""
    from test_distributed import free_port

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_strip_worker, args=(r, world, port, q, ((4, 0),))) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for rank, failures, info in res:
        assert not failures, "\n".join(failures[:5])
    assert all(i[4] == 0 for i in res[0][2])     # no emitting pass anywhere


# ---------------------------------------------------------------------------------------------------------------------
# facade (fs/fluid_simulator.py): the reference's create()/step()/field_to_numpy() contract and the state dump
# ---------------------------------------------------------------------------------------------------------------------
def test_facade_create_step_dump_and_resume(monkeypatch):
    """FluidSimulator.create(num, resolution, dt, dx, re, vor_eps, scheme) builds the reference's default object graph
    (RB-SOR omega=1.3 x2, fs/fluid_simulator.py:76-78); field_to_numpy() is the `d`-key dump (main.py:129-132);
    state_dict()/load_state_dict() resume a CIP run bit-identically (every physical buffer)."""
    from fake_fs2d import FakeFs2d
    from fs import _lib
    from fs.boundary_condition import build_scene
    from fs.fluid_simulator import DyeFluidSimulator, FluidSimulator
    from fs.pressure_updater import RedBlackSorPressureUpdater
    from fs.solver import CipMacSolver, DyeMacSolver

    FakeFs2d(_lib.load()).install(monkeypatch)
    res = 24
    dt, dx, re = 0.05 / res, 1.0 / res, 1e4
    sim = FluidSimulator.create(1, res, dt, dx, re, 5.0, "cip", device="cpu")
    assert isinstance(sim.solver, CipMacSolver) and isinstance(sim.solver.pressure_updater, RedBlackSorPressureUpdater)
    assert sim.solver.pressure_updater._n_iter == 2 and sim.solver.pressure_updater._relaxation_factor == 1.3
    assert sim.solver.resolution == (2 * res, res)
    const, mask = build_scene(1, 2 * res, res)
    ref = orc.OracleSolver(mask, const, dt, dx, re, "cip", 5.0, ("rbsor", 1.3, 2))
    for _ in range(3):
        sim.step()
        ref.update()
    out = sim.field_to_numpy()
    assert sorted(out) == ["p", "v"] and out["v"].shape == (2 * res, res, 2) and out["p"].dtype == np.float32
    assert_bitexact("v", out["v"], ref.v.current)
    assert_bitexact("p", out["p"], ref.p.current)
    # dump, continue, restore into a fresh simulator, continue: identical
    saved = {k: a.copy() for k, a in sim.state_dict().items()}
    for _ in range(2):
        sim.step()
    want = sim.state_dict()
    sim2 = FluidSimulator.create(1, res, dt, dx, re, 5.0, "cip", device="cpu")
    sim2.load_state_dict(saved)
    for _ in range(2):
        sim2.step()
    got = sim2.state_dict()
    assert sorted(got) == sorted(want)
    for k in want:
        assert_bitexact(f"resumed {k}", got[k], want[k])
    # dye facade: three fields in the dump, MAC solver for the non-CIP schemes
    dsim = DyeFluidSimulator.create(2, res, dt, dx, re, None, "upwind", device="cpu")
    assert isinstance(dsim.solver, DyeMacSolver) and dsim.solver.vorticity_confinement is None
    dsim.step()
    d = dsim.field_to_numpy()
    assert sorted(d) == ["dye", "p", "v"] and d["dye"].shape == (2 * res, res, 3)
    # render getters (fs/fluid_simulator.py:22-32, :111-119): (X, Y, 3) f32 images in rgb_buf, wall cells in the wall colour
    mask2 = build_scene(2, 2 * res, res)[1]
    v2, p2, dye2 = (f.to_numpy() for f in dsim.solver.get_fields())
    for getter, mode in ((dsim.get_norm_field, "norm"), (dsim.get_pressure_field, "pressure"),
                         (dsim.get_vorticity_field, "vorticity"), (dsim.get_dye_field, "dye")):
        img = getter()
        assert img is dsim.rgb_buf and img.to_numpy().shape == (2 * res, res, 3)
        assert_bitexact(f"render {mode}", img.to_numpy(), orc.render(v2, p2, dye2, mask2, dx, mode))
        assert np.array_equal(img.to_numpy()[mask2 == 1], np.broadcast_to(np.float32([0.5, 0.7, 0.5]), ((mask2 == 1).sum(), 3)))
    # error behaviour of the reference (fs/fluid_simulator.py:104-106, fs/boundary_condition.py:216-217)
    with pytest.raises(ValueError, match="Unknown scheme: bogus"):
        FluidSimulator.create(1, res, dt, dx, re, None, "bogus", device="cpu")
    with pytest.raises(NotImplementedError):
        FluidSimulator.create(9, res, dt, dx, re, None, "cip", device="cpu")


def test_constructor_injection_like_the_reference(monkeypatch):
    """The reference's plugin API is constructor injection: MacSolver(bc, pressure_updater, advect_function,
    dt, dx, re, vorticity_confinement) with fs.advection.advect_kk_scheme etc.; a user-defined PressureUpdater works too."""
    from fake_fs2d import FakeFs2d
    from fs import _lib
    from fs.advection import advect_kk_scheme, advect_upwind
    from fs.boundary_condition import get_boundary_condition
    from fs.pressure_updater import JacobiPressureUpdater, PressureUpdater
    from fs.solver import MacSolver
    from fs.vorticity_confinement import VorticityConfinement

    FakeFs2d(_lib.load()).install(monkeypatch)
    res = 16
    dt, dx, re = 0.05 / res, 1.0 / res, 100.0
    bc = get_boundary_condition(1, res, enable_dye=False, device="cpu")
    assert bc.get_resolution() == (2 * res, res)
    vc = VorticityConfinement(bc, dt, dx, 2.5)
    solver = MacSolver(bc, JacobiPressureUpdater(bc, dt, dx, n_iter=3), advect_kk_scheme, dt, dx, re, vc)
    solver.update()
    v, p = solver.get_fields()
    assert v.to_numpy().shape == (2 * res, res, 2) and p.to_numpy().shape == (2 * res, res)
    assert solver.is_wall(0, 0) and not solver.is_fluid_domain(0, 0) and solver.is_fluid_domain(res, res // 4)

    calls = []

    class CountingUpdater(PressureUpdater):        # a user plugin: called once per step with (p: DoubleBuffer, v_current)
        def update(self, p, v_current) -> None:
            calls.append((p, v_current))

    s2 = MacSolver(bc, CountingUpdater(bc, dt, dx), advect_upwind, dt, dx, re)
    s2.update(); s2.update()
    assert len(calls) == 2 and calls[0][0] is s2.p and calls[1][1] is s2.v.current
    with pytest.raises(TypeError):
        MacSolver(bc, CountingUpdater(bc, dt, dx), lambda *a: None, dt, dx, re)
