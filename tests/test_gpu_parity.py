"""GPU parity tests proper: the CUDA path (through the fs API -> ctypes -> C ABI of libfs2d.so) vs
the CPU oracle and vs the golden fixtures generated from the reference source.

Bar: BIT-EXACT fp32 (the library is compiled with -fmad=false and keeps the reference's literal
operation order), NaNs compare equal.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch
from conftest import GOLDEN, assert_bitexact

pytestmark = pytest.mark.gpu

BCS = (1, 2, 3, 4, 5)


@pytest.fixture(scope="module")
def env():
    from fs import _lib

    lib = _lib.load()
    assert lib.fs2d_device_ok(), lib.fs2d_last_error().decode()
    return lib


def fld(a: np.ndarray):
    from fs.double_buffer import Field

    f = Field(a.shape[:2], a.shape[2] if a.ndim == 3 else 1)
    f.from_numpy(a)
    return f


def make_fs(mask, const, dt, dx, re, scheme, vc, pressure):
    from fs.advection import advect_kk_scheme, advect_upwind
    from fs.boundary_condition import BoundaryCondition
    from fs.pressure_updater import JacobiPressureUpdater, RedBlackSorPressureUpdater
    from fs.solver import CipMacSolver, MacSolver
    from fs.vorticity_confinement import VorticityConfinement

    bc = BoundaryCondition(const, mask)
    vcf = VorticityConfinement(bc, dt, dx, vc) if vc is not None else None
    pu = (JacobiPressureUpdater(bc, dt, dx, pressure[1]) if pressure[0] == "jacobi"
          else RedBlackSorPressureUpdater(bc, dt, dx, pressure[1], pressure[2]))
    if scheme == "cip":
        return CipMacSolver(bc, pu, dt, dx, re, vcf)
    return MacSolver(bc, pu, advect_upwind if scheme == "upwind" else advect_kk_scheme, dt, dx, re, vcf)


def fs_state(s) -> dict:
    d = {"v_cur": s.v.current, "v_nxt": s.v.next, "p_cur": s.p.current, "p_nxt": s.p.next}
    if hasattr(s, "vx"):
        d.update(vx_cur=s.vx.current, vx_nxt=s.vx.next, vy_cur=s.vy.current, vy_nxt=s.vy.next)
    if s.vorticity_confinement is not None:
        d.update(vort=s.vorticity_confinement.vorticity, vort_abs=s.vorticity_confinement.vorticity_abs)
    return d


def load_fs_state(s, st: dict) -> None:
    for k, f in fs_state(s).items():
        f.from_numpy(st[k])


# ------------------------------------------------------------------------------------------------
# 1. every kernel, one launch, on the reference-generated fixtures
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num", BCS)
def test_each_kernel_matches_reference_fixture(env, num, kernels_golden, masks_small):
    from fs.solver import limit_field

    g = kernels_golden
    res, dt, dx, re, vcw = g["meta"]
    res = int(res)
    mask, const = masks_small[f"bc{num}_r{res}_mask"], masks_small[f"bc{num}_r{res}_const"]
    G = lambda k: g[f"bc{num}/{k}"]  # noqa: E731
    cip = make_fs(mask, const, dt, dx, re, "cip", vcw, ("jacobi", 1))
    bc, vcf, jac = cip._bc, cip.vorticity_confinement, cip.pressure_updater
    sor = make_fs(mask, const, dt, dx, re, "cip", None, ("rbsor", 1.3, 1)).pressure_updater

    f = fld(G("v")); bc.set_velocity_boundary_condition(f); assert_bitexact("vel_bc", f.to_numpy(), G("vel_bc"))
    f = fld(G("p")); bc.set_pressure_boundary_condition(f); assert_bitexact("p_bc", f.to_numpy(), G("p_bc"))
    for name in ("upwind", "kk"):
        s = make_fs(mask, const, dt, dx, re, name, None, ("jacobi", 1))
        f = fld(G("vn0")); s._update_velocities(f, fld(G("v")), fld(G("p")))
        assert_bitexact("mac_" + name, f.to_numpy(), G("mac_" + name))
    fn = fld(G("vn0")); cip._non_advection_phase(fn, fld(G("v")), fld(G("p")))
    assert_bitexact("nonadv", fn.to_numpy(), G("nonadv"))
    fxn, fyn = fld(G("vxn0")), fld(G("vyn0"))
    cip._non_advection_phase_grad(fxn, fyn, fld(G("vx")), fld(G("vy")), fld(G("v")), fn)
    assert_bitexact("nonadv_gx", fxn.to_numpy(), G("nonadv_gx")); assert_bitexact("nonadv_gy", fyn.to_numpy(), G("nonadv_gy"))
    a, b, c, fv = fld(G("vn0")), fld(G("vxn0")), fld(G("vyn0")), fld(G("v"))
    cip._advection_phase(a, b, c, fv, fld(G("vx")), fld(G("vy")), fv)
    assert_bitexact("cip_f", a.to_numpy(), G("cip_f")); assert_bitexact("cip_fx", b.to_numpy(), G("cip_fx"))
    assert_bitexact("cip_fy", c.to_numpy(), G("cip_fy"))
    a, b = fld(G("vxn0")), fld(G("vyn0")); cip._set_grad(a, b, fld(G("v")))
    assert_bitexact("grad_x", a.to_numpy(), G("grad_x")); assert_bitexact("grad_y", b.to_numpy(), G("grad_y"))
    vcf.vorticity.from_numpy(G("w0")); vcf.vorticity_abs.from_numpy(G("wa0"))
    vcf._calc_vorticity(fld(G("v")))
    assert_bitexact("vort", vcf.vorticity.to_numpy(), G("vort")); assert_bitexact("vort_abs", vcf.vorticity_abs.to_numpy(), G("vort_abs"))
    f = fld(G("vn0")); vcf._add_vorticity(f, fld(G("v"))); assert_bitexact("vort_add", f.to_numpy(), G("vort_add"))
    # the fused apply() kernel (calc + add in one pass) must leave the same three fields
    vcf.vorticity.from_numpy(G("w0")); vcf.vorticity_abs.from_numpy(G("wa0"))
    f = fld(G("vn0")); vcf._apply_fused(f, fld(G("v")))
    assert_bitexact("fused vort", vcf.vorticity.to_numpy(), G("vort")); assert_bitexact("fused vort_abs", vcf.vorticity_abs.to_numpy(), G("vort_abs"))
    assert_bitexact("fused vort_add", f.to_numpy(), G("vort_add"))
    f = fld(G("pn0")); jac._update(f, fld(G("p")), fld(G("v"))); assert_bitexact("jacobi", f.to_numpy(), G("jacobi"))
    f = fld(G("pn0")); sor._update(f, fld(G("p")), fld(G("v"))); assert_bitexact("rbsor", f.to_numpy(), G("rbsor"))
    f = fld(G("v") * np.float32(12.0)); limit_field(f, 10.0); assert_bitexact("limit", f.to_numpy(), G("limit"))


# ------------------------------------------------------------------------------------------------
# 2. N-step trajectories through Solver.update(): every physical buffer, vs the reference fixtures
# ------------------------------------------------------------------------------------------------
TRAJ = sorted(p.name for p in GOLDEN.glob("traj_[A-Z]_*.npz"))


@pytest.mark.parametrize("name", TRAJ)
def test_trajectory_matches_reference_fixture(env, name, masks_small):
    g = np.load(GOLDEN / name)
    num, res, vc = int(g["meta_num"]), int(g["meta_res"]), float(g["meta_vc"])
    kind, n_iter = str(g["meta_pressure"]), int(g["meta_n_iter"])
    pressure = ("jacobi", n_iter) if kind == "jacobi" else ("rbsor", 1.3, n_iter)
    s = make_fs(masks_small[f"bc{num}_r{res}_mask"], masks_small[f"bc{num}_r{res}_const"], float(g["meta_dt"]),
                float(g["meta_dx"]), float(g["meta_re"]), str(g["meta_scheme"]), None if vc < 0 else vc, pressure)
    load_fs_state(s, {k[3:]: g[k] for k in g.files if k.startswith("s0_")})
    for n in range(1, int(g["meta_steps"]) + 1):
        s.update()
        for k, f in fs_state(s).items():
            assert_bitexact(f"{name} step {n} {k}", f.to_numpy(), g[f"s{n}_{k}"])


# ------------------------------------------------------------------------------------------------
# 3. BASELINE configs at sizes the oracle finishes in seconds, vs the oracle
# ------------------------------------------------------------------------------------------------
CONFIGS = [
    # id, bc, res, dt, re, scheme, vc, pressure, steps
    ("cfg1_as_given", 1, 128, 0.005, 100.0, "upwind", 5.0, ("jacobi", 40), 3),       # unstable beyond 3 steps
    ("cfg1_stable", 1, 128, 0.0005, 100.0, "upwind", 5.0, ("jacobi", 40), 30),
    ("cfg2_r256", 2, 256, None, 1e4, "cip", 5.0, ("jacobi", 80), 6),
    ("cfg3_r256", 3, 256, None, 1e8, "cip", 10.0, ("jacobi", 100), 4),
    ("cfg5_r192", 5, 192, None, 1e6, "cip", 5.0, ("jacobi", 21), 5),               # odd n_iter, Y % 64 != 0
    ("kk_r100", 1, 100, None, 1000.0, "kk", 5.0, ("rbsor", 1.3, 2), 6),            # create() defaults, non-pow2 dx
    ("cip_r50_y_not_mult4", 4, 50, None, 1e4, "cip", None, ("jacobi", 7), 4),      # scalar Jacobi path (Y % 4 != 0)
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_config_trajectory_vs_oracle(env, cfg):
    from fs.boundary_condition import build_scene
    from oracle import oracle as orc

    _, num, res, dt, re, scheme, vc, pressure, steps = cfg
    dt = dt if dt is not None else 0.05 / res
    dx = 1.0 / res
    const, mask = build_scene(num, 2 * res, res)
    s = make_fs(mask, const, dt, dx, re, scheme, vc, pressure)
    ref = orc.OracleSolver(mask, const, dt, dx, re, scheme, vc, pressure)
    for n in range(1, steps + 1):
        s.update()
        ref.update()
        if n in (1, steps):
            got = fs_state(s)
            for k, a in ref.state().items():
                assert_bitexact(f"{cfg[0]} step {n} {k}", got[k].to_numpy(), a)


def test_facade_matches_reference_defaults(env):
    """FluidSimulator.create(...) == reference object graph (RB-SOR 1.3 x2) and the npz dump format."""
    from fs.boundary_condition import build_scene
    from fs.fluid_simulator import FluidSimulator
    from oracle import oracle as orc

    res = 64
    sim = FluidSimulator.create(5, res, 0.05 / res, 1.0 / res, 1e6, 5.0, "cip")
    const, mask = build_scene(5, 2 * res, res)
    ref = orc.OracleSolver(mask, const, 0.05 / res, 1.0 / res, 1e6, "cip", 5.0, ("rbsor", 1.3, 2))
    for _ in range(5):
        sim.step(); ref.update()
    out = sim.field_to_numpy()
    assert set(out) == {"v", "p"} and out["v"].shape == (2 * res, res, 2) and out["p"].shape == (2 * res, res)
    assert_bitexact("v", out["v"], ref.v.current); assert_bitexact("p", out["p"], ref.p.current)
    with pytest.raises(ValueError, match="Unknown scheme"):
        FluidSimulator.create(1, 32, 0.001, 1 / 32, 100.0, None, "weno")


# ------------------------------------------------------------------------------------------------
# 4. size-independent properties at BASELINE sizes + strip invariance of the dom row ranges
# ------------------------------------------------------------------------------------------------
def test_jacobi_row_range_invariance_and_literal_equivalence(env, res=512):
    """(a) sweeping rows [0,h) then [h,X) == one full sweep; (b) inline-BC sweep == in-place BC then a
    plain sweep; (c) fs2d_jacobi_update(n) == n literal reference iterations -- all bitwise, res=512
    (tests/test_kernels_emulated.py runs the same body at a smaller res)."""
    from fs import _lib
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.pressure_updater import JacobiPressureUpdater

    const, mask = build_scene(3, 2 * res, res)
    bc = BoundaryCondition(const, mask)
    rng = np.random.default_rng(5)
    p0 = rng.uniform(-1, 1, mask.shape).astype(np.float32)
    v = fld(rng.uniform(-1, 1, mask.shape + (2,)).astype(np.float32))
    dt, dx = 0.05 / res, 1.0 / res
    jac = JacobiPressureUpdater(bc, dt, dx, 9)
    full, halves = fld(p0 * 3), fld(p0 * 3)
    pc = fld(p0)
    src = jac._source(v)
    jac._sweep(full, pc, src, inline_bc=True)
    X = mask.shape[0]
    for r0, r1 in ((0, X // 3), (X // 3, X)):
        jac._sweep(halves, pc, src, inline_bc=True, dom=bc.dom.replace(r0=r0, r1=r1))
    assert_bitexact("row ranges", halves.to_numpy(), full.to_numpy())
    lit_in = fld(p0); bc.set_pressure_boundary_condition(lit_in)
    lit = fld(p0 * 3); jac._update(lit, lit_in, v)
    assert_bitexact("inline vs literal", full.to_numpy(), lit.to_numpy())
    # n sweeps: C loop vs literal Python loop, both physical buffers
    from fs.double_buffer import DoubleBuffer
    a, b = DoubleBuffer(mask.shape, 1), DoubleBuffer(mask.shape, 1)
    for db in (a, b):
        db.current.from_numpy(p0); db.next.from_numpy(p0[::-1].copy())
    jac.update(a, v)
    for _ in range(9):
        bc.set_pressure_boundary_condition(b.current); jac._update(b.next, b.current, v); b.swap()
    assert_bitexact("update cur", a.current.to_numpy(), b.current.to_numpy())
    assert_bitexact("update nxt", a.next.to_numpy(), b.next.to_numpy())


@pytest.mark.parametrize("num,res", [(2, 2048), (3, 1024)])
def test_full_size_step_vs_oracle(env, num, res):
    """BASELINE config 2 at its full size (and config 3's scene at res 1024): 2 CIP+VC steps with a
    reduced sweep count, every buffer bit-compared with the oracle (seconds of CPU time)."""
    from fs.boundary_condition import build_scene
    from oracle import oracle as orc

    dt, dx, re, vc, pressure = 0.05 / res, 1.0 / res, 1e4, 5.0, ("jacobi", 6)
    const, mask = build_scene(num, 2 * res, res)
    s = make_fs(mask, const, dt, dx, re, "cip", vc, pressure)
    ref = orc.OracleSolver(mask, const, dt, dx, re, "cip", vc, pressure)
    for _ in range(2):
        s.update(); ref.update()
    got = fs_state(s)
    for k, a in ref.state().items():
        assert_bitexact(f"bc{num} res{res} {k}", got[k].to_numpy(), a)


def test_properties_at_res_8192(env):
    """Size-independent properties on the res=8192 grid (16384 x 8192, BASELINE config 4):
    zero velocity => CIP advection is the identity; constant p & zero v is a Jacobi fixed point on
    cells whose stencil sees no outflow BC; limiter bounds |v| and leaves slow cells untouched; checksum of a strip-wise run ==
    checksum of the whole-grid run."""
    from fs import _lib
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.double_buffer import Field
    from fs.pressure_updater import JacobiPressureUpdater
    from fs.solver import limit_field

    res = 8192
    dt, dx = 0.05 / res, 1.0 / res
    const, mask = build_scene(2, 2 * res, res)
    bc = BoundaryCondition(const, mask)
    X, Y = mask.shape
    g = torch.Generator(device="cuda").manual_seed(1)

    def rnd(ch):
        f = Field((X, Y), ch)
        f.tensor.copy_(torch.rand(f.tensor.shape, device="cuda", generator=g) * 2 - 1)
        return f

    # CIP identity at zero velocity
    f, fx, fy = rnd(2), rnd(2), rnd(2)
    fn, fxn, fyn, zero = Field((X, Y), 2), Field((X, Y), 2), Field((X, Y), 2), Field((X, Y), 2)
    _lib.call("fs2d_cip_advect", fn.ptr(), fxn.ptr(), fyn.ptr(), f.ptr(), fx.ptr(), fy.ptr(), zero.ptr(),
              _lib.ptr(bc._bc_mask), bc.dom, dt, dx, dx**2, dx**3, _lib.stream())
    fluid = (bc._bc_mask == 0).unsqueeze(-1)
    for out, src in ((fn, f), (fxn, fx), (fyn, fy)):
        assert torch.equal(torch.where(fluid, out.tensor, 0), torch.where(fluid, src.tensor, 0))
        assert torch.count_nonzero(torch.where(fluid, 0, out.tensor)) == 0  # non-fluid cells untouched
    del fn, fxn, fyn, fx, fy
    # limiter idempotence
    f.tensor.mul_(30.0)
    before = f.tensor.clone()
    limit_field(f, 10.0)
    nrm, nrm0 = f.tensor.norm(dim=-1), before.norm(dim=-1)
    assert float(nrm.max()) <= 10.0 * (1 + 1e-6)
    small = nrm0 <= 10.0 * (1 - 1e-6)
    assert torch.equal(f.tensor[small], before[small])  # cells under the limit are untouched
    del before, nrm, nrm0, small, f
    # Jacobi: strip-wise == whole, and fixed point
    jac = JacobiPressureUpdater(bc, dt, dx, 1)
    pc, v = rnd(1), rnd(2)
    whole, strips = Field((X, Y), 1), Field((X, Y), 1)
    src = jac._source(v)
    jac._sweep(whole, pc, src, inline_bc=True)
    for k in range(8):
        jac._sweep(strips, pc, src, inline_bc=True, dom=bc.dom.replace(r0=k * X // 8, r1=(k + 1) * X // 8))
    assert torch.equal(whole.tensor, strips.tensor)
    pc.tensor.fill_(0.75); v.tensor.zero_()
    jac._sweep(whole, pc, jac._source(v), inline_bc=True)
    interior = torch.zeros_like(bc._bc_mask, dtype=torch.bool)
    interior[4:X - 4] = True
    ok = (bc._bc_mask == 0) & interior
    assert torch.equal(whole.tensor[ok], torch.full_like(whole.tensor[ok], 0.75))


# ------------------------------------------------------------------------------------------------
# 5. fused multi-iteration Jacobi (temporal blocking) == literal iterations, bitwise
# ------------------------------------------------------------------------------------------------
FUSED_CASES = [(1, 128, 64), (2, 256, 128), (3, 320, 160), (4, 200, 96), (5, 384, 192), (2, 1000, 512), (3, 2048, 1024)]


@pytest.mark.parametrize("listed", [True, False])
@pytest.mark.parametrize("num,X,Y", FUSED_CASES)
def test_fused_pass_equals_literal_iterations(env, num, X, Y, listed, t_list=(1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12), need=3):
    """listed: the pass walks the precomputed tile list of fs2d_fused_order (what the host layer does); not listed: it
    classifies its tiles on the fly (order = NULL)"""
    _fused_pass_check(num, X, Y, t_list=t_list, need=need, listed=listed)


def _fused_pass_check(num, X, Y, mask_override=None, t_list=(1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12), need=3, listed=True):
    from fs import _lib
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.pressure_updater import JacobiPressureUpdater

    const, mask = build_scene(num, X, Y)
    if mask_override is not None:
        mask = mask_override
    bc = BoundaryCondition(const, mask)
    rng = np.random.default_rng(X + num)
    p0 = rng.uniform(-1, 1, mask.shape).astype(np.float32)
    v = fld(rng.uniform(-1, 1, mask.shape + (2,)).astype(np.float32))
    dt, dx = 0.05 / Y, 1.0 / Y
    jac = JacobiPressureUpdater(bc, dt, dx, 1, fuse=0)
    src = jac._source(v)
    relaxed = torch.from_numpy((mask != 1)).to(src.tensor.device)
    checked = 0
    for T in t_list:
        if not bc.fused_ok(T):
            continue
        a, b = fld(p0), fld(p0)
        for _ in range(T):  # literal iterations
            bc.set_pressure_boundary_condition(a); jac._sweep(b, a, src, inline_bc=False); a, b = b, a
        fin, fout = fld(p0), fld(p0)
        order, n_order = bc.fused_order(T) if listed else (None, 0)
        _lib.call("fs2d_jacobi_fused", fout.ptr(), fin.ptr(), src.ptr(), _lib.ptr(bc._pcode), bc.dom, T, _lib.ptr(order), n_order,
                  _lib.stream())
        got, want = fout.tensor, a.tensor
        same = (got == want) | (got.isnan() & want.isnan())
        bad = relaxed & ~same
        assert not bool(bad.any()), f"bc{num} {X}x{Y} T={T}: {int(bad.sum())} relaxed cells differ, first {torch.nonzero(bad)[:8].tolist()}"
        assert torch.equal(fout.tensor[~relaxed], torch.from_numpy(p0).to(relaxed.device)[~relaxed])  # walls untouched
        checked += 1
    assert checked >= need
    return checked


@pytest.mark.parametrize("seed", range(4))
def test_fused_pass_random_obstacles(env, seed, size=(512, 256), t_list=(1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12)):
    """Random blocky obstacles (thick enough for the reach rule) scattered over a channel: many tiles mix open-fluid
    warps, wall faces, convex corners and global edges."""
    from fs.boundary_condition import build_scene

    X, Y = size[0] + 64 * seed, size[1] + 16 * seed
    _, mask = build_scene(1, X, Y)
    rng = np.random.default_rng(100 + seed)
    mask = mask.copy()
    for _ in range(30 + 10 * seed):
        h, w = rng.integers(5, 40, 2)
        i0, j0 = rng.integers(8, X - 48), rng.integers(8, Y - 48)
        mask[i0:i0 + h, j0:j0 + w] = 1
    _fused_pass_check(1, X, Y, mask_override=mask, t_list=t_list, need=1, listed=seed % 2 == 0)


@pytest.mark.parametrize("num,X,Y,n_iter", [(2, 256, 128, 80), (5, 384, 192, 21), (1, 128, 64, 7), (3, 320, 160, 13), (2, 96, 48, 3)])
def test_fused_update_equals_literal_update(env, num, X, Y, n_iter):
    """JacobiPressureUpdater.update with fused passes == without, both physical buffers, bitwise."""
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.double_buffer import DoubleBuffer
    from fs.pressure_updater import JacobiPressureUpdater

    const, mask = build_scene(num, X, Y)
    bc = BoundaryCondition(const, mask)
    rng = np.random.default_rng(n_iter)
    p0 = rng.uniform(-1, 1, mask.shape).astype(np.float32)
    p1 = rng.uniform(-1, 1, mask.shape).astype(np.float32)
    v = fld(rng.uniform(-1, 1, mask.shape + (2,)).astype(np.float32))
    dt, dx = 0.05 / Y, 1.0 / Y
    for other, expect_fused in ((p0, True), (p1, False)):   # equal / different never-written wall cells
        res = []
        for fuse in (5, 0):
            jac = JacobiPressureUpdater(bc, dt, dx, n_iter, fuse=fuse)
            db = DoubleBuffer(mask.shape, 1)
            db.current.from_numpy(p0); db.next.from_numpy(other)
            if fuse:
                assert (jac.fuse_mask(db) > 0) == (expect_fused and any(bc.fused_ok(t) for t in range(1, 6)))
            jac.update(db, v)
            res.append((db.current.to_numpy(), db.next.to_numpy()))
        assert_bitexact("cur", res[0][0], res[1][0])
        if other is p0:
            assert_bitexact("nxt", res[0][1], res[1][1])


# ------------------------------------------------------------------------------------------------
# 6. dye transport (DyeMacSolver / DyeCipMacSolver) vs the reference fixtures and the oracle
# ------------------------------------------------------------------------------------------------
DYE_TRAJ = sorted(p.name for p in GOLDEN.glob("traj_dye*.npz"))


def dye_state(s) -> dict:
    d = fs_state(s)
    d.update(dye_cur=s.dye.current, dye_nxt=s.dye.next)
    if hasattr(s, "dyex"):
        d.update(dyex_cur=s.dyex.current, dyex_nxt=s.dyex.next, dyey_cur=s.dyey.current, dyey_nxt=s.dyey.next)
    return d


@pytest.mark.parametrize("name", DYE_TRAJ)
def test_dye_trajectory_matches_reference_fixture(env, name):
    from fs.boundary_condition import DyeBoundaryCondition
    from fs.fluid_simulator import make_solver

    g = np.load(GOLDEN / name)
    sc = np.load(GOLDEN / "dye_scenes.npz")
    num, res, vc = int(g["meta_num"]), int(g["meta_res"]), float(g["meta_vc"])
    kind, n_iter = str(g["meta_pressure"]), int(g["meta_n_iter"])
    bc = DyeBoundaryCondition(sc[f"bc{num}_r{res}_const"], sc[f"bc{num}_r{res}_dye"], sc[f"bc{num}_r{res}_mask"])
    s = make_solver(bc, float(g["meta_dt"]), float(g["meta_dx"]), float(g["meta_re"]), None if vc < 0 else vc,
                    str(g["meta_scheme"]), pressure=kind, n_iter=n_iter, dye=True)
    for k, f in dye_state(s).items():
        f.from_numpy(g[f"s0_{k}"])
    for n in range(1, int(g["meta_steps"]) + 1):
        s.update()
        for k, f in dye_state(s).items():
            assert_bitexact(f"{name} step {n} {k}", f.to_numpy(), g[f"s{n}_{k}"])


@pytest.mark.parametrize("num,res,scheme", [(1, 128, "cip"), (2, 96, "kk"), (5, 100, "upwind"), (2, 320, "cip"), (3, 256, "cip")])
def test_dye_simulator_vs_oracle(env, num, res, scheme):
    """DyeFluidSimulator.create (main.py's default object graph) for a few steps vs the oracle.  res >= 256: grids wide enough
    for blocks that lie wholly inside the clamp window (the kernels' interior fast paths) and for RB-SOR chunks of 32 rows."""
    from fs.boundary_condition import build_scene
    from fs.fluid_simulator import DyeFluidSimulator
    from oracle import oracle as orc

    dt, dx, re, vc = 0.05 / res, 1.0 / res, 1e4, 5.0
    sim = DyeFluidSimulator.create(num, res, dt, dx, re, vc, scheme)
    const, mask, dye = build_scene(num, 2 * res, res, with_dye=True)
    ref = orc.OracleSolver(mask, const, dt, dx, re, scheme, vc, ("rbsor", 1.3, 2), bc_dye=dye)
    for _ in range(6 if res < 200 else 3):
        sim.step(); ref.update()
    out = sim.field_to_numpy()
    assert set(out) == {"v", "p", "dye"} and out["dye"].shape == (2 * res, res, 3)
    assert_bitexact("v", out["v"], ref.v.current); assert_bitexact("p", out["p"], ref.p.current)
    assert_bitexact("dye", out["dye"], ref.dye.current)
    assert float(out["dye"].max()) > 0.5   # dye actually entered the domain
    if res >= 200:      # every physical buffer, dye derivative fields included
        got = dye_state(sim.solver)
        for k, a in ref.state().items():
            assert_bitexact(f"bc{num} res {res} {k}", got[k].to_numpy(), a)


@pytest.mark.parametrize("vec", [1, 0])
@pytest.mark.parametrize("num,res,dx", [(2, 160, 1.0 / 160), (3, 136, 1.0 / 128)])
def test_dye_step_from_random_state_vs_oracle(env, num, res, dx, vec):
    """The default object graph stepped from a RANDOM state (velocities of both signs in every cell, non-zero dye and dye
    derivatives everywhere): the dye kernels' interior blocks see real data -- a fresh scene has dye only next to the inflow,
    which is edge-block territory.  dx: a non-power-of-two spacing (IEEE divisions) and a power of two (exact reciprocals)."""
    from fs.boundary_condition import build_scene
    from fs.fluid_simulator import DyeFluidSimulator
    from oracle import oracle as orc

    dt, re, vc = 0.02 / res, 3e3, 5.0
    sim = DyeFluidSimulator.create(num, res, dt, dx, re, vc, "cip")
    const, mask, dye = build_scene(num, 2 * res, res, with_dye=True)
    ref = orc.OracleSolver(mask, const, dt, dx, re, "cip", vc, ("rbsor", 1.3, 2), bc_dye=dye)
    assert env.fs2d_set_tuning(5, vec) == 0      # dye non-advection: four cells per thread (default) / one cell per thread
    try:
        _dye_random_steps(sim, ref, num, res)
    finally:
        env.fs2d_set_tuning(5, 1)


def _dye_random_steps(sim, ref, num, res):
    rng = np.random.default_rng(res)
    st = {}
    for k, a in ref.state().items():
        scale = {"v": 1.0, "p": 1.0, "vx": 20.0, "vy": 20.0, "dye": 1.0, "dyex": 30.0, "dyey": 30.0}.get(k.split("_")[0], 1.0)
        r = rng.uniform(-1.0, 1.0, a.shape).astype(np.float32) * np.float32(scale)
        st[k] = np.abs(r) if k.startswith("dye_") else r
    ref.load_state(st)
    got = dye_state(sim.solver)
    for k, a in st.items():
        got[k].from_numpy(a)
    for n in range(2):
        sim.step(); ref.update()
        got = dye_state(sim.solver)      # the double buffers have swapped
        for k, a in ref.state().items():
            assert_bitexact(f"bc{num} res {res} step {n} {k}", got[k].to_numpy(), a)


# ------------------------------------------------------------------------------------------------
# 7. render getters and full-state dump / restore
# ------------------------------------------------------------------------------------------------
def test_render_matches_reference_fixture(env):
    from fs.boundary_condition import DyeBoundaryCondition
    from fs.fluid_simulator import DyeFluidSimulator, make_solver

    g = np.load(GOLDEN / "render.npz")
    sc = np.load(GOLDEN / "dye_scenes.npz")
    for num, res in ((2, 16), (3, 20)):
        pre = f"bc{num}_r{res}/"
        bc = DyeBoundaryCondition(sc[f"bc{num}_r{res}_const"], sc[f"bc{num}_r{res}_dye"], sc[f"bc{num}_r{res}_mask"])
        sim = DyeFluidSimulator(make_solver(bc, 0.05 / res, 1.0 / res, 1e4, None, "cip", pressure="jacobi", n_iter=1, dye=True))
        s = sim.solver
        s.v.current.from_numpy(g[pre + "v"]); s.p.current.from_numpy(g[pre + "p"]); s.dye.current.from_numpy(g[pre + "dye"])
        assert_bitexact("norm", sim.get_norm_field().to_numpy(), g[pre + "norm"])
        assert_bitexact("pressure", sim.get_pressure_field().to_numpy(), g[pre + "pressure"])
        assert_bitexact("vorticity", sim.get_vorticity_field().to_numpy(), g[pre + "vorticity"])
        assert_bitexact("dye", sim.get_dye_field().to_numpy(), g[pre + "dye_img"])
        assert sim.rgb_buf.to_numpy().shape == (2 * res, res, 3)


def test_state_dict_roundtrip_resumes_bitwise(env):
    """A run restored from state_dict() continues bit-identically (the reference's npz dump cannot: no vx/vy)."""
    from fs.fluid_simulator import DyeFluidSimulator

    res = 64
    args = (2, res, 0.05 / res, 1.0 / res, 1e4, 5.0, "cip")
    a = DyeFluidSimulator.create(*args, pressure="jacobi", n_iter=9)
    for _ in range(4):
        a.step()
    saved = a.state_dict()
    for _ in range(3):
        a.step()
    b = DyeFluidSimulator.create(*args, pressure="jacobi", n_iter=9)
    b.load_state_dict(saved)
    for _ in range(3):
        b.step()
    for k, v in a.state_dict().items():
        assert_bitexact(k, b.state_dict()[k], v)


@pytest.mark.parametrize("scheme,vc,n_iter,dye", [("cip", 5.0, 12, False), ("cip", None, 7, True), ("upwind", 5.0, 4, False)])
def test_cuda_graph_stepping_is_bitwise_identical(env, scheme, vc, n_iter, dye):
    """FluidSimulator.enable_cuda_graph(): graph replay == eager stepping, every physical buffer, odd/even swaps."""
    from fs.fluid_simulator import DyeFluidSimulator, FluidSimulator

    res = 96
    cls = DyeFluidSimulator if dye else FluidSimulator
    args = (2, res, 0.05 / res, 1.0 / res, 1e4, vc, scheme)
    eager = cls.create(*args, pressure="jacobi", n_iter=n_iter)
    graph = cls.create(*args, pressure="jacobi", n_iter=n_iter)
    graph.enable_cuda_graph()
    eager.step()                                  # enable_cuda_graph() ran one warm-up step
    captured = graph._graphs
    for n in range(5):
        eager.step(); graph.step()
        se, sg = eager.state_dict(), graph.state_dict()
        for k in se:
            assert_bitexact(f"step {n} {k}", sg[k], se[k])
    assert graph._graphs is captured, "the graphs must be replayed, not re-captured, while nobody writes the buffers from the host"


def test_cuda_graph_default_path_replays_and_recaptures_after_host_writes(env):
    """The facade's defaults (RB-SOR x2, dye): step() replays the captured graphs (r02 regression: a never-cleared `dirty` flag
    made every step drop and re-capture them); a pressure buffer rewritten from the host drops them once, and stepping stays
    bit-identical to eager stepping."""
    from fs.fluid_simulator import DyeFluidSimulator

    res = 64
    args = (3, res, 0.05 / res, 1.0 / res, 1e4, 5.0, "cip")
    eager, graph = DyeFluidSimulator.create(*args), DyeFluidSimulator.create(*args)
    graph.enable_cuda_graph()
    eager.step()
    captured = graph._graphs
    for _ in range(3):
        eager.step(); graph.step()
    assert graph._graphs is captured
    rng = np.random.default_rng(5)
    p_new = rng.uniform(-1, 1, eager.solver.p.current.to_numpy().shape).astype(np.float32)
    for sim in (eager, graph):
        sim.solver.p.current.from_numpy(p_new)
    for n in range(3):
        eager.step(); graph.step()
        se, sg = eager.state_dict(), graph.state_dict()
        for k in se:
            assert_bitexact(f"after host write, step {n} {k}", sg[k], se[k])
    assert graph._graphs is not None and graph._graphs is not captured


# ------------------------------------------------------------------------------------------------
# 8. BASELINE full sizes
# ------------------------------------------------------------------------------------------------
def test_config3_full_size_step_vs_oracle(env):
    """BASELINE config 3 exactly (bc=3 res=4096 CIP Re=1e8 vc=10, 100 Jacobi sweeps): one step, every buffer
    bit-compared with the oracle."""
    from fs.boundary_condition import build_scene
    from oracle import oracle as orc

    res = 4096
    dt, dx, re, vc, pressure = 0.05 / res, 1.0 / res, 1e8, 10.0, ("jacobi", 100)
    const, mask = build_scene(3, 2 * res, res)
    s = make_fs(mask, const, dt, dx, re, "cip", vc, pressure)
    ref = orc.OracleSolver(mask, const, dt, dx, re, "cip", vc, pressure)
    s.update(); ref.update()
    got = fs_state(s)
    for k, a in ref.state().items():
        assert_bitexact(f"config3 {k}", got[k].to_numpy(), a)


def test_bench_workload_step_vs_oracle(env):
    """The EXACT workload bench.py times on one GPU -- bc=2 on the 8192 x 8192 shard of BASELINE config 4, CIP, Re=1e5, vc=5,
    dt=0.05/8192, 80 Jacobi sweeps, quiescent start -- two steps, every physical buffer bit-compared with the oracle
    (67 M cells: the oracle needs a few seconds per step on the box's host cores)."""
    from fs.boundary_condition import build_scene
    from oracle import oracle as orc
    import os

    orc.set_threads(os.cpu_count() or 1)
    X = Y = 8192
    dt, dx, re, vc, pressure = 0.05 / Y, 1.0 / Y, 1e5, 5.0, ("jacobi", 80)
    const, mask = build_scene(2, X, Y)
    s = make_fs(mask, const, dt, dx, re, "cip", vc, pressure)
    ref = orc.OracleSolver(mask, const, dt, dx, re, "cip", vc, pressure)
    assert any(t > 0 for t in s.pressure_updater.plan(s.p)), "the bench workload must run fused passes"
    for n in range(2):
        s.update(); ref.update()
        got = fs_state(s)
        for k, a in ref.state().items():
            assert_bitexact(f"bench workload step {n} {k}", got[k].to_numpy(), a)


def test_config5_grid_fused_equals_literal(env):
    """BASELINE config 5's grid on ONE GPU (bc=5 res=16384: 32768 x 16384 = 537 M cells): the fused Jacobi update
    (200 sweeps) equals 200 literal iterations bitwise -- size-independent property, no oracle needed."""
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.double_buffer import DoubleBuffer, Field
    from fs.pressure_updater import JacobiPressureUpdater

    res = 16384
    const, mask = build_scene(5, 2 * res, res)
    bc = BoundaryCondition(const, mask)
    del const, mask
    X, Y = 2 * res, res
    g = torch.Generator(device="cuda").manual_seed(3)
    v = Field((X, Y), 2)
    v.tensor.copy_(torch.rand(v.tensor.shape, device="cuda", generator=g) * 2 - 1)
    p0 = torch.rand((X, Y), device="cuda", generator=g) * 2 - 1
    res_cur = []
    for fuse in (8, 0):
        jac = JacobiPressureUpdater(bc, 0.05 / res, 1.0 / res, 200 if fuse else 200, fuse=fuse)
        db = DoubleBuffer((X, Y), 1)
        db.current.tensor.copy_(p0); db.next.tensor.copy_(p0)
        db.current.dirty = db.next.dirty = True
        if fuse:
            assert jac.fuse_mask(db) != 0
        jac.update(db, v)
        res_cur.append(db.current.tensor.clone())
        del db, jac
    same = (res_cur[0] == res_cur[1]) | (res_cur[0].isnan() & res_cur[1].isnan())
    assert bool(same.all()), f"{int((~same).sum())} cells differ"


# ------------------------------------------------------------------------------------------------
# 9. adversarial random masks: thin walls, inflow/outflow cells anywhere, wall-BC cells feeding inflow cells
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(6))
def test_random_mask_trajectory_vs_oracle(env, seed):
    """Every code path that depends on the mask (sparse BC tables, target-centric velocity BC, pcode, fused-pass
    validity analysis and its fall-backs) on masks the reference's scenes never produce; 3 CIP+VC steps with
    Jacobi (fuse auto) and 2 upwind steps with RB-SOR, all physical buffers vs the oracle."""
    from oracle import oracle as orc

    rng = np.random.default_rng(100 + seed)
    X, Y = 96, 64
    mask = np.zeros((X, Y), dtype=np.uint8)
    mask[:, :2] = 1; mask[:, -2:] = 1
    for _ in range(int(rng.integers(6, 14))):           # random wall blobs, some 1 cell thick
        i, j = int(rng.integers(4, X - 8)), int(rng.integers(2, Y - 6))
        mask[i:i + int(rng.integers(1, 7)), j:j + int(rng.integers(1, 7))] = 1
    mask[:2, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.8, 2, mask[:2, 2:-2])       # ragged inflow
    mask[-2:, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.7, 3, mask[-2:, 2:-2])     # ragged outflow
    for _ in range(4):                                    # stray inflow / outflow cells inside the domain
        mask[int(rng.integers(3, X - 3)), int(rng.integers(3, Y - 3))] = int(rng.integers(2, 4))
    const = np.zeros((X, Y, 2), dtype=np.float32)
    const[mask == 2] = (1.0, 0.0)
    res = Y
    dt, dx = 0.05 / res, 1.0 / res
    for scheme, vc, pressure in (("cip", 5.0, ("jacobi", 11)), ("upwind", None, ("rbsor", 1.3, 2))):
        s = make_fs(mask, const, dt, dx, 300.0, scheme, vc, pressure)
        ref = orc.OracleSolver(mask, const, dt, dx, 300.0, scheme, vc, pressure)
        init = {}
        for k, a in ref.state().items():
            scale = 0.05 / dx if k[:2] in ("vx", "vy") else 0.5
            init[k] = (rng.uniform(-1, 1, a.shape) * scale).astype(np.float32)
        init["p_nxt"] = init["p_cur"].copy()              # equal never-written cells: the fused path may engage
        ref.load_state(init)
        load_fs_state(s, init)
        for n in range(3 if scheme == "cip" else 2):
            s.update(); ref.update()
            got = fs_state(s)
            for k, a in ref.state().items():
                assert_bitexact(f"seed {seed} {scheme} step {n} {k}", got[k].to_numpy(), a)


# ------------------------------------------------------------------------------------------------
# 9. TMA-fed streaming stencil kernels (fs2d_stream.cu) == direct kernels, bitwise
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num,X,Y", [(2, 256, 128), (3, 200, 176), (5, 333, 208), (1, 64, 48)])
def test_stream_kernels_equal_direct_kernels(env, num, X, Y):
    """cip_nonadv / cip_advect through the TMA tile pipeline vs the one-cell-per-thread kernels: whole grid and a row
    window with clamp bounds inside the array (what an edge rank of a strip decomposition passes)."""
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.fluid_simulator import make_solver

    const, mask = build_scene(num, X, Y)
    bc = BoundaryCondition(const, mask)
    dt, dx = 0.05 / Y, 1.0 / 128      # power-of-two dx; Y is not (exercises partial tiles)
    rng = np.random.default_rng(X * Y)
    doms = [bc.dom, bc.dom.replace(r0=7, r1=X - 9, clo=3, chi=X - 4)]
    for dxv in (dx, 0.01):             # exact-reciprocal and true-division instantiations
        s = make_solver(bc, dt, dxv, 1e3, 5.0, "cip", pressure="jacobi", n_iter=1)
        vcf = s.vorticity_confinement
        init = {k: rng.uniform(-1, 1, getattr(s, k).current.tensor.shape).astype(np.float32) for k in ("v", "vx", "vy", "p")}
        for dom in doms:
            out = []
            for stream in (2, 0):
                env.fs2d_set_tuning(2, stream)
                try:
                    for k, a in init.items():
                        getattr(s, k).current.from_numpy(a); getattr(s, k).next.from_numpy(a[::-1].copy())
                    old = bc.dom
                    bc.dom = dom
                    s._non_advection_phase(s.v.next, s.v.current, s.p.current)
                    r1 = s.v.next.to_numpy()
                    s._advection_phase(s.v.next, s.vx.next, s.vy.next, s.v.current, s.vx.current, s.vy.current, s.v.current)
                    adv = (s.v.next.to_numpy(), s.vx.next.to_numpy(), s.vy.next.to_numpy())
                    vcf.vorticity.from_numpy(init["p"]); vcf.vorticity_abs.from_numpy(np.abs(init["p"][:, ::-1]).copy())
                    vcf._apply_fused(s.v.next, s.v.current)
                    out.append((r1,) + adv + (s.v.next.to_numpy(), vcf.vorticity.to_numpy(), vcf.vorticity_abs.to_numpy()))
                finally:
                    bc.dom = old
                    env.fs2d_set_tuning(2, 1)
            for name, a, b in zip(("nonadv", "adv f", "adv fx", "adv fy", "vort vn", "vort w", "vort |w|"), out[0], out[1]):
                assert_bitexact(f"{name} bc{num} dx={dxv} dom={dom.r0}:{dom.r1}", a, b)


@pytest.mark.parametrize("num,X,Y", [(1, 64, 8), (2, 256, 128), (3, 200, 132), (5, 333, 260)])
def test_vectorised_nonadv_equals_one_cell_kernel(env, num, X, Y):
    """cip_nonadv on four cells per thread (128-bit accesses, j-neighbours by shuffle; the default when Y % 4 == 0) vs the
    one-cell-per-thread kernel (fs2d_set_tuning(6, 0)): widths of two lanes, one full warp, a warp plus one lane, two warps
    plus one lane; whole grid and a row window with clamp bounds inside the array; both division instantiations."""
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.fluid_simulator import make_solver

    const, mask = build_scene(num, X, Y)
    bc = BoundaryCondition(const, mask)
    rng = np.random.default_rng(X + Y)
    doms = [bc.dom, bc.dom.replace(r0=7, r1=X - 9, clo=3, chi=X - 4)]
    for dxv in (1.0 / 128, 0.01):
        s = make_solver(bc, 0.05 / 128, dxv, 1e3, 5.0, "cip", pressure="jacobi", n_iter=1)
        init = {k: rng.uniform(-1, 1, getattr(s, k).current.tensor.shape).astype(np.float32) for k in ("v", "p")}
        for dom in doms:
            out = []
            for vec in (1, 0):
                assert env.fs2d_set_tuning(6, vec) == 0
                old = bc.dom
                try:
                    s.v.current.from_numpy(init["v"]); s.v.next.from_numpy(init["v"][::-1].copy()); s.p.current.from_numpy(init["p"])
                    bc.dom = dom
                    s._non_advection_phase(s.v.next, s.v.current, s.p.current)
                    out.append(s.v.next.to_numpy())
                finally:
                    bc.dom = old
                    env.fs2d_set_tuning(6, 1)
            assert_bitexact(f"nonadv bc{num} {X}x{Y} dx={dxv} dom={dom.r0}:{dom.r1}", out[0], out[1])
            if (np.asarray(mask)[dom.r0:dom.r1] != 1).any():
                assert not np.array_equal(out[0], init["v"][::-1])      # it did write


def test_fused_pass_split_into_interior_and_edge_launches(env, X=1000, Y=512):
    """fs2d_jacobi_fused on an interior row window + fs2d_jacobi_fused_part on the remaining tile rows == one launch
    (what the multi-rank host does to hide the halo exchange)."""
    from fs import _lib
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.halo import split_windows
    from fs.pressure_updater import JacobiPressureUpdater
    import ctypes

    const, mask = build_scene(2, X, Y)
    bc = BoundaryCondition(const, mask)
    rng = np.random.default_rng(7)
    p0 = rng.uniform(-1, 1, mask.shape).astype(np.float32)
    v = fld(rng.uniform(-1, 1, mask.shape + (2,)).astype(np.float32))
    jac = JacobiPressureUpdater(bc, 0.05 / Y, 1.0 / Y, 1, fuse=0)
    src = jac._source(v)
    for T in (4, 8):
        assert bc.fused_ok(T)
        rows, cols, hr, hc, tmax = (ctypes.c_int() for _ in range(5))
        _lib.call("fs2d_fused_tile", T, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(hr), ctypes.byref(hc), ctypes.byref(tmax))
        mid, m = split_windows(bc.dom, rows.value - 2 * hr.value, T)
        assert mid is not None
        whole, parts, fin = fld(p0), fld(p0), fld(p0)
        jac._fused(whole, fin, src, T)
        jac._fused(parts, fin, src, T, dom=mid)
        jac._fused(parts, fin, src, T, skip=(1, m - 1))
        assert torch.equal(whole.tensor, parts.tensor), f"T={T}: split launches differ from the single launch"
        with pytest.raises((ValueError, RuntimeError)):
            jac._fused(parts, fin, src, T, skip=(1, 10 ** 6))


# ------------------------------------------------------------------------------------------------
# 10. IEEE special cases of the guarded divisions / square root (fdiv_z) vs the oracle's plain C arithmetic
# ------------------------------------------------------------------------------------------------
def test_division_special_cases_match_the_oracle(env):
    """Zero, denormal, infinite and NaN operands in the vorticity-confinement normalisation (0/0, x/0 by underflow of
    the norm, 0/x, inf/inf), the viscous term (0/Re, denormal/Re) and the pressure source: the kernels that keep zero
    dividends out of div.rn.f32 must still produce the oracle's IEEE results (NaN == NaN)."""
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.fluid_simulator import make_solver
    from oracle import oracle as orc

    X, Y = 64, 48
    const, mask = build_scene(1, X, Y)
    bc = BoundaryCondition(const, mask)
    dt, dx, re, vcw = 0.05 / Y, 1.0 / 64, 1e4, 5.0
    s = make_solver(bc, dt, dx, re, vcw, "cip", pressure="jacobi", n_iter=1)
    vcf = s.vorticity_confinement
    rng = np.random.default_rng(3)
    special = np.array([0.0, -0.0, 1e-32, -1e-32, 1e-42, -3e-45, 1.0, -2.5, 3e38, -3e38, np.inf, -np.inf, np.nan, 1e-20, 7e-39],
                       dtype=np.float32)
    pick = lambda shape: special[rng.integers(0, len(special), shape)]  # noqa: E731
    for trial in range(6):
        v = pick(mask.shape + (2,)); p = pick(mask.shape); w = pick(mask.shape); wa = np.abs(pick(mask.shape))
        if trial % 2:   # smooth background with sparse special values: neighbours differ by tiny amounts
            v = np.where(rng.random(v.shape) < 0.2, v, np.float32(0.25)); wa = np.where(rng.random(wa.shape) < 0.3, wa, np.float32(0.5))
        v, p, w, wa = (np.ascontiguousarray(a, dtype=np.float32) for a in (v, p, w, wa))
        with np.errstate(all="ignore"):
            # vorticity force
            want = v[..., ::-1].copy(); orc.vort_add(want, v, w, wa, mask, dx, dt, vcw)
            vcf.vorticity.from_numpy(w); vcf.vorticity_abs.from_numpy(wa)
            got = fld(v[..., ::-1].copy()); vcf._add_vorticity(got, fld(v))
            assert_bitexact(f"vort_add special {trial}", got.to_numpy(), want)
            # non-advection phase (division by Re)
            want = v[..., ::-1].copy(); orc.cip_nonadv(want, v, p, mask, dt, dx, re)
            got = fld(v[..., ::-1].copy()); s._non_advection_phase(got, fld(v), fld(p))
            assert_bitexact(f"nonadv special {trial}", got.to_numpy(), want)
            # one Jacobi sweep from the source terms (division by 8 dt)
            pn_want = p[::-1].copy(); orc.jacobi_sweep(pn_want, p, v, mask, dt, dx)
            jac = s.pressure_updater
            pn = fld(p[::-1].copy()); jac._sweep(pn, fld(p), jac._source(fld(v)), inline_bc=False)
            assert_bitexact(f"jacobi special {trial}", pn.to_numpy(), pn_want)


# ------------------------------------------------------------------------------------------------
# 11. the two tails of fs2d_jacobi_update; tile lists
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num,X,Y,n_iter", [(2, 256, 128, 80), (5, 384, 192, 21), (1, 128, 64, 7), (3, 320, 160, 13), (2, 96, 48, 3),
                                            (4, 200, 96, 4), (1, 288, 352, 10)])
def test_two_literal_tail_equals_literal_update(env, num, X, Y, n_iter):
    """fs2d_set_tuning(4, 0): the update ends with two literal iterations instead of the default {fused pass emitting the BC
    values of its penultimate state, ONE literal iteration}; both physical buffers -- wall-BC cells included -- must equal
    n literal iterations either way (the default tail is what test_fused_update_equals_literal_update runs)."""
    env.fs2d_set_tuning(4, 0)
    try:
        test_fused_update_equals_literal_update(env, num, X, Y, n_iter)
    finally:
        env.fs2d_set_tuning(4, 1)


@pytest.mark.parametrize("num,X,Y", [(2, 1000, 512), (5, 768, 384), (3, 640, 320)])
def test_tile_list_classes(env, num, X, Y):
    """fs2d_fused_order: every tile of the pass is listed exactly once or dropped; dropped tiles have no relaxed / BC cell in
    their output region; the slow tiles come first; tiles listed as open fluid contain nothing but plain fluid cells."""
    from fs.boundary_condition import BoundaryCondition, build_scene
    import ctypes
    from fs import _lib

    const, mask = build_scene(num, X, Y)
    bc = BoundaryCondition(const, mask)
    pcode = bc._pcode.cpu().numpy()
    for T in (1, 4, 8, 12):
        rows, cols, hr, hc, tmax = (ctypes.c_int() for _ in range(5))
        _lib.call("fs2d_fused_tile", T, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(hr), ctypes.byref(hc), ctypes.byref(tmax))
        TI, TJ, HI, HJ = rows.value - 2 * hr.value, cols.value - 2 * hc.value, hr.value, hc.value
        order, n = bc.fused_order(T)
        _, n_slow, n_skip = bc._fused_orders[(T, bc.dom.r0, bc.dom.r1, bc.dom.clo, bc.dom.chi, 0, 0)][1:]
        ent = order.cpu().numpy()[:n]
        tiles_i, tiles_j = -(-X // TI), -(-Y // TJ)
        tiles, cls = ((ent >> 14) & 0x3FFF) * tiles_j + (ent & 0x3FFF), ent >> 28      # entry = col | row << 14 | class << 28
        assert n + n_skip == tiles_i * tiles_j and len(set(tiles.tolist())) == n
        assert (cls[:n_slow] == 1).all() and (cls[n_slow:] == 0).all()
        listed = set(tiles.tolist())
        for t in range(tiles_i * tiles_j):
            r0, c0 = (t // tiles_j) * TI, (t % tiles_j) * TJ
            out = pcode[r0:min(r0 + TI, X), c0:min(c0 + TJ, Y)] & 15
            if t not in listed:
                assert (out == 9).all(), f"dropped tile {t} has work to do"
        for t in tiles[cls == 0]:
            r0, c0 = (t // tiles_j) * TI - HI, (t % tiles_j) * TJ - HJ
            assert r0 > 0 and c0 > 0 and r0 + rows.value < X and c0 + cols.value < Y
            assert (pcode[r0:r0 + rows.value, c0:c0 + cols.value] == 0).all()


# ------------------------------------------------------------------------------------------------
# 12. red-black SOR: both colour passes in one kernel == the two reference passes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num,X,Y", [(1, 128, 64), (2, 256, 128), (3, 200, 176), (5, 384, 192), (4, 97, 80), (2, 1000, 512)])
def test_rbsor_fused_colours_equal_two_passes(env, num, X, Y):
    """fs2d_rbsor_iteration == fs2d_rbsor_pass(odd, pn <- pc) + fs2d_rbsor_pass(even, pn <- pn): whole grid and row windows with
    clamp bounds inside the array, random fields in BOTH buffers (stale values of never-written cells must survive), also on
    masks with fluid cells on the grid's first / last rows."""
    from fs import _lib
    from fs.boundary_condition import BoundaryCondition, build_scene
    from fs.pressure_updater import RedBlackSorPressureUpdater

    const, mask = build_scene(num, X, Y)
    rng = np.random.default_rng(num * 1000 + X)
    for variant in range(2):
        m = mask.copy()
        if variant == 1:      # fluid cells on the first / last rows and next to the first / last columns: every clamp of sample()
            m[:2, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.5, 0, m[:2, 2:-2])
            m[-2:, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.5, 0, m[-2:, 2:-2])
            m[4:-4, :2] = np.where(rng.random((X - 8, 2)) < 0.5, 0, m[4:-4, :2])
            m[4:-4, -2:] = np.where(rng.random((X - 8, 2)) < 0.5, 0, m[4:-4, -2:])
        bc = BoundaryCondition(const, m)
        sor = RedBlackSorPressureUpdater(bc, 0.05 / Y, 1.0 / Y, 1.3, 1)
        v = fld(rng.uniform(-1, 1, m.shape + (2,)).astype(np.float32))
        src = sor._source(v)
        pc0, pn0 = (rng.uniform(-1, 1, m.shape).astype(np.float32) for _ in range(2))
        doms = [bc.dom, bc.dom.replace(r0=5, r1=X - 11, clo=2, chi=X - 3), bc.dom.replace(r0=33, r1=34)]
        for dom in doms:
            res = []
            for fused in (True, False):
                pc, pn = fld(pc0), fld(pn0)
                if fused:
                    _lib.call("fs2d_rbsor_iteration", pn.ptr(), pc.ptr(), src.ptr(), _lib.ptr(bc._bc_mask), dom, 1.3, 1.0 - 1.3, _lib.stream())
                else:
                    sor._pass(pn, pc, src, 1, dom=dom)
                    sor._pass(pn, pn, src, 0, dom=dom)
                res.append(pn.to_numpy())
            assert_bitexact(f"bc{num} {X}x{Y} variant {variant} rows {dom.r0}:{dom.r1}", res[0], res[1])
