"""Host-side BC tables (fs/_bc_tables.py) vs the oracle's BC kernels, on CPU tensors."""
from __future__ import annotations

import numpy as np
import pytest
import torch
from conftest import assert_bitexact

from fs import _bc_tables as T
from fs.boundary_condition import build_scene
from fs.distributed import Partition
from oracle import oracle as orc


def apply_vel_table(v: np.ndarray, const: np.ndarray, t: dict) -> np.ndarray:
    """numpy model of k_vel_bc_gather/k_vel_bc_scatter (test-only)."""
    out = v.reshape(-1, 2).copy()
    flat, c = v.reshape(-1, 2), const.reshape(-1, 2)
    tgt, src, kind = t["tgt"].numpy().astype(np.int64), t["src"].numpy().astype(np.int64), t["kind"].numpy()
    vals = np.where((kind == 0)[:, None], -flat[src], c[tgt])
    of = np.stack([np.fmax(flat[src][:, 0], np.float32(0.05)), flat[tgt][:, 1]], axis=1)
    vals = np.where((kind == 2)[:, None], of, vals)
    out[tgt] = vals
    return out.reshape(v.shape)


def apply_p_table(p: np.ndarray, t: dict) -> np.ndarray:
    flat = p.reshape(-1)
    out = flat.copy()
    tgt, s0, s1, kind = (t[k].numpy().astype(np.int64) for k in ("tgt", "src0", "src1", "kind"))
    vals = np.where(kind == 0, flat[s0], np.where(kind == 1, (flat[s0] + flat[s1]) / np.float32(2.0), np.float32(0.0)))
    out[tgt] = vals.astype(np.float32)
    return out.reshape(p.shape)


def scenes():
    for num in (1, 2, 3, 4, 5):
        for res in (16, 24, 40):
            yield num, res


@pytest.mark.parametrize("num,res", list(scenes()))
def test_tables_equal_oracle_on_scenes(num, res):
    const, mask = build_scene(num, 2 * res, res)
    rng = np.random.default_rng(num * 100 + res)
    v = rng.uniform(-1, 1, mask.shape + (2,)).astype(np.float32)
    p = rng.uniform(-1, 1, mask.shape).astype(np.float32)
    tm = torch.from_numpy(mask)
    vt = T.velocity_table(tm)
    assert num == 3 or not vt["thin_walls"]  # tiny bc3 discs are 1-cell walls
    want = v.copy(); orc.vel_bc(want, mask, const)
    assert_bitexact("vel table", apply_vel_table(v, const, vt), want)
    pt = T.pressure_table(T.pressure_codes(tm))
    want = p.copy(); orc.p_bc(want, mask)
    assert_bitexact("p table", apply_p_table(p, pt), want)


def test_tables_equal_oracle_on_random_masks():
    rng = np.random.default_rng(11)
    for trial in range(60):
        X, Y = int(rng.integers(5, 28)), int(rng.integers(5, 28))
        mask = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(X, Y), p=[0.45, 0.45, 0.05, 0.05])
        const = rng.uniform(-1, 1, (X, Y, 2)).astype(np.float32)
        v = rng.uniform(-1, 1, (X, Y, 2)).astype(np.float32)
        p = rng.uniform(-1, 1, (X, Y)).astype(np.float32)
        tm = torch.from_numpy(mask)
        want = v.copy(); orc.vel_bc(want, mask, const)
        assert_bitexact(f"vel {trial}", apply_vel_table(v, const, T.velocity_table(tm)), want)
        want = p.copy(); orc.p_bc(want, mask)
        assert_bitexact(f"p {trial}", apply_p_table(p, T.pressure_table(T.pressure_codes(tm))), want)


def test_pcode_predicates():
    _, mask = build_scene(2, 64, 32)
    code = T.pressure_codes(torch.from_numpy(mask)).numpy()
    assert ((code >= 1) & (code <= 9)).sum() == (mask == 1).sum()
    assert ((code == T.PC_FLUID) == (mask == 0)).all()
    assert ((code == T.PC_INFLOW) == (mask == 2)).all() and ((code == T.PC_OUTFLOW) == (mask == 3)).all()
    # bc2 has never-written wall cells next to inflow/outflow cells (SURVEY T1 / DESIGN "stale")
    assert T.exposed_stale_cells(torch.from_numpy(code)).numel() > 0


def test_window_tables_are_local_and_checked():
    _, mask = build_scene(2, 64, 32)
    tm = torch.from_numpy(mask)
    part = Partition(64, rank=1, world=2, halo=2)
    g0, g1 = part.owned(); w0, w1 = part.window()
    t = T.velocity_table(tm, g0, min(g1, 64), w0, w1)
    assert t["n"] > 0 and int(t["tgt"].min()) >= 0 and int(t["src"].min()) >= 0
    with pytest.raises(ValueError):
        T.velocity_table(tm, 0, 64, 1, 64)  # source rows outside the window


def test_partition_arithmetic():
    for X, P in ((64, 1), (64, 2), (67, 4), (8192 * 8, 8)):
        rows = []
        for r in range(P):
            p = Partition(X, r, P, 0 if P == 1 else 2)
            g0, g1 = p.owned()
            rows += list(range(g0, g1))
            assert p.window() == (g0 - p.halo, g1 + p.halo)
        assert rows == list(range(X))
    with pytest.raises(ValueError):
        Partition(64, 2, 2, 2)
    with pytest.raises(ValueError):
        Partition(64, 0, 2, 1)


def test_fused_reach_bounds_the_stale_rows_below_a_strip():
    """fused_reach_ok(fresh_below=T): an inflow cell right below a strip's last row doubles the downward dependency reach
    (its BC value is p(i+1, j), fs/boundary_condition.py:62-63, so the relaxed cell above it reads two rows down), and
    after T iterations the cone touches row row1 + T -- beyond the T halo rows exchanged before the pass.  The tile
    (anchored at row0) extends further down, so only the strip-aware bound catches it."""
    X, Y, tile_rows, tile_cols = 96, 128, 64, 128
    for stray_inflow, expect_strip_ok in ((True, False), (False, True)):
        mask = np.zeros((X, Y), dtype=np.uint8)
        mask[:, :2] = 1; mask[:, -2:] = 1
        mask[60:64, 40:60] = 1                          # an ordinary thick obstacle: reach stays one row per iteration
        if stray_inflow:
            mask[48, 50] = 2                            # first row below the strip [0, 48)
        code = T.pressure_codes(torch.from_numpy(mask))
        for t in (1, 2, 3, 4):
            hj = (t + 3) & ~3
            assert T.fused_reach_ok(code, t, tile_rows, tile_cols, t, hj, 0, X)                     # single domain: fine
            assert T.fused_reach_ok(code, t, tile_rows, tile_cols, t, hj, 0, 48)                    # tile-only bound: blind to it
            assert T.fused_reach_ok(code, t, tile_rows, tile_cols, t, hj, 0, 48, fresh_below=t) == expect_strip_ok
        # brute force: poison everything below the fresh rows and compare T literal iterations, strip vs full domain
        rng = np.random.default_rng(int(stray_inflow))
        p0 = rng.uniform(-1, 1, (X, Y)).astype(np.float32)
        v0 = np.zeros((X, Y, 2), np.float32)
        t = 3
        full, part = p0.copy(), p0.copy()
        part[48 + t:] = np.nan
        for a in (full, part):
            for _ in range(t):
                orc.p_bc(a, mask)
                b = a.copy()
                orc.jacobi_sweep(b, a, v0, mask, 0.01, 1.0 / 64)
                a[...] = b
        same = np.array_equal(full[:48][mask[:48] != 1], part[:48][mask[:48] != 1], equal_nan=False)
        assert same == expect_strip_ok


# ---------------------------------------------------------------------------------------------------------------------
# a rank that builds only its strip of the scene (build_scene(rows=...), BoundaryCondition(row_offset=...)) must end up
# with exactly the tables of a build from the global arrays
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num,X,Y,world,halo", [(2, 320, 64, 3, 9), (5, 384, 96, 2, 13), (3, 400, 80, 4, 9), (1, 300, 96, 2, 4),
                                                (4, 96, 48, 2, 2), (5, 256, 128, 1, 0)])
def test_strip_built_boundary_condition_equals_global_build(num, X, Y, world, halo):
    from fs.boundary_condition import BoundaryCondition, DyeBoundaryCondition

    const, mask, dye = build_scene(num, X, Y, with_dye=True)
    for rank in range(world):
        part = Partition(X, rank, world, halo)
        a, b = BoundaryCondition.strip_rows(part)
        assert 0 <= a <= part.owned()[0] and part.owned()[1] <= b <= X and (b - a < X or world == 1 or X < 200)
        c_w, m_w, d_w = build_scene(num, X, Y, with_dye=True, rows=(a, b))
        assert np.array_equal(m_w, mask[a:b]) and np.array_equal(c_w, const[a:b]) and np.array_equal(d_w, dye[a:b])
        glob = DyeBoundaryCondition(const, dye, mask, device="cpu", partition=part)
        loc = DyeBoundaryCondition(c_w, d_w, m_w, device="cpu", partition=part, row_offset=a)
        for name in ("_bc_mask", "_pcode", "_bc_const", "_bc_dye", "_dye_tgt", "_exposed_stale"):
            assert torch.equal(getattr(glob, name), getattr(loc, name)), f"rank {rank}: {name} differs"
        for tname in ("_vel_table", "_p_table"):
            tg, tl = getattr(glob, tname), getattr(loc, tname)
            for k, v in tg.items():
                if isinstance(v, torch.Tensor):
                    assert torch.equal(v, tl[k]), f"rank {rank}: {tname}[{k}] differs"
                elif isinstance(v, dict):
                    for kk, vv in v.items():
                        assert (torch.equal(vv, tl[k][kk]) if isinstance(vv, torch.Tensor) else vv == tl[k][kk]), f"{tname}[{k}][{kk}]"
                else:
                    assert v == tl[k], f"rank {rank}: {tname}[{k}] {v} != {tl[k]}"
        assert [getattr(glob.dom, f) for f, _ in glob.dom._fields_] == [getattr(loc.dom, f) for f, _ in loc.dom._fields_]
        if Y % 16 == 0:
            for T in (3, 8) if rank % 2 == 0 else (12,):
                if world == 1 or halo >= T + 1:
                    assert glob.fused_ok(T) == loc.fused_ok(T), f"rank {rank}: fused_ok({T}) differs"
    with pytest.raises(ValueError):     # arrays that do not cover the rows the rank needs
        part = Partition(X, 0, max(world, 2), max(halo, 2))
        BoundaryCondition(const[5:], mask[5:], device="cpu", partition=part, row_offset=5)
