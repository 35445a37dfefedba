"""Differential test of every dense C-ABI entry point: the kernel SOURCES on the CUDA emulation (tests/cuda_emu) against the
oracle-backed model of the fs2d_dom contract (tests/fake_fs2d.py) -- random grid sizes (any Y, also odd), random masks,
random row windows [r0, r1) with clamp bounds [clo, chi] anywhere inside the local array (what the ranks of a strip
decomposition and the overlap scheme pass), power-of-two and general dx, all outputs bit for bit, and every row outside
[r0, r1) untouched.  The GPU parity tests use whole grids and two hand-picked windows; this is the exhaustive side.
"""
from __future__ import annotations

import ctypes
import sys

import numpy as np
import pytest
import torch
from conftest import REPO, assert_bitexact

sys.path.insert(0, str(REPO / "tests" / "cuda_emu"))


@pytest.fixture(scope="module")
def libs():
    import build_emu
    from fake_fs2d import FakeFs2d
    from fs import _lib

    emu = ctypes.CDLL(str(build_emu.build()))
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(emu, name)
        fn.restype, fn.argtypes = res, args
    return emu, FakeFs2d(_lib.load())


def _scene(rng, X, Y):
    mask = np.zeros((X, Y), np.uint8)
    mask[:, :2] = 1
    mask[:, -2:] = 1
    for _ in range(int(rng.integers(4, 16))):
        i, j = int(rng.integers(0, X - 3)), int(rng.integers(0, max(Y - 3, 1)))
        mask[i:i + int(rng.integers(1, 6)), j:j + int(rng.integers(1, 6))] = 1
    mask[:2, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.8, 2, mask[:2, 2:-2])
    mask[-1:, 2:-2] = np.where(rng.random((1, Y - 4)) < 0.7, 3, mask[-1:, 2:-2])
    return mask


class Case:
    """one random problem: arrays, a window, and helpers to run an entry point through both implementations"""

    def __init__(self, seed: int, libs) -> None:
        from fs import _bc_tables, _lib

        self.emu, self.fake = libs
        rng = self.rng = np.random.default_rng(seed)
        self.X, self.Y = int(rng.integers(20, 90)), int(rng.choice([16, 32, 48, 64, 80, 37, 50, 66, 21]))
        X, Y = self.X, self.Y
        self.mask = _scene(rng, X, Y)
        clo = int(rng.integers(0, 4))
        chi = X - 1 - int(rng.integers(0, 4))
        r0 = int(rng.integers(clo, clo + 6))
        r1 = int(rng.integers(max(r0, chi - 6), chi + 2))
        if seed % 4 == 0:
            clo, chi, r0, r1 = 0, X - 1, 0, X                 # the whole grid every fourth case
        if seed % 7 == 3:
            r1 = min(r0 + int(rng.integers(0, 3)), chi + 1)   # 0-2 row windows (edge windows of the overlap scheme)
        self.dom = _lib.Dom(X, Y, r0, r1, clo, chi, int(rng.integers(0, 5)))
        self.dx = 1.0 / 64 if seed % 2 else 0.013
        self.dt, self.re = 0.05 / 64, 300.0
        # like BoundaryCondition: the rows outside the clamp window do not exist globally -- wall filler in the mask, W_NONE in
        # pcode, and the codes of the window's rows are resolved from the window's own rows
        self.mask[:clo] = 1
        self.mask[chi + 1:] = 1
        self.pcode = np.full((X, Y), _bc_tables.PC_W_NONE, np.uint8)
        win = np.ascontiguousarray(self.mask[clo:chi + 1])
        self.pcode[clo:chi + 1] = _bc_tables.pack_pcode(_bc_tables.pressure_codes(torch.from_numpy(win))).numpy()

    def f(self, *shape, scale=1.0):
        return (self.rng.uniform(-1, 1, shape) * scale).astype(np.float32)

    def run(self, name: str, arrays: dict, outs: tuple, build_args, written: dict | None = None) -> None:
        """arrays: name -> ndarray; build_args(ptrs) -> argument tuple of the entry point; written: output -> (lo, hi) rows it
        may write when that is not [r0, r1)"""
        res = []
        for impl in ("emu", "fake"):
            ts = {k: torch.from_numpy(a.copy()) for k, a in arrays.items()}
            args = build_args({k: self.fake.ptr(t) for k, t in ts.items()})
            if impl == "emu":
                rc = getattr(self.emu, name)(*args)
                assert rc == 0, self.emu.fs2d_last_error().decode()
            else:
                self.fake.call(name, *args)
            res.append({k: ts[k].numpy() for k in outs})
        d = self.dom
        tag = f"{name} {self.X}x{self.Y} rows {d.r0}:{d.r1} clamp {d.clo}:{d.chi} dx={self.dx}"
        for k in outs:
            assert_bitexact(f"{tag}: {k}", res[0][k], res[1][k])
            lo, hi = (written or {}).get(k, (d.r0, d.r1))
            untouched = np.ones(self.X, bool)
            untouched[lo:hi] = False
            assert_bitexact(f"{tag}: {k} rows outside the window", res[0][k][untouched], arrays[k][untouched])


@pytest.mark.parametrize("seed", range(24))
def test_dense_entry_points_against_the_window_model(libs, seed):
    c = Case(seed, libs)
    X, Y, d, m = c.X, c.Y, c.dom, c.mask
    v, p = c.f(X, Y, 2), c.f(X, Y)
    fx, fy = c.f(X, Y, 2, scale=3.0), c.f(X, Y, 2, scale=3.0)
    dye, dyx, dyy = np.abs(c.f(X, Y, 3)), c.f(X, Y, 3, scale=3.0), c.f(X, Y, 3, scale=3.0)
    o2a, o2b, o2c, o1a, o1b = c.f(X, Y, 2), c.f(X, Y, 2), c.f(X, Y, 2), c.f(X, Y), np.abs(c.f(X, Y))
    o3a, o3b, o3c = c.f(X, Y, 3), c.f(X, Y, 3), c.f(X, Y, 3)
    dt, dx, re = c.dt, c.dx, c.re

    for scheme in (0, 1):
        c.run("fs2d_mac_update", dict(vn=o2a, vc=v, pc=p, mask=m), ("vn",),
              lambda q: (q["vn"], q["vc"], q["pc"], q["mask"], d, dt, dx, re, scheme, None))
        c.run("fs2d_dye_mac", dict(dn=o3a, dc=dye, vc=v, mask=m), ("dn",),
              lambda q: (q["dn"], q["dc"], q["vc"], q["mask"], d, dt, dx, scheme, None))
    c.run("fs2d_cip_nonadv", dict(fn=o2a, fc=v, pc=p, mask=m), ("fn",),
          lambda q: (q["fn"], q["fc"], q["pc"], q["mask"], d, dt, dx, re, None))
    c.run("fs2d_cip_nonadv_grad", dict(fxn=o2a, fyn=o2b, fxc=fx, fyc=fy, fc=v, fn=o2c, mask=m), ("fxn", "fyn"),
          lambda q: (q["fxn"], q["fyn"], q["fxc"], q["fyc"], q["fc"], q["fn"], q["mask"], d, 2.0 * dx, None))
    c.run("fs2d_cip_advect", dict(fn=o2a, fxn=o2b, fyn=o2c, fc=v, fxc=fx, fyc=fy, mask=m), ("fn", "fxn", "fyn"),
          lambda q: (q["fn"], q["fxn"], q["fyn"], q["fc"], q["fxc"], q["fyc"], q["fc"], q["mask"], d, dt, dx, dx**2, dx**3, None))
    c.run("fs2d_set_grad", dict(fx=o2a, fy=o2b, f=v), ("fx", "fy"), lambda q: (q["fx"], q["fy"], q["f"], d, dx, None))
    c.run("fs2d_vort_calc", dict(w=o1a, wabs=o1b, vc=v, mask=m), ("w", "wabs"),
          lambda q: (q["w"], q["wabs"], q["vc"], q["mask"], d, dx, None))
    c.run("fs2d_vort_add", dict(vn=o2a, vc=v, w=o1a, wabs=o1b, mask=m), ("vn",),
          lambda q: (q["vn"], q["vc"], q["w"], q["wabs"], q["mask"], d, dx, dt * 5.0, None))
    c.run("fs2d_vort_apply", dict(vn=o2a, w=o1a, wabs=o1b, vc=v, mask=m), ("vn", "w", "wabs"),
          lambda q: (q["vn"], q["w"], q["wabs"], q["vc"], q["mask"], d, dx, dt * 5.0, None))
    c.run("fs2d_limit", dict(v=v * np.float32(14.0)), ("v",), lambda q: (q["v"], d, 10.0, None))
    c.run("fs2d_clamp", dict(f=c.f(X, Y, 3, scale=2.0)), ("f",), lambda q: (q["f"], d, 3, 0.0, 1.0, None))
    c.run("fs2d_dye_nonadv", dict(dn=o3a, dc=dye, mask=m), ("dn",), lambda q: (q["dn"], q["dc"], q["mask"], d, dt, dx, re, None))
    c.run("fs2d_dye_nonadv_grad", dict(fxn=o3a, fyn=o3b, fxc=dyx, fyc=dyy, fc=dye, fn=o3c, mask=m), ("fxn", "fyn"),
          lambda q: (q["fxn"], q["fyn"], q["fxc"], q["fyc"], q["fc"], q["fn"], q["mask"], d, 2.0 * dx, None))
    c.run("fs2d_dye_cip_advect", dict(fn=o3a, fxn=o3b, fyn=o3c, fc=dye, fxc=dyx, fyc=dyy, v=v, mask=m), ("fn", "fxn", "fyn"),
          lambda q: (q["fn"], q["fxn"], q["fyn"], q["fc"], q["fxc"], q["fyc"], q["v"], q["mask"], d, dt, dx, dx**2, dx**3, None))
    c.run("fs2d_dye_set_grad", dict(fx=o3a, fy=o3b, f=dye), ("fx", "fy"), lambda q: (q["fx"], q["fy"], q["f"], d, dx, None))
    c.run("fs2d_render", dict(rgb=o3a, v=v, p=p, dye=dye, mask=m), ("rgb",),
          lambda q: (q["rgb"], q["v"], q["p"], q["dye"], q["mask"], d, dx, seed % 4, None))


@pytest.mark.parametrize("seed", range(24))
def test_pressure_entry_points_against_the_window_model(libs, seed):
    """source pre-pass + one sweep (plain and inline-BC) and the two colour passes of RB-SOR on random windows"""
    c = Case(100 + seed, libs)
    X, Y, d, m = c.X, c.Y, c.dom, c.mask
    v, p, pn = c.f(X, Y, 2), c.f(X, Y), c.f(X, Y)
    dt, dx = c.dt, c.dx
    # the source array of the real library holds (t2, t3), the model's holds the velocity rows: run source + sweep as one unit
    wide = d.replace(r0=max(d.r0 - 1, d.clo), r1=min(d.r1 + 1, d.chi + 1))   # the sweep reads the source of rows [r0, r1) only
    for inline_bc in (0, 1):
        res = []
        for impl in ("emu", "fake"):
            ts = {k: torch.from_numpy(a.copy()) for k, a in dict(pn=pn, pc=p, src=np.zeros((X, Y, 2), np.float32), vc=v, pcode=c.pcode).items()}
            q = {k: c.fake.ptr(t) for k, t in ts.items()}
            for name, args in (("fs2d_pressure_source", (q["src"], q["vc"], d, dt, dx, None)),
                               ("fs2d_jacobi_sweep", (q["pn"], q["pc"], q["src"], q["pcode"], d, inline_bc, None))):
                if impl == "emu":
                    assert getattr(c.emu, name)(*args) == 0, c.emu.fs2d_last_error().decode()
                else:
                    c.fake.call(name, *args)
            res.append(ts["pn"].numpy())
        assert_bitexact(f"jacobi sweep inline_bc={inline_bc} {X}x{Y} rows {d.r0}:{d.r1} clamp {d.clo}:{d.chi}", res[0], res[1])
        untouched = np.ones(X, bool)
        untouched[d.r0:d.r1] = False
        assert_bitexact("rows outside the window", res[0][untouched], pn[untouched])
    for parity in (1, 0):
        res = []
        for impl in ("emu", "fake"):
            ts = {k: torch.from_numpy(a.copy()) for k, a in dict(pn=pn, pc=p, src=np.zeros((X, Y, 2), np.float32), vc=v, mask=m).items()}
            q = {k: c.fake.ptr(t) for k, t in ts.items()}
            pc_arg = q["pc"] if parity == 1 else q["pn"]        # the even pass runs in place on pn (fs/pressure_updater.py:96)
            for name, args in (("fs2d_pressure_source", (q["src"], q["vc"], d, dt, dx, None)),
                               ("fs2d_rbsor_pass", (q["pn"], pc_arg, q["src"], q["mask"], d, 1.3, 1.0 - 1.3, parity, None))):
                if impl == "emu":
                    assert getattr(c.emu, name)(*args) == 0, c.emu.fs2d_last_error().decode()
                else:
                    c.fake.call(name, *args)
            res.append(ts["pn"].numpy())
        assert_bitexact(f"rbsor parity {parity} {X}x{Y} rows {d.r0}:{d.r1} gi0 {d.gi0}", res[0], res[1])
    res = []     # both colours in one kernel (fs2d_rbsor_iteration) vs the model's two passes
    for impl in ("emu", "fake"):
        ts = {k: torch.from_numpy(a.copy()) for k, a in dict(pn=pn, pc=p, src=np.zeros((X, Y, 2), np.float32), vc=v, mask=m).items()}
        q = {k: c.fake.ptr(t) for k, t in ts.items()}
        for name, args in (("fs2d_pressure_source", (q["src"], q["vc"], d, dt, dx, None)),
                           ("fs2d_rbsor_iteration", (q["pn"], q["pc"], q["src"], q["mask"], d, 1.3, 1.0 - 1.3, None))):
            if impl == "emu":
                assert getattr(c.emu, name)(*args) == 0, c.emu.fs2d_last_error().decode()
            else:
                c.fake.call(name, *args)
        res.append(ts["pn"].numpy())
    assert_bitexact(f"rbsor iteration {X}x{Y} rows {d.r0}:{d.r1} gi0 {d.gi0}", res[0], res[1])
    del wide
