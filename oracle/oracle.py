"""ctypes front-end of the CPU oracle (`fs2d_oracle.c`) + the reference's step orchestration.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.

`OracleSolver.update()` restates `CipMacSolver.update()` (/root/reference/fs/solver.py:192-227) and
`MacSolver.update()` (:79-89) including the physical double-buffer identities and swap counts
(SURVEY T1); `JacobiPressureUpdater.update` (fs/pressure_updater.py:56-60) and
`RedBlackSorPressureUpdater.update/_update` (:86-96).  Host-side constant folding follows SURVEY 8(c):
Python doubles folded first, then cast to fp32.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libfs2d_oracle.so"
_SRC = _HERE / "fs2d_oracle.c"
_LIB = None

VELOCITY_LIMIT = 10.0  # fs/solver.py:12


def build(force: bool = False) -> Path:
    """gcc-compile the C restatement (no FMA contraction, no fast-math)."""
    if force or not _SO.exists() or _SO.stat().st_mtime < _SRC.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        cmd = ["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC",
               "-shared", str(_SRC), "-o", str(_SO), "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        build()
        _LIB = ctypes.CDLL(str(_SO))
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


_f = ctypes.c_float
_i = ctypes.c_int


def f32(x: float) -> np.float32:
    return np.float32(x)


# ----------------------------------------------------------------------------- kernels
def vel_bc(v, mask, bc_const, target_centric: bool = False) -> None:
    X, Y = mask.shape
    fn = lib().orc_vel_bc_tc if target_centric else lib().orc_vel_bc
    fn(_p(v), _p(mask), _p(bc_const), _i(X), _i(Y))


def p_bc(p, mask, tmp=None) -> None:
    X, Y = mask.shape
    tmp = np.empty_like(p) if tmp is None else tmp
    lib().orc_p_bc(_p(p), _p(mask), _i(X), _i(Y), _p(tmp))


def mac_update(vn, vc, pc, mask, dt, dx, re, scheme: str) -> None:
    X, Y = mask.shape
    lib().orc_mac_update(_p(vn), _p(vc), _p(pc), _p(mask), _i(X), _i(Y), _f(f32(dt)), _f(f32(dx)), _f(f32(re)),
                         _i({"upwind": 0, "kk": 1}[scheme]))


def cip_nonadv(fn, fc, pc, mask, dt, dx, re) -> None:
    X, Y = mask.shape
    lib().orc_cip_nonadv(_p(fn), _p(fc), _p(pc), _p(mask), _i(X), _i(Y), _f(f32(dt)), _f(f32(dx)), _f(f32(re)))


def cip_nonadv_grad(fxn, fyn, fxc, fyc, fc, fn, mask, dx) -> None:
    X, Y = mask.shape
    lib().orc_cip_nonadv_grad(_p(fxn), _p(fyn), _p(fxc), _p(fyc), _p(fc), _p(fn), _p(mask), _i(X), _i(Y),
                              _f(f32(2.0 * dx)))


def cip_advect(fn, fxn, fyn, fc, fxc, fyc, v, mask, dt, dx) -> None:
    X, Y = mask.shape
    lib().orc_cip_advect(_p(fn), _p(fxn), _p(fyn), _p(fc), _p(fxc), _p(fyc), _p(v), _p(mask), _i(X), _i(Y),
                         _f(f32(dt)), _f(f32(dx)), _f(f32(dx**2)), _f(f32(dx**3)))


def set_grad(fx, fy, f, dx) -> None:
    X, Y = f.shape[:2]
    lib().orc_set_grad(_p(fx), _p(fy), _p(f), _i(X), _i(Y), _f(f32(dx)))


def vort_calc(w, wabs, vc, mask, dx) -> None:
    X, Y = mask.shape
    lib().orc_vort_calc(_p(w), _p(wabs), _p(vc), _p(mask), _i(X), _i(Y), _f(f32(dx)))


def vort_add(vn, vc, w, wabs, mask, dx, dt, weight) -> None:
    X, Y = mask.shape
    lib().orc_vort_add(_p(vn), _p(vc), _p(w), _p(wabs), _p(mask), _i(X), _i(Y), _f(f32(dx)), _f(f32(dt * weight)))


def jacobi_sweep(pn, pc, vc, mask, dt, dx) -> None:
    X, Y = mask.shape
    lib().orc_jacobi_sweep(_p(pn), _p(pc), _p(vc), _p(mask), _i(X), _i(Y), _f(f32(dt)), _f(f32(dx)))


def rbsor_pass(pn, pc, vc, mask, dt, dx, omega, parity) -> None:
    X, Y = mask.shape
    lib().orc_rbsor_pass(_p(pn), _p(pc), _p(vc), _p(mask), _i(X), _i(Y), _f(f32(dt)), _f(f32(dx)),
                         _f(f32(omega)), _f(f32(1.0 - omega)), _i(parity))


def limit(v, lim=VELOCITY_LIMIT) -> None:
    X, Y = v.shape[:2]
    lib().orc_limit(_p(v), _i(X), _i(Y), _f(f32(lim)))


# ----------------------------------------------------------------------------- dye kernels (C = 3)
def dye_bc(dye, bc_dye, mask) -> None:
    X, Y = mask.shape
    lib().orc_dye_bc(_p(dye), _p(bc_dye), _p(mask), _i(X), _i(Y))


def clamp(f, low=0.0, high=1.0) -> None:
    X, Y, C = f.shape
    lib().orc_clamp(_p(f), _i(X), _i(Y), _i(C), _f(f32(low)), _f(f32(high)))


def dye_mac(dn, dc, vc, mask, dt, dx, scheme: str) -> None:
    X, Y = mask.shape
    lib().orc_dye_mac(_p(dn), _p(dc), _p(vc), _p(mask), _i(X), _i(Y), _i(dn.shape[2]), _f(f32(dt)), _f(f32(dx)),
                      _i({"upwind": 0, "kk": 1}[scheme]))


def dye_nonadv(dn, dc, mask, dt, dx, re) -> None:
    X, Y = mask.shape
    lib().orc_dye_nonadv(_p(dn), _p(dc), _p(mask), _i(X), _i(Y), _i(dn.shape[2]), _f(f32(dt)), _f(f32(dx)), _f(f32(re)))


def cip_nonadv_grad_n(fxn, fyn, fxc, fyc, fc, fn, mask, dx) -> None:
    X, Y = mask.shape
    lib().orc_cip_nonadv_grad_n(_p(fxn), _p(fyn), _p(fxc), _p(fyc), _p(fc), _p(fn), _p(mask), _i(X), _i(Y),
                                _i(fc.shape[2]), _f(f32(2.0 * dx)))


def cip_advect_n(fn, fxn, fyn, fc, fxc, fyc, v, mask, dt, dx) -> None:
    X, Y = mask.shape
    lib().orc_cip_advect_n(_p(fn), _p(fxn), _p(fyn), _p(fc), _p(fxc), _p(fyc), _p(v), _p(mask), _i(X), _i(Y),
                           _i(fc.shape[2]), _f(f32(dt)), _f(f32(dx)), _f(f32(dx**2)), _f(f32(dx**3)))


def set_grad_n(fx, fy, f, dx) -> None:
    X, Y, C = f.shape
    lib().orc_set_grad_n(_p(fx), _p(fy), _p(f), _i(X), _i(Y), _i(C), _f(f32(dx)))


def render(v, p, dye, mask, dx, mode: str) -> np.ndarray:
    X, Y = mask.shape
    rgb = np.zeros((X, Y, 3), dtype=np.float32)
    lib().orc_render(_p(rgb), _p(v), _p(p), _p(dye) if dye is not None else None, _p(mask), _i(X), _i(Y), _f(f32(dx)),
                     _i({"norm": 0, "pressure": 1, "vorticity": 2, "dye": 3}[mode]))
    return rgb


def set_threads(n: int) -> int:
    """Size of the OpenMP team of the oracle kernels, set through the runtime (an OMP_NUM_THREADS read at load time cannot
    be trusted: torch.distributed.run exports OMP_NUM_THREADS=1 to every rank).  Returns the team size in effect."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    lib().orc_set_threads(_i(int(n)))
    return int(lib().orc_max_threads())


# ----------------------------------------------------------------------------- orchestration
class Buf:
    """fs/double_buffer.py:4-18 -- two physical arrays + reference swap."""

    def __init__(self, shape) -> None:
        self.current = np.zeros(shape, dtype=np.float32)
        self.next = np.zeros(shape, dtype=np.float32)

    def swap(self) -> None:
        self.current, self.next = self.next, self.current


class OracleSolver:
    def __init__(self, mask, bc_const, dt, dx, re, scheme="cip", vc=None, pressure=("jacobi", 2), bc_dye=None):
        self.mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self.bc_const = np.ascontiguousarray(bc_const, dtype=np.float32)
        self.dt, self.dx, self.re, self.scheme, self.vc, self.pressure = dt, dx, re, scheme, vc, pressure
        X, Y = self.mask.shape
        self.resolution = (X, Y)
        self.v = Buf((X, Y, 2))
        self.p = Buf((X, Y))
        self._tmp = np.empty((X, Y), dtype=np.float32)
        if scheme == "cip":
            self.vx, self.vy = Buf((X, Y, 2)), Buf((X, Y, 2))
            set_grad(self.vx.current, self.vy.current, self.v.current, dx)  # solver.py:190
        if vc is not None:
            self.vort = np.zeros((X, Y), dtype=np.float32)
            self.vort_abs = np.zeros((X, Y), dtype=np.float32)
        # dye variants: DyeMacSolver (fs/solver.py:110-161), DyeCipMacSolver (:335-401)
        self.bc_dye = None if bc_dye is None else np.ascontiguousarray(bc_dye, dtype=np.float32)
        if self.bc_dye is not None:
            self.dye = Buf((X, Y, 3))
            if scheme == "cip":
                self.dyex, self.dyey = Buf((X, Y, 3)), Buf((X, Y, 3))
                set_grad_n(self.dyex.current, self.dyey.current, self.dye.current, dx)  # solver.py:351

    # fs/pressure_updater.py:56-60 / :86-96
    def _pressure_update(self) -> None:
        p, v = self.p, self.v.current
        if self.pressure[0] == "jacobi":
            for _ in range(self.pressure[1]):
                p_bc(p.current, self.mask, self._tmp)
                jacobi_sweep(p.next, p.current, v, self.mask, self.dt, self.dx)
                p.swap()
        else:
            _, omega, n_iter = self.pressure
            for _ in range(n_iter):
                p_bc(p.current, self.mask, self._tmp)
                rbsor_pass(p.next, p.current, v, self.mask, self.dt, self.dx, omega, 1)
                rbsor_pass(p.next, p.next, v, self.mask, self.dt, self.dx, omega, 0)
                p.swap()

    def _vc_apply(self) -> None:
        # fs/vorticity_confinement.py:57-59 (+ caller's swap, solver.py:85-86 / :197-198)
        vort_calc(self.vort, self.vort_abs, self.v.current, self.mask, self.dx)
        vort_add(self.v.next, self.v.current, self.vort, self.vort_abs, self.mask, self.dx, self.dt, self.vc)
        self.v.swap()

    def update(self) -> None:
        m = self.mask
        vel_bc(self.v.current, m, self.bc_const, target_centric=True)
        if self.scheme == "cip":
            v, vx, vy, p = self.v, self.vx, self.vy, self.p
            cip_nonadv(v.next, v.current, p.current, m, self.dt, self.dx, self.re)
            cip_nonadv_grad(vx.next, vy.next, vx.current, vy.current, v.current, v.next, m, self.dx)
            v.swap(); vx.swap(); vy.swap()
            cip_advect(v.next, vx.next, vy.next, v.current, vx.current, vy.current, v.current, m, self.dt, self.dx)
            v.swap(); vx.swap(); vy.swap()
        else:
            mac_update(self.v.next, self.v.current, self.p.current, m, self.dt, self.dx, self.re, self.scheme)
            self.v.swap()
        if self.vc is not None:
            self._vc_apply()
        self._pressure_update()
        limit(self.v.current)
        if self.bc_dye is not None:
            self._update_dye()

    def _update_dye(self) -> None:
        m, dye = self.mask, self.dye
        dye_bc(dye.current, self.bc_dye, m)                                   # solver.py:149 / :366
        if self.scheme == "cip":                                              # solver.py:385-401
            dx_, dy_ = self.dyex, self.dyey
            dye_nonadv(dye.next, dye.current, m, self.dt, self.dx, self.re)
            cip_nonadv_grad_n(dx_.next, dy_.next, dx_.current, dy_.current, dye.current, dye.next, m, self.dx)
            dye.swap(); dx_.swap(); dy_.swap()
            cip_advect_n(dye.next, dx_.next, dy_.next, dye.current, dx_.current, dy_.current, self.v.current, m,
                         self.dt, self.dx)
            dye.swap(); dx_.swap(); dy_.swap()
        else:                                                                 # solver.py:150-151
            dye_mac(dye.next, dye.current, self.v.current, m, self.dt, self.dx, self.scheme)
            dye.swap()
        clamp(dye.current, 0.0, 1.0)                                          # solver.py:152 / :373

    def state(self) -> dict:
        d = {"v_cur": self.v.current, "v_nxt": self.v.next, "p_cur": self.p.current, "p_nxt": self.p.next}
        if self.scheme == "cip":
            d.update(vx_cur=self.vx.current, vx_nxt=self.vx.next, vy_cur=self.vy.current, vy_nxt=self.vy.next)
        if self.vc is not None:
            d.update(vort=self.vort, vort_abs=self.vort_abs)
        if self.bc_dye is not None:
            d.update(dye_cur=self.dye.current, dye_nxt=self.dye.next)
            if self.scheme == "cip":
                d.update(dyex_cur=self.dyex.current, dyex_nxt=self.dyex.next, dyey_cur=self.dyey.current,
                         dyey_nxt=self.dyey.next)
        return d

    def load_state(self, st: dict) -> None:
        for k, a in st.items():
            if k in ("vort", "vort_abs"):
                getattr(self, k)[...] = a
            else:
                name, which = k.split("_")
                getattr(getattr(self, name), "current" if which == "cur" else "next")[...] = a
