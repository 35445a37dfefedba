"""CPU stand-in for libfs2d.so, for HOST-LOGIC tests only (tests/test_host_logic.py).

The product has no CPU path: `fs._lib` refuses CPU tensors.  To exercise the Python host layer without a GPU -- the
solver orchestration and swap counts, the sparse BC tables, the Jacobi schedule, and above all the multi-rank strip
logic of fs/halo.py under gloo -- the tests install this object in place of the three `fs._lib` hooks (`call`, `ptr`,
`stream`).  Every fs2d_* entry point is then executed by the CPU oracle on NumPy views of CPU tensors, honouring the
fs2d_dom contract of include/fs2d.h (update rows [r0, r1), clamp reads to rows [clo, chi]).  Pure host functions of the
real library (fs2d_jacobi_plan, fs2d_fused_tile) are passed through to it.

This says NOTHING about the CUDA kernels (tests/test_gpu_parity.py does that on a B200); it checks that the host layer
issues the right calls, on the right physical buffers and row windows, with the right halo exchanges in between.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from oracle import oracle as orc

_f, _i, _p = ctypes.c_float, ctypes.c_int, orc._p

PC_FLUID, PC_W_NONE, PC_INFLOW, PC_OUTFLOW = 0, 9, 10, 11


def mask_from_pcode(pcode: np.ndarray) -> np.ndarray:
    """cell types back from the pressure codes (include/fs2d.h FS2D_PC_*): checks that pcode carries the mask"""
    code = pcode & 15
    m = np.ones(code.shape, dtype=np.uint8)
    m[code == PC_FLUID] = 0
    m[code == PC_INFLOW] = 2
    m[code == PC_OUTFLOW] = 3
    return m


class FakeFs2d:
    def __init__(self, real_lib) -> None:
        self.real = real_lib
        self.tensors: dict[int, torch.Tensor] = {}
        self.src_params: dict[int, tuple[float, float]] = {}
        self.trace: list[str] = []

    # ---- the three hooks -----------------------------------------------------------------------------------------
    def ptr(self, t: torch.Tensor | None):
        if t is None:
            return None
        assert not t.is_cuda and t.is_contiguous()
        self.tensors[t.data_ptr()] = t
        return t.data_ptr()

    @staticmethod
    def stream() -> int:
        return 0

    def call(self, name: str, *args) -> None:
        self.trace.append(name)
        if name in ("fs2d_jacobi_plan", "fs2d_fused_tile", "fs2d_set_tuning"):
            from fs import _lib

            _lib.check(getattr(self.real, name)(*args))
            return
        getattr(self, name)(*args)

    def install(self, monkeypatch) -> "FakeFs2d":
        from fs import _lib

        monkeypatch.setattr(_lib, "call", self.call)
        monkeypatch.setattr(_lib, "ptr", self.ptr)
        monkeypatch.setattr(_lib, "stream", self.stream)
        return self

    def install_plain(self) -> "FakeFs2d":
        """for spawned worker processes (no monkeypatch fixture)"""
        from fs import _lib

        _lib.call, _lib.ptr, _lib.stream = self.call, self.ptr, self.stream
        return self

    # ---- helpers -------------------------------------------------------------------------------------------------
    def a(self, ptr, d=None, chan: int | None = None) -> np.ndarray:
        """NumPy view (shared memory) of the tensor registered under `ptr`"""
        t = self.tensors[ptr]
        arr = t.numpy()
        if d is not None:
            assert arr.shape[0] == d.rows and arr.shape[1] == d.Y, (arr.shape, d.rows, d.Y)
        if chan is not None:
            assert (arr.shape[2] if arr.ndim == 3 else 1) == chan, (arr.shape, chan)
        return arr

    @staticmethod
    def _windowed(d, outs, body) -> None:
        """Run `body(window_slice, *window_copies_of_outs)` on rows [clo, chi] and commit rows [r0, r1) of the copies."""
        assert d.clo <= d.r0 <= d.r1 <= d.chi + 1, (d.clo, d.r0, d.r1, d.chi)
        w = slice(d.clo, d.chi + 1)
        tmps = [np.ascontiguousarray(o[w]).copy() for o in outs]
        body(w, *tmps)
        a, b = d.r0 - d.clo, d.r1 - d.clo
        for o, t in zip(outs, tmps):
            o[d.r0:d.r1] = t[a:b]

    @staticmethod
    def _c(x: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(x)

    # ---- sparse boundary conditions (tables of fs/_bc_tables.py) ------------------------------------------------------
    def fs2d_vel_bc(self, v, bc_const, tgt, src, kind, scratch, n, stream) -> None:
        if n == 0:
            return
        V, C = self.a(v).reshape(-1, 2), self.a(bc_const).reshape(-1, 2)
        t, s, k = (self.a(x)[:n].astype(np.int64) for x in (tgt, src, kind))
        assert len(np.unique(t)) == n, "velocity-BC targets must be unique (gather form)"
        out = V[t].copy()
        out[k == 0] = -V[s[k == 0]]
        out[k == 1] = C[t[k == 1]]
        out[k == 2, 0] = np.fmax(V[s[k == 2], 0], np.float32(0.05))
        V[t] = out

    def fs2d_pressure_bc(self, p, tgt, src0, src1, kind, scratch, n, stream) -> None:
        if n == 0:
            return
        P = self.a(p).reshape(-1)
        t, s0, s1, k = (self.a(x)[:n].astype(np.int64) for x in (tgt, src0, src1, kind))
        assert len(np.unique(t)) == n
        out = np.zeros(n, dtype=np.float32)
        out[k == 0] = P[s0[k == 0]]
        out[k == 1] = (P[s0[k == 1]] + P[s1[k == 1]]) / np.float32(2.0)
        P[t] = out

    def fs2d_dye_bc(self, dye, bc_dye, tgt, n, stream) -> None:
        if n == 0:
            return
        t = self.a(tgt)[:n].astype(np.int64)
        self.a(dye).reshape(-1, 3)[t] = self.a(bc_dye).reshape(-1, 3)[t]

    # ---- dense kernels -----------------------------------------------------------------------------------------------
    def fs2d_mac_update(self, vn, vc, pc, mask, d, dt, dx, re, scheme, stream) -> None:
        VC, PC, M = self.a(vc, d, 2), self.a(pc, d, 1), self.a(mask, d)
        self._windowed(d, [self.a(vn, d, 2)], lambda w, o: orc.lib().orc_mac_update(
            _p(o), _p(VC[w]), _p(PC[w]), _p(M[w]), _i(o.shape[0]), _i(d.Y), _f(dt), _f(dx), _f(re), _i(scheme)))

    def fs2d_cip_nonadv(self, fn, fc, pc, mask, d, dt, dx, re, stream) -> None:
        FC, PC, M = self.a(fc, d, 2), self.a(pc, d, 1), self.a(mask, d)
        self._windowed(d, [self.a(fn, d, 2)], lambda w, o: orc.lib().orc_cip_nonadv(
            _p(o), _p(FC[w]), _p(PC[w]), _p(M[w]), _i(o.shape[0]), _i(d.Y), _f(dt), _f(dx), _f(re)))

    def fs2d_cip_nonadv_grad(self, fxn, fyn, fxc, fyc, fc, fn, mask, d, two_dx, stream) -> None:
        XC, YC, FC, FN, M = self.a(fxc, d, 2), self.a(fyc, d, 2), self.a(fc, d, 2), self.a(fn, d, 2), self.a(mask, d)
        self._windowed(d, [self.a(fxn, d, 2), self.a(fyn, d, 2)], lambda w, ox, oy: orc.lib().orc_cip_nonadv_grad(
            _p(ox), _p(oy), _p(XC[w]), _p(YC[w]), _p(FC[w]), _p(FN[w]), _p(M[w]), _i(ox.shape[0]), _i(d.Y), _f(two_dx)))

    def fs2d_cip_advect(self, fn, fxn, fyn, fc, fxc, fyc, v, mask, d, dt, dx, dx2, dx3, stream) -> None:
        FC, XC, YC, V, M = self.a(fc, d, 2), self.a(fxc, d, 2), self.a(fyc, d, 2), self.a(v, d, 2), self.a(mask, d)
        self._windowed(d, [self.a(fn, d, 2), self.a(fxn, d, 2), self.a(fyn, d, 2)], lambda w, o, ox, oy: orc.lib().orc_cip_advect(
            _p(o), _p(ox), _p(oy), _p(FC[w]), _p(XC[w]), _p(YC[w]), _p(V[w]), _p(M[w]), _i(o.shape[0]), _i(d.Y), _f(dt),
            _f(dx), _f(dx2), _f(dx3)))

    def fs2d_set_grad(self, fx, fy, f, d, dx, stream) -> None:
        F = self.a(f, d, 2)
        self._windowed(d, [self.a(fx, d, 2), self.a(fy, d, 2)], lambda w, ox, oy: orc.lib().orc_set_grad(
            _p(ox), _p(oy), _p(F[w]), _i(ox.shape[0]), _i(d.Y), _f(dx)))

    def fs2d_vort_calc(self, w_, wabs, vc, mask, d, dx, stream) -> None:
        VC, M = self.a(vc, d, 2), self.a(mask, d)
        self._windowed(d, [self.a(w_, d, 1), self.a(wabs, d, 1)], lambda w, ow, oa: orc.lib().orc_vort_calc(
            _p(ow), _p(oa), _p(VC[w]), _p(M[w]), _i(ow.shape[0]), _i(d.Y), _f(dx)))

    def fs2d_vort_add(self, vn, vc, w_, wabs, mask, d, dx, dtw, stream) -> None:
        VC, W, A, M = self.a(vc, d, 2), self.a(w_, d, 1), self.a(wabs, d, 1), self.a(mask, d)
        self._windowed(d, [self.a(vn, d, 2)], lambda w, o: orc.lib().orc_vort_add(
            _p(o), _p(VC[w]), _p(W[w]), _p(A[w]), _p(M[w]), _i(o.shape[0]), _i(d.Y), _f(dx), _f(dtw)))

    def fs2d_vort_apply(self, vn, w_, wabs, vc, mask, d, dx, dtw, stream) -> None:
        """== vort_calc then vort_add; the curl of the rows next to [r0, r1) is recomputed from vc, not stored"""
        VC, M = self.a(vc, d, 2), self.a(mask, d)

        def body(w, o, ow, oa):
            n = _i(o.shape[0])
            orc.lib().orc_vort_calc(_p(ow), _p(oa), _p(VC[w]), _p(M[w]), n, _i(d.Y), _f(dx))
            orc.lib().orc_vort_add(_p(o), _p(VC[w]), _p(ow), _p(oa), _p(M[w]), n, _i(d.Y), _f(dx), _f(dtw))

        self._windowed(d, [self.a(vn, d, 2), self.a(w_, d, 1), self.a(wabs, d, 1)], body)

    def fs2d_limit(self, v, d, limit, stream) -> None:
        self._windowed(d, [self.a(v, d, 2)], lambda w, o: orc.lib().orc_limit(_p(o), _i(o.shape[0]), _i(d.Y), _f(limit)))

    def fs2d_clamp(self, f, d, channels, low, high, stream) -> None:
        self._windowed(d, [self.a(f, d, channels)], lambda w, o: orc.lib().orc_clamp(
            _p(o), _i(o.shape[0]), _i(d.Y), _i(channels), _f(low), _f(high)))

    def fs2d_render(self, rgb, v, p, dye, mask, d, dx, mode, stream) -> None:
        V, P, M = self.a(v, d, 2), self.a(p, d, 1), self.a(mask, d)
        D = self.a(dye, d, 3) if dye is not None else None
        self._windowed(d, [self.a(rgb, d, 3)], lambda w, o: orc.lib().orc_render(
            _p(o), _p(V[w]), _p(P[w]), _p(D[w]) if D is not None else None, _p(M[w]), _i(o.shape[0]), _i(d.Y), _f(dx), _i(mode)))

    # ---- pressure ----------------------------------------------------------------------------------------------------
    def fs2d_pressure_source(self, src, vc, d, dt, dx, stream) -> None:
        """The fake source array just carries the velocity rows the real kernel reads ([r0-1, r1+1) clamped); the sweeps
        evaluate predict_p from them (bit-identical by construction of k_p_source, DESIGN.md "Source pre-pass")."""
        S, V = self.a(src, d, 2), self.a(vc, d, 2)
        lo, hi = max(d.r0 - 1, d.clo), min(d.r1 + 1, d.chi + 1)
        S[lo:hi] = V[lo:hi]
        self.src_params[src] = (dt, dx)

    def _sweep_window(self, pn_w, pc_w, src_w, mask_w, src_ptr, Y) -> None:
        dt, dx = self.src_params[src_ptr]
        orc.lib().orc_jacobi_sweep(_p(pn_w), _p(self._c(pc_w)), _p(self._c(src_w)), _p(self._c(mask_w)), _i(pn_w.shape[0]), _i(Y),
                                   _f(dt), _f(dx))

    def fs2d_jacobi_sweep(self, pn, pc, src, pcode, d, inline_bc, stream) -> None:
        PC, S, M = self.a(pc, d, 1), self.a(src, d, 2), mask_from_pcode(self.a(pcode, d))

        def body(w, o):
            pcw = PC[w].copy()
            if inline_bc:
                orc.p_bc(pcw, self._c(M[w]))
            self._sweep_window(o, pcw, S[w], M[w], src, d.Y)

        self._windowed(d, [self.a(pn, d, 1)], body)

    def fs2d_jacobi_update(self, pa, pb, src, pcode, d, n_sweeps, tgt, src0, src1, kind, scratch, n_bc, fuse_mask, orders,
                           n_orders, final_in_b, stream) -> None:
        cur, nxt = pa, pb
        for _ in range(n_sweeps):   # the literal loop of fs/pressure_updater.py:56-60
            self.fs2d_pressure_bc(cur, tgt, src0, src1, kind, scratch, n_bc, stream)
            self.fs2d_jacobi_sweep(nxt, cur, src, pcode, d, 0, stream)
            cur, nxt = nxt, cur
        final_in_b._obj.value = int(cur == pb)

    def _fused_rows(self, p_out, p_in, src, pcode, d, T, row_sets, emit: bool = False) -> None:
        PC = self.a(pcode, d)
        PI, PO, S, M = self.a(p_in, d, 1), self.a(p_out, d, 1), self.a(src, d, 2), mask_from_pcode(PC)
        w = slice(d.clo, d.chi + 1)
        mw = self._c(M[w])
        cur = PI[w].copy()
        bc_of_penultimate = None
        for _ in range(T):
            orc.p_bc(cur, mw)                 # cur now holds the post-BC values of the state before this iteration
            bc_of_penultimate = cur.copy()
            nxt = cur.copy()
            self._sweep_window(nxt, cur, S[w], mw, src, d.Y)
            cur = nxt
        relaxed = mw != 1
        code = PC[w] & 15
        wall_bc = (code >= 1) & (code <= 8)
        for r0, r1 in row_sets:     # BC cells of p_out are neither read nor written (include/fs2d.h)
            a, b = r0 - d.clo, r1 - d.clo
            PO[r0:r1][relaxed[a:b]] = cur[a:b][relaxed[a:b]]
            if emit:                # fs2d_jacobi_fused_tail: BC values of the penultimate state into the wall-BC cells of p_in
                PI[r0:r1][wall_bc[a:b]] = bc_of_penultimate[a:b][wall_bc[a:b]]

    def fs2d_fused_order(self, pcode, d, T, skip_from, skip_n, order, cap, counts, stream) -> None:
        """the tile list only steers the CUDA kernel's work distribution; the stand-in computes whole row sets"""
        counts[0], counts[1], counts[2] = 1, 0, 0

    def fs2d_jacobi_fused(self, p_out, p_in, src, pcode, d, T, order, n_order, stream) -> None:
        self._fused_rows(p_out, p_in, src, pcode, d, T, [(d.r0, d.r1)])

    def fs2d_jacobi_fused_part(self, p_out, p_in, src, pcode, d, T, skip_from, skip_n, order, n_order, stream) -> None:
        rows, hr = ctypes.c_int(), ctypes.c_int()
        self.real.fs2d_fused_tile(T, ctypes.byref(rows), None, ctypes.byref(hr), None, None)
        ti = rows.value - 2 * hr.value
        k = (d.r1 - d.r0 + ti - 1) // ti
        sets = [(d.r0 + q * ti, min(d.r0 + (q + 1) * ti, d.r1)) for q in range(k) if not (skip_from <= q < skip_from + skip_n)]
        self._fused_rows(p_out, p_in, src, pcode, d, T, sets)

    def fs2d_jacobi_fused_tail(self, p_out, p_in, src, pcode, d, T, skip_from, skip_n, order, n_order, stream) -> None:
        rows, hr = ctypes.c_int(), ctypes.c_int()
        self.real.fs2d_fused_tile(T, ctypes.byref(rows), None, ctypes.byref(hr), None, None)
        ti = rows.value - 2 * hr.value
        k = (d.r1 - d.r0 + ti - 1) // ti
        sets = [(d.r0 + q * ti, min(d.r0 + (q + 1) * ti, d.r1)) for q in range(k) if not (skip_from <= q < skip_from + skip_n)]
        self._fused_rows(p_out, p_in, src, pcode, d, T, sets, emit=True)

    def fs2d_rbsor_iteration(self, pn, pc, src, mask, d, omega, one_minus_omega, stream) -> None:
        """by definition (include/fs2d.h): the odd pass pn <- pc, then the even pass pn <- pn"""
        self.fs2d_rbsor_pass(pn, pc, src, mask, d, omega, one_minus_omega, 1, stream)
        self.fs2d_rbsor_pass(pn, pn, src, mask, d, omega, one_minus_omega, 0, stream)

    def fs2d_rbsor_pass(self, pn, pc, src, mask, d, omega, one_minus_omega, parity, stream) -> None:
        PC, S, M = self.a(pc, d, 1), self.a(src, d, 2), self.a(mask, d)
        dt, dx = self.src_params[src]
        par = (parity + d.gi0 + d.clo) & 1      # colours follow the GLOBAL row index (include/fs2d.h)

        def body(w, o):
            pcw = o if pc == pn else self._c(PC[w])
            orc.lib().orc_rbsor_pass(_p(o), _p(pcw), _p(self._c(S[w])), _p(self._c(M[w])), _i(o.shape[0]), _i(d.Y), _f(dt), _f(dx),
                                     _f(omega), _f(one_minus_omega), _i(par))

        self._windowed(d, [self.a(pn, d, 1)], body)

    # ---- dye (3 channels) --------------------------------------------------------------------------------------------
    def fs2d_dye_mac(self, dn, dc, vc, mask, d, dt, dx, scheme, stream) -> None:
        DC, VC, M = self.a(dc, d, 3), self.a(vc, d, 2), self.a(mask, d)
        self._windowed(d, [self.a(dn, d, 3)], lambda w, o: orc.lib().orc_dye_mac(
            _p(o), _p(DC[w]), _p(VC[w]), _p(M[w]), _i(o.shape[0]), _i(d.Y), _i(3), _f(dt), _f(dx), _i(scheme)))

    def fs2d_dye_nonadv(self, dn, dc, mask, d, dt, dx, re, stream) -> None:
        DC, M = self.a(dc, d, 3), self.a(mask, d)
        self._windowed(d, [self.a(dn, d, 3)], lambda w, o: orc.lib().orc_dye_nonadv(
            _p(o), _p(DC[w]), _p(M[w]), _i(o.shape[0]), _i(d.Y), _i(3), _f(dt), _f(dx), _f(re)))

    def fs2d_dye_nonadv_grad(self, fxn, fyn, fxc, fyc, fc, fn, mask, d, two_dx, stream) -> None:
        XC, YC, FC, FN, M = self.a(fxc, d, 3), self.a(fyc, d, 3), self.a(fc, d, 3), self.a(fn, d, 3), self.a(mask, d)
        self._windowed(d, [self.a(fxn, d, 3), self.a(fyn, d, 3)], lambda w, ox, oy: orc.lib().orc_cip_nonadv_grad_n(
            _p(ox), _p(oy), _p(XC[w]), _p(YC[w]), _p(FC[w]), _p(FN[w]), _p(M[w]), _i(ox.shape[0]), _i(d.Y), _i(3), _f(two_dx)))

    def fs2d_dye_cip_advect(self, fn, fxn, fyn, fc, fxc, fyc, v, mask, d, dt, dx, dx2, dx3, stream) -> None:
        FC, XC, YC, V, M = self.a(fc, d, 3), self.a(fxc, d, 3), self.a(fyc, d, 3), self.a(v, d, 2), self.a(mask, d)
        self._windowed(d, [self.a(fn, d, 3), self.a(fxn, d, 3), self.a(fyn, d, 3)], lambda w, o, ox, oy: orc.lib().orc_cip_advect_n(
            _p(o), _p(ox), _p(oy), _p(FC[w]), _p(XC[w]), _p(YC[w]), _p(V[w]), _p(M[w]), _i(o.shape[0]), _i(d.Y), _i(3), _f(dt),
            _f(dx), _f(dx2), _f(dx3)))

    def fs2d_dye_set_grad(self, fx, fy, f, d, dx, stream) -> None:
        F = self.a(f, d, 3)
        self._windowed(d, [self.a(fx, d, 3), self.a(fy, d, 3)], lambda w, ox, oy: orc.lib().orc_set_grad_n(
            _p(ox), _p(oy), _p(F[w]), _i(ox.shape[0]), _i(d.Y), _i(3), _f(dx)))
