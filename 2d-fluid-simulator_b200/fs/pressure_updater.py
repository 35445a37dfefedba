"""Pressure relaxation operators (API of /root/reference/fs/pressure_updater.py).

`JacobiPressureUpdater.update` == n_iter x { set_pressure_boundary_condition(p.current);
_update(p.next, p.current, v); p.swap() }  (:56-60).  The sweep kernel recomputes the post-BC
neighbour pressures inline from `pcode`, so the BC pass only has to be materialised for the last
two sweeps (whose stored values remain observable) -- `fs2d_jacobi_update` does the
whole loop in C.  `RedBlackSorPressureUpdater` (:69-114) is what `FluidSimulator.create()`
instantiates (omega=1.3, n_iter=2).
"""
from __future__ import annotations

import ctypes
from abc import ABCMeta, abstractmethod

from fs import _lib
from fs.boundary_condition import BoundaryCondition
from fs.double_buffer import DoubleBuffer, Field


class PressureUpdater(metaclass=ABCMeta):
    def __init__(self, boundary_condition: BoundaryCondition, dt: float, dx: float) -> None:
        self._bc = boundary_condition
        self.dt = dt
        self.dx = dx
        self._src: Field | None = None

    @abstractmethod
    def update(self, p: DoubleBuffer, v_current: Field) -> None:
        pass

    def _source(self, v_current: Field, dom=None, first: bool = True) -> Field:
        """(t2, t3) velocity terms of predict_p (:23-38), one float2 per cell; v is constant during an
        update (:56-60) so this runs once per update instead of once per sweep.  first=False: a further row window of the
        same update (multi-rank overlap)."""
        bc = self._bc
        if self._src is None:
            self._src = Field(bc.get_resolution(), 2, bc.device, bc.halo)
        _lib.call("fs2d_pressure_source", self._src.ptr(), v_current.ptr(), dom or bc.dom, self.dt, self.dx, _lib.stream())
        return self._src


class JacobiPressureUpdater(PressureUpdater):
    """Jacobi method (:41-66).

    `fuse` (not in the reference): number of iterations computed per pass over HBM by the fused
    shared-memory kernel (fs2d_jacobi_fused); "auto" picks FUSE_DEFAULT when the mask qualifies, 0 disables.
    Results are bit-identical either way."""

    FUSE_DEFAULT = 8

    def __init__(self, boundary_condition: BoundaryCondition, dt: float, dx: float, n_iter: int,
                 fuse: int | str = "auto") -> None:
        super().__init__(boundary_condition, dt, dx)
        self._n_iter = int(n_iter)
        self._fuse_request = fuse
        self._fuse_t: int | None = None      # resolved lazily (needs the device tables)
        self._stale_checked: tuple | None = None

    def _sweep(self, p_next: Field, p_current: Field, src: Field, inline_bc: bool, dom=None) -> None:
        bc = self._bc
        _lib.call("fs2d_jacobi_sweep", p_next.ptr(), p_current.ptr(), src.ptr(), _lib.ptr(bc._pcode), dom or bc.dom,
                  int(inline_bc), _lib.stream())

    # the reference's kernel of the same name (:62-66): one sweep, BC already applied by the caller
    def _update(self, p_next: Field, p_current: Field, v_current: Field) -> None:
        self._sweep(p_next, p_current, self._source(v_current), inline_bc=False)

    def fuse_mask(self, p: DoubleBuffer) -> int:
        """Bit t set: a fused pass of t iterations is valid for this mask and these buffers (0 = literal only).
        Every pass size has its own tiling, so each is validated separately (BoundaryCondition.fused_ok)."""
        if self._fuse_t is None:
            req = self._fuse_request
            t_max = self.FUSE_DEFAULT if req == "auto" else int(req)
            self._fuse_t = sum(1 << t for t in range(1, t_max + 1) if self._bc.fused_ok(t)) if t_max > 0 else 0
        if self._fuse_t == 0 and self._bc.partition.world == 1:
            return 0
        key = frozenset((id(p.current), id(p.next)))   # the pair of physical buffers, whichever is current
        if self._stale_checked != key or p.current.dirty or p.next.dirty:
            # never-written wall cells must agree between the two physical buffers (DESIGN.md "stale cells")
            ok = self._bc.stale_cells_agree(p.current, p.next)
            if self._bc.partition.world > 1:
                # every rank must follow the same schedule (its halo exchanges pair up): agree on the verdict.
                # Setup-time only -- runs again only after user code rewrites a pressure buffer from the host.
                import torch
                import torch.distributed as dist

                mine = self._fuse_t if ok else 0
                bits = torch.tensor([(mine >> t) & 1 for t in range(13)], dtype=torch.int32, device=p.current.tensor.device)
                dist.all_reduce(bits, op=dist.ReduceOp.MIN)     # bitwise AND across ranks (NCCL has no BAND)
                self._agreed_mask = sum(int(b) << t for t, b in enumerate(bits.tolist()))
            else:
                self._agreed_mask = self._fuse_t if ok else 0
            self._stale_checked = key
            p.current.dirty = p.next.dirty = False
        return self._agreed_mask

    def _orders(self, mask: int):
        """Host arrays (indexed by pass size) of the tile lists fs2d_jacobi_update hands to its fused passes."""
        key = (mask, self._bc.dom.r0, self._bc.dom.r1)
        if getattr(self, "_orders_key", None) != key:
            ptrs, counts = (ctypes.c_void_p * 13)(), (ctypes.c_int * 13)()
            for t in range(1, 13):
                if (mask >> t) & 1:
                    order, n = self._bc.fused_order(t)
                    ptrs[t], counts[t] = _lib.ptr(order), n
            self._orders_key, self._orders_val = key, (ptrs, counts)
        return self._orders_val

    def plan(self, p: DoubleBuffer) -> list[int]:
        """Schedule of one update: fused pass sizes (> 0) and literal iterations (0), fs2d_jacobi_plan."""
        cap = self._n_iter + 4
        sizes, n = (ctypes.c_int * cap)(), ctypes.c_int(0)
        _lib.call("fs2d_jacobi_plan", self._n_iter, self.fuse_mask(p), sizes, cap, ctypes.byref(n))
        return list(sizes[:n.value])

    def _fused(self, p_next: Field, p_current: Field, src: Field, t: int, dom=None, skip=None, emit: bool = False) -> None:
        """One fused pass; dom: row window (default: the owned rows); skip = (first, n): leave these tile rows of the
        window's tiling to another launch (fs2d_jacobi_fused_part); emit: the experimental tail pass that also stores the BC
        values of its penultimate state into the wall cells of p_current (fs2d_jacobi_fused_tail)."""
        bc = self._bc
        order, n_order = bc.fused_order(t, dom, skip)
        if emit:
            first, n = skip if skip is not None else (0, 0)
            _lib.call("fs2d_jacobi_fused_tail", p_next.ptr(), p_current.ptr(), src.ptr(), _lib.ptr(bc._pcode), dom or bc.dom, t,
                      first, n, _lib.ptr(order), n_order, _lib.stream())
        elif skip is None:
            _lib.call("fs2d_jacobi_fused", p_next.ptr(), p_current.ptr(), src.ptr(), _lib.ptr(bc._pcode), dom or bc.dom, t,
                      _lib.ptr(order), n_order, _lib.stream())
        else:
            _lib.call("fs2d_jacobi_fused_part", p_next.ptr(), p_current.ptr(), src.ptr(), _lib.ptr(bc._pcode), dom or bc.dom,
                      t, skip[0], skip[1], _lib.ptr(order), n_order, _lib.stream())

    def update(self, p: DoubleBuffer, v_current: Field) -> None:
        bc = self._bc
        if bc.partition.world > 1:
            from fs.halo import jacobi_update_distributed

            jacobi_update_distributed(self, p, v_current)
            return
        src = self._source(v_current)
        t = bc._p_table
        final_in_b = ctypes.c_int(0)
        mask = self.fuse_mask(p)
        orders, n_orders = self._orders(mask)
        _lib.call("fs2d_jacobi_update", p.current.ptr(), p.next.ptr(), src.ptr(), _lib.ptr(bc._pcode), bc.dom,
                  self._n_iter, _lib.ptr(t["tgt"]), _lib.ptr(t["src0"]), _lib.ptr(t["src1"]), _lib.ptr(t["kind"]),
                  _lib.ptr(bc._scratch), t["n"], mask, orders, n_orders, ctypes.byref(final_in_b), _lib.stream())
        if final_in_b.value:
            p.swap()  # n_iter odd: same net effect as the reference's n_iter swaps


class RedBlackSorPressureUpdater(PressureUpdater):
    """Red-black SOR (:69-114): odd pass pn <- f(pc), even pass pn <- f(pn) (:96)."""

    def __init__(self, boundary_condition: BoundaryCondition, dt: float, dx: float, relaxation_factor: float,
                 n_iter: int) -> None:
        super().__init__(boundary_condition, dt, dx)
        self._n_iter = int(n_iter)
        self._relaxation_factor = relaxation_factor

    def _pass(self, pn: Field, pc: Field, src: Field, parity: int, dom=None) -> None:
        bc = self._bc
        w = self._relaxation_factor
        _lib.call("fs2d_rbsor_pass", pn.ptr(), pc.ptr(), src.ptr(), _lib.ptr(bc._bc_mask), dom or bc.dom, w, 1.0 - w,
                  parity, _lib.stream())

    def _update(self, p_next: Field, p_current: Field, v_current: Field, src: Field | None = None) -> None:
        src = src if src is not None else self._source(v_current)
        if self.fused_colours and p_next is not p_current:
            bc, w = self._bc, self._relaxation_factor      # both colour passes in one pass over HBM (fs2d_rbsor_iteration)
            _lib.call("fs2d_rbsor_iteration", p_next.ptr(), p_current.ptr(), src.ptr(), _lib.ptr(bc._bc_mask), bc.dom, w, 1.0 - w,
                      _lib.stream())
            return
        self._pass(p_next, p_current, src, 1)   # _update_pressures_odd  (:98-102)
        self._pass(p_next, p_next, src, 0)      # _update_pressures_even (:104-108), pc = pn

    #: one kernel per iteration instead of one per colour (same results bit for bit); the strips of a multi-rank run keep the
    #: two passes (fs/halo.py: the neighbour's odd cells are exchanged in between)
    fused_colours = True

    def update(self, p: DoubleBuffer, v_current: Field) -> None:
        if self._bc.partition.world > 1:
            from fs.halo import rbsor_update_distributed

            rbsor_update_distributed(self, p, v_current)
            return
        src = self._source(v_current)
        for _ in range(self._n_iter):
            self._bc.set_pressure_boundary_condition(p.current)
            self._update(p.next, p.current, v_current, src)
            p.swap()
