"""Vorticity confinement (API of /root/reference/fs/vorticity_confinement.py:8-59).

`apply(v)` computes the curl of v.current into `vorticity`/`vorticity_abs` (fluid cells, :27-32)
and writes v.next = v.current + dt*weight*clamp(N x omega) (fluid cells, :34-55).  It writes
v.next only -- the caller swaps (fs/solver.py:85-86, :197-198).
"""
from __future__ import annotations

from fs import _lib
from fs.boundary_condition import BoundaryCondition
from fs.double_buffer import DoubleBuffer, Field


class VorticityConfinement:
    def __init__(self, boundary_condition: BoundaryCondition, dt: float, dx: float, weight: float) -> None:
        self._bc = boundary_condition
        self.dt = dt
        self.dx = dx
        self.weight = weight
        self._resolution = boundary_condition.get_resolution()
        dev, halo = boundary_condition.device, boundary_condition.halo
        self.vorticity = Field(self._resolution, 1, dev, halo)
        self.vorticity_abs = Field(self._resolution, 1, dev, halo)

    def _calc_vorticity(self, vc: Field, dom=None) -> None:
        bc = self._bc
        _lib.call("fs2d_vort_calc", self.vorticity.ptr(), self.vorticity_abs.ptr(), vc.ptr(), _lib.ptr(bc._bc_mask),
                  dom or bc.dom, self.dx, _lib.stream())

    def _add_vorticity(self, vn: Field, vc: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_vort_add", vn.ptr(), vc.ptr(), self.vorticity.ptr(), self.vorticity_abs.ptr(),
                  _lib.ptr(bc._bc_mask), bc.dom, self.dx, self.dt * self.weight, _lib.stream())

    def _apply_fused(self, vn: Field, vc: Field) -> None:
        """_calc_vorticity + _add_vorticity in one pass over HBM (fs2d_vort_apply): same vorticity, vorticity_abs
        and vn, 25 instead of 42 bytes per cell."""
        bc = self._bc
        _lib.call("fs2d_vort_apply", vn.ptr(), self.vorticity.ptr(), self.vorticity_abs.ptr(), vc.ptr(),
                  _lib.ptr(bc._bc_mask), bc.dom, self.dx, self.dt * self.weight, _lib.stream())

    def apply(self, v: DoubleBuffer) -> None:
        if self._bc.partition.world > 1:
            from fs.halo import vorticity_apply_distributed

            vorticity_apply_distributed(self, v)
            return
        self._apply_fused(v.next, v.current)  # v.next only; the solver swaps
