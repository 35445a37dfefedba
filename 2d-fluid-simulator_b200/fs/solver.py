"""Time-step orchestrators (API of /root/reference/fs/solver.py).

`MacSolver.update()` (:79-89) and `CipMacSolver.update()` (:192-227) keep the reference's kernel
sequence, write predicates and swap counts exactly: each field is two physical device
arrays; kernels write only where their predicate holds.
"""
from __future__ import annotations

from abc import ABCMeta, abstractmethod

from fs import _lib
from fs.advection import AdvectionScheme
from fs.boundary_condition import BoundaryCondition, DyeBoundaryCondition
from fs.double_buffer import DoubleBuffer, Field
from fs.pressure_updater import PressureUpdater
from fs.vorticity_confinement import VorticityConfinement

VELOCITY_LIMIT = 10.0  # :12


class Solver(metaclass=ABCMeta):
    def __init__(self, boundary_condition: BoundaryCondition) -> None:
        self._bc = boundary_condition
        self.resolution = boundary_condition.get_resolution()

    @abstractmethod
    def update(self) -> None:
        pass

    @abstractmethod
    def get_fields(self):
        pass

    def is_wall(self, i: int, j: int) -> bool:
        return self._bc.is_wall(i, j)

    def is_fluid_domain(self, i: int, j: int) -> bool:
        return self._bc.is_fluid_domain(i, j)

    def _buffer(self, n_channel: int) -> DoubleBuffer:
        return DoubleBuffer(self.resolution, n_channel, self._bc.device, self._bc.halo)

    timing_events = None  # optional (start, stop) torch.cuda.Event pair recorded around the pressure update

    def _pressure_update(self) -> None:
        ev = self.timing_events
        if ev is not None:
            ev[0].record()
        self.pressure_updater.update(self.p, self.v.current)
        if ev is not None:
            ev[1].record()


def limit_field(field: Field, limit: float, dom=None, bc: BoundaryCondition | None = None) -> None:
    """:38-43 -- rescale |v| > limit to limit (all cells, in place)."""
    if dom is None:
        if bc is not None:
            dom = bc.dom
        else:
            X, Y = field.resolution
            dom = _lib.Dom(rows=X + 2 * field.halo, Y=Y, r0=field.halo, r1=field.halo + X, clo=0,
                           chi=X + 2 * field.halo - 1, gi0=0)
    _lib.call("fs2d_limit", field.ptr(), dom, limit, _lib.stream())


def clamp_field(field: Field, low: float, high: float, bc: BoundaryCondition | None = None) -> None:
    """:46-49 -- clamp every component to [low, high] (all cells, in place)."""
    if bc is not None:
        dom = bc.dom
    else:
        X, Y = field.resolution
        dom = _lib.Dom(rows=X + 2 * field.halo, Y=Y, r0=field.halo, r1=field.halo + X, clo=0, chi=X + 2 * field.halo - 1, gi0=0)
    _lib.call("fs2d_clamp", field.ptr(), dom, max(field.n, 1), low, high, _lib.stream())


class MacSolver(Solver):
    """Explicit MAC update with upwind / Kawamura-Kuwahara advection (:52-107)."""

    def __init__(self, boundary_condition: BoundaryCondition, pressure_updater: PressureUpdater,
                 advect_function: AdvectionScheme, dt: float, dx: float, re: float,
                 vorticity_confinement: VorticityConfinement | None = None) -> None:
        super().__init__(boundary_condition)
        if not isinstance(advect_function, AdvectionScheme):
            raise TypeError("advect_function must be fs.advection.advect_upwind or fs.advection.advect_kk_scheme")
        self._advect = advect_function
        self.dt, self.dx, self.re = dt, dx, re
        self.pressure_updater = pressure_updater
        self.vorticity_confinement = vorticity_confinement
        self.v = self._buffer(2)
        self.p = self._buffer(1)

    def update(self) -> None:
        if self._bc.partition.world > 1:
            from fs.halo import mac_update_distributed

            mac_update_distributed(self)
            return
        self._bc.set_velocity_boundary_condition(self.v.current)
        self._update_velocities(self.v.next, self.v.current, self.p.current)
        self.v.swap()
        if self.vorticity_confinement is not None:
            self.vorticity_confinement.apply(self.v)
            self.v.swap()
        self._pressure_update()
        limit_field(self.v.current, VELOCITY_LIMIT, bc=self._bc)

    def get_fields(self) -> tuple[Field, Field]:
        return self.v.current, self.p.current

    def _update_velocities(self, vn: Field, vc: Field, pc: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_mac_update", vn.ptr(), vc.ptr(), pc.ptr(), _lib.ptr(bc._bc_mask), bc.dom, self.dt, self.dx,
                  self.re, self._advect.code, _lib.stream())


class CipMacSolver(Solver):
    """CIP (constrained interpolation profile) solver carrying v and its spatial derivatives (:164-332)."""

    def __init__(self, boundary_condition: BoundaryCondition, pressure_updater: PressureUpdater, dt: float,
                 dx: float, re: float, vorticity_confinement: VorticityConfinement | None = None) -> None:
        super().__init__(boundary_condition)
        self.dt, self.dx, self.re = dt, dx, re
        self.pressure_updater = pressure_updater
        self.vorticity_confinement = vorticity_confinement
        self.v = self._buffer(2)
        self.vx = self._buffer(2)
        self.vy = self._buffer(2)
        self.p = self._buffer(1)
        self._set_grad(self.vx.current, self.vy.current, self.v.current)  # all-zero state: no halo needed

    def update(self) -> None:
        if self._bc.partition.world > 1:
            from fs.halo import cip_update_distributed

            cip_update_distributed(self)
            return
        self._bc.set_velocity_boundary_condition(self.v.current)
        self._update_velocities(self.v, self.vx, self.vy, self.p)
        if self.vorticity_confinement is not None:
            self.vorticity_confinement.apply(self.v)
            self.v.swap()
        self._pressure_update()
        limit_field(self.v.current, VELOCITY_LIMIT, bc=self._bc)

    def get_fields(self) -> tuple[Field, Field]:
        return self.v.current, self.p.current

    def _set_grad(self, fx: Field, fy: Field, f: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_set_grad", fx.ptr(), fy.ptr(), f.ptr(), bc.dom, self.dx, _lib.stream())

    def _update_velocities(self, v: DoubleBuffer, vx: DoubleBuffer, vy: DoubleBuffer, p: DoubleBuffer) -> None:
        self._non_advection_phase(v.next, v.current, p.current)
        self._non_advection_phase_grad(vx.next, vy.next, vx.current, vy.current, v.current, v.next)
        v.swap(); vx.swap(); vy.swap()
        self._advection_phase(v.next, vx.next, vy.next, v.current, vx.current, vy.current, v.current)
        v.swap(); vx.swap(); vy.swap()

    def _non_advection_phase(self, fn: Field, fc: Field, pc: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_cip_nonadv", fn.ptr(), fc.ptr(), pc.ptr(), _lib.ptr(bc._bc_mask), bc.dom, self.dt, self.dx,
                  self.re, _lib.stream())

    def _non_advection_phase_grad(self, fxn: Field, fyn: Field, fxc: Field, fyc: Field, fc: Field, fn: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_cip_nonadv_grad", fxn.ptr(), fyn.ptr(), fxc.ptr(), fyc.ptr(), fc.ptr(), fn.ptr(),
                  _lib.ptr(bc._bc_mask), bc.dom, 2.0 * self.dx, _lib.stream())

    def _advection_phase(self, fn: Field, fxn: Field, fyn: Field, fc: Field, fxc: Field, fyc: Field, v: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_cip_advect", fn.ptr(), fxn.ptr(), fyn.ptr(), fc.ptr(), fxc.ptr(), fyc.ptr(), v.ptr(),
                  _lib.ptr(bc._bc_mask), bc.dom, self.dt, self.dx, self.dx**2, self.dx**3, _lib.stream())


class DyeMacSolver(MacSolver):
    """MacSolver + passive dye advected with the same scheme (:110-161)."""

    def __init__(self, boundary_condition: DyeBoundaryCondition, pressure_updater: PressureUpdater,
                 advect_function: AdvectionScheme, dt: float, dx: float, re: float,
                 vorticity_confinement: VorticityConfinement | None = None) -> None:
        super().__init__(boundary_condition, pressure_updater, advect_function, dt, dx, re, vorticity_confinement)
        self.dye = self._buffer(3)

    def update(self) -> None:
        super().update()
        if self._bc.partition.world > 1:
            from fs.halo import dye_update_distributed

            dye_update_distributed(self)
            return
        self._bc.set_dye_boundary_condition(self.dye.current)
        self._update_dye(self.dye.next, self.dye.current, self.v.current)
        self.dye.swap()
        clamp_field(self.dye.current, 0.0, 1.0, bc=self._bc)

    def get_fields(self) -> tuple[Field, Field, Field]:
        return self.v.current, self.p.current, self.dye.current

    def _update_dye(self, dn: Field, dc: Field, vc: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_dye_mac", dn.ptr(), dc.ptr(), vc.ptr(), _lib.ptr(bc._bc_mask), bc.dom, self.dt, self.dx,
                  self._advect.code, _lib.stream())


class DyeCipMacSolver(CipMacSolver):
    """CipMacSolver + dye carried with CIP (dye and its two derivative fields, :335-401)."""

    def __init__(self, boundary_condition: DyeBoundaryCondition, pressure_updater: PressureUpdater, dt: float, dx: float,
                 re: float, vorticity_confinement: VorticityConfinement | None = None) -> None:
        super().__init__(boundary_condition, pressure_updater, dt, dx, re, vorticity_confinement)
        self.dye = self._buffer(3)
        self.dyex = self._buffer(3)
        self.dyey = self._buffer(3)
        bc = self._bc
        _lib.call("fs2d_dye_set_grad", self.dyex.current.ptr(), self.dyey.current.ptr(), self.dye.current.ptr(), bc.dom,
                  self.dx, _lib.stream())

    def update(self) -> None:
        super().update()
        if self._bc.partition.world > 1:
            from fs.halo import dye_update_distributed

            dye_update_distributed(self)
            return
        self._bc.set_dye_boundary_condition(self.dye.current)
        self._update_dye(self.dye, self.dyex, self.dyey, self.v)
        clamp_field(self.dye.current, 0.0, 1.0, bc=self._bc)

    def get_fields(self) -> tuple[Field, Field, Field]:
        return self.v.current, self.p.current, self.dye.current

    def _non_advection_phase_dye(self, dn: Field, dc: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_dye_nonadv", dn.ptr(), dc.ptr(), _lib.ptr(bc._bc_mask), bc.dom, self.dt, self.dx, self.re,
                  _lib.stream())

    def _dye_grad(self, dxn: Field, dyn: Field, dxc: Field, dyc: Field, dc: Field, dn: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_dye_nonadv_grad", dxn.ptr(), dyn.ptr(), dxc.ptr(), dyc.ptr(), dc.ptr(), dn.ptr(),
                  _lib.ptr(bc._bc_mask), bc.dom, 2.0 * self.dx, _lib.stream())

    def _dye_advect(self, dn: Field, dxn: Field, dyn: Field, dc: Field, dxc: Field, dyc: Field, v: Field) -> None:
        bc = self._bc
        _lib.call("fs2d_dye_cip_advect", dn.ptr(), dxn.ptr(), dyn.ptr(), dc.ptr(), dxc.ptr(), dyc.ptr(), v.ptr(),
                  _lib.ptr(bc._bc_mask), bc.dom, self.dt, self.dx, self.dx**2, self.dx**3, _lib.stream())

    def _update_dye(self, dye: DoubleBuffer, dyex: DoubleBuffer, dyey: DoubleBuffer, v: DoubleBuffer) -> None:
        self._non_advection_phase_dye(dye.next, dye.current)
        self._dye_grad(dyex.next, dyey.next, dyex.current, dyey.current, dye.current, dye.next)
        dye.swap(); dyex.swap(); dyey.swap()
        self._dye_advect(dye.next, dyex.next, dyey.next, dye.current, dyex.current, dyey.current, v.current)
        dye.swap(); dyex.swap(); dyey.swap()
