"""Facade (API of /root/reference/fs/fluid_simulator.py:12-108).

`FluidSimulator.create(num, resolution, dt, dx, re, vor_eps, scheme)` builds the same object graph
as the reference (:60-108): scene -> VorticityConfinement (or None) -> RedBlackSorPressureUpdater
(omega=1.3, n_iter=2, :76-78) -> solver by scheme; `ValueError("Unknown scheme: ...")` otherwise.
Extra keyword arguments (not in the reference) select the Jacobi updater used by the BASELINE
configs: `pressure="jacobi", n_iter=N`.
"""
from __future__ import annotations

import numpy as np
import numpy.typing as npt

from fs.advection import advect_kk_scheme, advect_upwind
from fs.boundary_condition import BoundaryCondition, get_boundary_condition
from fs.pressure_updater import JacobiPressureUpdater, PressureUpdater, RedBlackSorPressureUpdater
from fs.solver import CipMacSolver, DyeCipMacSolver, DyeMacSolver, MacSolver
from fs.vorticity_confinement import VorticityConfinement


def make_solver(boundary_condition: BoundaryCondition, dt: float, dx: float, re: float, vor_eps: float | None,
                scheme: str, pressure_updater: PressureUpdater | None = None, pressure: str = "rbsor",
                n_iter: int = 2, relaxation_factor: float = 1.3, dye: bool = False):
    vorticity_confinement = (VorticityConfinement(boundary_condition, dt, dx, vor_eps) if vor_eps is not None else None)
    if pressure_updater is None:
        if pressure == "rbsor":
            pressure_updater = RedBlackSorPressureUpdater(boundary_condition, dt, dx,
                                                          relaxation_factor=relaxation_factor, n_iter=n_iter)
        elif pressure == "jacobi":
            pressure_updater = JacobiPressureUpdater(boundary_condition, dt, dx, n_iter=n_iter)
        else:
            raise ValueError(f"Unknown pressure updater: {pressure}")
    cip, mac = (DyeCipMacSolver, DyeMacSolver) if dye else (CipMacSolver, MacSolver)
    if scheme == "cip":
        return cip(boundary_condition, pressure_updater, dt, dx, re, vorticity_confinement)
    if scheme == "upwind":
        return mac(boundary_condition, pressure_updater, advect_upwind, dt, dx, re, vorticity_confinement)
    if scheme == "kk":
        return mac(boundary_condition, pressure_updater, advect_kk_scheme, dt, dx, re, vorticity_confinement)
    msg = f"Unknown scheme: {scheme}"
    raise ValueError(msg)


class FluidSimulator:
    def __init__(self, solver: MacSolver | CipMacSolver) -> None:
        self._solver = solver

    def step(self) -> None:
        self._solver.update()

    def field_to_numpy(self) -> dict[str, npt.NDArray]:
        """{"v": (X, Y, 2) f32, "p": (X, Y) f32} -- the `d`-key dump format (main.py:129-132)."""
        fields = self._solver.get_fields()
        return {"v": fields[0].to_numpy(), "p": fields[1].to_numpy()}

    @property
    def solver(self):
        return self._solver

    @staticmethod
    def create(num: int, resolution: int, dt: float, dx: float, re: float, vor_eps: float | None, scheme: str,
               **kwargs) -> "FluidSimulator":
        bc_kw = {k: kwargs.pop(k) for k in ("device", "partition") if k in kwargs}
        if scheme not in ("cip", "upwind", "kk"):
            msg = f"Unknown scheme: {scheme}"
            raise ValueError(msg)
        boundary_condition = get_boundary_condition(num, resolution, enable_dye=False, **bc_kw)
        return FluidSimulator(make_solver(boundary_condition, dt, dx, re, vor_eps, scheme, **kwargs))


class DyeFluidSimulator(FluidSimulator):
    """(:111-176) the default simulator of main.py: velocity/pressure + three dye channels."""

    def field_to_numpy(self) -> dict[str, npt.NDArray]:
        fields = self._solver.get_fields()
        return {"v": fields[0].to_numpy(), "p": fields[1].to_numpy(), "dye": fields[2].to_numpy()}

    @staticmethod
    def create(num: int, resolution: int, dt: float, dx: float, re: float, vor_eps: float | None, scheme: str,
               **kwargs) -> "DyeFluidSimulator":
        bc_kw = {k: kwargs.pop(k) for k in ("device", "partition") if k in kwargs}
        if scheme not in ("cip", "upwind", "kk"):
            msg = f"Unknown scheme: {scheme}"
            raise ValueError(msg)
        boundary_condition = get_boundary_condition(num, resolution, enable_dye=True, **bc_kw)
        return DyeFluidSimulator(make_solver(boundary_condition, dt, dx, re, vor_eps, scheme, dye=True, **kwargs))
