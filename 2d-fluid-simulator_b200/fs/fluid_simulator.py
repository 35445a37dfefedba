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

from fs import _lib
from fs.double_buffer import Field

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
        self._rgb: Field | None = None   # image buffer (:16), allocated on first use: 12 B/cell

    def step(self) -> None:
        if self._graphs is not None:
            self._replay()
        else:
            self._solver.update()

    # -- CUDA-graph stepping (not in the reference): one graph launch per time step -------------------
    _graphs = None
    _graph_strips = False

    def _buffers(self) -> list:
        s = self._solver
        return [b for b in (getattr(s, n, None) for n in ("v", "vx", "vy", "p", "dye", "dyex", "dyey")) if b is not None]

    def enable_cuda_graph(self, strips: bool = False) -> None:
        """Capture the ~100-200 kernel launches of `solver.update()` into CUDA graphs and replay them in step().

        Small grids are launch-bound (a res=512 step is 150 launches of a few microseconds each).  The double
        buffers swap roles inside a step, so a step's kernel arguments repeat with period 1 or 2; one graph is
        captured per phase and step() replays them in turn, mirroring the swaps on the Python side so that
        `.current` / `.next` stay truthful.  Results are bit-identical (same kernels).  On row strips (world > 1) the
        halo SendRecvs are captured with the kernels (NCCL point-to-point operations are capturable; every rank must call
        this at the same point, like any collective): opt-in with strips=True."""
        import torch

        if self._solver._bc.partition.world > 1 and not strips:
            raise NotImplementedError("CUDA-graph stepping on row strips is opt-in: enable_cuda_graph(strips=True) on every rank")
        if self._graphs is not None:
            return
        self._graph_strips = strips
        self._solver.update()               # warm-up outside capture: lazy allocations, one-time validity checks
        torch.cuda.synchronize()
        for f in (self._solver.p.current, self._solver.p.next):
            f.dirty = False                 # whatever the pressure updater validates per host write, it has validated now (see _replay)
        start = [id(b.current) for b in self._buffers()]
        graphs = []
        for _ in range(2):                  # period <= 2: every buffer swaps a fixed number of times per step
            before = [id(b.current) for b in self._buffers()]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._solver.update()
            flips = [k for k, b in enumerate(self._buffers()) if id(b.current) != before[k]]
            graphs.append((g, flips))
            if [id(b.current) for b in self._buffers()] == start:
                break
        else:
            raise RuntimeError("buffer roles did not return to the start after two steps")
        # the captures executed the Python-side swaps without running any kernel: roll them back, then replay
        for g, flips in reversed(graphs):
            for k in flips:
                self._buffers()[k].swap()
        self._graphs, self._phase = graphs, 0

    def _replay(self) -> None:
        if any(b.current.dirty or b.next.dirty for b in self._buffers() if b is self._solver.p):
            # a pressure buffer was rewritten from the host (from_numpy / load_state_dict) since the capture: the captured
            # schedule may contain fused passes whose precondition (equal never-written wall cells) no longer holds.
            # Drop the graphs and capture again: the warm-up step of the capture IS this step (eager; the updater re-validates).
            self._graphs = None
            self.enable_cuda_graph(strips=self._graph_strips)
            return
        g, flips = self._graphs[self._phase]
        g.replay()
        bufs = self._buffers()
        for k in flips:
            bufs[k].swap()
        self._phase = (self._phase + 1) % len(self._graphs)

    # -- render getters (:22-32): (X, Y, 3) f32 images, wall cells in the wall colour -------------------
    @property
    def rgb_buf(self) -> Field:
        if self._rgb is None:
            bc = self._solver._bc
            self._rgb = Field(self._solver.resolution, 3, bc.device, bc.halo)
        return self._rgb

    def _render(self, mode: int) -> Field:
        s, bc = self._solver, self._solver._bc
        fields = s.get_fields()
        dye = fields[2].ptr() if len(fields) > 2 else None
        _lib.call("fs2d_render", self.rgb_buf.ptr(), fields[0].ptr(), fields[1].ptr(), dye, _lib.ptr(bc._bc_mask), bc.dom,
                  s.dx, mode, _lib.stream())
        return self.rgb_buf

    def get_norm_field(self) -> Field:
        return self._render(0)

    def get_pressure_field(self) -> Field:
        return self._render(1)

    def get_vorticity_field(self) -> Field:
        if self._solver._bc.partition.world > 1:
            from fs.halo import exchanger_for

            exchanger_for(self._solver._bc).exchange(self._solver.get_fields()[0], 1)
        return self._render(2)

    # -- full-state dump / restore: the `d`-key dump (main.py:129-132) cannot resume a CIP run
    #    because vx/vy are not in it; these carry every physical buffer.
    def state_dict(self) -> dict[str, npt.NDArray]:
        out = {}
        for name in ("v", "vx", "vy", "p", "dye", "dyex", "dyey"):
            buf = getattr(self._solver, name, None)
            if buf is not None:
                out[name + "_cur"], out[name + "_nxt"] = buf.current.to_numpy(), buf.next.to_numpy()
        vc = self._solver.vorticity_confinement
        if vc is not None:
            out["vort"], out["vort_abs"] = vc.vorticity.to_numpy(), vc.vorticity_abs.to_numpy()
        return out

    def load_state_dict(self, state: dict) -> None:
        for key, a in state.items():
            if key in ("vort", "vort_abs"):
                getattr(self._solver.vorticity_confinement, "vorticity" if key == "vort" else "vorticity_abs").from_numpy(a)
            else:
                name, which = key.rsplit("_", 1)
                getattr(getattr(self._solver, name), "current" if which == "cur" else "next").from_numpy(a)

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
        bc_kw = {k: kwargs.pop(k) for k in ("device", "partition", "obstacle_image") if k in kwargs}
        if num not in (1, 2, 3, 4, 5, 6):      # the reference builds the scene first (:70), so a bad scene number wins over a bad scheme
            raise NotImplementedError
        if scheme not in ("cip", "upwind", "kk"):
            msg = f"Unknown scheme: {scheme}"
            raise ValueError(msg)
        boundary_condition = get_boundary_condition(num, resolution, enable_dye=False, **bc_kw)
        return FluidSimulator(make_solver(boundary_condition, dt, dx, re, vor_eps, scheme, **kwargs))


class DyeFluidSimulator(FluidSimulator):
    """(:111-176) the default simulator of main.py: velocity/pressure + three dye channels."""

    def get_dye_field(self) -> Field:
        return self._render(3)

    def field_to_numpy(self) -> dict[str, npt.NDArray]:
        fields = self._solver.get_fields()
        return {"v": fields[0].to_numpy(), "p": fields[1].to_numpy(), "dye": fields[2].to_numpy()}

    @staticmethod
    def create(num: int, resolution: int, dt: float, dx: float, re: float, vor_eps: float | None, scheme: str,
               **kwargs) -> "DyeFluidSimulator":
        bc_kw = {k: kwargs.pop(k) for k in ("device", "partition", "obstacle_image") if k in kwargs}
        if num not in (1, 2, 3, 4, 5, 6):
            raise NotImplementedError
        if scheme not in ("cip", "upwind", "kk"):
            msg = f"Unknown scheme: {scheme}"
            raise ValueError(msg)
        boundary_condition = get_boundary_condition(num, resolution, enable_dye=True, **bc_kw)
        return DyeFluidSimulator(make_solver(boundary_condition, dt, dx, re, vor_eps, scheme, dye=True, **kwargs))
