"""Host-fed stepping: overlap PCIe traffic with compute (not in the reference, which keeps its state on the device).

`HostPipelinedStepper` serves callers whose state lives in HOST memory: every `submit(host_state)` uploads
(v, vx, vy, p) from pinned buffers, runs one `solver.update()` and downloads (v, p) -- the
`field_to_numpy()` payload -- into pinned result buffers.  Two solver instances (sharing one
BoundaryCondition) and three CUDA streams form a 3-stage pipeline, so the upload of request k+1 and the
download of request k-1 overlap the kernels of request k; `results(k)` waits for request k only.
Throughput is then bounded by max(H2D, compute, D2H) instead of their sum.
"""
from __future__ import annotations

import torch

from fs.fluid_simulator import make_solver


class HostPipelinedStepper:
    STATE = ("v", "vx", "vy", "p")      # uploaded per request (CIP state; vx/vy absent for MacSolver)
    RESULT = ("v", "p")                 # downloaded per request

    def __init__(self, boundary_condition, dt, dx, re, vor_eps, scheme, depth: int = 2, **solver_kw) -> None:
        self.solvers = [make_solver(boundary_condition, dt, dx, re, vor_eps, scheme, **solver_kw) for _ in range(depth)]
        self.state_names = [n for n in self.STATE if hasattr(self.solvers[0], n)]
        self.up, self.comp, self.down = (torch.cuda.Stream() for _ in range(3))
        self.out = [{n: torch.empty_like(getattr(s, n).current.owned(), device="cpu").pin_memory() for n in self.RESULT}
                    for s in self.solvers]
        self._down_done: list[torch.cuda.Event | None] = [None] * depth
        self._k = 0
        self.h2d_bytes = sum(getattr(self.solvers[0], n).current.owned().numel() * 4 for n in self.state_names)
        self.d2h_bytes = sum(t.numel() * 4 for t in self.out[0].values())

    def submit(self, host_state: dict) -> int:
        """Enqueue one step on `host_state` ({name: pinned CPU tensor shaped like the owned rows}); returns its ticket."""
        k, slot = self._k, self._k % len(self.solvers)
        s = self.solvers[slot]
        if self._down_done[slot] is not None:
            self.up.wait_event(self._down_done[slot])          # the slot's previous result has left the device
        with torch.cuda.stream(self.up):
            for n in self.state_names:
                f = getattr(s, n).current
                f.owned().copy_(host_state[n], non_blocking=True)
                f.dirty = True       # written from the host: the Jacobi updater re-checks its fused-pass precondition (stale wall cells)
            up_done = torch.cuda.Event()
            up_done.record()
        self.comp.wait_event(up_done)
        with torch.cuda.stream(self.comp):
            s.update()
            comp_done = torch.cuda.Event()
            comp_done.record()
        self.down.wait_event(comp_done)
        with torch.cuda.stream(self.down):
            fields = dict(zip(("v", "p"), s.get_fields()[:2]))
            for n in self.RESULT:
                self.out[slot][n].copy_(fields[n].owned(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self._down_done[slot] = ev
        self._k += 1
        return k

    def results(self, ticket: int) -> dict:
        """Pinned host tensors {"v", "p"} of request `ticket` (valid until `depth` more requests are submitted)."""
        slot = ticket % len(self.solvers)
        self._down_done[slot].synchronize()
        return self.out[slot]

    def synchronize(self) -> None:
        for ev in self._down_done:
            if ev is not None:
                ev.synchronize()
