"""Advection scheme selectors (API of /root/reference/fs/advection.py).

In the reference `advect_upwind` (:12-24) and `advect_kk_scheme` (:27-60) are per-cell `@ti.func`s
injected into `MacSolver(...)`.  Here they are importable sentinels: `MacSolver` maps them to the
scheme selector of `fs2d_mac_update` (include/fs2d.h); the stencils themselves live in
csrc/fs2d_kernels.cu (`advect_upwind`, `advect_kk` device functions).
"""
from __future__ import annotations

from fs import _lib


class AdvectionScheme:
    def __init__(self, name: str, code: int, radius: int) -> None:
        self.name, self.code, self.radius = name, code, radius

    def __call__(self, *args, **kwargs):
        raise TypeError(f"{self.name} is a device-side stencil selector; pass it to MacSolver(...) instead of calling it")

    def __repr__(self) -> str:
        return f"<advection scheme {self.name}>"


advect_upwind = AdvectionScheme("upwind", _lib.SCHEME_UPWIND, radius=1)
advect_kk_scheme = AdvectionScheme("kk", _lib.SCHEME_KK, radius=2)
