"""ctypes binding of libfs2d.so (C ABI declared in include/fs2d.h).

There is NO CPU fallback: every kernel call requires the sm_100a library and CUDA tensors, and
fails loudly otherwise.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_float, c_int, c_void_p
from pathlib import Path

import torch

PKG_ROOT = Path(__file__).resolve().parents[1]
LIB_PATH = PKG_ROOT / "lib" / "libfs2d.so"

FS2D_E_BADARG, FS2D_E_CUDA, FS2D_E_NCCL = -1, -2, -3
SCHEME_UPWIND, SCHEME_KK = 0, 1


class Dom(ctypes.Structure):
    """fs2d_dom (include/fs2d.h): local row-strip array description."""

    _fields_ = [("rows", c_int), ("Y", c_int), ("r0", c_int), ("r1", c_int), ("clo", c_int), ("chi", c_int),
                ("gi0", c_int)]

    def replace(self, **kw) -> "Dom":
        d = Dom(self.rows, self.Y, self.r0, self.r1, self.clo, self.chi, self.gi0)
        for k, v in kw.items():
            setattr(d, k, v)
        return d


_P = c_void_p
_SIGNATURES = {
    "fs2d_last_error": (c_char_p, []),
    "fs2d_version": (c_int, []),
    "fs2d_launch_count": (ctypes.c_ulonglong, []),
    "fs2d_device_ok": (c_int, []),
    "fs2d_set_tuning": (c_int, [c_int, c_int]),
    "fs2d_vel_bc": (c_int, [_P, _P, _P, _P, _P, _P, c_int, _P]),
    "fs2d_pressure_bc": (c_int, [_P, _P, _P, _P, _P, _P, c_int, _P]),
    "fs2d_mac_update": (c_int, [_P, _P, _P, _P, Dom, c_float, c_float, c_float, c_int, _P]),
    "fs2d_cip_nonadv": (c_int, [_P, _P, _P, _P, Dom, c_float, c_float, c_float, _P]),
    "fs2d_cip_nonadv_grad": (c_int, [_P, _P, _P, _P, _P, _P, _P, Dom, c_float, _P]),
    "fs2d_cip_advect": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, Dom, c_float, c_float, c_float, c_float, _P]),
    "fs2d_set_grad": (c_int, [_P, _P, _P, Dom, c_float, _P]),
    "fs2d_vort_calc": (c_int, [_P, _P, _P, _P, Dom, c_float, _P]),
    "fs2d_vort_add": (c_int, [_P, _P, _P, _P, _P, Dom, c_float, c_float, _P]),
    "fs2d_vort_apply": (c_int, [_P, _P, _P, _P, _P, Dom, c_float, c_float, _P]),
    "fs2d_pressure_source": (c_int, [_P, _P, Dom, c_float, c_float, _P]),
    "fs2d_jacobi_sweep": (c_int, [_P, _P, _P, _P, Dom, c_int, _P]),
    "fs2d_jacobi_update": (c_int, [_P, _P, _P, _P, Dom, c_int, _P, _P, _P, _P, _P, c_int, c_int, POINTER(_P), POINTER(c_int),
                                   POINTER(c_int), _P]),
    "fs2d_fused_order": (c_int, [_P, Dom, c_int, c_int, c_int, _P, c_int, POINTER(c_int), _P]),
    "fs2d_jacobi_plan": (c_int, [c_int, c_int, POINTER(c_int), c_int, POINTER(c_int)]),
    "fs2d_jacobi_fused": (c_int, [_P, _P, _P, _P, Dom, c_int, _P, c_int, _P]),
    "fs2d_jacobi_fused_part": (c_int, [_P, _P, _P, _P, Dom, c_int, c_int, c_int, _P, c_int, _P]),
    "fs2d_jacobi_fused_tail": (c_int, [_P, _P, _P, _P, Dom, c_int, c_int, c_int, _P, c_int, _P]),
    "fs2d_fused_tile": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "fs2d_rbsor_pass": (c_int, [_P, _P, _P, _P, Dom, c_float, c_float, c_int, _P]),
    "fs2d_rbsor_iteration": (c_int, [_P, _P, _P, _P, Dom, c_float, c_float, _P]),
    "fs2d_limit": (c_int, [_P, Dom, c_float, _P]),
    "fs2d_render": (c_int, [_P, _P, _P, _P, _P, Dom, c_float, c_int, _P]),
    "fs2d_dye_bc": (c_int, [_P, _P, _P, c_int, _P]),
    "fs2d_clamp": (c_int, [_P, Dom, c_int, c_float, c_float, _P]),
    "fs2d_dye_mac": (c_int, [_P, _P, _P, _P, Dom, c_float, c_float, c_int, _P]),
    "fs2d_dye_nonadv": (c_int, [_P, _P, _P, Dom, c_float, c_float, c_float, _P]),
    "fs2d_dye_nonadv_grad": (c_int, [_P, _P, _P, _P, _P, _P, _P, Dom, c_float, _P]),
    "fs2d_dye_cip_advect": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, Dom, c_float, c_float, c_float, c_float, _P]),
    "fs2d_dye_set_grad": (c_int, [_P, _P, _P, Dom, c_float, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> ctypes.CDLL:
    """Load libfs2d.so (built in-tree by __graft_entry__.build()); never falls back."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "The fs package has no CPU/PyTorch fallback.")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = load().fs2d_last_error().decode(errors="replace")
    if rc == FS2D_E_BADARG:
        raise ValueError(f"libfs2d: {msg}")
    raise RuntimeError(f"libfs2d error {rc}: {msg}")


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a contiguous CUDA tensor (raises on CPU tensors: no fallback)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libfs2d kernels need CUDA tensors (sm_100a); there is no CPU fallback")
    if not t.is_contiguous():
        raise ValueError("libfs2d kernels need contiguous tensors")
    return t.data_ptr()


def stream() -> int:
    """The current stream of the CURRENT device.  Fields live on the device that was current when their BoundaryCondition
    was built (BoundaryCondition refuses any other `device=`), so this is also the stream of the tensors' device."""
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))
