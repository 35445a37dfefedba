"""Device fields and the reference's DoubleBuffer (/root/reference/fs/double_buffer.py:4-18).

A `Field` is a PyTorch tensor used purely as a device buffer (fp32, row-major (X, Y) or AoS
(X, Y, C) exactly like Taichi's `ti.field` / `ti.Vector.field` and their `to_numpy()` layout).
In a row-strip decomposition the tensor holds the rank's owned rows plus `halo` rows on each
side; `to_numpy()`/`from_numpy()` always speak about the owned rows.
"""
from __future__ import annotations

import numpy as np
import torch


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("fs (B200 build) needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Field:
    def __init__(self, resolution: tuple[int, int], n_channel: int = 1, device=None, halo: int = 0,
                 dtype=torch.float32) -> None:
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self.n = int(n_channel)
        self.halo = int(halo)
        rows = self.resolution[0] + 2 * self.halo
        shape = (rows, self.resolution[1]) + ((self.n,) if self.n > 1 else ())
        self.tensor = torch.zeros(shape, dtype=dtype, device=device if device is not None else default_device())
        self.dirty = True  # set when user code writes the buffer from the host side

    # -- Taichi-field look-alikes used by callers of the reference API
    @property
    def shape(self) -> tuple[int, int]:
        return self.resolution

    def owned(self) -> torch.Tensor:
        h = self.halo
        return self.tensor[h:self.tensor.shape[0] - h] if h else self.tensor

    def to_numpy(self) -> np.ndarray:
        """A host COPY of the owned rows (like Taichi's to_numpy()); never a view of live memory."""
        t = self.owned().detach()
        return t.cpu().numpy() if t.is_cuda else t.numpy().copy()   # CPU tensors only occur in the host-logic tests

    def from_numpy(self, a: np.ndarray) -> None:
        src = torch.from_numpy(np.ascontiguousarray(a)).to(self.tensor.dtype)
        if tuple(src.shape) != tuple(self.owned().shape):
            raise ValueError(f"shape mismatch: field {tuple(self.owned().shape)} vs array {tuple(src.shape)}")
        self.owned().copy_(src)
        self.dirty = True

    def fill(self, value: float) -> None:
        self.tensor.fill_(value)
        self.dirty = True

    def ptr(self) -> int:
        from fs import _lib

        return _lib.ptr(self.tensor)


class DoubleBuffer:
    """Two physical device arrays + reference swap (identity of the arrays matters)."""

    def __init__(self, resolution: tuple[int, int], n_channel: int, device=None, halo: int = 0) -> None:
        self.current = Field(resolution, n_channel, device, halo)
        self.next = Field(resolution, n_channel, device, halo)

    def swap(self) -> None:
        self.current, self.next = self.next, self.current

    def reset(self) -> None:
        self.current.fill(0)
        self.next.fill(0)
