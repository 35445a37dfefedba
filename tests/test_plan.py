"""Jacobi schedule planner (fs2d_jacobi_plan): host logic, no GPU needed."""
from __future__ import annotations

import ctypes
import itertools

import pytest

from fs import _lib


def plan(n, mask, tail=1):
    lib = _lib.load()
    sizes, k = (ctypes.c_int * (n + 8))(), ctypes.c_int()
    assert lib.fs2d_set_tuning(4, tail) == 0      # 1 (default): {emitting fused pass, one literal iteration}; 0: two literal iterations
    try:
        assert lib.fs2d_jacobi_plan(n, mask, sizes, n + 8, ctypes.byref(k)) == 0
    finally:
        lib.fs2d_set_tuning(4, 1)
    return list(sizes[:k.value])


@pytest.mark.parametrize("n,mask", list(itertools.product([0, 1, 2, 3, 7, 21, 40, 80, 100, 200],
                                                          [0, 0b10, 0b100000, 0b111111110, 0b1010101010100, 0b1000000000000])))
@pytest.mark.parametrize("tail", [1, 0])
def test_plan_invariants(n, mask, tail):
    pl = plan(n, mask, tail)
    assert sum(t if t else 1 for t in pl) == n                    # covers exactly n iterations
    assert len(pl) % 2 == n % 2                                   # buffer-flip parity of the reference
    assert all(t == 0 or (mask >> t) & 1 for t in pl)             # only validated pass sizes
    if n >= 1:
        assert pl[-1] == 0                                        # the last iteration is always literal
    if tail == 0 or mask == 0:
        assert pl[-min(n, 2):] == [0] * min(n, 2)                 # ... and so is the one before it, unless the pass before emits
    if mask == 0:
        assert pl == [0] * n


def test_plan_prefers_large_passes():
    pl = plan(80, 0b111111110, tail=0)
    assert sum(1 for t in pl if t == 0) == 2 and max(pl) == 8 and len(pl) <= 12
    pl = plan(80, 0b111111110)
    assert sum(1 for t in pl if t == 0) == 1 and pl[-2] > 0 and max(pl) == 8 and len(pl) <= 12
