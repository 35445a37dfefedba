"""1-D row-strip decomposition of the X (slow) axis across the GPUs of one box.

`Partition` is pure index arithmetic (testable without a GPU).  Rank r owns global rows
[g0, g1); its local arrays hold the window [g0 - halo, g1 + halo) so that halo rows of the
neighbours' data sit in the same tensor as the owned rows.  Global edges keep the reference's
clamp-to-edge semantics (the window rows outside [0, X) exist in memory but are never read).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Partition:
    x_global: int
    rank: int = 0
    world: int = 1
    halo: int = 0
    #: optional strip boundaries (world + 1 increasing row indices from 0 to x_global): strips of unequal height, e.g. balanced
    #: by work instead of by rows (balanced_bounds); None = equal split
    bounds: tuple | None = None

    @staticmethod
    def single(x_global: int) -> "Partition":
        return Partition(x_global, 0, 1, 0)

    def __post_init__(self) -> None:
        if not (0 <= self.rank < self.world):
            raise ValueError(f"rank {self.rank} outside world of {self.world}")
        if self.world > 1 and self.halo < 2:
            raise ValueError("a multi-rank partition needs halo >= 2 (velocity BC / KK stencil reach +-2 rows)")
        if self.x_global < self.world * max(self.halo, 1):
            raise ValueError("strips thinner than the halo are not supported")
        if self.bounds is not None:
            b = tuple(int(x) for x in self.bounds)
            if len(b) != self.world + 1 or b[0] != 0 or b[-1] != self.x_global:
                raise ValueError(f"bounds must be {self.world + 1} row indices from 0 to {self.x_global}")
            if any(b[k + 1] - b[k] < max(self.halo, 1) for k in range(self.world)):
                raise ValueError("strips thinner than the halo are not supported")
            object.__setattr__(self, "bounds", b)

    def owned(self, rank: int | None = None) -> tuple[int, int]:
        """Global rows [g0, g1) owned by `rank` (balanced split, remainder to the low ranks)."""
        r = self.rank if rank is None else rank
        if self.bounds is not None:
            return self.bounds[r], self.bounds[r + 1]
        base, rem = divmod(self.x_global, self.world)
        g0 = r * base + min(r, rem)
        return g0, g0 + base + (1 if r < rem else 0)

    def window(self) -> tuple[int, int]:
        g0, g1 = self.owned()
        return g0 - self.halo, g1 + self.halo

    @property
    def has_lower(self) -> bool:  # a neighbour holding smaller row indices
        return self.rank > 0

    @property
    def has_upper(self) -> bool:
        return self.rank < self.world - 1


def balanced_bounds(row_weight, world: int, min_rows: int = 1) -> tuple:
    """Strip boundaries that give every rank about the same total weight: boundary k is the first row at which the running
    sum of `row_weight` reaches k / world of the total (every strip at least `min_rows` tall).  For scenes whose walls are
    unevenly distributed along i -- the fused Jacobi passes skip tiles that lie inside walls, so equal rows are not equal work."""
    w = np.asarray(row_weight, dtype=np.float64)
    X = int(w.shape[0])
    if world * min_rows > X:
        raise ValueError("not enough rows for that many strips")
    cum = np.concatenate(([0.0], np.cumsum(w)))
    bounds = [0]
    for k in range(1, world):
        r = int(np.searchsorted(cum, cum[-1] * k / world, side="left"))
        r = max(r, bounds[-1] + min_rows)
        r = min(r, X - (world - k) * min_rows)
        bounds.append(r)
    bounds.append(X)
    return tuple(bounds)
