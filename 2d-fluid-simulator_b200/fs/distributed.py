"""1-D row-strip decomposition of the X (slow) axis across the GPUs of one box.

`Partition` is pure index arithmetic (testable without a GPU).  Rank r owns global rows
[g0, g1); its local arrays hold the window [g0 - halo, g1 + halo) so that halo rows of the
neighbours' data sit in the same tensor as the owned rows.  Global edges keep the reference's
clamp-to-edge semantics (the window rows outside [0, X) exist in memory but are never read).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Partition:
    x_global: int
    rank: int = 0
    world: int = 1
    halo: int = 0

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

    def owned(self, rank: int | None = None) -> tuple[int, int]:
        """Global rows [g0, g1) owned by `rank` (balanced split, remainder to the low ranks)."""
        r = self.rank if rank is None else rank
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
