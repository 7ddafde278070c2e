"""Box and Discrete spaces (the two the classic-control hot path uses).

Reference: space/box.py:16-132, space/discrete.py:15-78.  Only the glue the hot path needs:
`flat_size`, `shape`, bounds (action clipping, algorithm/ppo.py:228-235) and `n`.
"""

from __future__ import annotations

import math

import numpy as np


class AbstractSpace:
    pass


class Box(AbstractSpace):
    def __init__(self, low, high, shape=None):
        low = np.asarray(low, dtype=np.float32)
        high = np.asarray(high, dtype=np.float32)
        if shape is not None:
            low = np.broadcast_to(low, shape).copy()
            high = np.broadcast_to(high, shape).copy()
        if low.shape != high.shape:
            raise ValueError("Box low and high must have the same shape.")
        self.low, self.high = low, high

    @property
    def shape(self) -> tuple[int, ...]:
        return self.low.shape

    @property
    def flat_size(self) -> int:
        return int(math.prod(self.shape))

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(x.shape == self.shape and (x >= self.low).all() and (x <= self.high).all())

    def __repr__(self):
        return f"Box(low={self.low}, high={self.high})"


class Discrete(AbstractSpace):
    def __init__(self, n: int):
        if n <= 0:
            raise ValueError("Discrete n must be positive.")
        self.n = int(n)

    @property
    def shape(self) -> tuple[int, ...]:
        return ()

    @property
    def flat_size(self) -> int:
        return 1

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(x.shape == () and 0 <= int(x) < self.n)

    def __repr__(self):
        return f"Discrete({self.n})"


__all__ = ["AbstractSpace", "Box", "Discrete"]
