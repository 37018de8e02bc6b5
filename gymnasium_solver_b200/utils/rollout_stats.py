"""Windowed / streaming statistics of the collector (reference: utils/rollout_stats.py:6-67)."""
from __future__ import annotations

import math
from collections import deque


class RollingWindow:
    """Fixed-length window with an O(1) running mean."""

    def __init__(self, maxlen: int):
        if maxlen <= 0:
            raise ValueError("RollingWindow maxlen must be > 0")
        self._items = deque(maxlen=int(maxlen))
        self._total = 0.0

    def append(self, value) -> None:
        if len(self._items) == self._items.maxlen:
            self._total -= float(self._items[0])
        self._items.append(value)
        self._total += float(value)

    def extend(self, values) -> None:
        for v in values:
            self.append(v)

    def mean(self) -> float:
        return self._total / len(self._items) if self._items else 0.0

    def __len__(self) -> int:
        return len(self._items)

    def __bool__(self) -> bool:
        return bool(self._items)


class RunningStats:
    """count / sum / sum-of-squares aggregates.  ``update`` takes an array; ``update_moments`` takes the
    (sum, sumsq, count) triple the engine's gs_moments kernel produces on device."""

    def __init__(self) -> None:
        self.count = 0
        self.sum = 0.0
        self.sum_squared = 0.0

    def update(self, values) -> None:
        import numpy as np

        v = np.asarray(values)
        if v.size == 0:
            return
        f = v.ravel().astype(np.float32, copy=False)
        self.update_moments(float(f.sum()), float((f * f).sum()), int(v.size))

    def update_moments(self, s: float, s2: float, n: int) -> None:
        if n <= 0:
            return
        self.count += int(n)
        self.sum += float(s)
        self.sum_squared += float(s2)

    def mean(self) -> float:
        return self.sum / self.count if self.count else 0.0

    def std(self) -> float:
        if not self.count:
            return 0.0
        m = self.mean()
        return math.sqrt(max(0.0, self.sum_squared / self.count - m * m))
