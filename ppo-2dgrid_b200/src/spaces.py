"""Minimal stand-ins for the two gymnasium spaces the reference touches (`action_space.n`,
`action_space.sample()`, `observation_space.shape`); gymnasium itself is not a dependency of the B200 build."""
from __future__ import annotations

import numpy as np


class Discrete:
    def __init__(self, n, seed=None):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return int(self._rng.integers(0, self.n))

    def contains(self, x):
        return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

    __contains__ = contains

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def __repr__(self):
        return f"Discrete({self.n})"


class Box:
    def __init__(self, low, high, shape, dtype=np.uint8):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"
