"""``spaces.Box`` for the 4-float observation / 2-float action spaces.

Uses gymnasium's Box when gymnasium is importable (so the reference's ``isinstance(space, spaces.Box)``
checks see the real class); otherwise a minimal stand-in with the attributes the reference reads
(low, high, shape, dtype, sample, contains, seed) — twoseriescstr.py:74-85."""
from __future__ import annotations

import numpy as np

# The stand-in is always defined; `_box_class()` picks gymnasium's Box whenever gymnasium is importable AT CALL TIME (a test harness may
# put a gymnasium package on sys.path after this module was imported).

class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self.bounded_below = np.isfinite(self.low)
        self.bounded_above = np.isfinite(self.high)
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    def seed(self, seed=None):
        ss = np.random.SeedSequence(seed)
        self._np_random = np.random.Generator(np.random.PCG64(ss))
        return ss.entropy

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed()
        return self._np_random

    def sample(self, mask=None):
        return self.np_random.uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def is_bounded(self, manner="both"):
        return True

    def __eq__(self, other):
        return (
            hasattr(other, "low")
            and self.shape == tuple(other.shape)
            and np.allclose(self.low, other.low)
            and np.allclose(self.high, other.high)
        )

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


def _box_class():
    try:  # pragma: no cover - depends on the host environment
        from gymnasium import spaces as gym_spaces

        return gym_spaces.Box
    except ImportError:
        return Box


def observation_space():
    return _box_class()(low=np.full(4, -1.0, np.float32), high=np.full(4, 1.0, np.float32), dtype=np.float32)


def action_space():
    return _box_class()(low=np.full(2, -1.0, np.float32), high=np.full(2, 1.0, np.float32), dtype=np.float32)
