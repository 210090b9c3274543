from collections import OrderedDict
from typing import Any, Optional, Sequence as _Seq

import numpy as np

from gymnasium.utils import seeding
from gymnasium.spaces import utils  # noqa: E402,F401  (populated below)


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    def __class_getitem__(cls, item):
        return cls

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed()
        return self._np_random

    def seed(self, seed=None):
        self._np_random, seed = seeding.np_random(seed)
        return seed

    def sample(self, mask=None):
        raise NotImplementedError

    def contains(self, x) -> bool:
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)

    @property
    def is_np_flattenable(self):
        return False


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            if np.isscalar(low) and np.isscalar(high):
                shape = (1,)
            else:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
        self.bounded_below = -np.inf < self.low
        self.bounded_above = np.inf > self.high
        super().__init__(shape, dtype, seed)

    @property
    def is_np_flattenable(self):
        return True

    def is_bounded(self, manner="both"):
        below = bool(np.all(self.bounded_below))
        above = bool(np.all(self.bounded_above))
        return {"both": below and above, "below": below, "above": above}[manner]

    def sample(self, mask=None):
        # bounded float Box: gymnasium draws uniform(low, high) from the space's own PCG64
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp = ~self.bounded_below & self.bounded_above
        low_b = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        sample[unbounded] = self.np_random.normal(size=unbounded[unbounded].shape)
        sample[low_b] = self.np_random.exponential(size=low_b[low_b].shape) + self.low[low_b]
        sample[upp] = -self.np_random.exponential(size=upp[upp].shape) + high[upp]
        sample[bounded] = self.np_random.uniform(low=self.low[bounded], high=high[bounded], size=bounded[bounded].shape)
        if self.dtype.kind in "iu":
            sample = np.floor(sample)
        return sample.astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(
            np.can_cast(x.dtype, self.dtype) and x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)
        )

    def __eq__(self, other):
        return (
            isinstance(other, Box)
            and self.shape == other.shape
            and self.dtype == other.dtype
            and np.allclose(self.low, other.low)
            and np.allclose(self.high, other.high)
        )

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)
        super().__init__((), np.int64, seed)

    def sample(self, mask=None):
        return np.int64(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None, start=None):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        self.start = np.zeros_like(self.nvec) if start is None else np.array(start, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self, mask=None):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype) + self.start

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x - self.start >= 0) and np.all(x - self.start < self.nvec))

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        self.n = n
        shape = (n,) if np.isscalar(n) else tuple(n)
        super().__init__(shape, np.int8, seed)

    def sample(self, mask=None):
        return self.np_random.integers(0, 2, size=self.shape, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))

    def __eq__(self, other):
        return isinstance(other, MultiBinary) and self.shape == other.shape


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kwargs):
        if spaces is None:
            spaces = {}
        if isinstance(spaces, (dict, OrderedDict)):
            spaces = OrderedDict(spaces)
        else:
            spaces = OrderedDict(spaces)
        spaces.update(kwargs)
        self.spaces = spaces
        super().__init__(None, None, seed)

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def values(self):
        return self.spaces.values()

    def __getitem__(self, key):
        return self.spaces[key]

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def sample(self, mask=None):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        return isinstance(x, dict) and all(k in x and self.spaces[k].contains(x[k]) for k in self.spaces)

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def sample(self, mask=None):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(s.contains(p) for s, p in zip(self.spaces, x))


class Sequence(Space):
    def __init__(self, space, seed=None, stack=False):
        self.feature_space = space
        super().__init__(None, None, seed)


class Text(Space):
    def __init__(self, max_length, min_length=1, charset=None, seed=None):
        self.max_length = max_length
        super().__init__((), None, seed)


class Graph(Space):
    def __init__(self, node_space, edge_space, seed=None):
        self.node_space = node_space
        self.edge_space = edge_space
        super().__init__(None, None, seed)


def _flatdim(space) -> int:
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return int(space.n)
    if isinstance(space, MultiDiscrete):
        return int(np.sum(space.nvec))
    if isinstance(space, MultiBinary):
        return int(np.prod(space.shape))
    if isinstance(space, Dict):
        return sum(_flatdim(s) for s in space.spaces.values())
    if isinstance(space, Tuple):
        return sum(_flatdim(s) for s in space.spaces)
    raise NotImplementedError(type(space))


utils.flatdim = _flatdim
