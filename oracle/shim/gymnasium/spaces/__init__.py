"""Minimal gymnasium.spaces (TEST-ONLY shim, see ../__init__.py)."""
from __future__ import annotations

import copy
from typing import Any, Generic, TypeVar

import numpy as np

from ..utils import seeding
from . import space as space  # noqa: F401  (gymnasium.spaces.space.Space is referenced by the reference)
from .space import Space

T = TypeVar("T")


def _short_repr(arr):
    arr = np.asarray(arr)
    if arr.size != 0 and np.min(arr) == np.max(arr):
        return str(np.min(arr))
    return str(arr)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is not None:
            shape = tuple(int(d) for d in shape)
        elif isinstance(low, np.ndarray):
            shape = low.shape
        elif isinstance(high, np.ndarray):
            shape = high.shape
        else:
            shape = (1,)
        low_a = np.full(shape, low, dtype=float) if np.isscalar(low) else np.asarray(low, dtype=float)
        high_a = np.full(shape, high, dtype=float) if np.isscalar(high) else np.asarray(high, dtype=float)
        low_a = np.broadcast_to(low_a, shape)
        high_a = np.broadcast_to(high_a, shape)
        self.bounded_below = -np.inf < low_a
        self.bounded_above = np.inf > high_a
        self.low = low_a.astype(dtype)
        self.high = high_a.astype(dtype)
        self.low_repr = _short_repr(self.low)
        self.high_repr = _short_repr(self.high)
        super().__init__(shape, dtype, seed)

    def is_bounded(self, manner="both"):
        below = bool(np.all(self.bounded_below))
        above = bool(np.all(self.bounded_above))
        return {"both": below and above, "below": below, "above": above}[manner]

    def sample(self, mask=None):
        high = self.high.astype(float)
        low = self.low.astype(float)
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp = ~self.bounded_below & self.bounded_above
        lowb = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        sample[unbounded] = self.np_random.normal(size=unbounded[unbounded].shape)
        sample[lowb] = self.np_random.exponential(size=lowb[lowb].shape) + low[lowb]
        sample[upp] = -self.np_random.exponential(size=upp[upp].shape) + high[upp]
        sample[bounded] = self.np_random.uniform(low=low[bounded], high=high[bounded], size=bounded[bounded].shape)
        return sample.astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low_repr}, {self.high_repr}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return (
            isinstance(other, Box)
            and self.shape == other.shape
            and self.dtype == other.dtype
            and np.array_equal(self.low, other.low)
            and np.array_equal(self.high, other.high)
        )


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = np.int64(n)
        self.start = np.int64(start)
        super().__init__((), np.int64, seed)

    def sample(self, mask=None):
        return self.start + self.np_random.integers(self.n)

    def contains(self, x):
        return bool(self.start <= int(x) < self.start + self.n)

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None, start=None):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        self.start = np.zeros(self.nvec.shape, dtype=dtype) if start is None else np.array(start, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self, mask=None):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype) + self.start

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(self.start <= x) and np.all(x - self.start < self.nvec))

    def __repr__(self):
        return f"MultiDiscrete({self.nvec})"

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        shape = tuple(int(d) for d in n) if isinstance(n, (tuple, list, np.ndarray)) else (int(n),)
        self.n = n
        super().__init__(shape, np.int8, seed)

    def sample(self, mask=None):
        return self.np_random.integers(low=0, high=2, size=self.shape, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))


class Dict(Space, Generic[T]):
    def __init__(self, spaces=None, seed=None, **spaces_kwargs):
        if spaces is None:
            spaces = {}
        if isinstance(spaces, Dict):
            spaces = spaces.spaces
        self.spaces = dict(spaces)
        self.spaces.update(spaces_kwargs)
        super().__init__(None, None, seed)

    def __getitem__(self, key):
        return self.spaces[key]

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __contains__(self, x):  # gymnasium semantics: membership of a sample
        return self.contains(x)

    def sample(self, mask=None):
        return {k: s.sample() for k, s in self.spaces.items()}

    def contains(self, x):
        return isinstance(x, dict) and x.keys() == self.spaces.keys() and all(
            self.spaces[k].contains(v) for k, v in x.items()
        )

    def seed(self, seed=None):
        super().seed(seed if not isinstance(seed, dict) else None)
        for i, s in enumerate(self.spaces.values()):
            s.seed(None if seed is None or isinstance(seed, dict) else int(self.np_random.integers(2**31)))

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {s}" for k, s in self.spaces.items()) + ")"

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces


class Tuple(Space, Generic[T]):
    def __init__(self, spaces=(), seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)

    def sample(self, mask=None):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return isinstance(x, (tuple, list)) and len(x) == len(self.spaces) and all(
            s.contains(v) for s, v in zip(self.spaces, x)
        )

    def __repr__(self):
        return "Tuple(" + ", ".join(str(s) for s in self.spaces) + ")"

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces
