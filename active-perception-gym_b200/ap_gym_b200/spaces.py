"""Observation / action / prediction spaces reported by the engine.

When `gymnasium` is importable its classes are used, so the spaces are the very objects a reference user
expects (`ap_gym/image_space.py:9-62`, `ap_gym/logit_space.py:4-23`,
`ap_gym/active_perception_env.py:27-59`).  gymnasium is absent from the build image, so a small
self-contained subset (shape / dtype / low / high / n / nvec / spaces, `contains`, `sample`) stands in.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as _gym

    _HAVE_GYM = True
except Exception:  # noqa: BLE001
    _gym = None
    _HAVE_GYM = False

if _HAVE_GYM:
    Space = _gym.spaces.Space
    Box = _gym.spaces.Box
    Discrete = _gym.spaces.Discrete
    MultiDiscrete = _gym.spaces.MultiDiscrete
    Dict = _gym.spaces.Dict
    try:
        AutoresetMode = _gym.vector.AutoresetMode
        NEXT_STEP = AutoresetMode.NEXT_STEP
    except AttributeError:  # gymnasium < 1.1
        NEXT_STEP = "NextStep"
    # the engine's environments ARE gymnasium vector envs when gymnasium is there: VectorWrapper and the reference's
    # wrappers assert isinstance(env, gymnasium.vector.VectorEnv)
    VectorEnvBase = _gym.vector.VectorEnv
else:
    VectorEnvBase = object

    class Space:
        def __init__(self, shape=None, dtype=None, seed=None):
            self._shape = None if shape is None else tuple(int(s) for s in shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        @property
        def shape(self):
            return self._shape

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def __contains__(self, x):
            return self.contains(x)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            super().__init__(shape, dtype, seed)
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape).astype(self.dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            bounded = np.isfinite(self.low) & np.isfinite(self.high)
            u = self._rng.uniform(lo, hi, self.shape)
            g = self._rng.standard_normal(self.shape)
            return np.where(bounded, u, g).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        def __eq__(self, o):
            return isinstance(o, Box) and self.shape == o.shape and self.dtype == o.dtype and \
                np.array_equal(self.low, o.low) and np.array_equal(self.high, o.high)

    class Discrete(Space):
        def __init__(self, n, seed=None):
            super().__init__((), np.int64, seed)
            self.n = np.int64(n)

        def contains(self, x):
            return bool(0 <= int(x) < self.n)

        def sample(self):
            return self._rng.integers(self.n)

        def __repr__(self):
            return f"Discrete({self.n})"

        def __eq__(self, o):
            return isinstance(o, Discrete) and self.n == o.n

    class MultiDiscrete(Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec = np.asarray(nvec, dtype=dtype)
            super().__init__(self.nvec.shape, dtype, seed)

        def contains(self, x):
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.all(x >= 0) and np.all(x < self.nvec))

        def sample(self):
            return (self._rng.random(self.shape) * self.nvec).astype(self.dtype)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec})"

        def __eq__(self, o):
            return isinstance(o, MultiDiscrete) and np.array_equal(self.nvec, o.nvec)

    class Dict(Space):
        def __init__(self, spaces=None, seed=None):
            super().__init__(None, None, seed)
            self.spaces = dict(spaces or {})

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def values(self):
            return self.spaces.values()

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def contains(self, x):
            return isinstance(x, dict) and x.keys() == self.spaces.keys() and all(
                self.spaces[k].contains(v) for k, v in x.items())

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {s}" for k, s in self.spaces.items()) + ")"

        def __eq__(self, o):
            return isinstance(o, Dict) and self.spaces == o.spaces

    NEXT_STEP = "NextStep"


class ImageSpace(Box):
    """Box with shape (*batch, height, width, channels) in [0, 1]  (ap_gym/image_space.py:9-62)."""

    def __init__(self, width, height, channels, batch_shape=(), dtype=np.float32, seed=None, low=0.0, high=1.0):
        super().__init__(low, high, (*batch_shape, height, width, channels), dtype, seed)

    height = property(lambda self: self.shape[-3])
    width = property(lambda self: self.shape[-2])
    channels = property(lambda self: self.shape[-1])
    batch_shape = property(lambda self: self.shape[:-3])

    def __repr__(self):
        return f"ImageSpace({self.shape}, {self.dtype})"


class LogitSpace(Box):
    """Unbounded Box holding class logits (ap_gym/logit_space.py:4-23)."""

    def __repr__(self):
        return f"LogitSpace({self.shape}, {self.dtype})"


class ActivePerceptionActionSpace(Dict):
    """Dict{"action", "prediction"} (ap_gym/active_perception_env.py:27-59)."""

    def __init__(self, inner_action_space, prediction_space, seed=None):
        super().__init__({"action": inner_action_space, "prediction": prediction_space}, seed=seed)

    @property
    def inner_action_space(self):
        return self["action"]

    @property
    def prediction_space(self):
        return self["prediction"]


def sparse_target_spaces(single_target_space, n: int):
    """`<id>-sparse` ids (SparsifyVectorWrapper, ap_gym/sparsify_wrapper.py:49-59): the prediction target becomes
    Dict{"target": <space>, "weight": Box(0, 1, (), float32)}; returns (single, batched)."""
    single = Dict({"target": single_target_space, "weight": Box(0, 1, (), np.float32)})
    t = single_target_space
    if isinstance(t, Discrete):
        bt = MultiDiscrete(np.full((n,), t.n))
    else:
        bt = batch_box(t, n)
    batched = Dict({"target": bt, "weight": Box(0, 1, (n,), np.float32)})
    return single, batched


def batch_box(space: Box, n: int, cls=None, **kw) -> Box:
    """gymnasium.vector.utils.batch_space for Box-like spaces: tile low/high along a new leading axis."""
    low = np.tile(space.low, (n,) + (1,) * space.low.ndim)
    high = np.tile(space.high, (n,) + (1,) * space.high.ndim)
    if cls is ImageSpace:
        return ImageSpace(space.shape[-2], space.shape[-3], space.shape[-1], (n,) + tuple(space.shape[:-3]),
                          space.dtype, None, low, high)
    return (cls or Box)(low, high, low.shape, space.dtype)
