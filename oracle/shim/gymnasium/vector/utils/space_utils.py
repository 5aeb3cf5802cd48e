from __future__ import annotations

import copy
from functools import singledispatch

import numpy as np

from ...spaces import Box, Dict, Discrete, MultiBinary, MultiDiscrete, Space, Tuple


@singledispatch
def batch_space(space, n: int = 1):
    raise TypeError(f"The space provided to `batch_space` is not a gymnasium Space instance, type: {type(space)}")


@batch_space.register(Box)
def _batch_space_box(space: Box, n: int = 1):
    repeats = tuple([n] + [1] * space.low.ndim)
    low, high = np.tile(space.low, repeats), np.tile(space.high, repeats)
    return Box(low=low, high=high, dtype=space.dtype, seed=copy.deepcopy(space.np_random))


@batch_space.register(Discrete)
def _batch_space_discrete(space: Discrete, n: int = 1):
    return MultiDiscrete(np.full((n,), space.n, dtype=space.dtype), dtype=space.dtype,
                         seed=copy.deepcopy(space.np_random), start=np.full((n,), space.start, dtype=space.dtype))


@batch_space.register(MultiDiscrete)
def _batch_space_multidiscrete(space: MultiDiscrete, n: int = 1):
    repeats = tuple([n] + [1] * space.nvec.ndim)
    low = np.tile(space.start, repeats)
    high = low + np.tile(space.nvec, repeats) - 1
    return Box(low=low, high=high, dtype=space.dtype, seed=copy.deepcopy(space.np_random))


@batch_space.register(MultiBinary)
def _batch_space_multibinary(space: MultiBinary, n: int = 1):
    return Box(low=0, high=1, shape=(n,) + space.shape, dtype=space.dtype, seed=copy.deepcopy(space.np_random))


@batch_space.register(Tuple)
def _batch_space_tuple(space: Tuple, n: int = 1):
    return Tuple(tuple(batch_space(s, n=n) for s in space.spaces), seed=copy.deepcopy(space.np_random))


@batch_space.register(Dict)
def _batch_space_dict(space: Dict, n: int = 1):
    return Dict({k: batch_space(s, n=n) for k, s in space.items()}, seed=copy.deepcopy(space.np_random))


def iterate(space, items):
    if isinstance(space, Dict):
        keys = list(space.keys())
        its = [iterate(space[k], items[k]) for k in keys]
        for vals in zip(*its):
            yield {k: v for k, v in zip(keys, vals)}
    elif isinstance(space, Tuple):
        if len(space.spaces) == 0:
            raise TypeError("cannot iterate an empty Tuple space")
        for vals in zip(*[iterate(s, it) for s, it in zip(space.spaces, items)]):
            yield tuple(vals)
    else:
        yield from iter(items)


def create_empty_array(space, n: int = 1, fn=np.zeros):
    if isinstance(space, Dict):
        return {k: create_empty_array(s, n=n, fn=fn) for k, s in space.items()}
    if isinstance(space, Tuple):
        return tuple(create_empty_array(s, n=n, fn=fn) for s in space.spaces)
    return fn((n,) + tuple(space.shape), dtype=space.dtype)


def concatenate(space, items, out):
    if isinstance(space, Dict):
        return {k: concatenate(s, [item[k] for item in items], out[k]) for k, s in space.items()}
    if isinstance(space, Tuple):
        return tuple(concatenate(s, [item[i] for item in items], out[i]) for i, s in enumerate(space.spaces))
    return np.stack(items, axis=0, out=out)
