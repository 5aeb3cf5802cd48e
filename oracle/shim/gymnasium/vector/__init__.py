"""Minimal gymnasium.vector (TEST-ONLY shim): VectorEnv, VectorWrapper, SyncVectorEnv (NEXT_STEP autoreset)."""
from __future__ import annotations

import copy
import enum
from typing import Any, Generic, TypeVar

import numpy as np

from ..utils import seeding
from . import utils
from .utils import batch_space, concatenate, create_empty_array, iterate

ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")
ArrayType = TypeVar("ArrayType")


class AutoresetMode(enum.Enum):
    NEXT_STEP = "NextStep"
    SAME_STEP = "SameStep"
    DISABLED = "Disabled"


class VectorEnv(Generic[ObsType, ActType, ArrayType]):
    metadata: dict[str, Any] = {}
    spec = None
    render_mode = None
    closed = False
    observation_space = None
    action_space = None
    single_observation_space = None
    single_action_space = None
    num_envs: int
    _np_random = None
    _np_random_seed = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, self._np_random_seed = seeding.np_random(seed)

    def step(self, actions):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self, **kwargs):
        if self.closed:
            return
        self.close_extras(**kwargs)
        self.closed = True

    def close_extras(self, **kwargs):
        pass

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, self._np_random_seed = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value
        self._np_random_seed = -1

    @property
    def unwrapped(self):
        return self

    def _add_info(self, vector_infos, env_info, env_num):
        """gymnasium VectorEnv._add_info: dict-of-arrays with `_key` validity masks."""
        for key, value in env_info.items():
            if isinstance(value, dict):
                array = self._add_info(vector_infos.get(key, {}), value, env_num)
            else:
                if key not in vector_infos:
                    if type(value) in [int, float, bool] or issubclass(type(value), np.number):
                        array = np.zeros(self.num_envs, dtype=type(value))
                    elif isinstance(value, np.ndarray):
                        array = np.zeros((self.num_envs, *value.shape), dtype=value.dtype)
                    else:
                        array = np.full(self.num_envs, fill_value=None, dtype=object)
                else:
                    array = vector_infos[key]
                array[env_num] = value
            array_mask = vector_infos.get(f"_{key}", np.zeros(self.num_envs, dtype=np.bool_))
            array_mask[env_num] = True
            vector_infos[key], vector_infos[f"_{key}"] = array, array_mask
        return vector_infos


class VectorWrapper(VectorEnv):
    def __init__(self, env):
        self.env = env
        assert isinstance(env, VectorEnv)
        self._observation_space = None
        self._action_space = None
        self._single_observation_space = None
        self._single_action_space = None
        self._metadata = None

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, actions):
        return self.env.step(actions)

    def render(self):
        return self.env.render()

    def close(self, **kwargs):
        return self.env.close(**kwargs)

    def close_extras(self, **kwargs):
        return self.env.close_extras(**kwargs)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def _space_prop(name):  # noqa: N805
        private = f"_{name}"

        def getter(self):
            value = getattr(self, private)
            return getattr(self.env, name) if value is None else value

        def setter(self, space):
            setattr(self, private, space)

        return property(getter, setter)

    observation_space = _space_prop("observation_space")
    action_space = _space_prop("action_space")
    single_observation_space = _space_prop("single_observation_space")
    single_action_space = _space_prop("single_action_space")
    metadata = _space_prop("metadata")
    del _space_prop

    @property
    def num_envs(self):
        return self.env.num_envs

    @property
    def np_random(self):
        return self.env.np_random

    @np_random.setter
    def np_random(self, value):
        self.env.np_random = value

    @property
    def spec(self):
        return self.env.spec

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def closed(self):
        return self.env.closed

    @closed.setter
    def closed(self, value):
        self.env.closed = value


class SyncVectorEnv(VectorEnv):
    """gymnasium.vector.SyncVectorEnv restated (gymnasium 1.1, sync_vector_env.py): a python loop over
    sub-envs with per-env NEXT_STEP autoreset; `reset(seed=s)` seeds sub-env i with s+i; rewards f64."""

    def __init__(self, env_fns, copy: bool = True, observation_mode="same", autoreset_mode=AutoresetMode.NEXT_STEP):
        self.copy = copy
        self.env_fns = env_fns
        self.observation_mode = observation_mode
        self.autoreset_mode = autoreset_mode
        assert autoreset_mode == AutoresetMode.NEXT_STEP, "shim: only NEXT_STEP is restated"
        self.envs = [env_fn() for env_fn in env_fns]
        self.num_envs = len(self.envs)
        self.metadata = dict(self.envs[0].metadata)
        self.metadata["autoreset_mode"] = self.autoreset_mode
        self.render_mode = self.envs[0].render_mode
        self.single_action_space = self.envs[0].action_space
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.single_observation_space = self.envs[0].observation_space
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        self._env_obs = [None for _ in range(self.num_envs)]
        self._observations = create_empty_array(self.single_observation_space, n=self.num_envs)
        self._rewards = np.zeros((self.num_envs,), dtype=np.float64)
        self._terminations = np.zeros((self.num_envs,), dtype=np.bool_)
        self._truncations = np.zeros((self.num_envs,), dtype=np.bool_)
        self._autoreset_envs = np.zeros((self.num_envs,), dtype=np.bool_)

    @property
    def np_random_seed(self):
        return tuple(env.get_wrapper_attr("np_random_seed") for env in self.envs)

    @property
    def np_random(self):
        return tuple(env.np_random for env in self.envs)

    def reset(self, *, seed=None, options=None):
        if seed is None:
            seed = [None for _ in range(self.num_envs)]
        elif isinstance(seed, (int, np.integer)):
            seed = [int(seed) + i for i in range(self.num_envs)]
        assert len(seed) == self.num_envs
        infos = {}
        for i, (env, single_seed) in enumerate(zip(self.envs, seed)):
            self._env_obs[i], env_info = env.reset(seed=single_seed, options=options)
            infos = self._add_info(infos, env_info, i)
        self._terminations = np.zeros((self.num_envs,), dtype=np.bool_)
        self._truncations = np.zeros((self.num_envs,), dtype=np.bool_)
        self._autoreset_envs = np.zeros((self.num_envs,), dtype=np.bool_)
        self._observations = concatenate(self.single_observation_space, self._env_obs, self._observations)
        return (copy.deepcopy(self._observations) if self.copy else self._observations), infos

    def step(self, actions):
        actions = iterate(self.action_space, actions)
        infos = {}
        for i, action in enumerate(actions):
            if self._autoreset_envs[i]:
                self._env_obs[i], env_info = self.envs[i].reset()
                self._rewards[i] = 0.0
                self._terminations[i] = False
                self._truncations[i] = False
            else:
                (
                    self._env_obs[i],
                    self._rewards[i],
                    self._terminations[i],
                    self._truncations[i],
                    env_info,
                ) = self.envs[i].step(action)
            infos = self._add_info(infos, env_info, i)
        self._observations = concatenate(self.single_observation_space, self._env_obs, self._observations)
        self._autoreset_envs = np.logical_or(self._terminations, self._truncations)
        return (
            copy.deepcopy(self._observations) if self.copy else self._observations,
            np.copy(self._rewards),
            np.copy(self._terminations),
            np.copy(self._truncations),
            infos,
        )

    def render(self):
        return tuple(env.render() for env in self.envs)

    def close_extras(self, **kwargs):
        for env in getattr(self, "envs", []):
            env.close()


class AsyncVectorEnv(VectorEnv):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("shim: AsyncVectorEnv (multiprocessing) is not restated")
