from __future__ import annotations

from typing import Any, Generic, TypeVar

from .utils import seeding

ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")
WrapperObsType = TypeVar("WrapperObsType")
WrapperActType = TypeVar("WrapperActType")


class Env(Generic[ObsType, ActType]):
    metadata: dict[str, Any] = {"render_modes": []}
    render_mode = None
    spec = None
    action_space = None
    observation_space = None
    _np_random = None
    _np_random_seed = None

    def step(self, action):
        raise NotImplementedError

    def reset(self, *, seed=None, options=None):
        # gymnasium.Env.reset: (re)seed only when a seed is given
        if seed is not None:
            self._np_random, self._np_random_seed = seeding.np_random(seed)

    def render(self):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, self._np_random_seed = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value
        self._np_random_seed = -1

    def has_wrapper_attr(self, name):
        return hasattr(self, name)

    def get_wrapper_attr(self, name):
        return getattr(self, name)


class Wrapper(Env, Generic[WrapperObsType, WrapperActType, ObsType, ActType]):
    def __init__(self, env):
        self.env = env
        self._action_space = None
        self._observation_space = None
        self._metadata = None
        self._cached_spec = None

    def step(self, action):
        return self.env.step(action)

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def spec(self):
        return self.env.spec

    @property
    def action_space(self):
        if self._action_space is None:
            return self.env.action_space
        return self._action_space

    @action_space.setter
    def action_space(self, space):
        self._action_space = space

    @property
    def observation_space(self):
        if self._observation_space is None:
            return self.env.observation_space
        return self._observation_space

    @observation_space.setter
    def observation_space(self, space):
        self._observation_space = space

    @property
    def metadata(self):
        if self._metadata is None:
            return self.env.metadata
        return self._metadata

    @metadata.setter
    def metadata(self, value):
        self._metadata = value

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def np_random(self):
        return self.env.np_random

    @np_random.setter
    def np_random(self, value):
        self.env.np_random = value

    @property
    def _np_random(self):
        raise AttributeError("Can't access `_np_random` of a wrapper, use `.unwrapped._np_random`.")

    def get_wrapper_attr(self, name):
        if name in dir(self):
            return getattr(self, name)
        return self.env.get_wrapper_attr(name)
