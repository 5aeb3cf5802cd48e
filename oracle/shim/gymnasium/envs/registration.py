"""Minimal gymnasium.envs.registration (TEST-ONLY shim)."""
from __future__ import annotations

import copy
import enum
import importlib
import re
from dataclasses import dataclass, field
from typing import Any, Callable

from .. import error

EnvCreator = Callable[..., Any]
VectorEnvCreator = Callable[..., Any]

ENV_ID_RE = re.compile(r"^(?:(?P<namespace>[\w:-]+)\/)?(?:(?P<name>[\w:.-]+?))(?:-v(?P<version>\d+))?$")


class VectorizeMode(enum.Enum):
    ASYNC = "async"
    SYNC = "sync"
    VECTOR_ENTRY_POINT = "vector_entry_point"


@dataclass
class WrapperSpec:
    name: str
    entry_point: str
    kwargs: dict | None


@dataclass
class EnvSpec:
    id: str
    entry_point: EnvCreator | str | None = None
    reward_threshold: float | None = None
    nondeterministic: bool = False
    max_episode_steps: int | None = None
    order_enforce: bool = True
    disable_env_checker: bool = False
    kwargs: dict = field(default_factory=dict)
    additional_wrappers: tuple = ()
    vector_entry_point: VectorEnvCreator | str | None = None


registry: dict[str, EnvSpec] = {}


def parse_env_id(env_id: str):
    match = ENV_ID_RE.fullmatch(env_id)
    if not match:
        raise error.Error(f"Malformed environment ID: {env_id}")
    ns, name, version = match.group("namespace", "name", "version")
    return ns, name, (int(version) if version is not None else None)


def get_env_id(ns, name, version):
    full_name = name
    if ns is not None:
        full_name = f"{ns}/{name}"
    if version is not None:
        full_name = f"{full_name}-v{version}"
    return full_name


def load_env_creator(name: str):
    mod_name, attr_name = name.split(":")
    return getattr(importlib.import_module(mod_name), attr_name)


def register(id, entry_point=None, reward_threshold=None, nondeterministic=False, max_episode_steps=None,
             order_enforce=True, disable_env_checker=False, additional_wrappers=(), vector_entry_point=None,
             kwargs=None):
    registry[id] = EnvSpec(
        id=id, entry_point=entry_point, reward_threshold=reward_threshold, nondeterministic=nondeterministic,
        max_episode_steps=max_episode_steps, order_enforce=order_enforce, disable_env_checker=disable_env_checker,
        kwargs=dict(kwargs or {}), additional_wrappers=tuple(additional_wrappers),
        vector_entry_point=vector_entry_point,
    )


def spec(env_id):
    if env_id not in registry:
        raise error.NameNotFound(f"Environment `{env_id}` doesn't exist.")
    return registry[env_id]


def make(id, max_episode_steps=None, disable_env_checker=None, **kwargs):
    """gym.make: entry_point(**kwargs) -> [PassiveEnvChecker, OrderEnforcing: transparent here] ->
    additional_wrappers in order (innermost first)."""
    env_spec = id if isinstance(id, EnvSpec) else spec(id)
    env_kwargs = {**copy.copy(env_spec.kwargs), **kwargs}
    creator = env_spec.entry_point
    if isinstance(creator, str):
        creator = load_env_creator(creator)
    env = creator(**env_kwargs)
    try:
        env.unwrapped.spec = copy.copy(env_spec)
    except AttributeError:
        pass
    if max_episode_steps is not None or env_spec.max_episode_steps is not None:
        raise NotImplementedError("shim: gymnasium.wrappers.TimeLimit is not provided")
    for wrapper_spec in env_spec.additional_wrappers:
        env = load_env_creator(wrapper_spec.entry_point)(env=env, **(wrapper_spec.kwargs or {}))
    return env


def make_vec(id, num_envs=1, vectorization_mode=None, vector_kwargs=None, wrappers=None, **kwargs):
    from ..vector import SyncVectorEnv

    env_spec = id if isinstance(id, EnvSpec) else spec(id)
    vector_kwargs = dict(vector_kwargs or {})
    if vectorization_mode is None:
        vectorization_mode = (
            VectorizeMode.VECTOR_ENTRY_POINT if env_spec.vector_entry_point is not None else VectorizeMode.SYNC
        )
    elif isinstance(vectorization_mode, str):
        vectorization_mode = VectorizeMode(vectorization_mode)

    if vectorization_mode == VectorizeMode.VECTOR_ENTRY_POINT:
        if len(env_spec.additional_wrappers) > 0:
            raise error.Error("Cannot use `vector_entry_point` vectorization mode with additional_wrappers")
        creator = env_spec.vector_entry_point
        if isinstance(creator, str):
            creator = load_env_creator(creator)
        env = creator(num_envs=num_envs, **{**env_spec.kwargs, **kwargs}, **vector_kwargs)
    elif vectorization_mode == VectorizeMode.SYNC:
        def create_single_env():
            single = make(env_spec, **kwargs)
            for w in wrappers or ():
                single = w(single)
            return single

        env = SyncVectorEnv(env_fns=[create_single_env for _ in range(num_envs)], **vector_kwargs)
    else:
        raise NotImplementedError("shim: AsyncVectorEnv is not provided")
    try:
        env.unwrapped.spec = copy.copy(env_spec)
    except AttributeError:
        pass
    return env
