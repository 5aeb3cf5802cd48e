"""Registry + type hooks that let the reference (`ap_gym` 0.5.0) hand out the B200 engine.

    import ap_gym, ap_gym_b200
    ap_gym_b200.register_envs(replace=True)               # every image / LIDAR id now vectorises through the engine
    env = ap_gym.make_vec("MNIST-v0", num_envs=65536)     # -> ap_gym_b200.ImageClassificationVectorEnv

What the reference does at these places (file:line under /root/reference/ap_gym):
  envs/registration.py:87-142     `register(id, entry_point, vector_entry_point, additional_wrappers, kwargs)` and its
                                  `<id>-sparse` twin (`SparsifyVectorWrapper` around the vector entry point)
  envs/registration.py:145-183    image ids: `vector_entry_point = ImageClassificationVectorEnv | ImageLocalizationVectorEnv`
                                  wrapped in the `*VectorLogWrapper`; kwargs = {"image_perception_config": ...}
  envs/registration.py:319-356    LIDAR ids: single-env entry point + TimeLimit(100) + regression LogWrapper, vectorised by
                                  gymnasium's SyncVectorEnv (:753-767); kwargs = {"dataset", "static_map"}
  active_perception_vector_env.py:311-320  `ensure_active_perception_vector_env`: isinstance(BaseActivePerceptionVectorEnv)
The engine emits the LogWrappers' `info["stats"]` itself (fused episode statistics), applies TimeLimit and the `-sparse`
reward rule in the kernels, so the hook installs the bare engine classes as vector entry points.

Nothing here is imported by `import ap_gym_b200`: the reference (and gymnasium) are only touched when a hook is called.
"""
from __future__ import annotations

from functools import partial
from typing import Any, Callable


def enable_reference_compat() -> bool:
    """Registers the engine's env classes as virtual subclasses of the reference's `BaseActivePerceptionVectorEnv`
    (an ABC), so that `ensure_active_perception_vector_env(env)` returns the env itself and the reference's vector
    wrappers accept it.  Needs `ap_gym` to be importable; returns False otherwise."""
    try:
        from ap_gym.active_perception_vector_env import BaseActivePerceptionVectorEnv
    except Exception:  # noqa: BLE001  (ap_gym or one of its dependencies is missing)
        return False
    from .image_env import ImageClassificationVectorEnv, ImageLocalizationVectorEnv
    from .lidar_env import LIDARLocalization2DVectorEnv

    for cls in (ImageClassificationVectorEnv, ImageLocalizationVectorEnv, LIDARLocalization2DVectorEnv):
        BaseActivePerceptionVectorEnv.register(cls)
    return True


def _closure_values(fn, depth=0):
    """Everything reachable through the closure cells / partial arguments of a callable (the reference hides the env
    class behind two lambdas and, for `-sparse`, a wrapper factory)."""
    if fn is None or depth > 6:
        return
    if isinstance(fn, partial):
        yield from _closure_values(fn.func, depth + 1)
        for a in list(fn.args) + list(fn.keywords.values()):
            yield a
            if callable(a):
                yield from _closure_values(a, depth + 1)
        return
    for cell in getattr(fn, "__closure__", None) or ():
        try:
            v = cell.cell_contents
        except ValueError:
            continue
        yield v
        if callable(v) and not isinstance(v, type):
            yield from _closure_values(v, depth + 1)


def _image_task_of(spec) -> str | None:
    """"cls" / "loc" for the reference's image ids, None for everything else."""
    if "image_perception_config" not in (spec.kwargs or {}):
        return None
    from ap_gym.envs.image_classification import ImageClassificationVectorEnv as RefCls
    from ap_gym.envs.image_localization import ImageLocalizationVectorEnv as RefLoc

    for v in _closure_values(spec.vector_entry_point):
        if v is RefCls:
            return "cls"
        if v is RefLoc:
            return "loc"
    return None


def _is_sparse(spec) -> bool:
    return spec.id.rsplit("-v", 1)[0].endswith("-sparse")


def _lidar_step_limit(spec, default=100) -> int:
    for w in spec.additional_wrappers or ():
        kw = getattr(w, "kwargs", None) or {}
        if "max_episode_steps" in kw:
            return int(kw["max_episode_steps"])
    return default


def engine_entry_point(spec, **engine_kwargs) -> Callable[..., Any] | None:
    """The engine-backed `vector_entry_point` for a registered reference spec, or None when the id is not on the
    engine's path (LightDark, hide-and-seek, ...)."""
    from .image_env import ImageClassificationVectorEnv, ImageLocalizationVectorEnv
    from .lidar_env import LIDARLocalization2DVectorEnv

    sparse = _is_sparse(spec)
    task = _image_task_of(spec)
    if task is not None:
        cls = ImageClassificationVectorEnv if task == "cls" else ImageLocalizationVectorEnv

        def make_image(num_envs, image_perception_config, render_mode="rgb_array", **kw):
            return cls(num_envs, image_perception_config, render_mode, sparse=sparse, **{**engine_kwargs, **kw})

        return make_image
    kwargs = spec.kwargs or {}
    if "dataset" in kwargs and "static_map" in kwargs and spec.vector_entry_point is None:
        step_limit = _lidar_step_limit(spec)

        def make_lidar(num_envs, dataset, static_map=False, render_mode="rgb_array", **kw):
            return LIDARLocalization2DVectorEnv(num_envs, dataset, render_mode, static_map=static_map,
                                                max_episode_steps=step_limit, sparse=sparse, **{**engine_kwargs, **kw})

        return make_lidar
    return None


def _folded_entry_point(spec):
    """Single-env entry point with the spec's `additional_wrappers` applied inside (gymnasium refuses
    `vector_entry_point` vectorisation while `additional_wrappers` is non-empty; the reference uses the same fold for its
    `-sparse` ids, registration.py:120-131)."""
    import gymnasium as gym

    entry, wrappers = spec.entry_point, tuple(spec.additional_wrappers or ())
    if not wrappers:
        return entry

    def make_single(*args, **kwargs):
        creator = gym.envs.registration.load_env_creator(entry) if isinstance(entry, str) else entry
        env = creator(*args, **kwargs)
        for w in wrappers:
            env = gym.envs.registration.load_env_creator(w.entry_point)(env=env, **(w.kwargs or {}))
        return env

    return make_single


def register_envs(namespace: str | None = "b200", replace: bool = False, **engine_kwargs) -> list[str]:
    """Registers the engine behind the reference's ids (call after `import ap_gym`, which registers them).

    namespace="b200"   adds `b200/<Name>-v0` next to every image / LIDAR id of the reference
    replace=True       ALSO swaps the `vector_entry_point` of the reference ids themselves, so that
                       `ap_gym.make_vec("MNIST-v0", num_envs=N)` returns the engine
    engine_kwargs      passed to every engine constructor (device=, check_errors=, ...)
    Returns the ids that now resolve to the engine.  LIDAR ids vectorise through `vector_entry_point` after this call
    (gymnasium picks it over SyncVectorEnv when `vectorization_mode` is None or "vector_entry_point")."""
    import gymnasium as gym

    enable_reference_compat()
    done = []
    for env_id, spec in list(gym.registry.items()):
        if getattr(spec, "namespace", None) not in (None, ""):
            continue
        entry = engine_entry_point(spec, **engine_kwargs)
        if entry is None:
            continue
        if namespace:
            new_id = f"{namespace}/{env_id}"
            if new_id not in gym.registry:
                gym.register(id=new_id, entry_point=_folded_entry_point(spec), vector_entry_point=entry,
                             kwargs=dict(spec.kwargs or {}), max_episode_steps=spec.max_episode_steps,
                             additional_wrappers=(), disable_env_checker=True)
            done.append(new_id)
        if replace:
            spec.entry_point = _folded_entry_point(spec)
            spec.additional_wrappers = ()
            spec.vector_entry_point = entry
            done.append(env_id)
    return done
