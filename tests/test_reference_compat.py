"""CPU, needs /root/reference: the drop-in claim under gymnasium (test-only shim) and the reference's own wrappers.

The engine classes cannot step without a GPU, so a stub subclass (spaces and bookkeeping of the real class, a fake step)
stands in for the handle; everything else -- the gymnasium.vector.VectorEnv subclassing, the ABC registration behind
`ensure_active_perception_vector_env`, the reference's VectorLogWrapper / VectorToSingleWrapper / SparsifyVectorWrapper on
top, and the registry hook that makes `ap_gym.make_vec(id)` return the engine -- is the shipped code.  Runs in a
subprocess: the shim has to be importable BEFORE ap_gym_b200 is imported."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.reference
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
ROOT = sys.argv[1]
sys.path[:0] = [ROOT + "/oracle/shim", "/root/reference", ROOT + "/active-perception-gym_b200", ROOT]
import numpy as np
import gymnasium as gym
import ap_gym
import ap_gym_b200 as apg
from ap_gym_b200 import image_env, lidar_env, _capi
from ap_gym.active_perception_vector_env import BaseActivePerceptionVectorEnv, ensure_active_perception_vector_env

# ---- 1. class relations
for cls in (apg.ImageClassificationVectorEnv, apg.ImageLocalizationVectorEnv, apg.LIDARLocalization2DVectorEnv):
    assert issubclass(cls, gym.vector.VectorEnv), cls
assert apg.enable_reference_compat()
for cls in (apg.ImageClassificationVectorEnv, apg.ImageLocalizationVectorEnv, apg.LIDARLocalization2DVectorEnv):
    assert issubclass(cls, BaseActivePerceptionVectorEnv), cls


# ---- 2. a handle-less stand-in: the real class's spaces / properties, a fake step
class StubCls(apg.ImageClassificationVectorEnv):
    def __init__(self, num_envs, image_perception_config, render_mode="rgb_array", *, sparse=False, **kw):
        cfg = image_perception_config
        self.created_with = dict(num_envs=num_envs, sparse=sparse, kw=kw)
        self.num_envs, self._config, self._render_mode, self.sparse = num_envs, cfg, render_mode, bool(sparse)
        self._S0, self._S1 = cfg.sensor_size
        self._C, self._K = 1, int(cfg.dataset.num_classes)
        self._H = self._W = 28
        self._closed, self._spec, self._t = False, None, 0
        self._build_spaces()

    def reset(self, *, seed=None, options=None):
        self._t = 0
        self._labels = np.arange(self.num_envs) % self._K
        return self.observation_space.sample(), {"index": np.arange(self.num_envs)}

    def step(self, action):
        p = np.asarray(action["prediction"], np.float32)
        assert p.shape == (self.num_envs, self._K) and np.asarray(action["action"]).shape == (self.num_envs, 2)
        reset_step = self._t >= self._config.step_limit
        self._t = 0 if reset_step else self._t + 1
        term = np.full(self.num_envs, self._t >= self._config.step_limit)
        inner = self.loss_fn._inner if self.sparse else self.loss_fn
        loss = inner(p, self._labels, (self.num_envs,)).astype(np.float32)
        base = np.zeros(self.num_envs, np.float32)
        target = self._labels
        if self.sparse:
            w = term.astype(np.float32)
            target = {"target": self._labels, "weight": w}
            loss = loss * w
        info = {"index": np.arange(self.num_envs), "base_reward": base, "prediction": {"target": target, "loss": loss}}
        return self.observation_space.sample(), base - loss, term, np.zeros(self.num_envs, bool), info

    def close(self, **kw):
        self._closed = self.closed = True


ds = ap_gym.envs.image.CircleSquareDataset(image_shape=(28, 28), show_gradient=True)
env = StubCls(6, ap_gym.envs.image.ImagePerceptionConfig(dataset=ds))
assert isinstance(env, gym.vector.VectorEnv) and isinstance(env, BaseActivePerceptionVectorEnv)
assert ensure_active_perception_vector_env(env) is env          # not re-wrapped as a pseudo env
assert env.metadata["autoreset_mode"] == gym.vector.AutoresetMode.NEXT_STEP
assert env.prediction_space.shape == (6, 2) and env.single_prediction_space.shape == (2,)
assert env.single_prediction_target_space.n == 2 and env.inner_action_space.shape == (6, 2)

# the reference's own wrappers accept it and run on top of it
from ap_gym import ActiveClassificationVectorLogWrapper
wrapped = ActiveClassificationVectorLogWrapper(env)
wrapped.reset(seed=0)
rng = np.random.default_rng(0)
for s in range(17):
    act = {"action": rng.uniform(-1, 1, (6, 2)).astype(np.float32), "prediction": rng.standard_normal((6, 2)).astype(np.float32)}
    obs, rew, te, tr, info = wrapped.step(act)
    if s == 15:
        assert te.all() and "stats" in info and "avg_accuracy" in info["stats"]["scalar"]
assert wrapped.loss_fn is env.loss_fn and wrapped.prediction_target_space == env.prediction_target_space

# `-sparse`: Dict target space + weighted loss, consistent with the reward (SparsifyVectorWrapper contract)
senv = StubCls(4, ap_gym.envs.image.ImagePerceptionConfig(dataset=ds), sparse=True)
assert isinstance(senv.loss_fn, apg.WeightedLossFn)
assert set(senv.single_prediction_target_space.spaces) == {"target", "weight"}
assert set(senv.prediction_target_space.spaces) == {"target", "weight"}
senv.reset(seed=0)
for s in range(17):
    act = {"action": np.zeros((4, 2), np.float32), "prediction": rng.standard_normal((4, 2)).astype(np.float32)}
    obs, rew, te, tr, info = senv.step(act)
    again = senv.loss_fn(act["prediction"], info["prediction"]["target"], (4,))
    np.testing.assert_allclose(info["base_reward"] - rew, again, rtol=1e-6, atol=1e-7)

# ---- 3. the registry hook: make_vec(id) hands out the engine class
image_env.ImageClassificationVectorEnv = StubCls          # what engine_entry_point() resolves (no GPU here)
created = {}


class StubLidar:
    def __init__(self, num_envs, dataset, render_mode="rgb_array", **kw):
        created.update(num_envs=num_envs, dataset=dataset, kw=kw)
        self.num_envs = num_envs

    unwrapped = property(lambda self: self)


BaseActivePerceptionVectorEnv.register(StubLidar)
lidar_env.LIDARLocalization2DVectorEnv = StubLidar
ids = apg.register_envs(namespace="b200", replace=True, check_errors="never")
assert "CircleSquare-v0" in ids and "b200/CircleSquare-v0" in ids and "LIDARLocRooms-v0" in ids
assert "CircleSquare-v0-sparse" in ids or "CircleSquare-sparse-v0" in ids
assert not any("LightDark" in i for i in ids)
e1 = ap_gym.make_vec("CircleSquare-v0", num_envs=5)
assert type(e1) is StubCls and e1.created_with == dict(num_envs=5, sparse=False, kw={"check_errors": "never"})
assert e1.spec is not None and e1.spec.max_episode_steps == 16
e2 = ap_gym.make_vec("b200/CircleSquare-v0", num_envs=3)
assert type(e2) is StubCls
sparse_id = "CircleSquare-sparse-v0" if "CircleSquare-sparse-v0" in ids else "CircleSquare-v0-sparse"
e3 = ap_gym.make_vec(sparse_id, num_envs=2)
assert type(e3) is StubCls and e3.sparse
e4 = gym.make_vec("LIDARLocRooms-v0", num_envs=7)
assert type(e4) is StubLidar and created["num_envs"] == 7 and created["kw"]["static_map"] is False
assert created["kw"]["max_episode_steps"] == 100 and created["kw"]["sparse"] is False
# the single-env path of a LIDAR id keeps its TimeLimit + LogWrapper (folded into the entry point)
single = gym.make("LIDARLocRooms-v0")
names = []
w = single
while hasattr(w, "env"):
    names.append(type(w).__name__)
    w = w.env
assert "TimeLimit" in names and any("Log" in n for n in names), names
print("COMPAT OK")
'''


def test_engine_envs_are_drop_ins_for_the_reference_api():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "COMPAT OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


HIDE_AND_SEEK = r'''
import sys
ROOT = sys.argv[1]
sys.path[:0] = [ROOT + "/oracle/shim", "/root/reference", ROOT + "/active-perception-gym_b200", ROOT]
import numpy as np
import ap_gym
from ap_gym.envs.image import CircleSquareDataset, ImagePerceptionConfig
from ap_gym.envs.image_classification import ImageClassificationVectorEnv
from ap_gym.envs.circle_square_catch_or_flee import CircleSquareHideAndSeekVectorWrapper as RefW
from ap_gym_b200 import CircleSquareHideAndSeekVectorWrapper as W

ds = CircleSquareDataset(image_shape=(28, 28), show_gradient=True)


class EngineSurface:  # what the wrapper touches of an engine env, fed by the reference's own vector env
    def __init__(self, inner):
        self.inner, self.config, self._W, self._H = inner, inner.config, 28, 28
        self.prediction_space = inner.prediction_space

    def reset(self, **kw):
        return self.inner.reset(**kw)

    def step(self, a):
        return self.inner.step(a)


ref = RefW(ImageClassificationVectorEnv(8, ImagePerceptionConfig(dataset=ds)))
mine = W(EngineSurface(ImageClassificationVectorEnv(8, ImagePerceptionConfig(dataset=ds))))
# (the reference's own mask_prediction=True variant raises KeyError in reset(): its info has no "prediction" there)
masked = W(EngineSurface(ImageClassificationVectorEnv(8, ImagePerceptionConfig(dataset=ds))), mask_prediction=True)
ref.reset(seed=3)
mine.reset(seed=3)
_, minfo = masked.reset(seed=3)
assert minfo["prediction"]["target"] == ()
rng = np.random.default_rng(0)
for s in range(20):
    a = {"action": rng.uniform(-1, 1, (8, 2)).astype(np.float32), "prediction": rng.standard_normal((8, 2)).astype(np.float32)}
    o1, r1, t1, _, i1 = ref.step(a)
    o2, r2, t2, _, i2 = mine.step(a)
    o3, r3, t3, _, i3 = masked.step(a)
    assert np.allclose(r1, r2, rtol=1e-6, atol=1e-7), s
    assert np.allclose(i1["base_reward"], i2["base_reward"], rtol=1e-6, atol=1e-7)
    # no prediction: reward = base reward (movement cost + hide-and-seek term), identical term as above
    assert np.allclose(r3, i3["base_reward"]) and np.allclose(i3["base_reward"], i2["base_reward"], rtol=1e-6, atol=1e-7)
    assert i3["prediction"]["target"] == ()
print("HIDE OK")
'''


def test_hide_and_seek_reward_term_matches_the_reference_wrapper():
    r = subprocess.run([sys.executable, "-c", HIDE_AND_SEEK, ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "HIDE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
