"""Drop-in vectorised LIDAR localisation environment backed by the fused sm_100a step kernel.

`LIDARLocalization2DVectorEnv(num_envs, dataset, ...)` is the engine's equivalent of what
`ap_gym.make_vec("LIDARLoc*-v0", num_envs)` assembles from python objects (ap_gym/envs/registration.py:319-356,
753-767): a gymnasium SyncVectorEnv (NEXT_STEP per-env autoreset, env i seeded with seed+i) over
`TimeLimit(LIDARLocalization2DEnv(...), 100, issue_termination=True)`.  The constructor keeps the reference's
keyword arguments (ap_gym/envs/lidar_localization2d.py:122-132) and adds `num_envs` in front.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Literal

import numpy as np

from . import _capi
from .floor_maps import FloorMapBank, procedural_spec
from .loss_fn import WeightedLossFn, make_mse_loss_fn_and_target_space
from .spaces import (NEXT_STEP, ActivePerceptionActionSpace, Box, Dict, ImageSpace, VectorEnvBase, batch_box,
                     sparse_target_spaces)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _void(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lidar_directions(beam_count: int, lidar_range: float) -> np.ndarray:
    """The reference's beam table (lidar_localization2d.py:180-187), float32 (B, 2)."""
    ang = np.linspace(-np.pi, np.pi, beam_count, dtype=np.float32, endpoint=False)
    unscaled = np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    return np.ascontiguousarray(unscaled * lidar_range, dtype=np.float32)


class LIDARLocalization2DVectorEnv(VectorEnvBase):
    """A `gymnasium.vector.VectorEnv` when gymnasium is importable.  Returned CUDA tensors are DLPack views of
    engine-owned buffers: valid until the next step / reset / rollout, freed by close() (`copy_outputs=True` returns
    private copies)."""

    metadata: dict[str, Any] = {"render_modes": ["rgb_array"], "render_fps": 4, "autoreset_mode": NEXT_STEP}
    closed = False

    def __init__(self, num_envs: int, dataset, render_mode: Literal["rgb_array"] = "rgb_array",
                 static_map: bool = False, lidar_beam_count: int = 8, lidar_range: float = 5,
                 static_map_index: int = 0, prefetch: bool = True, prefetch_buffer_size: int = 128, *,
                 max_episode_steps: int = 100, device: int = 0, env_offset: int = 0, emit_cells: bool = False,
                 check_errors: Literal["step", "lazy", "never"] = "lazy", output: Literal["torch", "numpy"] = "torch",
                 sparse: bool = False, copy_outputs: bool = False):
        """`sparse=True` = the reference's `<id>-sparse` registration (SparsifyWrapper around every env,
        sparsify_wrapper.py:95-150): the loss only enters the reward of the step that terminates an env's episode and
        info["prediction"]["target"] becomes {"target": ..., "weight": float32(terminated)}."""
        if render_mode not in self.metadata["render_modes"]:
            raise ValueError(f"Invalid render mode: {render_mode}")
        self._render_mode = render_mode
        self.num_envs = int(num_envs)
        self._device = int(device)
        self._static = bool(static_map)
        # procedural datasets (len == 2**32: ours or the reference's FloorMapDatasetRooms / FloorMapDatasetMaze): every
        # (auto)reset draws its dataset index and generates the map on the device; a static map is generated once
        self._procedural = procedural_spec(dataset)
        self._dataset = dataset
        if self._procedural is not None:
            kind, kw = self._procedural
            gen = FloorMapBank.generate_rooms if kind == "rooms" else FloorMapBank.generate_maze
            if self._static:
                self._bank = gen([int(static_map_index)], device=int(device), **kw)
                static_map_index = 0
                self._procedural = None
            else:
                self._bank = None
                self._W, self._H = int(kw["width"]), int(kw["height"])
        else:
            self._bank = FloorMapBank.from_dataset(dataset)
        self._B = int(lidar_beam_count)
        self._range = lidar_range
        self._T = int(max_episode_steps)
        if self._bank is not None:
            self._W, self._H = int(self._bank.map_width), int(self._bank.map_height)
        self._check_errors = check_errors
        self._numpy_mode = output == "numpy"
        self._steps_since_poll = 0
        self._emit = bool(emit_cells)
        self._lib = _capi.lib()
        cfg = _capi.LidarCfg(device=self._device, num_envs=self.num_envs, width=self._W, height=self._H,
                             beam_count=self._B, lidar_range=float(lidar_range), step_limit=self._T,
                             static_map=int(self._static), static_map_index=int(static_map_index),
                             emit_cells=int(self._emit), env_offset=int(env_offset),
                             sparse_reward=int(bool(sparse)))
        self.sparse = bool(sparse)
        h = C.c_void_p()
        _capi.check(self._lib.apg_lidar_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._closed = False
        if self._procedural is not None:
            kind, kw = self._procedural
            _capi.check(self._lib.apg_lidar_set_map_generator(
                self._h, 1 if kind == "rooms" else 2, int(kw.get("max_rooms", 0)), int(kw.get("door_width", 0)),
                float(kw.get("branching_prob", 1.0))))
        else:
            _capi.check(self._lib.apg_lidar_set_map_bank(self._h, _void(self._bank.bits), len(self._bank)))
        self._dirs = lidar_directions(self._B, lidar_range)
        _capi.check(self._lib.apg_lidar_set_beams(self._h, _void(self._dirs)))
        self._views = self._make_views()
        self._has_reset = False
        self._copy = bool(copy_outputs)
        self._cuda_stream = 0
        self.spec = None

        obs = {"lidar": Box(low=0, high=1, shape=(self._B,), dtype=np.float32),
               "odometry": Box(low=-1, high=1, shape=(2,), dtype=np.float32)}
        if not self._static:
            obs["map"] = ImageSpace(width=self._W, height=self._H, channels=1)
        obs["time_step"] = Box(low=-1.0, high=1.0, shape=(), dtype=np.float32)
        self.single_observation_space = Dict(obs)
        self.observation_space = Dict({k: batch_box(s, self.num_envs, ImageSpace if isinstance(s, ImageSpace) else None)
                                       for k, s in obs.items()})
        inner = Box(low=-1, high=1, shape=(2,), dtype=np.float32)
        pred = Box(low=-1, high=1, shape=(2,))
        self.single_action_space = ActivePerceptionActionSpace(inner, pred)
        self.action_space = ActivePerceptionActionSpace(batch_box(inner, self.num_envs), batch_box(pred, self.num_envs))
        self.loss_fn, self.single_prediction_target_space = make_mse_loss_fn_and_target_space(2, -1, 1)
        self.prediction_target_space = batch_box(self.single_prediction_target_space, self.num_envs)
        if self.sparse:  # SparsifyWrapper / SparsifyVectorWrapper (sparsify_wrapper.py:49-60): weighted loss + Dict target
            self.single_prediction_target_space, self.prediction_target_space = sparse_target_spaces(
                self.single_prediction_target_space, self.num_envs)
            self.loss_fn = WeightedLossFn(self.loss_fn)

    prediction_space = property(lambda self: self.action_space["prediction"])
    inner_action_space = property(lambda self: self.action_space["action"])
    single_prediction_space = property(lambda self: self.single_action_space["prediction"])
    single_inner_action_space = property(lambda self: self.single_action_space["action"])
    render_mode = property(lambda self: self._render_mode)
    dataset = property(lambda self: self._bank if self._bank is not None else self._dataset)
    unwrapped = property(lambda self: self)

    def _make_views(self):
        b = _capi.LidarBuffers()
        _capi.check(self._lib.apg_lidar_get_buffers(self._h, C.byref(b)))
        N, B, dev = self.num_envs, self._B, self._device
        t = _capi.dlpack_tensor
        v = {
            "lidar": t(b.d_lidar, dev, "float", 32, (N, B)),
            "odometry": t(b.d_odometry, dev, "float", 32, (N, 2)),
            "time_step": t(b.d_time_step, dev, "float", 32, (N,)),
            "reward": t(b.d_reward, dev, "float", 32, (N,)),
            "base_reward": t(b.d_base_reward, dev, "float", 32, (N,)),
            "loss": t(b.d_loss, dev, "float", 32, (N,)),
            "target": t(b.d_target, dev, "float", 32, (N, 2)),
            "terminated": t(b.d_terminated, dev, "bool", 8, (N,)),
            "truncated": t(b.d_truncated, dev, "bool", 8, (N,)),
            "stepped": t(b.d_stepped, dev, "bool", 8, (N,)),
            "map_idx": t(b.d_map_idx, dev, "int", 64, (N,)),
            "pos": t(b.d_pos, dev, "float", 32, (N, 2)),
            "initial_pos": t(b.d_initial_pos, dev, "float", 32, (N, 2)),
        }
        for k in ("ep_avg1", "ep_final1", "ep_avg2", "ep_final2", "ep_return"):
            v[k] = t(getattr(b, "d_" + k), dev, "float", 32, (N,))
        if b.d_map:
            v["map"] = t(b.d_map, dev, "float", 32, (N, self._H, self._W, 1))
        if b.d_cells:
            v["cells"] = t(b.d_cells, dev, "int", 32, (N, B, 2))
            v["dists"] = t(b.d_dists, dev, "float", 32, (N, B))
        return v

    def _out(self, name, as_numpy):
        v = self._views[name]
        if as_numpy:
            return v.cpu().numpy()
        return v.clone() if self._copy else v

    def _use_current_stream(self, tensor):
        """Queue on torch's current stream: inputs produced on a side stream are complete, outputs ordered for readers."""
        import torch

        st = int(torch.cuda.current_stream(tensor.device).cuda_stream)
        if st != self._cuda_stream:
            _capi.check(self._lib.apg_lidar_set_stream(self._h, C.c_void_p(st)))
            self._cuda_stream = st

    def _obs(self, as_numpy):
        keys = ["lidar", "odometry"] + ([] if self._static else ["map"]) + ["time_step"]
        return {k: self._out(k, as_numpy) for k in keys}

    def _poll(self):
        flags = C.c_uint32(0)
        _capi.check(self._lib.apg_lidar_poll_error(self._h, C.byref(flags), 1))
        self._steps_since_poll = 0
        if flags.value & _capi.FLAG_NAN_ACTION:
            raise ValueError("NaN values detected in action.")
        if flags.value & _capi.FLAG_NAN_PRED:
            raise ValueError("NaN values detected in prediction.")
        if flags.value & _capi.FLAG_MAPGEN:
            raise ValueError("a procedural map exceeded the device generator's limits (door_width >= room width?)")

    def _maybe_poll(self, host_mode):
        self._steps_since_poll += 1
        if self._check_errors == "never":
            return
        if self._check_errors == "step" or host_mode or self._steps_since_poll >= 256:
            self._poll()

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        if seed is None and self._has_reset:
            # every env resets, its own Generators continue (what SyncVectorEnv.reset() without a seed does)
            _capi.check(self._lib.apg_lidar_reset_all(self._h))
        else:
            if seed is None:  # like gymnasium's lazily seeded np_random: entropy from the OS
                seed = int(np.random.SeedSequence().entropy % (2**63))
            _capi.check(self._lib.apg_lidar_reset(self._h, C.c_uint64(int(seed))))
        self._has_reset = True
        as_numpy = self._numpy_mode or bool(options and options.get("numpy", False))
        return self._obs(as_numpy), {"map_idx": self._out("map_idx", as_numpy)}

    def step(self, action):
        a, p = action["action"], action["prediction"]
        host_mode = not _is_torch(a)
        N = self.num_envs
        if host_mode:
            a = np.ascontiguousarray(a, dtype=np.float32)
            p = np.ascontiguousarray(p, dtype=np.float32)
            if a.shape != (N, 2) or p.shape != (N, 2):
                raise ValueError(f"bad action/prediction shapes {a.shape} {p.shape}")
            reward = np.empty(N, np.float32)
            term = np.empty(N, np.bool_)
            trunc = np.empty(N, np.bool_)
            lidar = np.empty((N, self._B), np.float32)
            _capi.check(self._lib.apg_lidar_step_host(self._h, _void(a), _void(p), _void(reward), _void(term),
                                                      _void(trunc), _void(lidar)))
            self._maybe_poll(True)
            obs = self._obs(True)
            obs["lidar"] = lidar
            terminated, truncated = term, trunc
        else:
            if not (a.is_cuda and p.is_cuda and a.is_contiguous() and p.is_contiguous()
                    and str(a.dtype) == "torch.float32" and str(p.dtype) == "torch.float32"):
                raise ValueError("torch inputs must be contiguous float32 CUDA tensors")
            if tuple(a.shape) != (N, 2) or tuple(p.shape) != (N, 2):
                raise ValueError(f"bad action/prediction shapes {tuple(a.shape)} {tuple(p.shape)}")
            self._use_current_stream(a)
            _capi.check(self._lib.apg_lidar_step(self._h, a.data_ptr(), p.data_ptr()))
            self._maybe_poll(False)
            obs = self._obs(False)
            reward = self._views["reward"]
            terminated, truncated = self._views["terminated"], self._views["truncated"]
        stepped = self._out("stepped", host_mode)
        info = {
            "base_reward": self._out("base_reward", host_mode), "_base_reward": stepped,
            "prediction": {"target": self._out("target", host_mode), "_target": stepped,
                           "loss": self._out("loss", host_mode), "_loss": stepped},
            "_prediction": stepped,
            "map_idx": self._out("map_idx", host_mode), "_map_idx": ~stepped,
            # ActiveRegressionLogWrapper's info["stats"]: valid for the envs whose `terminated` flag is set
            "stats": {"scalar": {"avg_euclidean_distance": self._out("ep_avg1", host_mode),
                                 "final_euclidean_distance": self._out("ep_final1", host_mode),
                                 "avg_mse": self._out("ep_avg2", host_mode), "final_mse": self._out("ep_final2", host_mode),
                                 "return": self._out("ep_return", host_mode)},
                      "_scalar": terminated},
        }
        if self.sparse:
            w = terminated.astype(np.float32) if host_mode else terminated.to(dtype=self._views["reward"].dtype)
            info["prediction"]["target"] = {"target": info["prediction"]["target"], "weight": w}
        return obs, reward, terminated, truncated, info

    def episode_stats(self, reduce: bool = True, clear: bool = False) -> dict[str, float]:
        """Means over all episodes finished so far (see ImageClassificationVectorEnv.episode_stats)."""
        import torch

        from .distributed import reduce_stats

        h = (C.c_double * _capi.NUM_STATS)()
        _capi.check(self._lib.apg_lidar_get_stats(self._h, h, int(clear)))
        t = torch.tensor(list(h), dtype=torch.float64, device=torch.device("cuda", self._device))
        if reduce:
            reduce_stats(t)
        t = t.cpu().tolist()
        ep = max(t[0], 1.0)
        return {"episodes": t[0], "avg_euclidean_distance": t[1] / ep, "final_euclidean_distance": t[2] / ep,
                "avg_mse": t[3] / ep, "final_mse": t[4] / ep, "return": t[5] / ep}

    def rollout(self, actions, predictions, rewards=None, dones=None):
        n_steps, N = int(actions.shape[0]), self.num_envs

        def _chk(t, shape, dtypes, what):
            if t is None:
                return
            if not (_is_torch(t) and t.is_cuda and t.is_contiguous() and str(t.dtype) in dtypes and tuple(t.shape) == shape):
                raise ValueError(f"rollout: {what} must be a contiguous CUDA tensor of shape {shape} and dtype "
                                 f"{' or '.join(dtypes)}")

        _chk(actions, (n_steps, N, 2), ("torch.float32",), "actions")
        _chk(predictions, (n_steps, N, 2), ("torch.float32",), "predictions")
        _chk(rewards, (n_steps, N), ("torch.float32",), "rewards")
        _chk(dones, (n_steps, N), ("torch.bool", "torch.uint8"), "dones")
        self._use_current_stream(actions)
        _capi.check(self._lib.apg_lidar_rollout(
            self._h, actions.data_ptr(), predictions.data_ptr(), int(actions.shape[0]),
            None if rewards is None else rewards.data_ptr(), None if dones is None else dones.data_ptr()))

    def state(self):
        out = {"pos": self._views["pos"], "initial_pos": self._views["initial_pos"]}
        for k in ("cells", "dists"):
            if k in self._views:
                out[k] = self._views[k]
        return out

    def sync(self):
        _capi.check(self._lib.apg_lidar_sync(self._h))

    def render(self):
        raise NotImplementedError("render() is host-side tooling outside the hot path (DESIGN.md, out of scope)")

    def close(self, **kwargs):
        if not getattr(self, "_closed", True):
            self._closed = True
            self.closed = True
            self._views = {}
            self._lib.apg_lidar_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
