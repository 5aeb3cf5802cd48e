"""Drop-in image vector environments backed by the fused sm_100a step kernel.

Mirrors the reference's public surface for this path:
  ImagePerceptionConfig            ap_gym/envs/image/image_perception_module.py:20-34
  ImageClassificationVectorEnv     ap_gym/envs/image_classification.py:22-168
  ImageLocalizationVectorEnv       ap_gym/envs/image_localization.py:24-254
  ActivePerceptionVectorEnv.step   ap_gym/active_perception_vector_env.py:84-111

`step({"action": (N,2) f32, "prediction": (N,K)|(N,2) f32})` accepts CUDA torch tensors (zero-copy, fully
asynchronous, returns DLPack views of engine-owned device buffers that stay valid until the next step) or
numpy arrays (copied through the C-ABI host call; numpy comes back).  Everything between the two calls is one
kernel launch; there is no CPU implementation behind it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Literal, Sequence

import numpy as np

from . import _capi
from .datasets import ImageBank, bank_from_dataset
from .loss_fn import CrossEntropyLossFn, WeightedLossFn, make_mse_loss_fn_and_target_space
from .reset_stream import ImageResetStream
from .spaces import (NEXT_STEP, ActivePerceptionActionSpace, Box, Dict, Discrete, ImageSpace, LogitSpace,
                     MultiDiscrete, VectorEnvBase, batch_box, sparse_target_spaces)


@dataclass(frozen=True)
class ImagePerceptionConfig:
    """Same fields and defaults as the reference dataclass (image_perception_module.py:20-34)."""

    dataset: Any
    sensor_size: tuple[int, int] = (5, 5)
    sensor_scale: float = 1.0
    max_step_length: float | Sequence[float] = 0.2
    step_limit: int = 16
    display_visitation: bool = True
    render_unvisited_opacity: float = 0.0
    render_visited_opacity: float = 0.3
    prefetch_buffer_size: int = 128
    prefetch: bool = True
    unique_sampling_max_grid_cell_size_rel = 0.2
    unique_sampling_top_k: int = 10
    randomly_invert_labels: bool = False


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _void(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _ImageVectorEnvBase(VectorEnvBase):
    """Common part of the two image environments.  A `gymnasium.vector.VectorEnv` when gymnasium is importable (the
    reference's wrappers and `gymnasium.vector.VectorWrapper` check for it); `ap_gym_b200.enable_reference_compat()`
    additionally registers the classes with the reference's `BaseActivePerceptionVectorEnv` ABC.

    Aliasing contract: `step()` / `reset()` with CUDA tensors return DLPack VIEWS of engine-owned device buffers; the
    next `step()` / `reset()` / `rollout()` overwrites them and `close()` frees them (the reference returns fresh
    arrays).  Construct with `copy_outputs=True` to get private copies instead."""

    metadata: dict[str, Any] = {"render_modes": ["rgb_array"], "render_fps": 2, "autoreset_mode": NEXT_STEP}
    _TASK = _capi.TASK_CLASSIFICATION
    closed = False

    def __init__(self, num_envs: int, image_perception_config: ImagePerceptionConfig,
                 render_mode: Literal["rgb_array"] = "rgb_array", *, device: int = 0,
                 env_offset: int = 0, global_num_envs: int | None = None, track_visited: bool = True,
                 check_errors: Literal["step", "lazy", "never"] = "lazy",
                 output: Literal["torch", "numpy"] = "torch", payload_ring: int = 4,
                 blend: Literal["float64", "float32"] = "float64", sparse: bool = False, copy_outputs: bool = False):
        """`sparse=True` is the reference's `<id>-sparse` registration (SparsifyVectorWrapper, sparsify_wrapper.py:23-88):
        reward = base_reward - loss * weight with weight = terminated, info["prediction"]["target"] becomes
        {"target": ..., "weight": float32(terminated)}, `loss_fn` becomes WeightedLossFn(loss_fn) and the prediction
        target spaces Dict{"target", "weight"}."""
        if render_mode not in self.metadata["render_modes"]:
            raise ValueError(f"Unsupported render mode: {render_mode}")
        if blend not in ("float64", "float32"):
            raise ValueError("blend must be 'float64' (the reference's arithmetic, bit-exact) or 'float32'")
        self.blend = blend
        self.sparse = bool(sparse)
        cfg = image_perception_config
        self._config = cfg
        self._render_mode = render_mode
        self.num_envs = int(num_envs)
        self._device = int(device)
        self._global_n = int(global_num_envs) if global_num_envs is not None else self.num_envs
        self._slice = slice(int(env_offset), int(env_offset) + self.num_envs)
        if self._slice.stop > self._global_n:
            raise ValueError("env_offset + num_envs exceeds global_num_envs")
        self._check_errors = check_errors
        self._numpy_mode = output == "numpy"
        self._steps_since_poll = 0
        self._lib = _capi.lib()

        bank: ImageBank = bank_from_dataset(cfg.dataset)
        self._bank = bank
        M, H, W, Cn = bank.images.shape
        self._H, self._W, self._C = H, W, Cn
        self._K = int(bank.num_classes)
        S0, S1 = (int(s) for s in cfg.sensor_size)
        self._S0, self._S1 = S0, S1
        msl = np.ones(2) * np.array(cfg.max_step_length)
        if msl.shape != (2,):
            raise ValueError("max_step_length must be a scalar or a pair")
        eff = np.array(cfg.sensor_size) * cfg.sensor_scale
        if np.any(np.array([H, W]) < eff):
            raise ValueError(
                f"Image size {(H, W)} cannot be smaller than effective sensor size {tuple(eff)}.")

        ccfg = _capi.ImageCfg(
            device=self._device, num_envs=self.num_envs, task=self._TASK, height=H, width=W, channels=Cn,
            sensor_h=S0, sensor_w=S1, sensor_scale=float(cfg.sensor_scale),
            max_step_length=(C.c_double * 2)(float(msl[0]), float(msl[1])), step_limit=int(cfg.step_limit),
            num_classes=self._K, bank_dtype=_capi.BANK_U8 if bank.is_u8 else _capi.BANK_F32,
            randomly_invert_labels=int(cfg.randomly_invert_labels), track_visited=int(track_visited),
            payload_ring=int(payload_ring), blend_f32=int(blend == "float32"), sparse_reward=int(bool(sparse)))
        self._ring = max(1, min(int(payload_ring) if payload_ring > 0 else _capi.PAYLOAD_RING_DEFAULT,
                                _capi.PAYLOAD_RING_MAX))
        h = C.c_void_p()
        _capi.check(self._lib.apg_image_create(C.byref(ccfg), C.byref(h)))
        self._h = h
        _capi.check(self._lib.apg_image_set_bank(self._h, _void(bank.images), _void(bank.labels), M, 0))
        self._track = bool(track_visited)
        self._stream: ImageResetStream | None = None
        self._views = self._make_views()
        self._build_spaces()
        self._closed = False
        self._copy = bool(copy_outputs)
        self._spec = None
        self._cuda_stream = 0       # the CUDA stream the engine queues on (0 = legacy default stream)
        self._rng_exposed = False   # env.np_random was handed out: its state goes back to the engine before stepping

    # ------------------------------------------------------------------------------------------ plumbing
    def _buffers(self) -> _capi.ImageBuffers:
        b = _capi.ImageBuffers()
        _capi.check(self._lib.apg_image_get_buffers(self._h, C.byref(b)))
        return b

    def _make_views(self):
        b = self._buffers()
        N, S0, S1, Cn, T = self.num_envs, self._S0, self._S1, self._C, int(self._config.step_limit)
        dev = self._device
        t = _capi.dlpack_tensor
        v = {
            "glimpse": t(b.d_glimpse, dev, "float", 32, (N, S0, S1, Cn)),
            "glimpse_pos": t(b.d_glimpse_pos, dev, "float", 32, (N, 2)),
            "time_step": t(b.d_time_step, dev, "float", 32, (N,)),
            "reward": t(b.d_reward, dev, "float", 32, (N,)),
            "base_reward": t(b.d_base_reward, dev, "float", 32, (N,)),
            "loss": t(b.d_loss, dev, "float", 32, (N,)),
            "terminated": t(b.d_terminated, dev, "bool", 8, (N,)),
            "truncated": t(b.d_truncated, dev, "bool", 8, (N,)),
            "index": t(b.d_index, dev, "int", 64, (N,)),
            "pos": t(b.d_pos, dev, "float", 64, (N, 2)),
        }
        for k in ("ep_avg1", "ep_final1", "ep_avg2", "ep_final2", "ep_return"):
            v[k] = t(getattr(b, "d_" + k), dev, "float", 32, (N,))
        v["first_correct"] = t(b.d_first_correct, dev, "int", 32, (N,))
        v["last_incorrect"] = t(b.d_last_incorrect, dev, "int", 32, (N,))
        if b.d_target_glimpse:
            v["target_glimpse"] = t(b.d_target_glimpse, dev, "float", 32, (N, S0, S1, Cn))
            v["target_xy"] = t(b.d_target_xy, dev, "float", 32, (N, 2))
            v["cur_target_xy"] = t(b.d_cur_target_xy, dev, "float", 32, (N, 2))
        if b.d_target_label:
            v["target_label"] = t(b.d_target_label, dev, "int", 32, (N,))
        if b.d_inverted_label:
            v["inverted_label"] = t(b.d_inverted_label, dev, "int", 32, (N,))
        if b.d_visited:
            v["visited"] = t(b.d_visited, dev, "int", 32, (N, self._H, (self._W + 31) // 32))
            v["stamp_xy"] = t(b.d_stamp_xy, dev, "int", 32, (T, 2, N))
            v["stamp_q"] = t(b.d_stamp_q, dev, "float", 32, (T, N))
        return v

    def _build_spaces(self):
        S0, S1, Cn, N = self._S0, self._S1, self._C, self.num_envs
        obs = {
            "glimpse": ImageSpace(S1, S0, Cn, dtype=np.float32),
            "glimpse_pos": Box(-1, 1, (2,), np.float32),
            "time_step": Box(-1, 1, (), np.float32),
        }
        if self._config.randomly_invert_labels:
            obs["inverted_label"] = Discrete(3)
        if self._TASK == _capi.TASK_LOCALIZATION:
            obs["target_glimpse"] = ImageSpace(S1, S0, Cn, dtype=np.float32)
        self.single_observation_space = Dict(obs)
        batched = {}
        for k, s in obs.items():
            if isinstance(s, ImageSpace):
                batched[k] = batch_box(s, N, ImageSpace)
            elif isinstance(s, Discrete):
                batched[k] = MultiDiscrete(np.full((N,), s.n))
            else:
                batched[k] = batch_box(s, N)
        self.observation_space = Dict(batched)
        inner = Box(-np.ones(2, dtype=np.float32), np.ones(2, dtype=np.float32))
        if self._TASK == _capi.TASK_CLASSIFICATION:
            pred = LogitSpace(-np.inf, np.inf, shape=(self._K,))
            self.single_prediction_target_space = Discrete(self._K)
            self.prediction_target_space = MultiDiscrete([self._K] * N)
            self.loss_fn = CrossEntropyLossFn(num_classes=self._K).normalized
            batched_pred = batch_box(pred, N, LogitSpace)
        else:
            pred = Box(low=-1, high=1, shape=(2,))
            self.loss_fn, self.single_prediction_target_space = make_mse_loss_fn_and_target_space(2, -1, 1)
            self.prediction_target_space = batch_box(self.single_prediction_target_space, N)
            batched_pred = batch_box(pred, N)
        self.single_action_space = ActivePerceptionActionSpace(inner, pred)
        self.action_space = ActivePerceptionActionSpace(batch_box(inner, N), batched_pred)
        if self.sparse:  # SparsifyVectorWrapper.__init__ (sparsify_wrapper.py:49-60)
            self.single_prediction_target_space, self.prediction_target_space = sparse_target_spaces(
                self.single_prediction_target_space, N)
            self.loss_fn = WeightedLossFn(self.loss_fn)

    # the reference's derived properties (active_perception_vector_env.py:52-66)
    prediction_space = property(lambda self: self.action_space["prediction"])
    inner_action_space = property(lambda self: self.action_space["action"])
    single_prediction_space = property(lambda self: self.single_action_space["prediction"])
    single_inner_action_space = property(lambda self: self.single_action_space["action"])
    render_mode = property(lambda self: self._render_mode)
    config = property(lambda self: self._config)
    unwrapped = property(lambda self: self)

    @property
    def spec(self):
        return self._spec

    @spec.setter
    def spec(self, spec):
        """Like the reference (image_classification.py:155-159): the episode length is the config's step limit."""
        if spec is not None:
            import copy

            spec = copy.copy(spec)
            try:
                spec.max_episode_steps = int(self._config.step_limit)
            except AttributeError:
                pass
        self._spec = spec

    @property
    def np_random(self):
        """The env-level Generator (g0 of SURVEY 3.4) in its CURRENT state.  Draws made from it are handed back to the
        engine before the next step / rollout, so the reference's stream position is kept either way."""
        if self._stream is None:
            self._seed(None)
        self._pull_streams()
        self._rng_exposed = True
        return self._stream.g0

    def _before_device_work(self, tensor=None):
        """Queue on torch's current stream (so that inputs produced on a side stream are complete and outputs are
        ordered for their consumers) and hand back Generator states the caller may have advanced."""
        if self._rng_exposed and getattr(self, "_engine_has_streams", False):
            self._push_streams()
            self._rng_exposed = False
        if tensor is not None:
            import torch

            st = int(torch.cuda.current_stream(tensor.device).cuda_stream)
            if st != self._cuda_stream:
                _capi.check(self._lib.apg_image_set_stream(self._h, C.c_void_p(st)))
                self._cuda_stream = st

    def _seed(self, seed):
        self._stream = ImageResetStream(seed, self._global_n, len(self._bank), self._slice)
        self._engine_has_streams = False

    def _push_streams(self):
        """Hand the three numpy Generators to the engine: from here on it draws every autoreset payload itself
        (bit-identical to numpy, csrc/apg_rng.h) so no python runs on the autoreset path."""
        st = [_capi.Pcg64State.from_numpy(g) for g in (self._stream.g0, self._stream.g1, self._stream.g2)]
        _capi.check(self._lib.apg_image_set_streams(self._h, C.byref(st[0]), C.byref(st[1]), C.byref(st[2]),
                                                    self._global_n, self._slice.start))
        self._engine_has_streams = True

    def _pull_streams(self):
        """Bring the numpy Generators up to date with the draws the engine made since _push_streams."""
        if getattr(self, "_engine_has_streams", False):
            st = [_capi.Pcg64State() for _ in range(3)]
            _capi.check(self._lib.apg_image_get_streams(self._h, C.byref(st[0]), C.byref(st[1]), C.byref(st[2])))
            for s_, g in zip(st, (self._stream.g0, self._stream.g1, self._stream.g2)):
                s_.to_numpy(g)

    # ------------------------------------------------------------------------------------------ errors
    def _poll(self):
        flags = C.c_uint32(0)
        _capi.check(self._lib.apg_image_poll_error(self._h, C.byref(flags), 1))
        self._steps_since_poll = 0
        f = flags.value
        if f & _capi.FLAG_NAN_PRED:
            raise ValueError("NaN values detected in prediction.")
        if f & _capi.FLAG_NAN_ACTION:
            raise ValueError("NaN values detected in action.")
        if f & _capi.FLAG_OUT_OF_GRID:
            raise ValueError("One of the requested xi is out of bounds")

    def _maybe_poll(self, host_mode: bool):
        self._steps_since_poll += 1
        if self._check_errors == "never":
            return
        if self._check_errors == "step" or host_mode or self._steps_since_poll >= 256:
            self._poll()

    # ------------------------------------------------------------------------------------------ reset
    def _explicit_targets(self, options, idx, pos) -> np.ndarray | None:
        return None

    def _autoreset_targets(self) -> np.ndarray | None:
        return None

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        if seed is not None or self._stream is None:
            self._seed(seed)
        else:
            self._pull_streams()
        self._rng_exposed = False
        idx, inv, pos = self._stream.draw_episode(bool(self._config.randomly_invert_labels))
        target = self._explicit_targets(options, idx, pos)
        _capi.check(self._lib.apg_image_reset(self._h, _void(idx), _void(pos), _void(target), _void(inv)))
        self._push_streams()
        self._last_prediction = None
        as_numpy = bool(options and options.get("numpy", False)) or self._numpy_mode
        return self._obs(as_numpy), {"index": self._out("index", as_numpy)}

    def queue_autoreset_payloads(self, count: int = 1):
        """Draw and upload the payloads of the next `count` autoresets ahead of time (bounded by the payload ring);
        without it the engine draws each payload when its autoreset step is queued."""
        _capi.check(self._lib.apg_image_prefetch_payloads(self._h, int(count)))

    # ------------------------------------------------------------------------------------------ step
    def _out(self, name: str, as_numpy: bool):
        v = self._views[name]
        if as_numpy:
            return v.cpu().numpy()
        return v.clone() if self._copy else v

    def _obs(self, as_numpy: bool):
        keys = ["glimpse", "glimpse_pos", "time_step"]
        if self._config.randomly_invert_labels:
            keys.append("inverted_label")
        if self._TASK == _capi.TASK_LOCALIZATION:
            keys.append("target_glimpse")
        return {k: self._out(k, as_numpy) for k in keys}

    def _pred_width(self) -> int:
        return self._K if self._TASK == _capi.TASK_CLASSIFICATION else 2

    def step(self, action):
        a, p = action["action"], action["prediction"]
        host_mode = not _is_torch(a)
        want_a, want_p = (self.num_envs, 2), (self.num_envs, self._pred_width())
        if host_mode:
            a = np.ascontiguousarray(a, dtype=np.float32)
            p = np.ascontiguousarray(p, dtype=np.float32)
            if a.shape != want_a or p.shape != want_p:
                raise ValueError(f"bad action/prediction shapes {a.shape} {p.shape}, expected {want_a} {want_p}")
            self._before_device_work()
            N = self.num_envs
            reward = np.empty(N, np.float32)
            term = np.empty(N, np.bool_)
            trunc = np.empty(N, np.bool_)
            glimpse = np.empty((N, self._S0, self._S1, self._C), np.float32)
            _capi.check(self._lib.apg_image_step_host(self._h, _void(a), _void(p), _void(reward), _void(term),
                                                      _void(trunc), _void(glimpse)))
            self._maybe_poll(True)
            obs = self._obs(True)
            obs["glimpse"] = glimpse
            terminated, truncated = term, trunc
        else:
            if not (a.is_cuda and p.is_cuda and a.is_contiguous() and p.is_contiguous()
                    and a.dtype == p.dtype and str(a.dtype) == "torch.float32"):
                raise ValueError("torch inputs must be contiguous float32 CUDA tensors")
            if tuple(a.shape) != want_a or tuple(p.shape) != want_p:
                raise ValueError(f"bad action/prediction shapes {tuple(a.shape)} {tuple(p.shape)}, expected {want_a} {want_p}")
            self._before_device_work(a)
            _capi.check(self._lib.apg_image_step(self._h, a.data_ptr(), p.data_ptr()))
            self._maybe_poll(False)
            obs = self._obs(False)
            reward = self._views["reward"]
            terminated, truncated = self._views["terminated"], self._views["truncated"]
        info = {
            "index": self._out("index", host_mode),
            "base_reward": self._out("base_reward", host_mode),
            "prediction": {"target": self._out(self._TARGET_VIEW, host_mode), "loss": self._out("loss", host_mode)},
        }
        if self.sparse:
            w = terminated.astype(np.float32) if host_mode else terminated.to(dtype=self._views["reward"].dtype)
            info["prediction"]["target"] = {"target": info["prediction"]["target"], "weight": w}
        if self._buffers().needs_reset:  # this step ended every episode: the LogWrapper's info["stats"]
            info["stats"] = self._stats_info(host_mode, terminated)
        return obs, reward, terminated, truncated, info

    def _stats_info(self, as_numpy, done):
        m1, m2 = self._METRICS
        scalar = {f"avg_{m1}": self._out("ep_avg1", as_numpy), f"final_{m1}": self._out("ep_final1", as_numpy),
                  f"avg_{m2}": self._out("ep_avg2", as_numpy), f"final_{m2}": self._out("ep_final2", as_numpy),
                  "return": self._out("ep_return", as_numpy)}
        if self._TASK == _capi.TASK_CLASSIFICATION:
            fc, li = self._out("first_correct", as_numpy), self._out("last_incorrect", as_numpy)
            scalar.update({"first_correct": fc, "_first_correct": fc >= 0, "last_incorrect": li, "_last_incorrect": li >= 0})
        return {"scalar": scalar, "_scalar": done}

    def episode_stats(self, reduce: bool = True, clear: bool = False) -> dict[str, float]:
        """Means over all episodes finished so far (this shard, or all ranks with `reduce=True` inside an initialised
        torch.distributed group: ONE all-reduce of 8 float64 -- the only collective of the data path)."""
        import torch

        from .distributed import reduce_stats

        h = (C.c_double * _capi.NUM_STATS)()
        _capi.check(self._lib.apg_image_get_stats(self._h, h, int(clear)))
        t = torch.tensor(list(h), dtype=torch.float64, device=torch.device("cuda", self._device))
        if reduce:
            reduce_stats(t)
        t = t.cpu().tolist()
        ep = max(t[0], 1.0)
        m1, m2 = self._METRICS
        return {"episodes": t[0], f"avg_{m1}": t[1] / ep, f"final_{m1}": t[2] / ep, f"avg_{m2}": t[3] / ep,
                f"final_{m2}": t[4] / ep, "return": t[5] / ep}

    def rollout(self, actions, predictions, rewards=None, dones=None):
        """`n_steps` consecutive steps from stacked CUDA tensors (n_steps, N, 2) / (n_steps, N, K|2); per-step
        rewards (f32) / dones (bool) land in the optional (n_steps, N) tensors.  Asynchronous."""
        n_steps = int(actions.shape[0])
        N, K = self.num_envs, self._pred_width()

        def _chk(t, shape, dtypes, what):
            if t is None:
                return
            if not (_is_torch(t) and t.is_cuda and t.is_contiguous() and str(t.dtype) in dtypes and tuple(t.shape) == shape):
                raise ValueError(f"rollout: {what} must be a contiguous CUDA tensor of shape {shape} and dtype "
                                 f"{' or '.join(dtypes)}")

        _chk(actions, (n_steps, N, 2), ("torch.float32",), "actions")
        _chk(predictions, (n_steps, N, K), ("torch.float32",), "predictions")
        _chk(rewards, (n_steps, N), ("torch.float32",), "rewards")
        _chk(dones, (n_steps, N), ("torch.bool", "torch.uint8"), "dones")
        self._before_device_work(actions)
        _capi.check(self._lib.apg_image_rollout(
            self._h, actions.data_ptr(), predictions.data_ptr(), n_steps,
            None if rewards is None else rewards.data_ptr(), None if dones is None else dones.data_ptr()))

    # ------------------------------------------------------------------------------------------ state access
    def sync(self):
        _capi.check(self._lib.apg_image_sync(self._h))

    def state(self) -> dict[str, Any]:
        """Device views of the engine state (float64 positions, visited bitmask, overlay stamp log)."""
        b = self._buffers()
        out = {"pos": self._views["pos"], "elapsed": b.elapsed, "needs_reset": bool(b.needs_reset),
               "stamp_count": b.stamp_count}
        for k in ("visited", "stamp_xy", "stamp_q", "cur_target_xy"):
            if k in self._views:
                out[k] = self._views[k]
        return out

    def overlay_maps(self, envs: Sequence[int] | None = None):
        """Rebuilds the reference's render-resolution overlay (`visitation_counts` int32, `prediction_quality_map`
        float32; image_perception_module.py:219-234) of the current episode from the on-device stamp log."""
        st = self.state()
        envs = np.arange(self.num_envs) if envs is None else np.asarray(envs)
        H, W = self._H, self._W
        Wr = max(128, self._S1)
        rs = Wr / W
        Hr = int(round(rs * H))
        size = np.round(np.flip(np.array(self._config.sensor_size) * self._config.sensor_scale * rs)).astype(np.int32)
        counts = np.zeros((len(envs), Hr, Wr), np.int32)
        quality = np.zeros((len(envs), Hr, Wr), np.float32)
        n_st = st["stamp_count"]
        if n_st == 0:
            return counts, quality
        xy = st["stamp_xy"][:n_st].cpu().numpy()[:, :, envs]
        q = st["stamp_q"][:n_st].cpu().numpy()[:, envs]
        rows = np.arange(len(envs))[:, None, None]
        for s in range(n_st):
            xr = xy[s, 0][:, None] + np.arange(size[0]) - size[0] // 2
            yr = xy[s, 1][:, None] + np.arange(size[1]) - size[1] // 2
            coords = (rows, np.clip(yr, 0, Hr - 1)[:, :, None], np.clip(xr, 0, Wr - 1)[:, None, :])
            counts[coords] += 1
            quality[coords] = q[s][:, None, None]
        return counts, quality

    def _render_geometry(self):
        H, W = self._H, self._W
        Wr = max(128, self._S1)
        rs = Wr / W
        Hr = int(round(rs * H))
        return Wr, Hr, rs

    def _to_render_coords(self, pos_norm):
        """image_perception_module.py:406-413: normalised (x, y) -> render pixels."""
        Wr, Hr, rs = self._render_geometry()
        eff = np.array(self._config.sensor_size) * self._config.sensor_scale
        lim = (np.array([self._W, self._H]) - 1) / 2 - (eff - 1) / 2
        return np.asarray(pos_norm, dtype=np.float64) * lim * rs + np.array([Wr, Hr]) / 2

    def render(self, envs: Sequence[int] | None = None, return_pil_imgs: bool = False):
        """ImagePerceptionModule.render (image_perception_module.py:333-401) rebuilt from the engine state: the bank
        image of every environment, the visitation / prediction-quality overlay of the current episode (device stamp
        log) and the sensor rectangle.  Host-side (PIL), not on the step path; `envs` limits the work to a subset."""
        from PIL import Image, ImageDraw

        from .style import COLOR_AGENT, quality_color

        cfg = self._config
        envs = np.arange(self.num_envs) if envs is None else np.asarray(envs)
        Wr, Hr, rs = self._render_geometry()
        self.sync()
        idx = self._views["index"].cpu().numpy()[envs]
        imgs = self._bank.images[idx]
        imgs = imgs.astype(np.float32) / np.float32(255) if self._bank.is_u8 else imgs
        if self._C == 1:
            imgs = imgs[..., 0]
        counts, quality = self.overlay_maps(envs)
        visited = counts > 0
        overlay = (visited[..., None] * np.concatenate(
            [quality_color(quality), np.full_like(quality[..., None], int(255 * cfg.render_visited_opacity))], axis=-1)
            + ~visited[..., None] * (0, 0, 0, int(255 * cfg.render_unvisited_opacity))).round().astype(np.uint8)
        pos = self._to_render_coords(self._views["pos"].cpu().numpy()[envs])
        size = np.array(cfg.sensor_size) * cfg.sensor_scale * rs
        border = max(1, int(round(1 / 128 * Wr)))
        out = []
        for img, tl, br, ol in zip(imgs, pos - size / 2, pos + size / 2, overlay):
            rgb = Image.fromarray((img * 255).astype(np.uint8)).resize((Wr, Hr), resample=Image.Resampling.NEAREST).convert("RGB")
            if cfg.display_visitation:
                alpha = ol[..., -1:] / 255
                rgb = Image.fromarray((np.array(rgb) * (1 - alpha) + alpha * ol[..., :-1]).astype(np.uint8))
            draw = ImageDraw.Draw(rgb, "RGBA")
            rect = np.concatenate([tl, br])
            draw.rectangle(tuple(rect + border), outline=(0, 0, 0, 80), width=border)
            draw.rectangle(tuple(rect), outline=COLOR_AGENT, width=border)
            out.append(rgb)
        return out if return_pil_imgs else np.asarray(out)

    def close(self, **kwargs):
        if not getattr(self, "_closed", True):
            self._closed = True
            self.closed = True
            self._views = {}
            self._lib.apg_image_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class ImageClassificationVectorEnv(_ImageVectorEnvBase):
    """ap_gym.envs.image_classification.ImageClassificationVectorEnv on the B200 engine."""

    _TASK = _capi.TASK_CLASSIFICATION
    _TARGET_VIEW = "target_label"
    _METRICS = ("correct_label_prob", "accuracy")

    def _explicit_targets(self, options, idx, pos):
        return None


class ImageLocalizationVectorEnv(_ImageVectorEnvBase):
    """ap_gym.envs.image_localization.ImageLocalizationVectorEnv on the B200 engine.

    Explicit `reset()` targets follow the reference (`sample_unique_glimpse_positions`,
    image_perception_module.py:253-292): the uniqueness search -- O(729^2 * 300) per 64x64x3 image, 1.2 s on a CPU
    core -- runs on the GPU (`apg_image_unique_topk`) and is cached per bank image; the random pick and jitter come
    from the module Generator exactly like in the reference.  `options={"prediction_target": (N,2) float32}` overrides
    the targets; `target_sampling="uniform"` skips the search and draws U(-1,1) (the Generator is still advanced
    exactly as the reference would, so later episodes stay in lock-step)."""

    _TASK = _capi.TASK_LOCALIZATION
    _TARGET_VIEW = "target_xy"
    _METRICS = ("euclidean_distance", "mse")

    def __init__(self, *args, target_sampling: Literal["unique", "uniform"] = "unique", **kwargs):
        super().__init__(*args, **kwargs)
        if target_sampling not in ("unique", "uniform"):
            raise ValueError("target_sampling must be 'unique' or 'uniform'")
        self._target_sampling = target_sampling
        cfg = self._config
        eff = np.array(cfg.sensor_size) * cfg.sensor_scale
        lim = (np.array([self._W, self._H]) - 1) / 2 - (eff - 1) / 2
        self._max_cell = eff / lim * cfg.unique_sampling_max_grid_cell_size_rel
        cnt = np.ceil(2 / self._max_cell)
        self._grid = np.ascontiguousarray(np.stack(
            np.meshgrid(np.linspace(-1, 1, int(cnt[0])), np.linspace(-1, 1, int(cnt[1])), indexing="ij"),
            axis=-1).reshape(-1, 2))
        self._topk_cache: dict[int, np.ndarray] = {}

    def unique_topk(self, bank_indices: np.ndarray) -> np.ndarray:
        """(n, top_k) grid indices of the most unique glimpse positions of the given bank images (cached)."""
        k = int(self._config.unique_sampling_top_k)
        bank_indices = np.asarray(bank_indices, dtype=np.int64)
        missing = np.array(sorted({int(i) for i in bank_indices} - self._topk_cache.keys()), dtype=np.int64)
        if len(missing):
            out = np.empty((len(missing), k), dtype=np.int32)
            _capi.check(self._lib.apg_image_unique_topk(self._h, _void(missing), len(missing), _void(self._grid),
                                                        len(self._grid), k, _void(out)))
            for i, row in zip(missing, out):
                self._topk_cache[int(i)] = row
        return np.stack([self._topk_cache[int(i)] for i in bank_indices])

    def _explicit_targets(self, options, idx, pos):
        cfg = self._config
        sel, jitter = self._stream.draw_unique_selection(cfg.unique_sampling_top_k, self._max_cell)
        if options and options.get("prediction_target") is not None:
            t = np.ascontiguousarray(options["prediction_target"], dtype=np.float32)
            if t.shape != (self.num_envs, 2):
                raise ValueError("options['prediction_target'] must have shape (num_envs, 2)")
            return t
        if self._target_sampling == "uniform":
            return np.ascontiguousarray(np.clip(jitter / self._max_cell, -1, 1).astype(np.float32))
        top = self.unique_topk(idx)
        base = self._grid[top[np.arange(self.num_envs), sel]]
        return np.ascontiguousarray(np.clip(base + jitter, -1, 1).astype(np.float32))

    def _autoreset_targets(self):
        return self._stream.draw_autoreset_targets()

    def step(self, action):
        out = super().step(action)
        self._last_prediction = action["prediction"]
        return out

    def render(self, envs: Sequence[int] | None = None):
        """ImageLocalizationVectorEnv.render (image_localization.py:183-220): the module's rendering plus the target
        rectangle and the last prediction."""
        from PIL import ImageDraw

        from .style import COLOR_PRED

        envs = np.arange(self.num_envs) if envs is None else np.asarray(envs)
        imgs = super().render(envs, return_pil_imgs=True)
        Wr, Hr, rs = self._render_geometry()
        size = np.array(self._config.sensor_size) * self._config.sensor_scale * rs
        border = max(1, int(round(1 / 128 * Wr)))
        targets = self._views["cur_target_xy"].cpu().numpy()[envs]
        last = getattr(self, "_last_prediction", None)
        if last is not None:
            last = (last.detach().cpu().numpy() if _is_torch(last) else np.asarray(last))[envs]
        for i, (img, target) in enumerate(zip(imgs, targets)):
            draw = ImageDraw.Draw(img, "RGBA")
            t = self._to_render_coords(target)
            draw.rectangle((tuple(t - size / 2), tuple(t + size / 2)), outline=COLOR_PRED + (100,), width=border)
            if last is not None:
                lp = self._to_render_coords(last[i])
                rect = np.concatenate([lp - size / 2, lp + size / 2])
                draw.rectangle(tuple(rect), outline=COLOR_PRED, width=border)
                draw.rectangle(tuple(rect + border), outline=(0, 0, 0, 80), width=border)
        return np.asarray(imgs)
