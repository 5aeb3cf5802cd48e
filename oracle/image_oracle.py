"""CPU ORACLE (test infrastructure, NOT product code) for the image hot path of ap_gym 0.5.0.

A dependency-free numpy restatement of what the reference computes per vector step for
`ImageClassificationVectorEnv` / `ImageLocalizationVectorEnv`.  Only `tests/`, `__graft_entry__.smoke()`
and the CPU-baseline legs of `bench.py` may import this module; the product path never does.

Parity status: PINNED.  `tests/test_oracle_image.py` checks this file (a) against the golden vectors in
`tests/golden/image_*.npz`, which `tests/golden/make_golden.py` produced by running the UNMODIFIED
reference (`/root/reference`, through the test-only gymnasium shim in `oracle/shim`) with real numpy/scipy,
and (b), when `/root/reference` is present, against the live reference.

Reference lines restated (all paths relative to /root/reference/ap_gym):
  RNG chain ............ envs/image_classification.py:144-149, envs/image/image_perception_module.py:105-118,
                         envs/dataset/dataset_iterator.py:35-59
  module reset ......... envs/image/image_perception_module.py:120-189
  module step .......... envs/image/image_perception_module.py:191-217
  project_sphere ....... util.py:94-97
  overlay stamp ........ envs/image/image_perception_module.py:219-234, 441-445, 409-413
  get_glimpse .......... envs/image/image_perception_module.py:294-331 (+ scipy 1.18.1
                         interpolate/_rgi.py:520-549 `_evaluate_linear`, `find_indices`)
  unique targets ....... envs/image/image_perception_module.py:253-292
  classification env ... envs/image_classification.py:109-127, active_classification_env.py:47-65
  localization env ..... envs/image_localization.py:131-181, active_regression_env.py:29-52,79-110
  loss / reward ........ loss_fn.py:69-110,212-220,261-267, active_perception_vector_env.py:84-111
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------------------------------------
# loss functions (loss_fn.py)
# ----------------------------------------------------------------------------------------------------------
def log_softmax_f32(x: np.ndarray) -> np.ndarray:
    """scipy.special.log_softmax(x, axis=-1) for float32 input (stays float32)."""
    x = np.asarray(x)
    x_max = np.amax(x, axis=-1, keepdims=True)
    x_max = np.where(np.isfinite(x_max), x_max, 0).astype(x.dtype)
    tmp = x - x_max
    s = np.sum(np.exp(tmp), axis=-1, keepdims=True)
    with np.errstate(divide="ignore"):
        out = np.log(s)
    return tmp - out


def softmax_f32(x: np.ndarray) -> np.ndarray:
    """scipy.special.softmax(x, axis=-1)."""
    x = np.asarray(x)
    x_max = np.amax(x, axis=-1, keepdims=True)
    e = np.exp(x - x_max)
    return e / np.sum(e, axis=-1, keepdims=True)


def ce_normalized(logits: np.ndarray, target: np.ndarray, num_classes: int) -> np.ndarray:
    """CrossEntropyLossFn(K).normalized (loss_fn.py:69-83,100-110,212-220): f32 CE times the np.float64
    scale 1/ln K plus offset -0.0 -> float64."""
    ce = -np.take_along_axis(log_softmax_f32(logits), target[..., None].astype(np.int64), axis=-1)[..., 0]
    upper = np.log(num_classes)
    scale = 1 / (upper - 0.0)
    offset = -0.0 * scale
    return ce * scale + offset


def mse_normalized(pred: np.ndarray, target: np.ndarray, low: float = -1.0, high: float = 1.0) -> np.ndarray:
    """MSELossFn(std).normalized (active_regression_env.py:29-52, loss_fn.py:261-267): python-float scale
    (3.0 for [-1, 1]) keeps float32."""
    target_std = (high - low) / np.sqrt(12)
    blind = float(np.mean(target_std**2))
    scale = 1 / (blind - 0.0)
    offset = -0.0 * scale
    return np.mean((pred - target) ** 2, axis=-1) * scale + offset


def project_sphere(x: np.ndarray, radius: float = 1.0) -> np.ndarray:
    """util.py:94-97, float32 in -> float32 out."""
    magnitude = np.sqrt(np.sum(x * x, axis=-1, keepdims=True))
    direction = x / np.maximum(magnitude, radius)
    return np.where(magnitude > radius, direction * radius, x)


# ----------------------------------------------------------------------------------------------------------
# glimpse resampling (image_perception_module.py:294-331 + scipy RGI linear)
# ----------------------------------------------------------------------------------------------------------
def tap_indices(coord: np.ndarray, length: int):
    """scipy `find_indices` on the grid g[k] = k - (length-1)/2: the interval index i0 with
    g[i0] <= coord < g[i0+1], clamped to [0, length-2], and the normalised distance coord - g[i0]."""
    off = (length - 1) / 2
    k = np.floor(coord + off).astype(np.int64)
    # exact comparisons against the grid (the float sum above may round across a grid value)
    k = np.where((k - off) > coord, k - 1, k)
    k = np.where((k + 1 - off) <= coord, k + 1, k)
    k = np.clip(k, 0, length - 2)
    w = (coord - (k - off)) / 1.0
    return k.astype(np.int32), w


def glimpse(images: np.ndarray, pos_norm: np.ndarray, sensor_size, sensor_scale: float):
    """images (N,H,W,C) f32, pos_norm (N,2) [x,y] any float dtype -> glimpse (N,S0,S1,C) f32 and the tap
    indices (i0 (N,S0), j0 (N,S1)) that must match the engine bit-exactly."""
    N, H, W, C = images.shape
    S0, S1 = sensor_size
    eff = np.array(sensor_size) * sensor_scale
    lim = (np.array([W, H]) - 1) / 2 - (eff - 1) / 2  # (lim_x, lim_y); x uses S0, y uses S1 like the reference
    c = pos_norm * lim  # (N,2) float64
    off0 = (np.arange(S0) - (S0 - 1) / 2) * sensor_scale
    off1 = (np.arange(S1) - (S1 - 1) / 2) * sensor_scale
    ys = c[:, 1, None] + off0[None]  # (N,S0)
    xs = c[:, 0, None] + off1[None]  # (N,S1)
    if np.any(ys < -(H - 1) / 2) or np.any(ys > (H - 1) / 2) or np.any(xs < -(W - 1) / 2) or np.any(xs > (W - 1) / 2):
        raise ValueError("One of the requested xi is out of bounds")
    i0, wy = tap_indices(ys, H)
    j0, wx = tap_indices(xs, W)
    n = np.arange(N)[:, None, None]
    I0, J0 = i0[:, :, None], j0[:, None, :]
    wy_, wx_ = wy[:, :, None, None], wx[:, None, :, None]
    v = np.zeros((N, S0, S1, C))
    v = v + images[n, I0, J0] * ((1 - wy_) * (1 - wx_))
    v = v + images[n, I0, J0 + 1] * ((1 - wy_) * wx_)
    v = v + images[n, I0 + 1, J0] * (wy_ * (1 - wx_))
    v = v + images[n, I0 + 1, J0 + 1] * (wy_ * wx_)
    return v.clip(0, 1).astype(np.float32), i0, j0


def render_geometry(H: int, W: int, sensor_size, sensor_scale: float):
    """Render-resolution geometry of the visitation overlay (image_perception_module.py:167-171)."""
    render_width = max(128, sensor_size[1])
    rs = render_width / W
    render_height = int(round(rs * H))
    return render_width, render_height, rs


def overlay_stamp(pos_norm: np.ndarray, H: int, W: int, sensor_size, sensor_scale: float):
    """Integer rectangle centre (N,2) [x,y] and size (2,) of the stamp (image_perception_module.py:219-224)."""
    Wr, Hr, rs = render_geometry(H, W, sensor_size, sensor_scale)
    eff = np.array(sensor_size) * sensor_scale
    lim = (np.array([W, H]) - 1) / 2 - (eff - 1) / 2
    pos = pos_norm * lim * rs + np.array([Wr, Hr]) / 2
    size = eff * rs
    centre = np.round(pos).astype(np.int32)
    size_i = np.round(np.flip(size)).astype(np.int32)
    return centre, size_i


def apply_stamp(counts: np.ndarray, quality: np.ndarray, centre: np.ndarray, size_i: np.ndarray, pq: np.ndarray):
    """counts[coords] += 1 (duplicates count once) and quality[coords] = clip(pq, 0, 1)."""
    N, Hr, Wr = counts.shape
    x_range = centre[:, 0, None] + np.arange(size_i[0]) - size_i[0] // 2
    y_range = centre[:, 1, None] + np.arange(size_i[1]) - size_i[1] // 2
    coords = (
        np.arange(N)[:, None, None],
        np.clip(y_range, 0, Hr - 1)[:, :, None],
        np.clip(x_range, 0, Wr - 1)[:, None, :],
    )
    counts[coords] += 1
    quality[coords] = np.clip(pq[:, None, None], 0, 1)


def footprint_mask(i0: np.ndarray, j0: np.ndarray, H: int, W: int) -> np.ndarray:
    """Engine-defined image-resolution visited bitmask of ONE glimpse: bit (r, c) is set for every tap row
    r in {i0[k], i0[k]+1} and tap column c in {j0[l], j0[l]+1}.  Returns uint32 (N, H, ceil(W/32))."""
    N = i0.shape[0]
    words = (W + 31) // 32
    rows = np.zeros((N, H), dtype=bool)
    cols = np.zeros((N, W), dtype=bool)
    n = np.arange(N)[:, None]
    rows[n, i0] = True
    rows[n, i0 + 1] = True
    cols[n, j0] = True
    cols[n, j0 + 1] = True
    colbits = np.zeros((N, words), dtype=np.uint32)
    for c in range(W):
        colbits[:, c // 32] |= cols[:, c].astype(np.uint32) << np.uint32(c % 32)
    return np.where(rows[:, :, None], colbits[:, None, :], np.uint32(0)).astype(np.uint32)


# ----------------------------------------------------------------------------------------------------------
# reset-path RNG chain
# ----------------------------------------------------------------------------------------------------------
class ImageResetStream:
    """The three numpy Generators of the image reset path and the order in which they are consumed."""

    def __init__(self, seed: int, num_envs: int, bank_len: int):
        self.g0 = np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(seed))))  # env-level
        s1 = self.g0.integers(0, 2**32 - 1, endpoint=True)
        self.g1 = np.random.default_rng(s1)  # module-level
        s2 = self.g1.integers(0, 2**32 - 1, endpoint=True)
        self.g2 = np.random.default_rng(s2)  # dataset batch iterator
        self.num_envs = num_envs
        self.bank_len = bank_len

    def draw_episode(self, randomly_invert_labels: bool = False):
        idx = self.g2.integers(0, self.bank_len, self.num_envs)
        inv = None
        if randomly_invert_labels:
            inv = self.g1.integers(0, 2, size=self.num_envs) == 1
        pos = self.g1.uniform(-1, 1, size=(self.num_envs, 2))
        return idx, inv, pos

    def draw_autoreset_targets(self):
        return self.g0.uniform(-1, 1, (self.num_envs, 2)).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------
# the environments
# ----------------------------------------------------------------------------------------------------------
class ImageEnvOracle:
    """Restates ImageClassificationVectorEnv (task='cls') / ImageLocalizationVectorEnv (task='loc').

    `bank_images` is the float32 (M,H,W,C) image bank the reference dataset would return
    (uint8 banks are converted with the reference's `astype(float32) / 255`), `bank_labels` int (M,)."""

    def __init__(self, bank_images, bank_labels, num_envs, task="cls", num_classes=None, sensor_size=(5, 5),
                 sensor_scale=1.0, max_step_length=0.2, step_limit=16, randomly_invert_labels=False,
                 unique_sampling_top_k=10, track_overlay=True):
        bank_images = np.asarray(bank_images)
        if bank_images.dtype == np.uint8:
            bank_images = bank_images.astype(np.float32) / 255
        if bank_images.ndim == 3:
            bank_images = bank_images[..., None]
        assert bank_images.dtype == np.float32
        self.bank = bank_images
        self.labels = np.asarray(bank_labels).astype(np.int32)
        self.N = int(num_envs)
        self.task = task
        self.K = int(num_classes) if num_classes is not None else int(self.labels.max()) + 1
        self.sensor_size = tuple(sensor_size)
        self.sensor_scale = sensor_scale
        self.msl = np.ones(2) * np.array(max_step_length)
        self.T = int(step_limit)
        self.invert = bool(randomly_invert_labels)
        self.top_k = unique_sampling_top_k
        self.track_overlay = track_overlay
        _, self.H, self.W, self.C = self.bank.shape
        self.stream = None
        self.counts = self.quality = None

    # -- helpers --------------------------------------------------------------------------------------
    def _module_reset(self):
        idx, inv, pos = self.stream.draw_episode(self.invert)
        self.index = idx
        self.images = self.bank[idx]
        labels = self.labels[idx]
        self.labels_inverted = inv
        if self.invert:
            labels = np.where(inv, self.K - labels - 1, labels)
        self.cur_labels = labels
        eff = np.array(self.sensor_size) * self.sensor_scale
        if np.any(np.array([self.H, self.W]) < eff):
            raise ValueError("Image size cannot be smaller than effective sensor size")
        self.pos = pos
        self.t = 0
        obs = self._obs()
        if self.track_overlay:
            Wr, Hr, _ = render_geometry(self.H, self.W, self.sensor_size, self.sensor_scale)
            self.counts = np.zeros((self.N, Hr, Wr), dtype=np.int32)
            self.quality = np.zeros((self.N, Hr, Wr), dtype=np.float32)
        self.visited = self.last_footprint.copy()
        self.prev_done = np.zeros(self.N, dtype=bool)
        return obs, {"index": self.index}

    def _obs(self):
        g, i0, j0 = glimpse(self.images, self.pos, self.sensor_size, self.sensor_scale)
        self.tap_i0, self.tap_j0 = i0, j0
        self.last_footprint = footprint_mask(i0, j0, self.H, self.W)
        out = {
            "glimpse": g,
            "glimpse_pos": self.pos.astype(np.float32),
            "time_step": np.full(self.N, (self.t / self.T) * 2 - 1, np.float32),
        }
        if self.invert:
            out["inverted_label"] = (
                np.full(self.N, 2) if self.t > 0 else self.labels_inverted.astype(np.int32)
            )
        return out

    def _module_step(self, action, pq):
        if np.any(np.isnan(pq)):
            raise ValueError("NaN values detected in prediction.")
        centre, size_i = overlay_stamp(self.pos, self.H, self.W, self.sensor_size, self.sensor_scale)
        self.last_stamp = (centre, np.clip(pq, 0, 1).astype(np.float32))
        if self.track_overlay:
            apply_stamp(self.counts, self.quality, centre, size_i, pq)
        if np.any(self.prev_done):
            if not np.all(self.prev_done):
                raise NotImplementedError("Partial reset is not supported.")
            obs, info = self._module_reset()
            terminated = False
            base = np.zeros(self.N)
        else:
            if np.any(np.isnan(action)):
                raise ValueError("NaN values detected in action.")
            step = self.msl * project_sphere(action)
            self.pos = np.clip(self.pos + step, -1, 1)
            base = -np.sqrt(np.sum(action * action, axis=-1)) * 1e-3
            info = {"index": self.index}
            self.t += 1
            terminated = self.t >= self.T
            obs = self._obs()
            self.visited = self.visited | self.last_footprint
        term = np.full(self.N, terminated)
        trunc = np.zeros(self.N, dtype=bool)
        self.prev_done = term | trunc
        return obs, base, term, trunc, info

    def _unique_targets(self):
        """sample_unique_glimpse_positions (image_perception_module.py:253-292)."""
        eff = np.array(self.sensor_size) * self.sensor_scale
        lim = (np.array([self.W, self.H]) - 1) / 2 - (eff - 1) / 2
        sensor_size_norm = eff / lim
        max_cell = sensor_size_norm * 0.2
        cnt = np.ceil(2 / max_cell)
        grid = np.stack(
            np.meshgrid(np.linspace(-1, 1, int(cnt[0])), np.linspace(-1, 1, int(cnt[1])), indexing="ij"), axis=-1
        ).reshape(-1, 2)
        P = len(grid)
        top = np.empty((self.N, self.top_k), dtype=np.int64)
        for n in range(self.N):  # per env to bound memory
            g, _, _ = glimpse(np.broadcast_to(self.images[n], (P, self.H, self.W, self.C)), grid,
                              self.sensor_size, self.sensor_scale)
            diff = np.mean((g[None] - g[:, None]) ** 2, axis=(-3, -2, -1))
            uniq = np.min(diff + np.diag(np.full(P, np.inf)), axis=-1)
            top[n] = np.argsort(-uniq)[: self.top_k]
        sel = self.stream.g1.integers(0, self.top_k, size=self.N)
        base = grid[top[np.arange(self.N), sel]]
        return np.clip(base + self.stream.g1.uniform(-max_cell, max_cell, (self.N, 2)), -1, 1)

    # -- public surface ---------------------------------------------------------------------------------
    def reset(self, seed=None, explicit_targets=None):
        if seed is not None:
            self.stream = ImageResetStream(seed, self.N, len(self.bank))
        assert self.stream is not None, "oracle: always seed the first reset"
        obs, info = self._module_reset()
        if self.task == "loc":
            if explicit_targets is None:
                self.target = self._unique_targets().astype(np.float32)
            else:
                # caller supplies the targets (skips the O(P^2) search) but the module Generator must still
                # advance exactly as sample_unique_glimpse_positions would (image_perception_module.py:278-289)
                self.stream.g1.integers(0, self.top_k, size=self.N)
                self.stream.g1.uniform(-1.0, 1.0, (self.N, 2))
                self.target = np.asarray(explicit_targets, dtype=np.float32).copy()
            self.env_prev_done = np.zeros(self.N, dtype=bool)
            obs["target_glimpse"], _, _ = glimpse(self.images, self.target, self.sensor_size, self.sensor_scale)
        return obs, info

    def step(self, action, prediction):
        action = np.asarray(action)
        prediction = np.asarray(prediction)
        if self.task == "cls":
            pq = softmax_f32(prediction)[np.arange(self.N), self.cur_labels]
            obs, base, term, trunc, info = self._module_step(action, pq)
            target = self.cur_labels  # AFTER the module step: new labels on the reset step
            loss = ce_normalized(prediction, target, self.K)
        else:
            target = self.target.copy()  # the OLD target is what is scored / returned
            if np.any(self.env_prev_done):
                self.target[self.env_prev_done] = self.stream.draw_autoreset_targets()[self.env_prev_done]
            pq = 1 - np.sqrt(np.sum((prediction - self.target) ** 2, axis=-1)) / np.sqrt(4)
            obs, base, term, trunc, info = self._module_step(action, pq)
            self.env_prev_done = term | trunc
            obs["target_glimpse"], _, _ = glimpse(self.images, self.target, self.sensor_size, self.sensor_scale)
            loss = mse_normalized(prediction, target)
        info = dict(info)
        info.update({"base_reward": base, "prediction": {"target": target, "loss": loss}})
        return obs, base - loss, term, trunc, info
