"""CircleSquareHideAndSeek reward term on top of the engine's classification env.

Reference: `CircleSquareHideAndSeekVectorWrapper` (ap_gym/envs/circle_square_catch_or_flee.py:18-118), registered as
`CircleSquareHideAndSeek-v0` / `CircleSquareHideAndSeekNoPrediction-v0` (ap_gym/envs/registration.py:470-510):

    additional_reward = (2 * label - 1) * || glimpse_pos - (object_position / sensor_pos_lim - 1) ||        (float64)
    info["base_reward"] += additional_reward;  reward += additional_reward
    mask_prediction=True: the prediction is ignored (zeros are fed), reward = base_reward, target = ()

The object position and class of every bank image are properties of the dataset, so they are gathered once at
construction into a device table indexed by `info["index"]`; per step the term is four element-wise tensor operations
on the engine's outputs (device tensors stay on the device, numpy stays numpy).
"""
from __future__ import annotations

import numpy as np

from .loss_fn import ZeroLossFn


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class CircleSquareHideAndSeekVectorWrapper:
    def __init__(self, env, mask_prediction: bool = False):
        ds = env.config.dataset
        if not hasattr(ds, "get_object_position_and_label"):
            raise TypeError("the environment's dataset must provide get_object_position_and_label (CircleSquareDataset)")
        self.env = env
        self._mask = bool(mask_prediction)
        cfg = env.config
        eff = np.array(cfg.sensor_size) * cfg.sensor_scale
        lim = (np.array([env._W, env._H]) - 1) / 2 - (eff - 1) / 2      # sensor_pos_lim_pixels (x, y)
        positions, labels = ds.get_object_position_and_label(np.arange(len(ds)))
        self._pos_norm = np.flip(np.asarray(positions), axis=-1) / lim - 1   # normalize_coords(flip(positions)) - 1
        self._sign = np.asarray(labels) * 2 - 1
        self._dev = None
        if self._mask:
            self.loss_fn = ZeroLossFn()

    def __getattr__(self, item):
        return getattr(self.env, item)

    def _tables(self, like):
        if not _is_torch(like):
            return self._pos_norm, self._sign
        if self._dev is None or self._dev[0].device != like.device:
            import torch

            self._dev = (torch.as_tensor(self._pos_norm, device=like.device), torch.as_tensor(self._sign, device=like.device))
        return self._dev

    def additional_reward(self, glimpse_pos, index):
        """(2 label - 1) * ||glimpse_pos - object_position_norm|| in float64."""
        pos_norm, sign = self._tables(glimpse_pos)
        if _is_torch(glimpse_pos):
            idx = index.long()
            d = (glimpse_pos.double() - pos_norm[idx]).pow(2).sum(-1).sqrt()
            return sign[idx].double() * d
        d = np.linalg.norm(np.asarray(glimpse_pos) - pos_norm[index], axis=-1)
        return sign[index] * d

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        info = dict(info)
        if self._mask:
            info["prediction"] = {"target": ()}
        return obs, info

    def step(self, actions):
        if self._mask:
            p = actions["action"].new_zeros(self.env.prediction_space.shape) if _is_torch(actions["action"]) else \
                np.zeros(self.env.prediction_space.shape, np.float32)
            actions = {"action": actions["action"], "prediction": p}
        obs, reward, terminated, truncated, info = self.env.step(actions)
        info = dict(info)
        info["prediction"] = dict(info["prediction"])
        extra = self.additional_reward(obs["glimpse_pos"], info["index"])
        if _is_torch(reward):
            info["base_reward"] = info["base_reward"] + extra.to(info["base_reward"].dtype)
            reward = info["base_reward"].clone() if self._mask else reward + extra.to(reward.dtype)
        else:
            info["base_reward"] = (info["base_reward"] + extra).astype(info["base_reward"].dtype)
            reward = info["base_reward"].copy() if self._mask else (reward + extra).astype(reward.dtype)
        if self._mask:
            info["prediction"]["target"] = ()
        return obs, reward, terminated, truncated, info

    def close(self, **kwargs):
        return self.env.close(**kwargs)
