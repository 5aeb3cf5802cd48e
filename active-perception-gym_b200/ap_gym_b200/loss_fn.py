"""Loss-function objects exposed as `env.loss_fn` (API mirror of ap_gym/loss_fn.py:25-349).

The step kernels compute the same normalised losses on device; these objects exist so that callers can
evaluate the loss of a policy's prediction themselves (`.numpy`, `.torch`), read `lower_bound` /
`blind_guessing_expected_value`, or build `.normalized` variants exactly like with the reference.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np


class LossFn(ABC):
    @abstractmethod
    def numpy(self, prediction, target, batch_shape=()):
        ...

    def torch(self, prediction, target, batch_shape=()):
        raise NotImplementedError("Loss function is not implemented for torch.")

    def jax(self, prediction, target, batch_shape=()):
        raise NotImplementedError("Loss function is not implemented for jax.")

    def __call__(self, prediction, target, batch_shape=()):
        return self.numpy(prediction, target, batch_shape)

    # bounds -------------------------------------------------------------------------------------------
    def _lower_bound(self):
        return -np.inf

    def _blind_guessing_expected_value(self):
        return None

    @property
    def lower_bound(self):
        return self._lower_bound()

    @property
    def blind_guessing_expected_value(self):
        return self._blind_guessing_expected_value()

    @property
    def normalized(self):
        """(loss - lower_bound) / (blind_guess - lower_bound)   (ap_gym/loss_fn.py:69-83)."""
        hi, lo = self.blind_guessing_expected_value, self.lower_bound
        if hi is None:
            raise ValueError("Cannot normalize loss function without blind guessing expected value.")
        if hi <= lo:
            raise ValueError(
                "Cannot normalize loss function when blind guessing expected value is not greater than lower bound.")
        scale = 1 / (hi - lo)
        return LossFnAffineTransformation(self, scale, -lo * scale)


class LossFnAffineTransformation(LossFn):
    def __init__(self, inner_loss_fn: LossFn, scale, offset):
        self._inner, self._scale, self._offset = inner_loss_fn, scale, offset

    scale = property(lambda self: self._scale)
    offset = property(lambda self: self._offset)
    inner = property(lambda self: self._inner)

    def numpy(self, prediction, target, batch_shape=()):
        return self._inner.numpy(prediction, target, batch_shape=batch_shape) * self._scale + self._offset

    def torch(self, prediction, target, batch_shape=()):
        return self._inner.torch(prediction, target, batch_shape=batch_shape) * self._scale + self._offset

    def jax(self, prediction, target, batch_shape=()):
        return self._inner.jax(prediction, target, batch_shape=batch_shape) * self._scale + self._offset

    def _lower_bound(self):
        return self._inner.lower_bound * self._scale + self._offset

    def _blind_guessing_expected_value(self):
        v = self._inner.blind_guessing_expected_value
        return None if v is None else v * self._scale + self._offset


class ZeroLossFn(LossFn):
    def numpy(self, prediction, target, batch_shape=()):
        return np.zeros(batch_shape, dtype=np.float32)

    def torch(self, prediction, target, batch_shape=()):
        import torch

        return torch.zeros(batch_shape)

    def _lower_bound(self):
        return 0.0

    def _blind_guessing_expected_value(self):
        return 0.0


class LambdaLossFn(LossFn):
    def __init__(self, np, torch=None, jax=None, lower_bound=-float("inf"), blind_guessing_expected_value=None):
        self._np, self._torch, self._jax = np, torch, jax
        self._lo, self._blind = lower_bound, blind_guessing_expected_value

    def numpy(self, prediction, target, batch_shape=()):
        return self._np(prediction, target, batch_shape)

    def torch(self, prediction, target, batch_shape=()):
        if self._torch is None:
            raise NotImplementedError("Loss function is not implemented for torch.")
        return self._torch(prediction, target, batch_shape)

    def jax(self, prediction, target, batch_shape=()):
        if self._jax is None:
            raise NotImplementedError("Loss function is not implemented for jax.")
        return self._jax(prediction, target, batch_shape)

    def _lower_bound(self):
        return self._lo

    def _blind_guessing_expected_value(self):
        return self._blind


def _np_log_softmax(x):
    m = np.amax(x, axis=-1, keepdims=True)
    m = np.where(np.isfinite(m), m, 0).astype(x.dtype)
    z = x - m
    return z - np.log(np.sum(np.exp(z), axis=-1, keepdims=True))


class CrossEntropyLossFn(LossFn):
    """-log_softmax(logits)[target]; blind guess = ln K  (ap_gym/loss_fn.py:204-247)."""

    def __init__(self, num_classes=None):
        self._k = num_classes

    def numpy(self, prediction, target, batch_shape=()):
        prediction = np.asarray(prediction)
        target = np.asarray(target)
        return -np.take_along_axis(_np_log_softmax(prediction), target[..., None].astype(np.int64), axis=-1)[..., 0]

    def torch(self, prediction, target, batch_shape=()):
        import torch

        lsm = torch.nn.functional.log_softmax(prediction, dim=-1)
        return -torch.take_along_dim(lsm, target[..., None].long(), dim=-1)[..., 0]

    def _lower_bound(self):
        return 0.0

    def _blind_guessing_expected_value(self):
        return None if self._k is None else np.log(self._k)


class MSELossFn(LossFn):
    """mean((prediction - target)^2); blind guess = mean(target_std^2)  (ap_gym/loss_fn.py:250-289)."""

    def __init__(self, target_std=None):
        self._blind = None if target_std is None else float(np.mean(np.asarray(target_std) ** 2))

    def numpy(self, prediction, target, batch_shape=()):
        return np.mean((np.asarray(prediction) - np.asarray(target)) ** 2, axis=-1)

    def torch(self, prediction, target, batch_shape=()):
        return ((prediction - target) ** 2).mean(dim=-1)

    def _lower_bound(self):
        return 0.0

    def _blind_guessing_expected_value(self):
        return self._blind


class WeightedLossFn(LossFn):
    """inner(prediction, target["target"]) * target["weight"]  (ap_gym/loss_fn.py:292-349)."""

    def __init__(self, inner_loss_fn, min_weight=0.0, average_weight=None):
        self._inner, self._min_w, self._avg_w = inner_loss_fn, min_weight, average_weight

    def numpy(self, prediction, target, batch_shape=()):
        return self._inner.numpy(prediction, target["target"], batch_shape) * target["weight"]

    def torch(self, prediction, target, batch_shape=()):
        return self._inner.torch(prediction, target["target"], batch_shape) * target["weight"]

    def _lower_bound(self):
        return self._min_w * self._inner.lower_bound

    def _blind_guessing_expected_value(self):
        v = self._inner.blind_guessing_expected_value
        return None if v is None or self._avg_w is None else self._avg_w * v


def make_mse_loss_fn_and_target_space(target_dim, prediction_low=None, prediction_high=None, target_std=None):
    """ap_gym/active_regression_env.py:29-52: uniform-target std (high-low)/sqrt(12) -> normalised MSE."""
    from .spaces import Box

    space = Box(low=prediction_low, high=prediction_high, shape=(target_dim,))
    if target_std is None and np.isfinite(prediction_low) and np.isfinite(prediction_high):
        target_std = (prediction_high - prediction_low) / np.sqrt(12)
    fn = MSELossFn(target_std=target_std)
    return (fn.normalized if target_std is not None else fn), space
