"""CPU ORACLE data sources (test infrastructure, NOT product code).

* `circle_square_bank`  restates /root/reference/ap_gym/envs/image/circle_square_dataset.py:9-112
  (`CircleSquareDataset(show_gradient, image_shape, object_extents)` -> all `len(dataset)` images as the
  float32 (M,H,W,1) array `ImageClassificationDataset.get_data_point_batch` would return, plus labels).
  Pinned: tests compare its sha256 with the one of the reference-generated bank stored in the goldens.
* `synthetic_u8_bank`   the synthetic stand-ins for MNIST / CIFAR10 / TinyImageNet (no network for the real
  datasets): `default_rng(seed).integers(0, 256, (M,H,W,C), uint8)`, labels `integers(0, K, M)`
  (SURVEY.md section 8d).
"""
from __future__ import annotations

import hashlib

import numpy as np


def circle_square_bank(image_shape=(28, 28), show_gradient=True, object_extents=8):
    H0, H1 = image_shape
    coords = np.stack(np.meshgrid(np.arange(H0), np.arange(H1), indexing="ij"), axis=-1)
    max_vals = [2, H1, H0]
    M = int(np.prod(max_vals))
    imgs = np.empty((M, H0, H1, 1), dtype=np.float32)
    labels = np.empty(M, dtype=np.int32)
    max_dist = np.sqrt(np.sum(np.array(image_shape) ** 2))
    for idx in range(M):
        rem = idx
        vals = []
        for mv in max_vals:  # _unpack
            v = rem % mv
            vals.append(v)
            rem = (rem - v) // mv
        label, pos_x, pos_y = vals
        position = np.stack([pos_y, pos_x], axis=-1)
        if show_gradient:
            img = 1 - np.linalg.norm(position - coords, axis=-1) / max_dist
        else:
            img = np.zeros(image_shape)
        if label == 0:
            img[
                (position[0] - object_extents / 2 <= coords[:, :, 0])
                & (coords[:, :, 0] <= position[0] + object_extents / 2)
                & (position[1] - object_extents / 2 <= coords[:, :, 1])
                & (coords[:, :, 1] <= position[1] + object_extents / 2)
            ] = 1.0
        else:
            img[np.linalg.norm(position - coords, axis=-1) <= object_extents / 2] = 1.0
        imgs[idx] = img[:, :, None].astype(np.float32)
        labels[idx] = label
    return imgs, labels


def synthetic_u8_bank(seed: int, M: int, H: int, W: int, C: int, K: int):
    rng = np.random.default_rng(seed)
    imgs = rng.integers(0, 256, (M, H, W, C), dtype=np.uint8)
    labels = rng.integers(0, K, M).astype(np.int32)
    return imgs, labels


def sha256_of(arr: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()
