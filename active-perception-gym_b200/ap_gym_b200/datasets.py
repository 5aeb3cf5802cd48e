"""Image banks: the device-resident replacement of the reference's per-reset dataset gather
(ap_gym/envs/dataset/*, ap_gym/envs/image/image_classification_dataset.py:54-98).

Any object with the reference's dataset surface works as `ImagePerceptionConfig.dataset`:
`load()`, `__len__`, `__getitem__(indices) -> (images float32 (n,H,W,C) in [0,1], labels int32)`,
`num_classes`.  The whole dataset is read ONCE at construction and uploaded; banks whose pixels are exactly
k/255 (every uint8-sourced dataset) are stored as uint8 and converted in-kernel with the reference's
`float32(u8) / 255`, which is bit-identical and moves a quarter of the bytes.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ImageBank:
    images: np.ndarray  # (M, H, W, C) uint8 or float32, C-contiguous
    labels: np.ndarray  # (M,) int32
    num_classes: int

    @property
    def is_u8(self) -> bool:
        return self.images.dtype == np.uint8

    def __len__(self):
        return len(self.images)


def _quantise_if_lossless(imgs_f32: np.ndarray):
    q = np.rint(imgs_f32 * 255.0)
    if np.all((q >= 0) & (q <= 255)):
        q8 = q.astype(np.uint8)
        if np.array_equal(q8.astype(np.float32) / 255, imgs_f32):
            return q8
    return None


def bank_from_dataset(dataset, chunk: int = 8192) -> ImageBank:
    if isinstance(dataset, ImageBank):
        return dataset
    if hasattr(dataset, "as_bank"):
        return dataset.as_bank()
    if hasattr(dataset, "load"):
        dataset.load()
    M = len(dataset)
    imgs, labels = [], []
    for s in range(0, M, chunk):
        idx = np.arange(s, min(M, s + chunk))
        im, lb = dataset[idx]
        im = np.asarray(im)
        if im.ndim == 3:
            im = im[..., None]
        imgs.append(im)
        labels.append(np.asarray(lb).astype(np.int32))
    images = np.ascontiguousarray(np.concatenate(imgs))
    labels = np.ascontiguousarray(np.concatenate(labels))
    if images.dtype != np.uint8:
        images = images.astype(np.float32, copy=False)
        q = _quantise_if_lossless(images)
        if q is not None:
            images = q
    return ImageBank(np.ascontiguousarray(images), labels, int(dataset.num_classes))


class ArrayImageDataset:
    """A finite in-memory dataset with the reference's dataset surface."""

    def __init__(self, images: np.ndarray, labels: np.ndarray, num_classes: int):
        images = np.asarray(images)
        if images.ndim == 3:
            images = images[..., None]
        if images.dtype != np.uint8:
            images = images.astype(np.float32)
        self._images = np.ascontiguousarray(images)
        self._labels = np.ascontiguousarray(np.asarray(labels).astype(np.int32))
        self._k = int(num_classes)

    def load(self):
        pass

    def __len__(self):
        return len(self._images)

    @property
    def num_classes(self):
        return self._k

    def _to_f32(self, im):
        return im.astype(np.float32) / 255 if im.dtype == np.uint8 else im

    def __getitem__(self, item):
        if isinstance(item, (list, tuple, np.ndarray)):
            item = np.asarray(item)
            return self._to_f32(self._images[item]), self._labels[item]
        return self._to_f32(self._images[int(item)]), int(self._labels[int(item)])

    def as_bank(self) -> ImageBank:
        bank = ImageBank(self._images, self._labels, self._k)
        if not bank.is_u8:
            q = _quantise_if_lossless(self._images)
            if q is not None:
                bank = ImageBank(q, self._labels, self._k)
        return bank


class SyntheticImageDataset(ArrayImageDataset):
    """Synthetic stand-in for the HuggingFace datasets (no network): uniform random uint8 pixels and labels,
    `default_rng(seed).integers(0, 256, (M,H,W,C), uint8)` / `integers(0, K, M)`."""

    SHAPES = {"mnist": (60000, 28, 28, 1, 10), "cifar10": (50000, 32, 32, 3, 10),
              "tinyimagenet": (100000, 64, 64, 3, 200)}

    def __init__(self, kind: str = "mnist", length: int | None = None, seed: int = 0):
        M, H, W, C, K = self.SHAPES[kind]
        M = int(length) if length is not None else M
        rng = np.random.default_rng(seed)
        images = rng.integers(0, 256, (M, H, W, C), dtype=np.uint8)
        labels = rng.integers(0, K, M)
        super().__init__(images, labels, K)
