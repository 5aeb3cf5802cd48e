"""Shared helpers for the parity tests (golden loading, case table, comparisons)."""
from __future__ import annotations

import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_mk = _load("make_golden", os.path.join(GOLDEN, "make_golden.py"))
IMAGE_CASES = _mk.IMAGE_CASES
make_bank = _mk.make_bank


def load_golden(kind: str, name: str):
    return np.load(os.path.join(GOLDEN, f"{kind}_{name}.npz"), allow_pickle=False)


def image_oracle_for(case):
    from oracle.image_oracle import ImageEnvOracle

    imgs, labels = make_bank(case["bank"])
    return ImageEnvOracle(imgs, labels, case["N"], task=case["task"], num_classes=case["K"], **case["cfg"]), imgs, labels


def assert_close(a, b, rtol=1e-5, atol=1e-6, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]!r} want {b[i]!r} (err {err[i]:.3e})")

_mkl = _load("make_golden_lidar", os.path.join(GOLDEN, "make_golden_lidar.py"))
LIDAR_CASES = _mkl.LIDAR_CASES
load_bank = _mkl.load_bank


def lidar_oracle_for(case):
    from oracle.lidar_oracle import LidarVecOracle

    maps = load_bank(case["bank"], case["M"])
    return LidarVecOracle(maps, case["N"], beam_count=case["B"], static_map=case["static"],
                          static_map_index=case.get("static_idx", 0)), maps
