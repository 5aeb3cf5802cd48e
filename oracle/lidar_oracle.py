"""CPU ORACLE (test infrastructure, NOT product code) for the vectorised LIDAR localisation environment.

Restates what `ap_gym.make_vec("LIDARLoc*-v0", num_envs=N, ...)` computes per vector step:
  gymnasium SyncVectorEnv (NEXT_STEP autoreset, seeds s+i, f64 rewards)      un-vendored, restated in oracle/shim
  TimeLimit(100, issue_termination=True) + time_step obs                     ap_gym/time_limit.py:113-152
  ActivePerceptionEnv.step (reward = base - loss, info replaced)             ap_gym/active_perception_env.py:101-121
  LIDARLocalization2DEnv.reset / _step / __get_obs / __lidar_scan            ap_gym/envs/lidar_localization2d.py:238-536
  reset RNG chain (per-env Generators)                                       ap_gym/envs/lidar_localization2d.py:303-311,547-557,
                                                                             ap_gym/envs/dataset/dataset_iterator.py:10-32
The per-ray / per-env arithmetic is the C file oracle/lidar_oracle.c (built with gcc by oracle/build_oracle.py).

PARITY STATUS: pinned against the unmodified reference running on the exact-geometry shapely stand-in
(tests/golden/lidar_*.npz); UNPINNED against real GEOS, which is absent (see lidar_oracle.c header).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .build_oracle import build

_lib = None


def _l():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.apg_oracle_scan_one.restype = C.c_float
        _lib.apg_oracle_scan_one.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                             C.c_float, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_maps(maps: np.ndarray) -> np.ndarray:
    """(M,H,W) bool (True = wall) -> (M,H,ceil(W/32)) uint32, bit x&31 of word x>>5."""
    maps = np.asarray(maps, dtype=bool)
    M, H, W = maps.shape
    words = (W + 31) // 32
    out = np.zeros((M, H, words), dtype=np.uint32)
    for x in range(W):
        out[:, :, x >> 5] |= maps[:, :, x].astype(np.uint32) << np.uint32(x & 31)
    return out


def beam_table(beam_count: int, lidar_range: float) -> np.ndarray:
    """lidar_directions (lidar_localization2d.py:180-187): float32 (B,2)."""
    ang = np.linspace(-np.pi, np.pi, beam_count, dtype=np.float32, endpoint=False)
    unscaled = np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    return np.ascontiguousarray(unscaled * lidar_range)


def scan(bits: np.ndarray, W: int, H: int, pos: np.ndarray, targets: np.ndarray):
    """__lidar_scan for one map: returns (distances f32 (n,), contact cells i32 (n,2))."""
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    targets = np.ascontiguousarray(targets, dtype=np.float32)
    n = len(targets)
    dist = np.empty(n, np.float32)
    cells = np.empty((n, 2), np.int32)
    _l().apg_oracle_scan(_p(bits), W, H, bits.shape[-1], _p(pos), _p(targets), n, _p(dist), _p(cells))
    return dist, cells


class LidarVecOracle:
    def __init__(self, maps: np.ndarray | None, num_envs: int, beam_count: int = 8, lidar_range: float = 5,
                 step_limit: int = 100, static_map: bool = False, static_map_index: int = 0, dataset_fn=None,
                 map_shape: tuple[int, int] | None = None):
        """`maps`: a finite bank (M,H,W), or None with `dataset_fn(idx) -> (H,W) bool` and `map_shape=(H,W)`: the
        reference's procedural datasets of len 2**32 (every reset fetches dataset[integers(0, 2**32)],
        dataset_iterator.py:27-32); each environment then owns its current map."""
        self.N = int(num_envs)
        self.dataset_fn = dataset_fn
        if dataset_fn is not None:
            self.H, self.W = map_shape
            self.M = 2 ** 32
            self.maps = np.zeros((self.N, self.H, self.W), dtype=bool)   # per-environment current maps
            self.bank = pack_maps(self.maps)
        else:
            self.maps = np.asarray(maps, dtype=bool)
            self.M, self.H, self.W = self.maps.shape
            self.bank = pack_maps(self.maps)
        self.words = self.bank.shape[-1]
        self.B = int(beam_count)
        self.range = float(lidar_range)
        self.T = int(step_limit)
        self.static = bool(static_map)
        self.static_idx = int(static_map_index)
        self.dirs = beam_table(self.B, lidar_range)
        N, B = self.N, self.B
        self.pos = np.zeros((N, 2), np.float32)
        self.init = np.zeros((N, 2), np.float32)
        self.aliased = np.zeros(N, np.int32)
        self.elapsed = np.zeros(N, np.int32)
        self.map_idx = np.zeros(N, np.int64)
        self.autoreset = np.zeros(N, bool)
        self.lidar = np.zeros((N, B), np.float32)
        self.odo = np.zeros((N, 2), np.float32)
        self.cells = np.zeros((N, B, 2), np.int32)
        self.dists = np.zeros((N, B), np.float32)
        self.g0 = [None] * N
        self.g2 = [None] * N

    # -- reset path (lidar_localization2d.py:293-315) -----------------------------------------------------
    def _reset_env(self, i: int):
        if not self.static:
            self.map_idx[i] = self.g2[i].integers(0, self.M)
        else:
            self.map_idx[i] = self.static_idx
        if self.dataset_fn is not None:
            self.maps[i] = self.dataset_fn(int(self.map_idx[i]))
            self.bank[i] = pack_maps(self.maps[i][None])[0]
            m = self.maps[i]
        else:
            m = self.maps[self.map_idx[i]]
        rows, cols = np.where(m == 0)
        k = self.g0[i].integers(0, len(rows))
        self.pos[i] = np.array([cols[k], rows[k]], dtype=np.float32) + 0.5
        self.init[i] = self.pos[i]
        self.aliased[i] = 1
        self.elapsed[i] = 0

    def _c_idx(self):
        """bank rows of the environments for the C batch calls"""
        return np.arange(self.N, dtype=np.int64) if self.dataset_fn is not None else self.map_idx

    def _obs_dict(self):
        obs = {"lidar": self.lidar.copy(), "odometry": self.odo.copy()}
        if not self.static:
            obs["map"] = (self.maps[self._c_idx()][..., None].astype(np.float32) / 255)
        obs["time_step"] = (2.0 * self.elapsed / self.T - 1.0).astype(np.float32)
        return obs

    def _scan_obs(self, mask):
        _l().apg_oracle_obs_batch(_p(self.bank), self.W, self.H, self.words, _p(self._c_idx()), self.N,
                                  _p(mask.astype(np.uint8)), _p(self.pos), _p(self.init), _p(self.dirs), self.B,
                                  C.c_float(self.range), _p(self.lidar), _p(self.odo), _p(self.cells), _p(self.dists))

    def reset(self, seed: int):
        for i in range(self.N):
            self.g0[i] = np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(seed) + i)))
            if not self.static:
                self.g2[i] = np.random.default_rng(self.g0[i].integers(0, 2**32, endpoint=True))
            self._reset_env(i)
        self.autoreset[:] = False
        self._scan_obs(np.ones(self.N, bool))
        return self._obs_dict(), {"map_idx": self.map_idx.copy()}

    def step(self, action: np.ndarray, prediction: np.ndarray):
        N = self.N
        action = np.ascontiguousarray(action, dtype=np.float32)
        prediction = np.ascontiguousarray(prediction, dtype=np.float32)
        if np.any(np.isnan(action[~self.autoreset])):
            raise ValueError("NaN values detected in action.")
        if np.any(np.isnan(prediction[~self.autoreset])):
            raise ValueError("NaN values detected in prediction.")
        stepped = ~self.autoreset
        target = np.zeros((N, 2), np.float32)
        base = np.zeros(N, np.float32)
        loss = np.zeros(N, np.float32)
        oob = np.zeros(N, np.uint8)
        for i in np.nonzero(self.autoreset)[0]:
            self._reset_env(i)
        _l().apg_oracle_step_batch(_p(self.bank), self.W, self.H, self.words, _p(self._c_idx()), N,
                                   _p(stepped.astype(np.uint8)), _p(self.pos), _p(self.init), _p(self.aliased),
                                   _p(action), _p(prediction), _p(self.dirs), self.B, C.c_float(self.range),
                                   _p(self.lidar), _p(self.odo), _p(self.cells), _p(self.dists), _p(target), _p(base),
                                   _p(loss), _p(oob))
        if self.autoreset.any():
            self._scan_obs(self.autoreset)
        self.elapsed[stepped] += 1
        terminated = stepped & ((oob != 0) | (self.elapsed >= self.T))
        truncated = np.zeros(N, bool)
        reward = np.where(stepped, (base - loss).astype(np.float64), 0.0)
        info = {"base_reward": base, "target": target, "loss": loss, "stepped": stepped.copy(),
                "map_idx": self.map_idx.copy(), "cells": self.cells.copy(), "dists": self.dists.copy()}
        self.autoreset = terminated | truncated
        return self._obs_dict(), reward, terminated, truncated, info
