"""Adversarial single-ray golden vectors: the reference's own `__lidar_scan` (on the exact-geometry shapely
stand-in) evaluated on degenerate rays -- exact diagonals through lattice corners from cell centres, start
points on grid lines, end points on grid lines, axis-parallel rays along grid lines -- plus random rays.

    python tests/golden/make_golden_scan.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))


def build_rays(rng, W, H, n_random=1500):
    pos, tgt = [], []

    def add(p, t):
        pos.append(p)
        tgt.append(t)

    R = 5.0
    for B in (8, 64):
        ang = np.linspace(-np.pi, np.pi, B, dtype=np.float32, endpoint=False)
        dirs = np.stack([np.cos(ang), np.sin(ang)], -1) * R
        for _ in range(12):  # cell centres: exact diagonal beams hit lattice corners
            p = (rng.integers(1, [W - 1, H - 1]) + 0.5).astype(np.float32)
            for d in dirs:
                add(p, p + d)
    for _ in range(300):  # start on a vertical / horizontal grid line or a lattice point
        p = rng.uniform(1, [W - 1, H - 1]).astype(np.float32)
        k = rng.integers(0, 3)
        if k in (0, 2):
            p[0] = np.round(p[0])
        if k in (1, 2):
            p[1] = np.round(p[1])
        d = rng.uniform(-3, 3, 2).astype(np.float32)
        if rng.random() < 0.4:
            d[rng.integers(0, 2)] = 0  # axis-parallel, possibly along the grid line itself
        if d[0] == 0 and d[1] == 0:
            d[0] = 1
        add(p, p + d)
    for _ in range(300):  # end exactly on a grid line / lattice point
        p = rng.uniform(1, [W - 1, H - 1]).astype(np.float32)
        t = rng.uniform(1, [W - 1, H - 1]).astype(np.float32)
        k = rng.integers(0, 3)
        if k in (0, 2):
            t[0] = np.round(t[0])
        if k in (1, 2):
            t[1] = np.round(t[1])
        add(p, t)
    for _ in range(200):  # exact diagonals between lattice-aligned points
        p = (rng.integers(1, [W - 1, H - 1]) + rng.choice([0.0, 0.25, 0.5])).astype(np.float32)
        s = rng.choice([-1.0, 1.0], 2).astype(np.float32) * np.float32(rng.integers(1, 5))
        add(p, p + s)
    for _ in range(n_random):
        p = rng.uniform(0.5, [W - 0.5, H - 0.5]).astype(np.float32)
        d = rng.uniform(-5, 5, 2).astype(np.float32)
        add(p, p + d)
    return np.array(pos, np.float32), np.array(tgt, np.float32)


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, ROOT)
    import ap_gym  # noqa: F401
    from ap_gym.envs.floor_map import FloorMapDataset
    from ap_gym.envs.lidar_localization2d import LIDARLocalization2DEnv

    rng = np.random.default_rng(42)
    W = H = 14
    maps = rng.random((6, H, W)) < np.array([0.15, 0.3, 0.45, 0.6, 0.3, 0.45])[:, None, None]
    maps[:, 0, :] = maps[:, -1, :] = maps[:, :, 0] = maps[:, :, -1] = True
    maps[4:, 0, 3:6] = False  # holes in the border: rays may leave the map

    class Bank(FloorMapDataset):
        def __init__(self):
            super().__init__(W, H)

        def get_data_point(self, idx):
            return maps[int(idx)]

        def get_data_point_batch(self, idx):
            return maps[np.asarray(idx)]

        def _get_length(self):
            return len(maps)

    rec = {"maps": maps}
    for m in range(len(maps)):
        env = LIDARLocalization2DEnv(Bank(), static_map=True, static_map_index=m)
        pos, tgt = build_rays(rng, W, H)
        dist = np.empty(len(pos), np.float32)
        cells = np.empty((len(pos), 2), np.int32)
        for i in range(len(pos)):
            d, c = env._LIDARLocalization2DEnv__lidar_scan(pos[i], tgt[i][None])
            dist[i], cells[i] = d[0], c[0]
        rec[f"pos{m}"], rec[f"tgt{m}"], rec[f"dist{m}"], rec[f"cells{m}"] = pos, tgt, dist, cells
        print("map", m, "rays", len(pos), "hits", int((cells[:, 0] >= 0).sum()), "zero-dist", int((dist == 0).sum()))
    np.savez_compressed(os.path.join(OUT, "lidar_scan_adversarial.npz"), **rec)


if __name__ == "__main__":
    main()
