"""Generates tests/golden/lidar_*.npz and the floor-map banks tests/golden/maps_*.npz by running the
UNMODIFIED reference (/root/reference) through the test-only gymnasium shim and the exact-geometry shapely
stand-in (oracle/shim).  Run here:  python tests/golden/make_golden.py lidar [case ...]

The map banks are produced by the reference's own generators (FloorMapDatasetRooms / FloorMapDatasetMaze,
ap_gym/envs/floor_map/*.py) and stored bit-packed; they are also what bench.py steps on the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))

MAP_BANKS = {
    # name: (generator, kwargs, number of maps [dataset indices 0..M-1])
    "rooms32": ("rooms", {}, 4096),
    "maze21": ("maze", {}, 1024),
    "maze63": ("maze", dict(width=63, height=63), 256),
}

LIDAR_CASES = {
    "rooms_b64_n8": dict(bank="rooms32", M=64, N=8, B=64, steps=215, seed=0, static=False),
    "rooms_b8_static_n4": dict(bank="rooms32", M=16, N=4, B=8, steps=120, seed=7, static=True, static_idx=3),
    "maze21_b64_n6": dict(bank="maze21", M=32, N=6, B=64, steps=130, seed=3, static=False),
    "maze63_b256_n2": dict(bank="maze63", M=8, N=2, B=256, steps=110, seed=5, static=False),
    "maze21_b12_n3": dict(bank="maze21", M=8, N=3, B=12, steps=105, seed=11, static=False),
}


def load_bank(name: str, M: int | None = None) -> np.ndarray:
    """(M,H,W) bool from the committed bit-packed bank."""
    z = np.load(os.path.join(OUT, f"maps_{name}.npz"))
    bits, W = z["bits"], int(z["width"])
    M = len(bits) if M is None else M
    x = np.arange(W)
    return ((bits[:M, :, x >> 5] >> (x & 31).astype(np.uint32)) & 1).astype(bool)


def make_banks(names):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, ROOT)
    from ap_gym.envs.floor_map import FloorMapDatasetMaze, FloorMapDatasetRooms
    from oracle.lidar_oracle import pack_maps

    for name in names:
        kind, kw, M = MAP_BANKS[name]
        ds = FloorMapDatasetRooms(**kw) if kind == "rooms" else FloorMapDatasetMaze(**kw)
        sys.setrecursionlimit(20000)
        maps = np.stack([ds.get_data_point(i) for i in range(M)])
        np.savez_compressed(os.path.join(OUT, f"maps_{name}.npz"), bits=pack_maps(maps),
                            width=np.int32(maps.shape[2]), height=np.int32(maps.shape[1]))
        print("bank", name, maps.shape, "wall fraction %.3f" % maps.mean())


def run_case(name, case):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, ROOT)
    import ap_gym
    from ap_gym.envs.floor_map import FloorMapDataset
    from ap_gym.envs.lidar_localization2d import LIDARLocalization2DEnv

    maps = load_bank(case["bank"], case["M"])

    class BankDataset(FloorMapDataset):
        """Finite map bank through the reference's dataset interface."""

        def __init__(self, maps):
            self._maps = maps
            super().__init__(maps.shape[2], maps.shape[1])

        def get_data_point(self, idx):
            return self._maps[int(idx)]

        def get_data_point_batch(self, idx):
            return self._maps[np.asarray(idx)]

        def _get_length(self):
            return len(self._maps)

    # instrumentation only: record every __lidar_scan call (the reference source is untouched)
    calls = []
    orig = LIDARLocalization2DEnv._LIDARLocalization2DEnv__lidar_scan

    def recording_scan(self, pos, target_pos, eps=1e-3):
        d, c = orig(self, pos, target_pos, eps)
        calls.append((id(self), np.array(pos, copy=True), np.array(target_pos, copy=True), d.copy(), c.copy()))
        return d, c

    LIDARLocalization2DEnv._LIDARLocalization2DEnv__lidar_scan = recording_scan
    try:
        env_id = {"rooms32": "LIDARLocRooms", "maze21": "LIDARLocMaze", "maze63": "LIDARLocMaze"}[case["bank"]]
        env_id += ("Static" if case["static"] else "") + "-v0"
        kwargs = dict(dataset=BankDataset(maps), lidar_beam_count=case["B"])
        if case["static"]:
            kwargs["static_map_index"] = case["static_idx"]
        env = ap_gym.make_vec(env_id, num_envs=case["N"], **kwargs)
        N, B, steps = case["N"], case["B"], case["steps"]
        sub = [e.unwrapped for e in env.unwrapped.envs]
        ids = {id(s): i for i, s in enumerate(sub)}
        rng = np.random.default_rng(2000 + case["seed"])
        obs, info = env.reset(seed=case["seed"])
        rec = {"reset_" + k: np.asarray(v) for k, v in obs.items() if k != "map"}
        if "map" in obs:
            rec["reset_map"] = np.packbits(obs["map"][..., 0] > 0, axis=-1)
        rec["reset_map_idx"] = np.asarray(info.get("map_idx", np.full(N, case.get("static_idx", 0))), dtype=np.int64)
        rec["reset_pos"] = np.stack([s._LIDARLocalization2DEnv__pos.copy() for s in sub])
        calls.clear()
        keys = ["lidar", "odometry", "time_step", "reward", "terminated", "truncated", "base_reward", "target", "loss",
                "stepped", "pos", "cells", "dists", "map_idx", "n_scan_calls"]
        out = {k: [] for k in keys}
        actions = np.empty((steps, N, 2), np.float32)
        preds = np.empty((steps, N, 2), np.float32)
        heading = rng.uniform(-np.pi, np.pi, N)
        for s in range(steps):
            # persistent headings make the agents run into walls and slide along them
            heading += rng.normal(0, 0.6, N)
            speed = rng.uniform(0.0, 1.4, N)
            a = (np.stack([np.cos(heading), np.sin(heading)], -1) * speed[:, None]).astype(np.float32)
            if s % 9 == 4:
                a[: max(1, N // 3)] = 0
            if s % 6 == 2:
                a[:, rng.integers(0, 2)] = 0  # axis-parallel moves
            p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
            actions[s], preds[s] = a, p
            calls.clear()
            o, r, te, tr, inf = env.step({"action": a, "prediction": p})
            out["lidar"].append(o["lidar"].copy())
            out["odometry"].append(o["odometry"].copy())
            out["time_step"].append(o["time_step"].copy())
            out["reward"].append(np.asarray(r, np.float64))
            out["terminated"].append(te.copy())
            out["truncated"].append(tr.copy())
            stepped = np.asarray(inf.get("_base_reward", np.zeros(N, bool)), bool)
            out["stepped"].append(stepped)
            out["base_reward"].append(np.asarray(inf.get("base_reward", np.zeros(N)), np.float32))
            pr = inf.get("prediction", {})
            out["target"].append(np.asarray(pr.get("target", np.zeros((N, 2))), np.float32))
            out["loss"].append(np.asarray(pr.get("loss", np.zeros(N)), np.float32))
            out["pos"].append(np.stack([s_._LIDARLocalization2DEnv__pos.copy() for s_ in sub]))
            cells = np.full((N, B, 2), -9, np.int32)
            dists = np.zeros((N, B), np.float32)
            ncalls = np.zeros(N, np.int32)
            for (sid, _, tp, d, c) in calls:
                i = ids[sid]
                ncalls[i] += 1
                if len(tp) == B:  # the observation scan is the last B-target call of the env
                    cells[i], dists[i] = c, d
            out["cells"].append(cells)
            out["dists"].append(dists)
            out["n_scan_calls"].append(ncalls)
            out["map_idx"].append(np.array([s_._LIDARLocalization2DEnv__map_idx for s_ in sub], dtype=np.int64))
        env.close()
    finally:
        LIDARLocalization2DEnv._LIDARLocalization2DEnv__lidar_scan = orig
    rec.update({"actions": actions, "preds": preds})
    rec.update({k: np.stack(v) for k, v in out.items()})
    np.savez_compressed(os.path.join(OUT, f"lidar_{name}.npz"), **rec)
    print(name, "ok", {k: v.shape for k, v in rec.items()})


def main(argv):
    need = [n for n in MAP_BANKS if not os.path.exists(os.path.join(OUT, f"maps_{n}.npz"))]
    if need or "--banks" in argv:
        make_banks(need if "--banks" not in argv else list(MAP_BANKS))
    for name, case in LIDAR_CASES.items():
        if [a for a in argv if not a.startswith("--")] and name not in argv:
            continue
        run_case(name, case)


if __name__ == "__main__":
    main(sys.argv[1:])
