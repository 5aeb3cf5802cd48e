"""CPU: the LIDAR oracle (oracle/lidar_oracle.c + .py) against golden traces recorded from the UNMODIFIED
reference running on the exact-geometry shapely stand-in (tests/golden/make_golden_lidar.py)."""
import numpy as np
import pytest

from helpers import LIDAR_CASES, assert_close, lidar_oracle_for, load_golden


@pytest.mark.parametrize("name", list(LIDAR_CASES))
def test_lidar_oracle_matches_reference_golden(name):
    case = LIDAR_CASES[name]
    g = load_golden("lidar", name)
    env, maps = lidar_oracle_for(case)
    obs, info = env.reset(seed=case["seed"])
    assert np.array_equal(info["map_idx"], g["reset_map_idx"])
    assert np.array_equal(env.pos, g["reset_pos"])
    assert np.array_equal(obs["lidar"], g["reset_lidar"])
    assert np.array_equal(obs["odometry"], g["reset_odometry"])
    assert np.array_equal(obs["time_step"], g["reset_time_step"])
    if not case["static"]:
        assert np.array_equal(np.packbits(obs["map"][..., 0] > 0, axis=-1), g["reset_map"])
        assert set(np.unique(obs["map"])) <= {np.float32(0), np.float32(1) / np.float32(255)}
    for s in range(case["steps"]):
        o, r, te, tr, inf = env.step(g["actions"][s], g["preds"][s])
        st = g["stepped"][s]
        assert np.array_equal(inf["stepped"], st), f"step {s}: which envs autoreset"
        assert np.array_equal(env.pos, g["pos"][s]), f"step {s} pos"
        assert np.array_equal(inf["map_idx"], g["map_idx"][s])
        assert np.array_equal(te, g["terminated"][s]) and np.array_equal(tr, g["truncated"][s])
        assert np.array_equal(inf["cells"], g["cells"][s]), f"step {s} hit cells"
        assert np.array_equal(inf["dists"], g["dists"][s]), f"step {s} raw distances"
        assert np.array_equal(o["lidar"], g["lidar"][s]), f"step {s} lidar"
        assert np.array_equal(o["odometry"], g["odometry"][s]), f"step {s} odometry"
        assert np.array_equal(o["time_step"], g["time_step"][s])
        assert np.array_equal(inf["target"][st], g["target"][s][st])
        assert np.array_equal(inf["base_reward"][st], g["base_reward"][s][st])
        assert np.array_equal(inf["loss"][st], g["loss"][s][st])
        assert np.array_equal(r, g["reward"][s])


def test_scan_matches_reference_on_adversarial_rays():
    """Single rays through lattice corners, along grid lines, starting/ending on grid lines (every branch of
    lidar_localization2d.py:502-533) -- bit-exact distances and cells vs the reference on the shapely stand-in."""
    import os

    from helpers import GOLDEN
    from oracle.lidar_oracle import pack_maps, scan

    g = np.load(os.path.join(GOLDEN, "lidar_scan_adversarial.npz"))
    maps = g["maps"]
    bank = pack_maps(maps)
    H, W = maps.shape[1:]
    for m in range(len(maps)):
        pos, tgt = g[f"pos{m}"], g[f"tgt{m}"]
        for i in range(len(pos)):
            d, c = scan(bank[m], W, H, pos[i], tgt[i][None])
            assert d[0] == g[f"dist{m}"][i] and np.array_equal(c[0], g[f"cells{m}"][i]), (
                f"map {m} ray {i}: pos {pos[i]} -> {tgt[i]}: got {d[0]} {c[0]}, "
                f"reference {g[f'dist{m}'][i]} {g[f'cells{m}'][i]}")


@pytest.mark.reference
@pytest.mark.parametrize("env_id,kind", [("LIDARLocRooms-v0", "rooms"), ("LIDARLocMaze-v0", "maze")])
def test_procedural_dataset_mode_matches_live_reference(env_id, kind):
    """The reference's own registration (FloorMapDatasetRooms() / FloorMapDatasetMaze(), len 2**32, per-episode
    dataset[integers(0, 2**32)]) run unmodified under the test-only shims, against the oracle in dataset mode with the
    iterative generators of oracle/floor_map_oracle.py: dataset indices, maps, positions, observations, rewards across an
    autoreset."""
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle", "shim"))
    sys.path.insert(0, "/root/reference")
    import ap_gym

    from oracle.floor_map_oracle import maze, rooms
    from oracle.lidar_oracle import LidarVecOracle

    sys.setrecursionlimit(20000)
    N, steps, seed = 3, 104, 5
    env = ap_gym.make_vec(env_id, num_envs=N)
    B = env.single_observation_space["lidar"].shape[0]
    H, W = env.single_observation_space["map"].shape[:2]
    fn = (lambda i: rooms(i, W, H)) if kind == "rooms" else (lambda i: maze(i, W, H))
    oracle = LidarVecOracle(None, N, beam_count=B, dataset_fn=fn, map_shape=(H, W))
    ro, rinfo = env.reset(seed=seed)
    oo, oinfo = oracle.reset(seed=seed)
    assert np.array_equal(np.asarray(rinfo["map_idx"], dtype=np.int64), oinfo["map_idx"])
    assert np.array_equal(ro["map"], oo["map"]) and np.array_equal(ro["lidar"], oo["lidar"])
    rng = np.random.default_rng(1)
    for s in range(steps):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        ro, rr, rte, rtr, rinfo = env.step({"action": a, "prediction": p})
        oo, orw, ote, otr, oinf = oracle.step(a, p)
        if "map_idx" in rinfo:  # only the environments that were reset by this step report it (gymnasium `_key` masks)
            m = np.asarray(rinfo.get("_map_idx", np.ones(N, bool)), dtype=bool)
            assert np.array_equal(np.asarray(rinfo["map_idx"], dtype=np.int64)[m], oinf["map_idx"][m]), f"step {s} indices"
        for k in ("lidar", "odometry", "map", "time_step"):
            assert np.array_equal(ro[k], oo[k]), f"step {s} {k}"
        assert np.array_equal(rte, ote) and np.array_equal(rtr, otr)
        assert np.allclose(rr, orw, rtol=0, atol=1e-7)
    env.close()
