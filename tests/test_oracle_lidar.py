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
