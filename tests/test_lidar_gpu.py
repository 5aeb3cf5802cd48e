"""GPU parity tests of the fused LIDAR step (through the C-ABI / ctypes layer) against (a) golden traces
recorded from the unmodified reference on the exact-geometry shapely stand-in and (b) the C oracle on seeded
inputs.  Bars: bit-exact hit cells, collision outcomes (positions), flags, map choices; float distances,
odometry and rewards are compared bit-exactly as well (same IEEE op order), which implies the 1e-5 bar."""
import numpy as np
import pytest

from helpers import LIDAR_CASES, lidar_oracle_for, load_bank, load_golden

pytestmark = pytest.mark.gpu


def _np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


def _engine(case, maps, **kw):
    import ap_gym_b200 as apg

    return apg.LIDARLocalization2DVectorEnv(
        case["N"], apg.FloorMapBank.from_maps(maps), static_map=case["static"], lidar_beam_count=case["B"],
        static_map_index=case.get("static_idx", 0), emit_cells=True, check_errors="step", **kw)


@pytest.mark.parametrize("mode", ["numpy", "torch"])
@pytest.mark.parametrize("name", list(LIDAR_CASES))
def test_engine_matches_reference_golden(name, mode):
    import torch

    case = LIDAR_CASES[name]
    g = load_golden("lidar", name)
    maps = load_bank(case["bank"], case["M"])
    env = _engine(case, maps)
    obs, info = env.reset(seed=case["seed"])
    assert np.array_equal(_np(info["map_idx"]), g["reset_map_idx"])
    assert np.array_equal(_np(env.state()["pos"]), g["reset_pos"])
    assert np.array_equal(_np(obs["lidar"]), g["reset_lidar"])
    assert np.array_equal(_np(obs["odometry"]), g["reset_odometry"])
    assert np.array_equal(_np(obs["time_step"]), g["reset_time_step"])
    if not case["static"]:
        assert np.array_equal(np.packbits(_np(obs["map"])[..., 0] > 0, axis=-1), g["reset_map"])
        assert set(np.unique(_np(obs["map"]))) <= {np.float32(0), np.float32(1) / np.float32(255)}
    for s in range(case["steps"]):
        a, p = g["actions"][s], g["preds"][s]
        if mode == "torch":
            a, p = torch.from_numpy(a).cuda(), torch.from_numpy(p).cuda()
        o, r, te, tr, inf = env.step({"action": a, "prediction": p})
        st = g["stepped"][s]
        assert np.array_equal(_np(inf["_base_reward"]), st), f"step {s}: which envs autoreset"
        assert np.array_equal(_np(env.state()["pos"]), g["pos"][s]), f"step {s} pos"
        assert np.array_equal(_np(inf["map_idx"]), g["map_idx"][s])
        assert np.array_equal(_np(te), g["terminated"][s]) and np.array_equal(_np(tr), g["truncated"][s])
        assert np.array_equal(_np(env.state()["cells"]), g["cells"][s]), f"step {s} hit cells"
        assert np.array_equal(_np(env.state()["dists"]), g["dists"][s]), f"step {s} raw distances"
        assert np.array_equal(_np(o["lidar"]), g["lidar"][s]), f"step {s} lidar"
        assert np.array_equal(_np(o["odometry"]), g["odometry"][s]), f"step {s} odometry"
        assert np.array_equal(_np(o["time_step"]), g["time_step"][s])
        assert np.array_equal(_np(inf["prediction"]["target"])[st], g["target"][s][st])
        assert np.array_equal(_np(inf["base_reward"])[st], g["base_reward"][s][st])
        assert np.array_equal(_np(inf["prediction"]["loss"])[st], g["loss"][s][st])
        assert np.array_equal(_np(r).astype(np.float64), g["reward"][s])
    env.close()


# GEOS-DEPENDENT RAYS.  The reference's result for a ray that grazes a lattice corner (zero-length intersection pieces:
# Point / MultiPoint / GeometryCollection, lidar_localization2d.py:512-528) is whatever shapely's GEOS returns; engine and
# oracle follow the reference's type dispatch under EXACT geometry, which no real GEOS has confirmed here
# ("GEOS unpinned").  Such rays only occur on the four exact-diagonal beams of the scan right after a reset.  Setting
# APG_TEST_MASK_GEOS_RAYS=1 removes those beams of reset observations from the bit-exact assertions below, should a
# real shapely ever disagree; by default they are compared like every other ray.
import os

MASK_GEOS_RAYS = os.environ.get("APG_TEST_MASK_GEOS_RAYS", "0") == "1"


def _beam_mask(B, reset_rows):
    """(N, B) bool: beams that take part in the comparison (all, unless the GEOS switch masks the exact diagonals of
    freshly reset environments)."""
    keep = np.ones((len(reset_rows), B), bool)
    if MASK_GEOS_RAYS and B % 8 == 0:
        diag = np.array([B // 8 * k for k in (1, 3, 5, 7)])
        keep[np.ix_(np.flatnonzero(reset_rows), diag)] = False
    return keep


@pytest.mark.parametrize("bank,B,N,steps", [("rooms32", 64, 2048, 230), ("maze21", 8, 1024, 120), ("maze63", 256, 256, 110),
                                            ("maze63", 256, 2304, 104)])
def test_engine_matches_oracle_seeded(bank, B, N, steps):
    _seeded_vs_oracle(bank, B, N, steps)


@pytest.mark.parametrize("bank,B,N,steps,beam_filter", [("rooms32", 64, 1024, 120, "0"), ("rooms32", 64, 1024, 120, "1"),
                                                        ("maze63", 256, 256, 60, "1"), ("maze21", 12, 512, 60, "1")])
def test_beam_filter_on_and_off(bank, B, N, steps, beam_filter, monkeypatch):
    """The beam pre-filter (csrc/apg_lidar_filter.cuh: rays that cannot touch a wall skip the exact walk) is chosen per
    map bank; forced on and off here, both must reproduce the oracle bit for bit."""
    monkeypatch.setenv("APG_LIDAR_FILTER", beam_filter)
    _seeded_vs_oracle(bank, B, N, steps)


def _seeded_vs_oracle(bank, B, N, steps):
    """Larger seeded runs (thousands of envs, >2 episodes incl. the device-side RNG resets) vs the C oracle; the last
    case is the LIDARLocMaze benchmark shape (63x63 mazes, 256 beams) at N = 2304 across one autoreset.  GEOS unpinned:
    see MASK_GEOS_RAYS above."""
    import ap_gym_b200 as apg
    from oracle.lidar_oracle import LidarVecOracle

    maps = load_bank(bank, 128)
    oracle = LidarVecOracle(maps, N, beam_count=B)
    env = apg.LIDARLocalization2DVectorEnv(N, apg.FloorMapBank.from_maps(maps), lidar_beam_count=B, emit_cells=True,
                                           output="numpy")
    oo, oinfo = oracle.reset(seed=123)
    eo, einfo = env.reset(seed=123)
    assert np.array_equal(einfo["map_idx"], oinfo["map_idx"])
    assert np.array_equal(eo["lidar"], oo["lidar"])
    rng = np.random.default_rng(9)
    heading = rng.uniform(-np.pi, np.pi, N)
    for s in range(steps):
        heading += rng.normal(0, 0.5, N)
        speed = rng.uniform(0, 1.3, N)
        a = (np.stack([np.cos(heading), np.sin(heading)], -1) * speed[:, None]).astype(np.float32)
        p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        oo, orw, ote, otr, oinf = oracle.step(a, p)
        eo, erw, ete, etr, einf = env.step({"action": a, "prediction": p})
        assert np.array_equal(_np(env.state()["pos"]), oracle.pos), f"step {s} pos"
        keep = _beam_mask(B, ~einf["_base_reward"])  # rows that were reset by this step
        assert np.array_equal(_np(env.state()["cells"])[keep], oinf["cells"][keep]), f"step {s} cells"
        assert np.array_equal(eo["lidar"][keep], oo["lidar"][keep]), f"step {s} lidar"
        assert np.array_equal(eo["odometry"], oo["odometry"])
        assert np.array_equal(eo["time_step"], oo["time_step"])
        assert np.array_equal(ete, ote) and np.array_equal(etr, otr)
        assert np.array_equal(einf["map_idx"], oinf["map_idx"])
        assert np.array_equal(erw.astype(np.float64), orw)
    env.close()


def test_unseeded_reset_continues_the_streams():
    import ap_gym_b200 as apg
    from oracle.lidar_oracle import LidarVecOracle

    maps = load_bank("rooms32", 32)
    N, B = 64, 8
    oracle = LidarVecOracle(maps, N, beam_count=B)
    env = apg.LIDARLocalization2DVectorEnv(N, apg.FloorMapBank.from_maps(maps), lidar_beam_count=B, output="numpy")
    oracle.reset(seed=5)
    env.reset(seed=5)
    rng = np.random.default_rng(0)
    for s in range(7):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        oracle.step(a, a)
        env.step({"action": a, "prediction": a})
    # SyncVectorEnv.reset() without a seed: every sub-env resets, Generators continue
    for i in range(N):
        oracle._reset_env(i)
    oracle.autoreset[:] = False
    oracle._scan_obs(np.ones(N, bool))
    eo, einfo = env.reset()
    assert np.array_equal(einfo["map_idx"], oracle.map_idx)
    assert np.array_equal(_np(env.state()["pos"]), oracle.pos)
    assert np.array_equal(eo["lidar"], oracle.lidar)
    env.close()


def test_adversarial_rays_through_the_engine():
    """The degenerate single rays of tests/golden/lidar_scan_adversarial.npz, cast as *beams*: an env is placed
    by choosing a one-free-cell... not possible for arbitrary positions, so instead the move ray is used:
    action = target - pos from a known position is not controllable either; the engine-level check therefore
    goes through full traces above, and the ray-level check compares the device scan with the oracle on the
    beams of many random-walk steps (every branch is counted by the oracle test)."""
    import ap_gym_b200 as apg
    from oracle.lidar_oracle import LidarVecOracle

    # dense 14x14 maps with holes in the border: rays leave the map, many multi-piece intersections
    g = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "lidar_scan_adversarial.npz"))
    maps = g["maps"][:3]
    N, B = 512, 64
    oracle = LidarVecOracle(maps, N, beam_count=B)
    env = apg.LIDARLocalization2DVectorEnv(N, apg.FloorMapBank.from_maps(maps), lidar_beam_count=B, emit_cells=True,
                                           output="numpy")
    oracle.reset(seed=1)
    env.reset(seed=1)
    rng = np.random.default_rng(3)
    for s in range(60):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        if s % 4 == 0:
            a = np.round(a * 2) / 2  # exact half-cell moves land on grid lines and lattice points
        p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        _, _, _, _, oinf = oracle.step(a, p)
        env.step({"action": a, "prediction": p})
        assert np.array_equal(_np(env.state()["pos"]), oracle.pos), f"step {s} pos"
        assert np.array_equal(_np(env.state()["cells"]), oinf["cells"]), f"step {s} cells"
        assert np.array_equal(_np(env.state()["dists"]), oinf["dists"]), f"step {s} dists"
    env.close()


def test_lidar_full_size_properties():
    """BASELINE config 3 size (N=262144, 64 beams, rooms): determinism, shard invariance (env_offset), episode
    length 100 + 1 autoreset step, value ranges, reward identity."""
    import torch
    import ap_gym_b200 as apg

    N, B = 262144, 64
    bank = apg.FloorMapBank.from_maps(load_bank("rooms32", 1024))
    env = apg.LIDARLocalization2DVectorEnv(N, bank, lidar_beam_count=B)
    lo, hi = 5000, 5000 + 8192
    shard = apg.LIDARLocalization2DVectorEnv(hi - lo, bank, lidar_beam_count=B, env_offset=lo)
    o, info = env.reset(seed=17)
    so, sinfo = shard.reset(seed=17)
    assert torch.equal(o["lidar"][lo:hi], so["lidar"]) and torch.equal(info["map_idx"][lo:hi], sinfo["map_idx"])
    gen = torch.Generator(device="cuda").manual_seed(0)
    for s in range(103):
        a = torch.rand((N, 2), generator=gen, device="cuda") * 2 - 1
        p = torch.rand((N, 2), generator=gen, device="cuda") * 2 - 1
        o, r, te, tr, inf = env.step({"action": a, "prediction": p})
        so, sr, ste, _, _ = shard.step({"action": a[lo:hi].contiguous(), "prediction": p[lo:hi].contiguous()})
        assert torch.equal(o["lidar"][lo:hi], so["lidar"]) and torch.equal(r[lo:hi], sr) and torch.equal(te[lo:hi], ste)
        assert bool(te.all()) == (s == 99) and not bool(tr.any())
        assert float(o["lidar"].min()) >= 0.0 and float(o["lidar"].max()) <= 1.0
        assert float(o["odometry"].abs().max()) <= 1.0
        st = inf["_base_reward"]
        assert bool(st.all()) == (s != 100)
        assert torch.equal(r[st], (inf["base_reward"] - inf["prediction"]["loss"])[st])
        if s == 100:
            assert float(r.abs().max()) == 0.0 and float(o["time_step"].max()) == -1.0
    env.close()
    shard.close()


def test_host_step_chunks_and_mapped_outputs():
    """apg_lidar_step_host with page-locked result buffers (written by the kernel, chunked pipeline) against the same
    call with pageable buffers (device-to-host copies)."""
    import ctypes as C
    import torch
    import ap_gym_b200 as apg
    from ap_gym_b200 import _capi

    maps = load_bank("rooms32", 64)
    N, B = 70000, 16
    lib = _capi.lib()
    rng = np.random.default_rng(1)
    acts = rng.uniform(-1, 1, (6, N, 2)).astype(np.float32)
    preds = rng.uniform(-1, 1, (6, N, 2)).astype(np.float32)
    results = []
    for pinned in (False, True):
        env = apg.LIDARLocalization2DVectorEnv(N, apg.FloorMapBank.from_maps(maps), lidar_beam_count=B)
        env.reset(seed=3)
        alloc = (lambda *sh, dt: torch.empty(sh, dtype=dt).pin_memory()) if pinned else (lambda *sh, dt: torch.empty(sh, dtype=dt))
        r, te, tr, li = alloc(N, dt=torch.float32), alloc(N, dt=torch.uint8), alloc(N, dt=torch.uint8), alloc(N, B, dt=torch.float32)
        out = []
        for s in range(6):
            _capi.check(lib.apg_lidar_step_host(env._h, acts[s].ctypes.data_as(C.c_void_p), preds[s].ctypes.data_as(C.c_void_p),
                                                r.data_ptr(), te.data_ptr(), tr.data_ptr(), li.data_ptr()))
            out.append((r.clone().numpy(), te.clone().numpy(), li.clone().numpy()))
        results.append(out)
        env.close()
    for s in range(6):
        for x, y in zip(results[0][s], results[1][s]):
            assert np.array_equal(x, y), f"step {s}"


@pytest.mark.parametrize("kind,B,N,steps", [("rooms", 64, 384, 208), ("maze", 12, 256, 104)])
def test_procedural_dataset_mode_matches_oracle(kind, B, N, steps):
    """Procedural-dataset mode (apg_lidar_set_map_generator): every (auto)reset draws dataset[integers(0, 2**32)] from the
    env's DatasetIterator stream and generates the map on the device.  Against the oracle in dataset mode (pinned to the
    live reference's LIDARLocRooms-v0 / LIDARLocMaze-v0 in tests/test_oracle_lidar.py): dataset indices, map observations,
    positions, scans, rewards across one or two autoresets; a shard with env_offset reproduces its slice."""
    import ap_gym_b200 as apg
    from oracle.floor_map_oracle import maze, rooms
    from oracle.lidar_oracle import LidarVecOracle

    if kind == "rooms":
        ds, fn, H, W = apg.FloorMapDatasetRooms(), (lambda i: rooms(i, 32, 32)), 32, 32
    else:
        ds, fn, H, W = apg.FloorMapDatasetMaze(), (lambda i: maze(i, 21, 21)), 21, 21
    oracle = LidarVecOracle(None, N, beam_count=B, dataset_fn=fn, map_shape=(H, W))
    env = apg.LIDARLocalization2DVectorEnv(N, ds, lidar_beam_count=B, emit_cells=True, output="numpy", check_errors="step")
    half = N // 2
    shard = apg.LIDARLocalization2DVectorEnv(N - half, ds, lidar_beam_count=B, output="numpy", env_offset=half)
    oo, oinfo = oracle.reset(seed=77)
    eo, einfo = env.reset(seed=77)
    so, sinfo = shard.reset(seed=77)
    assert np.array_equal(einfo["map_idx"], oinfo["map_idx"]) and einfo["map_idx"].max() > 2**31
    assert np.array_equal(sinfo["map_idx"], oinfo["map_idx"][half:])
    assert np.array_equal(eo["map"], oo["map"]) and np.array_equal(eo["lidar"], oo["lidar"])
    rng = np.random.default_rng(3)
    heading = rng.uniform(-np.pi, np.pi, N)
    for s in range(steps):
        heading += rng.normal(0, 0.5, N)
        speed = rng.uniform(0, 1.3, N)
        a = (np.stack([np.cos(heading), np.sin(heading)], -1) * speed[:, None]).astype(np.float32)
        p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        oo, orw, ote, otr, oinf = oracle.step(a, p)
        eo, erw, ete, etr, einf = env.step({"action": a, "prediction": p})
        so, srw, ste, strc, sinf = shard.step({"action": a[half:], "prediction": p[half:]})
        assert np.array_equal(einf["map_idx"], oinf["map_idx"]), f"step {s} dataset indices"
        assert np.array_equal(_np(env.state()["pos"]), oracle.pos), f"step {s} pos"
        keep = _beam_mask(B, ~einf["_base_reward"])
        assert np.array_equal(eo["lidar"][keep], oo["lidar"][keep]), f"step {s} lidar"
        assert np.array_equal(eo["map"], oo["map"]), f"step {s} map observation"
        assert np.array_equal(eo["odometry"], oo["odometry"]) and np.array_equal(eo["time_step"], oo["time_step"])
        assert np.array_equal(ete, ote) and np.array_equal(erw.astype(np.float64), orw)
        assert np.array_equal(so["lidar"], eo["lidar"][half:]) and np.array_equal(srw, erw[half:])
        assert np.array_equal(sinf["map_idx"], einf["map_idx"][half:])
    env.close()
    shard.close()
