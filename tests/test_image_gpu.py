"""GPU parity tests of the fused image step (through the C-ABI / ctypes layer) against
(a) the golden vectors recorded from the unmodified reference and (b) the numpy oracle on seeded inputs.

Bars (BASELINE.json north_star): bit-exact glimpse pixel coordinates (tap indices, checked through the visited
bitmask and the overlay stamps), positions, flags, targets, indices; the float64-blended glimpse pixels are
compared bit-exactly too (same IEEE op order, no FMA); losses/rewards within rtol 1e-5 (+1e-6 abs) because
numpy's SIMD exp/log are not correctly rounded."""
import numpy as np
import pytest

from helpers import IMAGE_CASES, assert_close, image_oracle_for, load_golden, make_bank

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


def _engine_for(case, imgs, labels, **kw):
    import ap_gym_b200 as apg

    ds = apg.ArrayImageDataset(imgs, labels, case["K"])
    cfg = apg.ImagePerceptionConfig(dataset=ds, **case["cfg"])
    cls = apg.ImageClassificationVectorEnv if case["task"] == "cls" else apg.ImageLocalizationVectorEnv
    return cls(case["N"], cfg, **kw)


def _np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


BLENDS = ["float64", "float32"]


def _assert_obs(got, want, key, blend, msg):
    """blend="float64" (default engine mode): everything bit-exact.  blend="float32" (opt-in fast mode): the resampled
    pixels must meet the north-star bar for float pixels -- 1e-5 RELATIVE, no absolute slack -- everything else
    (glimpse position, time step, labels) stays bit-exact."""
    got, want = _np(got), np.asarray(want)
    if blend == "float32" and key in ("glimpse", "target_glimpse"):
        assert got.dtype == want.dtype and got.shape == want.shape, msg
        assert np.all(np.abs(got.astype(np.float64) - want) <= 1e-5 * np.abs(want)), msg
    else:
        assert np.array_equal(got, want), msg


@pytest.mark.parametrize("blend", BLENDS)
@pytest.mark.parametrize("mode", ["numpy", "torch"])
@pytest.mark.parametrize("name", list(IMAGE_CASES))
def test_engine_matches_reference_golden(name, mode, blend):
    import torch

    case = IMAGE_CASES[name]
    g = load_golden("image", name)
    imgs, labels = make_bank(case["bank"])
    env = _engine_for(case, imgs, labels, check_errors="step", blend=blend)
    opts = {"prediction_target": g["reset_target"]} if case["task"] == "loc" else None
    obs, info = env.reset(seed=case["seed"], options=opts)
    for k in obs:
        _assert_obs(obs[k], g["reset_" + k], k, blend, f"reset obs {k}")
    assert np.array_equal(_np(info["index"]), g["reset_index"])
    for s in range(case["steps"]):
        a, p = g["actions"][s], g["preds"][s]
        if mode == "torch":
            a, p = torch.from_numpy(a).cuda(), torch.from_numpy(p).cuda()
        o, r, te, tr, inf = env.step({"action": a, "prediction": p})
        for k in o:
            _assert_obs(o[k], g["obs_" + k][s], k, blend, f"step {s} obs {k}")
        assert np.array_equal(_np(env.state()["pos"]), g["pos64"][s]), f"step {s} float64 position"
        assert np.array_equal(_np(te), g["terminated"][s]) and np.array_equal(_np(tr), g["truncated"][s])
        assert np.array_equal(_np(inf["index"]), g["index"][s])
        assert np.array_equal(_np(inf["prediction"]["target"]), g["target"][s])
        assert np.array_equal(_np(inf["base_reward"]).astype(np.float64), g["base_reward"][s])
        assert_close(_np(inf["prediction"]["loss"]), g["loss"][s], RTOL, ATOL, f"step {s} loss")
        assert_close(_np(r), g["reward"][s], RTOL, ATOL, f"step {s} reward")
        counts, quality = env.overlay_maps([0, 1])
        assert np.array_equal(counts.astype(np.uint8), g["counts"][s]), f"step {s} overlay counts"
        assert_close(quality, g["quality"][s], RTOL, ATOL, f"step {s} overlay quality")
    env.close()


@pytest.mark.parametrize("name", ["mnistlike_cls_n32", "tinyloc_n4", "scaled_cls_n8"])
def test_engine_matches_oracle_visited_mask_and_taps(name):
    case = IMAGE_CASES[name]
    g = load_golden("image", name)
    oracle, imgs, labels = image_oracle_for(case)
    env = _engine_for(case, imgs, labels)
    tgt = g["reset_target"] if case["task"] == "loc" else None
    oracle.reset(seed=case["seed"], explicit_targets=tgt)
    env.reset(seed=case["seed"], options={"prediction_target": tgt} if tgt is not None else None)
    assert np.array_equal(_np(env.state()["visited"]).view(np.uint32), oracle.visited)
    for s in range(case["steps"]):
        oracle.step(g["actions"][s], g["preds"][s])
        env.step({"action": g["actions"][s], "prediction": g["preds"][s]})
        assert np.array_equal(_np(env.state()["visited"]).view(np.uint32), oracle.visited), f"step {s} visited mask"
    env.close()


def test_engine_matches_oracle_seeded_mnist_2048():
    """MNIST-shaped bank, N=2048, 40 steps (2 autoresets) of seeded random actions vs the oracle."""
    from oracle.datasets import synthetic_u8_bank
    from oracle.image_oracle import ImageEnvOracle
    import ap_gym_b200 as apg

    N, K = 2048, 10
    imgs, labels = synthetic_u8_bank(seed=7, M=4096, H=28, W=28, C=1, K=K)
    oracle = ImageEnvOracle(imgs, labels, N, task="cls", num_classes=K, track_overlay=False)
    env = apg.ImageClassificationVectorEnv(
        N, apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, K)), output="numpy")
    o0, _ = oracle.reset(seed=11)
    e0, _ = env.reset(seed=11)
    assert np.array_equal(e0["glimpse"], o0["glimpse"])
    rng = np.random.default_rng(5)
    for s in range(40):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32) * np.float32(1.3)
        p = (rng.standard_normal((N, K)) * 4).astype(np.float32)
        oo, orw, ote, otr, oinf = oracle.step(a, p)
        eo, erw, ete, etr, einf = env.step({"action": a, "prediction": p})
        for k in oo:
            assert np.array_equal(eo[k], oo[k]), f"step {s} obs {k}"
        assert np.array_equal(ete, ote) and np.array_equal(etr, otr)
        assert np.array_equal(einf["prediction"]["target"], oinf["prediction"]["target"])
        assert np.array_equal(einf["index"], oinf["index"])
        assert_close(erw, orw, RTOL, ATOL, f"step {s} reward")
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("N,K", [(192, 4), (200, 4), (200, 10), (64, 7)])
def test_rollout_kernel_instantiations_match_oracle(N, K):
    """The four instantiations of the thread-per-environment kernel: class count 10 / runtime, whole warps only
    (N % 64 == 0: unrolled cooperative copies) / ragged batch -- 36 steps (2 autoresets) of seeded inputs vs the oracle,
    through step() and through one rollout window."""
    import torch

    from oracle.datasets import synthetic_u8_bank
    from oracle.image_oracle import ImageEnvOracle
    import ap_gym_b200 as apg

    imgs, labels = synthetic_u8_bank(seed=3, M=512, H=28, W=28, C=1, K=K)
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, K))
    oracle = ImageEnvOracle(imgs, labels, N, task="cls", num_classes=K, track_overlay=False)
    env = apg.ImageClassificationVectorEnv(N, cfg, output="numpy")
    env2 = apg.ImageClassificationVectorEnv(N, cfg)
    o0, _ = oracle.reset(seed=21)
    e0, _ = env.reset(seed=21)
    env2.reset(seed=21)
    assert np.array_equal(e0["glimpse"], o0["glimpse"])
    rng = np.random.default_rng(9)
    S = 36
    acts = (rng.uniform(-1, 1, (S, N, 2)) * 1.3).astype(np.float32)
    preds = (rng.standard_normal((S, N, K)) * 4).astype(np.float32)
    rew = np.zeros((S, N), np.float32)
    for s in range(S):
        oo, orw, ote, otr, oinf = oracle.step(acts[s], preds[s])
        eo, erw, ete, etr, einf = env.step({"action": acts[s], "prediction": preds[s]})
        for k in oo:
            assert np.array_equal(eo[k], oo[k]), f"step {s} obs {k}"
        assert np.array_equal(ete, ote) and np.array_equal(einf["prediction"]["target"], oinf["prediction"]["target"])
        assert_close(erw, orw, RTOL, ATOL, f"step {s} reward")
        rew[s] = erw
    r2 = torch.empty((S, N), dtype=torch.float32, device="cuda")
    d2 = torch.empty((S, N), dtype=torch.uint8, device="cuda")
    env2.rollout(torch.from_numpy(acts).cuda(), torch.from_numpy(preds).cuda(), rewards=r2, dones=d2)
    assert np.array_equal(r2.cpu().numpy(), rew)
    o1 = env.step({"action": acts[0], "prediction": preds[0]})[0]
    o2 = env2.step({"action": torch.from_numpy(acts[0]).cuda(), "prediction": torch.from_numpy(preds[0]).cuda()})[0]
    for k in o1:
        assert np.array_equal(o2[k].cpu().numpy(), o1[k]), k
    env.close()
    env2.close()


@pytest.mark.parametrize("geom", [
    dict(H=64, W=64, C=1, sensor_size=(36, 36), sensor_scale=1.0),   # more sensor rows than lanes per env, 2 chunks
    dict(H=40, W=48, C=3, sensor_size=(14, 14), sensor_scale=1.0),   # wide sensor: 42 floats per row, two chunks
    dict(H=70, W=72, C=3, sensor_size=(10, 10), sensor_scale=1.0),   # image wider than 64 px: 3 mask words
    dict(H=33, W=35, C=1, sensor_size=(5, 5), sensor_scale=2.5),     # scaled sensor: non-consecutive tap rows
    dict(H=28, W=28, C=3, sensor_size=(5, 5), sensor_scale=1.0),     # unrolled 5x5x3 path
    dict(H=40, W=40, C=1, sensor_size=(10, 10), sensor_scale=1.0),   # unrolled 10x10x1 path
])
@pytest.mark.parametrize("blend", BLENDS)
@pytest.mark.parametrize("task", ["cls", "loc"])
def test_engine_matches_oracle_odd_geometries(geom, task, blend):
    """Sensor / image geometries that exercise every code path of the kernel (generic row loop, column chunks,
    wide visited masks, clamped and scaled taps) against the oracle: bit-exact pixels, masks, positions."""
    from oracle.datasets import synthetic_u8_bank
    from oracle.image_oracle import ImageEnvOracle
    import ap_gym_b200 as apg

    N, K = 96, 5
    imgs, labels = synthetic_u8_bank(seed=21, M=64, H=geom["H"], W=geom["W"], C=geom["C"], K=K)
    kw = dict(sensor_size=geom["sensor_size"], sensor_scale=geom["sensor_scale"], step_limit=6, max_step_length=0.5)
    oracle = ImageEnvOracle(imgs, labels, N, task=task, num_classes=K, track_overlay=False, **kw)
    cls = apg.ImageClassificationVectorEnv if task == "cls" else apg.ImageLocalizationVectorEnv
    env = cls(N, apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, K), **kw), output="numpy",
              blend=blend)
    rng = np.random.default_rng(4)
    tgt = rng.uniform(-1, 1, (N, 2)).astype(np.float32) if task == "loc" else None
    oo, _ = oracle.reset(seed=2, explicit_targets=tgt)
    eo, _ = env.reset(seed=2, options={"prediction_target": tgt} if task == "loc" else None)
    for k in oo:
        _assert_obs(eo[k], oo[k], k, blend, f"reset obs {k}")
    for s in range(16):
        a = (rng.uniform(-1, 1, (N, 2)) * 1.5).astype(np.float32)
        p = rng.standard_normal((N, K if task == "cls" else 2)).astype(np.float32)
        oo, orw, ote, _, oinf = oracle.step(a, p)
        eo, erw, ete, _, einf = env.step({"action": a, "prediction": p})
        for k in oo:
            _assert_obs(eo[k], oo[k], k, blend, f"step {s} obs {k}")
        assert np.array_equal(_np(env.state()["pos"]), oracle.pos)
        assert np.array_equal(_np(env.state()["visited"]).view(np.uint32), oracle.visited), f"step {s} visited"
        assert np.array_equal(ete, ote)
        assert_close(erw, orw, RTOL, ATOL, f"step {s} reward")
    env.close()


@pytest.mark.parametrize("name", ["tinyloc_n4", "mnistloc_n6"])
def test_unique_glimpse_targets_match_reference(name):
    """Explicit reset() of the localisation env: the GPU uniqueness search + the module Generator's pick reproduce
    the reference's sample_unique_glimpse_positions targets (image_perception_module.py:253-292) bit-exactly."""
    import ap_gym_b200 as apg

    case = IMAGE_CASES[name]
    g = load_golden("image", name)
    imgs, labels = make_bank(case["bank"])
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, case["K"]), **case["cfg"])
    env = apg.ImageLocalizationVectorEnv(case["N"], cfg, output="numpy")
    obs, info = env.reset(seed=case["seed"])
    assert np.array_equal(_np(env.state()["cur_target_xy"]), g["reset_target"]), "unique-glimpse targets"
    assert np.array_equal(obs["target_glimpse"], g["reset_target_glimpse"])
    for s in range(20):  # crosses an autoreset: later episodes stay in lock-step with the reference
        o, r, te, tr, inf = env.step({"action": g["actions"][s], "prediction": g["preds"][s]})
        assert np.array_equal(o["target_glimpse"], g["obs_target_glimpse"][s])
        assert np.array_equal(inf["prediction"]["target"], g["target"][s])
    env.close()


@pytest.mark.parametrize("flag", ["APG_IMG_GENERIC", "APG_IMG_FORCE_SLOW"])
def test_tma_kernel_matches_staged_kernel_and_its_general_path(monkeypatch, flag):
    """TinyImageNetLoc geometry (64x64x3 bank, 10x10 sensor): the TMA-tile kernel (default, variant 3) against (a) the
    generic staged-tile kernel (APG_IMG_GENERIC=1) and (b) its own general per-output branch, which real inputs reach
    with probability ~1e-15 (APG_IMG_FORCE_SLOW=1 sends every environment through it): bit-identical observations,
    rewards, visited masks and statistics over two episodes, target glimpse (second set) included."""
    import ap_gym_b200 as apg
    from oracle.datasets import synthetic_u8_bank

    imgs, labels = synthetic_u8_bank(seed=9, M=64, H=64, W=64, C=3, K=5)
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, 5), sensor_size=(10, 10), step_limit=6)
    N = 1000
    rng = np.random.default_rng(1)
    tgt = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
    envs = []
    for v in ("0", "1"):
        monkeypatch.setenv(flag, v)
        envs.append(apg.ImageLocalizationVectorEnv(N, cfg, output="numpy"))
    monkeypatch.delenv(flag)
    assert envs[0]._lib.apg_image_variant(envs[0]._h) == 3
    assert envs[1]._lib.apg_image_variant(envs[1]._h) == (0 if flag == "APG_IMG_GENERIC" else 3)
    obs = [e.reset(seed=4, options={"prediction_target": tgt})[0] for e in envs]
    for k in obs[0]:
        assert np.array_equal(obs[0][k], obs[1][k]), k
    for s in range(15):
        a = (rng.uniform(-1, 1, (N, 2)) * 1.5).astype(np.float32)  # saturates at the image border regularly
        p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        out = [e.step({"action": a, "prediction": p}) for e in envs]
        for k in out[0][0]:
            assert np.array_equal(out[0][0][k], out[1][0][k]), f"step {s} obs {k}"
        assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
        assert np.array_equal(_np(envs[0].state()["visited"]), _np(envs[1].state()["visited"]))
    assert envs[0].episode_stats(reduce=False) == envs[1].episode_stats(reduce=False)
    for e in envs:
        e.close()


def test_tinyloc_benchmark_shape_matches_oracle_n4096():
    """The BENCHMARKED TinyImageNetLoc kernel (uint8 64x64x3 bank, 10x10x3 sensor, T=16) against the oracle at N=4096
    over 38 steps = two autoresets (both glimpse sets, the head kernel, persistent warps with several environments
    each), plus a shard of the same global batch (env_offset) that must reproduce its slice."""
    from oracle.datasets import synthetic_u8_bank
    from oracle.image_oracle import ImageEnvOracle
    import ap_gym_b200 as apg

    N, K, M = 4096, 200, 512
    imgs, labels = synthetic_u8_bank(seed=31, M=M, H=64, W=64, C=3, K=K)
    kw = dict(sensor_size=(10, 10))
    oracle = ImageEnvOracle(imgs, labels, N, task="loc", num_classes=K, track_overlay=False, **kw)
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, K), **kw)
    env = apg.ImageLocalizationVectorEnv(N, cfg, output="numpy")
    assert env._lib.apg_image_variant(env._h) == 3
    lo, hi = 1000, 1000 + 1111
    shard = apg.ImageLocalizationVectorEnv(hi - lo, cfg, output="numpy", env_offset=lo, global_num_envs=N)
    rng = np.random.default_rng(8)
    tgt = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
    oo, _ = oracle.reset(seed=12, explicit_targets=tgt)
    eo, _ = env.reset(seed=12, options={"prediction_target": tgt})
    so, _ = shard.reset(seed=12, options={"prediction_target": tgt[lo:hi]})
    for k in oo:
        assert np.array_equal(eo[k], oo[k]), f"reset obs {k}"
        assert np.array_equal(so[k], oo[k][lo:hi]), f"reset obs {k} (shard)"
    for s in range(38):
        a = (rng.uniform(-1, 1, (N, 2)) * 1.4).astype(np.float32)
        p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        oo, orw, ote, otr, oinf = oracle.step(a, p)
        eo, erw, ete, etr, einf = env.step({"action": a, "prediction": p})
        so, srw, ste, _, sinf = shard.step({"action": a[lo:hi], "prediction": p[lo:hi]})
        for k in oo:
            assert np.array_equal(eo[k], oo[k]), f"step {s} obs {k}"
            assert np.array_equal(so[k], oo[k][lo:hi]), f"step {s} obs {k} (shard)"
        assert np.array_equal(_np(env.state()["pos"]), oracle.pos)
        assert np.array_equal(_np(env.state()["visited"]).view(np.uint32), oracle.visited), f"step {s} visited"
        assert np.array_equal(ete, ote) and np.array_equal(etr, otr) and np.array_equal(ste, ote[lo:hi])
        assert np.array_equal(einf["prediction"]["target"], oinf["prediction"]["target"])
        assert np.array_equal(einf["index"], oinf["index"]) and np.array_equal(sinf["index"], oinf["index"][lo:hi])
        assert np.array_equal(erw, orw.astype(np.float32)) or np.allclose(erw, orw, rtol=RTOL, atol=ATOL), f"step {s} reward"
        assert np.array_equal(srw, erw[lo:hi])
    env.close()
    shard.close()


def test_error_conventions():
    import ap_gym_b200 as apg
    from oracle.datasets import synthetic_u8_bank

    imgs, labels = synthetic_u8_bank(seed=1, M=16, H=28, W=28, C=1, K=10)
    ds = apg.ArrayImageDataset(imgs, labels, 10)
    env = apg.ImageClassificationVectorEnv(4, apg.ImagePerceptionConfig(dataset=ds), check_errors="step")
    with pytest.raises(Exception):
        env.step({"action": np.zeros((4, 2), np.float32), "prediction": np.zeros((4, 10), np.float32)})  # before reset
    env.reset(seed=0)
    a = np.zeros((4, 2), np.float32)
    p = np.zeros((4, 10), np.float32)
    a[1, 0] = np.nan
    with pytest.raises(ValueError, match="NaN values detected in action"):
        env.step({"action": a, "prediction": p})
    a[1, 0] = 0
    p[2, 3] = np.nan
    with pytest.raises(ValueError, match="NaN values detected in prediction"):
        env.step({"action": a, "prediction": p})
    with pytest.raises(ValueError, match="Unsupported render mode"):
        apg.ImageClassificationVectorEnv(4, apg.ImagePerceptionConfig(dataset=ds), render_mode="human")
    small = apg.ArrayImageDataset(np.zeros((2, 4, 4, 1), np.uint8), np.zeros(2, np.int32), 2)
    with pytest.raises(ValueError, match="cannot be smaller than effective sensor size"):
        apg.ImageClassificationVectorEnv(2, apg.ImagePerceptionConfig(dataset=small))
    env.close()


def test_full_size_properties_mnist_65536():
    """BASELINE config 2 size: properties that need no oracle run -- determinism, shard invariance, the
    17-step autoreset cycle, zero-action idempotence, reward identity, value ranges."""
    import torch
    import ap_gym_b200 as apg

    N, K, T = 65536, 10, 16
    ds = apg.SyntheticImageDataset("mnist", length=8192, seed=0)
    cfg = apg.ImagePerceptionConfig(dataset=ds)
    env = apg.ImageClassificationVectorEnv(N, cfg)
    lo, hi = 1000, 1000 + 4096
    shard = apg.ImageClassificationVectorEnv(hi - lo, cfg, env_offset=lo, global_num_envs=N)
    gen = torch.Generator(device="cuda").manual_seed(0)
    o, _ = env.reset(seed=3)
    so, _ = shard.reset(seed=3)
    assert torch.equal(o["glimpse"][lo:hi], so["glimpse"])
    prev_glimpse = o["glimpse"].clone()
    for s in range(2 * (T + 1)):
        a = torch.rand((N, 2), generator=gen, device="cuda") * 2 - 1
        if s % 5 == 2:
            a.zero_()
        p = torch.randn((N, K), generator=gen, device="cuda")
        o, r, te, tr, inf = env.step({"action": a, "prediction": p})
        so, sr, ste, _, sinf = shard.step({"action": a[lo:hi].contiguous(), "prediction": p[lo:hi].contiguous()})
        cyc = s % (T + 1)
        assert bool(te.all()) == (cyc == T - 1) and not bool(tr.any())
        assert torch.equal(o["glimpse"][lo:hi], so["glimpse"]) and torch.equal(r[lo:hi], sr)
        assert torch.equal(inf["index"][lo:hi], sinf["index"])
        assert float(o["glimpse"].min()) >= 0.0 and float(o["glimpse"].max()) <= 1.0
        assert float(o["glimpse_pos"].abs().max()) <= 1.0
        if s % 5 == 2 and cyc != T:
            assert torch.equal(o["glimpse"], prev_glimpse), "zero action must not move the sensor"
        assert torch.allclose(r, inf["base_reward"] - inf["prediction"]["loss"], rtol=1e-6, atol=1e-6)
        ce = torch.nn.functional.cross_entropy(p, inf["prediction"]["target"].long(), reduction="none") / np.log(K)
        assert torch.allclose(inf["prediction"]["loss"], ce, rtol=1e-5, atol=1e-6)
        prev_glimpse = o["glimpse"].clone()
    env.close()
    shard.close()


def test_rollout_equals_stepping():
    import torch
    import ap_gym_b200 as apg

    N, K, steps = 512, 10, 40
    ds = apg.SyntheticImageDataset("mnist", length=1024, seed=2)
    cfg = apg.ImagePerceptionConfig(dataset=ds)
    gen = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand((steps, N, 2), generator=gen, device="cuda") * 2 - 1
    preds = torch.randn((steps, N, K), generator=gen, device="cuda")
    e1 = apg.ImageClassificationVectorEnv(N, cfg)
    e2 = apg.ImageClassificationVectorEnv(N, cfg)
    e1.reset(seed=9)
    e2.reset(seed=9)
    rew = torch.empty((steps, N), device="cuda")
    done = torch.empty((steps, N), dtype=torch.bool, device="cuda")
    e1.rollout(acts[:20], preds[:20], rew[:20], done[:20])
    e1.rollout(acts[20:], preds[20:], rew[20:], done[20:])
    for s in range(steps):
        o, r, te, tr, _ = e2.step({"action": acts[s], "prediction": preds[s]})
        assert torch.equal(r, rew[s]) and torch.equal(te, done[s])
    e1.sync()
    assert torch.equal(e1.step({"action": acts[0], "prediction": preds[0]})[0]["glimpse"],
                       e2.step({"action": acts[0], "prediction": preds[0]})[0]["glimpse"])
    e1.close()
    e2.close()


def test_rollout_equals_stepping_tinyloc_shape():
    """The TMA kernel steps a whole rollout window inside ONE launch (a warp takes its environments through every step
    before it moves on): windows of 1 / 20 / 30 steps with autoresets inside must reproduce step-by-step execution --
    per-step rewards and done flags, final observations (both glimpse sets), visited masks, statistics."""
    import torch
    import ap_gym_b200 as apg
    from oracle.datasets import synthetic_u8_bank

    N, steps = 3000, 51
    imgs, labels = synthetic_u8_bank(seed=13, M=128, H=64, W=64, C=3, K=5)
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, 5), sensor_size=(10, 10))
    gen = torch.Generator(device="cuda").manual_seed(2)
    acts = (torch.rand((steps, N, 2), generator=gen, device="cuda") * 2 - 1) * 1.3
    preds = torch.rand((steps, N, 2), generator=gen, device="cuda") * 2 - 1
    tgt = np.random.default_rng(0).uniform(-1, 1, (N, 2)).astype(np.float32)
    e1 = apg.ImageLocalizationVectorEnv(N, cfg)
    e2 = apg.ImageLocalizationVectorEnv(N, cfg)
    assert e1._lib.apg_image_variant(e1._h) == 3
    e1.reset(seed=9, options={"prediction_target": tgt})
    e2.reset(seed=9, options={"prediction_target": tgt})
    rew = torch.empty((steps, N), device="cuda")
    done = torch.empty((steps, N), dtype=torch.bool, device="cuda")
    e1.rollout(acts[:1], preds[:1], rew[:1], done[:1])
    e1.rollout(acts[1:21], preds[1:21], rew[1:21], done[1:21])
    e1.rollout(acts[21:], preds[21:], rew[21:], done[21:])
    for s in range(steps):
        o, r, te, tr, _ = e2.step({"action": acts[s], "prediction": preds[s]})
        assert torch.equal(r, rew[s]) and torch.equal(te, done[s]), f"step {s}"
    e1.sync()
    for k in ("glimpse", "target_glimpse", "glimpse_pos", "time_step"):
        assert torch.equal(e1._views[k], e2._views[k]), k
    assert torch.equal(e1.state()["visited"], e2.state()["visited"]) and torch.equal(e1.state()["pos"], e2.state()["pos"])
    assert e1.episode_stats(reduce=False) == pytest.approx(e2.episode_stats(reduce=False), rel=1e-12)
    e1.close()
    e2.close()


def test_mnist_benchmark_shape_matches_oracle_n65536():
    """BASELINE config 2 at FULL size (N = 65536, 28x28x1 uint8 bank, K = 10) against the oracle for 19 steps: 16 real
    steps, the autoreset step and two steps of the next episode, through the persistent rollout kernel (one launch per
    step here).  Bit-exact observations / flags / targets / indices, rewards within the float bar."""
    from oracle.datasets import synthetic_u8_bank
    from oracle.image_oracle import ImageEnvOracle
    import ap_gym_b200 as apg

    N, K = 65536, 10
    imgs, labels = synthetic_u8_bank(seed=3, M=8192, H=28, W=28, C=1, K=K)
    oracle = ImageEnvOracle(imgs, labels, N, task="cls", num_classes=K, track_overlay=False)
    env = apg.ImageClassificationVectorEnv(
        N, apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, K)), output="numpy")
    assert env._lib.apg_image_variant(env._h) == 4
    o0, _ = oracle.reset(seed=21)
    e0, _ = env.reset(seed=21)
    assert np.array_equal(e0["glimpse"], o0["glimpse"])
    rng = np.random.default_rng(6)
    for s in range(19):
        a = (rng.uniform(-1, 1, (N, 2)) * 1.2).astype(np.float32)
        p = (rng.standard_normal((N, K)) * 3).astype(np.float32)
        oo, orw, ote, otr, oinf = oracle.step(a, p)
        eo, erw, ete, etr, einf = env.step({"action": a, "prediction": p})
        for k in oo:
            assert np.array_equal(eo[k], oo[k]), f"step {s} obs {k}"
        assert np.array_equal(ete, ote) and np.array_equal(etr, otr)
        assert np.array_equal(einf["prediction"]["target"], oinf["prediction"]["target"])
        assert np.array_equal(einf["index"], oinf["index"])
        assert_close(erw, orw, RTOL, ATOL, f"step {s} reward")
    assert np.array_equal(_np(env.state()["visited"]).view(np.uint32), oracle.visited)
    env.close()


def test_hypothesis_seeds_and_geometries():
    """Property test: for random seeds, image / sensor geometries, scales, step lengths and tasks the engine
    reproduces the oracle bit-exactly (pixels, positions, masks, flags) -- whichever kernel the geometry selects."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from oracle.datasets import synthetic_u8_bank
    from oracle.image_oracle import ImageEnvOracle
    import ap_gym_b200 as apg

    @settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(seed=st.integers(0, 2**31 - 1), H=st.integers(12, 70), W=st.integers(12, 70), C=st.sampled_from([1, 3]),
           S=st.integers(2, 11), scale=st.sampled_from([1.0, 1.0, 1.25, 2.0]), task=st.sampled_from(["cls", "loc"]),
           msl=st.sampled_from([0.2, 0.5, 1.0]), T=st.integers(2, 6), u8=st.booleans())
    def check(seed, H, W, C, S, scale, task, msl, T, u8):
        # (sensor_scale < 1 lets the reference itself sample outside the image grid at the border -> ValueError: not drawn)
        if S * scale > min(H, W):
            return
        N, K = 37, 4
        imgs, labels = synthetic_u8_bank(seed=seed % 1000, M=16, H=H, W=W, C=C, K=K)
        if not u8:
            imgs = (imgs.astype(np.float32) / np.float32(255)) ** 2  # a float32 bank that is not k/255
        kw = dict(sensor_size=(S, S), sensor_scale=scale, step_limit=T, max_step_length=msl)
        oracle = ImageEnvOracle(imgs, labels, N, task=task, num_classes=K, track_overlay=False, **kw)
        cls = apg.ImageClassificationVectorEnv if task == "cls" else apg.ImageLocalizationVectorEnv
        env = cls(N, apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, K), **kw), output="numpy")
        rng = np.random.default_rng(seed)
        tgt = rng.uniform(-1, 1, (N, 2)).astype(np.float32) if task == "loc" else None
        oo, _ = oracle.reset(seed=seed, explicit_targets=tgt)
        eo, _ = env.reset(seed=seed, options={"prediction_target": tgt} if task == "loc" else None)
        for k in oo:
            assert np.array_equal(eo[k], oo[k]), f"reset obs {k}"
        for s in range(2 * T + 3):
            a = (rng.uniform(-1, 1, (N, 2)) * 1.6).astype(np.float32)
            p = rng.standard_normal((N, K if task == "cls" else 2)).astype(np.float32)
            oo, orw, ote, _, oinf = oracle.step(a, p)
            eo, erw, ete, _, einf = env.step({"action": a, "prediction": p})
            for k in oo:
                assert np.array_equal(eo[k], oo[k]), f"step {s} obs {k}"
            assert np.array_equal(_np(env.state()["pos"]), oracle.pos)
            assert np.array_equal(_np(env.state()["visited"]).view(np.uint32), oracle.visited), f"step {s} visited"
            assert np.array_equal(ete, ote)
            assert_close(erw, orw, RTOL, ATOL, f"step {s} reward")
        env.close()

    check()


def test_sparse_env_exposes_weighted_loss_and_dict_target_spaces():
    """`sparse=True` = SparsifyVectorWrapper (sparsify_wrapper.py:49-88): loss_fn is WeightedLossFn, the target spaces are
    Dict{"target", "weight"} and the reward the kernel computes equals base_reward - loss_fn(prediction, info target)."""
    import ap_gym_b200 as apg
    from oracle.datasets import synthetic_u8_bank

    imgs, labels = synthetic_u8_bank(seed=2, M=64, H=28, W=28, C=1, K=10)
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, 10))
    for cls, pw in ((apg.ImageClassificationVectorEnv, 10), (apg.ImageLocalizationVectorEnv, 2)):
        env = cls(64, cfg, sparse=True, output="numpy")
        assert isinstance(env.loss_fn, apg.WeightedLossFn)
        assert set(env.single_prediction_target_space.spaces) == {"target", "weight"}
        assert set(env.prediction_target_space.spaces) == {"target", "weight"}
        rng = np.random.default_rng(0)
        env.reset(seed=1, options={"prediction_target": rng.uniform(-1, 1, (64, 2)).astype(np.float32)} if pw == 2 else None)
        for s in range(18):
            a = rng.uniform(-1, 1, (64, 2)).astype(np.float32)
            p = rng.standard_normal((64, pw)).astype(np.float32)
            o, r, te, tr, info = env.step({"action": a, "prediction": p})
            tgt = info["prediction"]["target"]
            assert np.array_equal(tgt["weight"], te.astype(np.float32))
            want = info["base_reward"] - env.loss_fn(p, tgt, (64,))
            assert_close(r, want, RTOL, ATOL, f"step {s} sparse reward")
        env.close()


def test_prefetched_payloads_do_not_move_the_reference_streams():
    """Payloads drawn ahead of time advance the engine's copies of the Generators; reading them back (np_random, an
    unseeded reset) drops the prefetched payloads and rewinds, so the env continues exactly like one that never
    prefetched."""
    import ap_gym_b200 as apg
    from oracle.datasets import synthetic_u8_bank

    imgs, labels = synthetic_u8_bank(seed=5, M=100, H=28, W=28, C=1, K=10)
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, 10))
    a_env = apg.ImageClassificationVectorEnv(33, cfg, output="numpy", payload_ring=4)
    b_env = apg.ImageClassificationVectorEnv(33, cfg, output="numpy", payload_ring=4)
    a_env.reset(seed=7)
    b_env.reset(seed=7)
    rng = np.random.default_rng(0)
    for s in range(5):
        act = {"action": rng.uniform(-1, 1, (33, 2)).astype(np.float32), "prediction": rng.standard_normal((33, 10)).astype(np.float32)}
        a_env.step(act)
        b_env.step(act)
    a_env.queue_autoreset_payloads(3)
    st_a, st_b = a_env.np_random.bit_generator.state, b_env.np_random.bit_generator.state
    assert st_a == st_b
    oa, ia = a_env.reset()   # unseeded: the streams continue
    ob, ib = b_env.reset()
    assert np.array_equal(ia["index"], ib["index"]) and np.array_equal(oa["glimpse"], ob["glimpse"])
    for s in range(20):      # crosses an autoreset drawn by the engine
        act = {"action": rng.uniform(-1, 1, (33, 2)).astype(np.float32), "prediction": rng.standard_normal((33, 10)).astype(np.float32)}
        xa, xb = a_env.step(act), b_env.step(act)
        assert np.array_equal(xa[0]["glimpse"], xb[0]["glimpse"]) and np.array_equal(xa[4]["index"], xb[4]["index"])
    a_env.close()
    b_env.close()


@pytest.mark.parametrize("shape", ["mnist", "tinyloc"])
def test_host_step_chunks_and_mapped_outputs(shape, monkeypatch):
    """apg_image_step_host: the chunked pipeline (1 vs 4 chunks) and page-locked vs pageable result buffers give
    identical rewards / flags / observations."""
    import ctypes as C
    import torch
    import ap_gym_b200 as apg
    from ap_gym_b200 import _capi
    from oracle.datasets import synthetic_u8_bank

    N = 40000
    if shape == "mnist":
        imgs, labels = synthetic_u8_bank(seed=5, M=256, H=28, W=28, C=1, K=10)
        mk = lambda: apg.ImageClassificationVectorEnv(N, apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, 10)))
        pw, P = 10, 25
    else:
        imgs, labels = synthetic_u8_bank(seed=5, M=64, H=64, W=64, C=3, K=5)
        mk = lambda: apg.ImageLocalizationVectorEnv(
            N, apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, 5), sensor_size=(10, 10)),
            target_sampling="uniform")
        pw, P = 2, 300
    lib = _capi.lib()
    results = []
    rng = np.random.default_rng(3)
    acts = rng.uniform(-1, 1, (19, N, 2)).astype(np.float32)
    preds = rng.standard_normal((19, N, pw)).astype(np.float32)
    for chunks, pinned in (("1", False), ("4", True)):
        monkeypatch.setenv("APG_HOST_CHUNKS", chunks)
        env = mk()
        env.reset(seed=2)
        alloc = (lambda *sh, dt: torch.empty(sh, dtype=dt).pin_memory()) if pinned else (lambda *sh, dt: torch.empty(sh, dtype=dt))
        r, te, tr, g = alloc(N, dt=torch.float32), alloc(N, dt=torch.uint8), alloc(N, dt=torch.uint8), alloc(N, P, dt=torch.float32)
        out = []
        for s in range(19):
            _capi.check(lib.apg_image_step_host(env._h, acts[s].ctypes.data_as(C.c_void_p), preds[s].ctypes.data_as(C.c_void_p),
                                                r.data_ptr(), te.data_ptr(), tr.data_ptr(), g.data_ptr()))
            out.append((r.clone().numpy(), te.clone().numpy(), tr.clone().numpy(), g.clone().numpy()))
        results.append(out)
        env.close()
    for s in range(19):
        for x, y in zip(results[0][s], results[1][s]):
            assert np.array_equal(x, y), f"step {s}"
