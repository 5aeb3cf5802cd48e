"""GPU: the fused episode-statistics accumulators against the reference's LogWrapper formulas
(active_classification_env.py:149-197, active_regression_env.py:141-157,204-227, util.py:18-83), evaluated with
numpy on the golden traces; tests/test_stats_reference.py pins those formulas to the live reference."""
import numpy as np
import pytest

from helpers import IMAGE_CASES, LIDAR_CASES, assert_close, load_bank, load_golden, make_bank
from stats_formulas import classification_episode_stats, regression_episode_stats

pytestmark = pytest.mark.gpu


def _np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


@pytest.mark.parametrize("name", ["circlesquare_n64", "mnistlike_cls_n32", "tinyloc_n4", "mnistloc_n6"])
def test_image_episode_stats(name):
    import ap_gym_b200 as apg

    case = IMAGE_CASES[name]
    g = load_golden("image", name)
    imgs, labels = make_bank(case["bank"])
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, case["K"]), **case["cfg"])
    cls = apg.ImageClassificationVectorEnv if case["task"] == "cls" else apg.ImageLocalizationVectorEnv
    env = cls(case["N"], cfg, output="numpy")
    env.reset(seed=case["seed"], options={"prediction_target": g["reset_target"]} if case["task"] == "loc" else None)
    T = cfg.step_limit
    n_done = 0
    sums = np.zeros(5)
    for s in range(case["steps"]):
        _, r, te, _, inf = env.step({"action": g["actions"][s], "prediction": g["preds"][s]})
        if te.all():
            lo = s - (T - 1)
            fn = classification_episode_stats if case["task"] == "cls" else regression_episode_stats
            want = fn(g["preds"][lo:s + 1], g["target"][lo:s + 1], case["K"])
            sc = inf["stats"]["scalar"]
            for k, v in want.items():
                if k.startswith(("first", "last")):
                    assert np.array_equal(sc[k][sc["_" + k]], v[v >= 0]) and np.array_equal(sc["_" + k], v >= 0), k
                else:
                    assert_close(sc[k], v, 1e-5, 1e-6, f"{name} step {s} {k}")
            ret = g["reward"][lo:s + 1].sum(0)
            assert_close(sc["return"], ret, 1e-5, 1e-5, "episode return")
            n_done += case["N"]
            keys = [k for k in want if not k.startswith(("first", "last"))]
            sums[:4] += [want[k].astype(np.float64).sum() for k in keys]
            sums[4] += ret.sum()
        else:
            assert "stats" not in inf
    st = env.episode_stats(reduce=False)
    assert st["episodes"] == n_done
    vals = [v for k, v in st.items() if k != "episodes"]
    assert_close(np.array(vals), sums / n_done, 1e-5, 1e-6, "shard-wide episode statistics")
    env.close()


@pytest.mark.parametrize("name", ["rooms_b64_n8", "maze21_b12_n3"])
def test_lidar_episode_stats(name):
    import ap_gym_b200 as apg

    case = LIDAR_CASES[name]
    g = load_golden("lidar", name)
    maps = load_bank(case["bank"], case["M"])
    env = apg.LIDARLocalization2DVectorEnv(case["N"], apg.FloorMapBank.from_maps(maps), static_map=case["static"],
                                           lidar_beam_count=case["B"], output="numpy")
    env.reset(seed=case["seed"])
    start = np.zeros(case["N"], int)
    n_done = 0
    for s in range(case["steps"]):
        _, r, te, _, inf = env.step({"action": g["actions"][s], "prediction": g["preds"][s]})
        for i in np.nonzero(te)[0]:
            want = regression_episode_stats(g["preds"][start[i]:s + 1, i:i + 1], g["target"][start[i]:s + 1, i:i + 1], 0)
            sc = inf["stats"]["scalar"]
            for k, v in want.items():
                assert_close(sc[k][i:i + 1], v, 1e-5, 1e-6, f"{name} step {s} env {i} {k}")
            assert_close(sc["return"][i:i + 1], g["reward"][start[i]:s + 1, i].sum(keepdims=True), 1e-5, 1e-5, "return")
            start[i] = s + 2  # the next step is the autoreset step
            n_done += 1
    assert env.episode_stats(reduce=False)["episodes"] == n_done
    env.close()


# ---------------------------------------------------------------------------------------------------------------------
# `-sparse` ids (SparsifyVectorWrapper / SparsifyWrapper, sparsify_wrapper.py:60-88, 137-150):
# reward = base_reward - loss * weight, weight = float32(terminated); everything else as in the dense env.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("task", ["cls", "loc"])
def test_sparse_image_env_matches_dense_run(task):
    import ap_gym_b200 as apg
    from oracle.datasets import synthetic_u8_bank

    N, K, T = 193, 7, 5
    imgs, labels = synthetic_u8_bank(seed=9, M=200, H=28, W=28, C=1, K=K)
    cls = apg.ImageClassificationVectorEnv if task == "cls" else apg.ImageLocalizationVectorEnv
    kw = dict(target_sampling="uniform") if task == "loc" else {}
    envs = []
    for sparse in (False, True):
        cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, K), step_limit=T)
        env = cls(N, cfg, output="numpy", sparse=sparse, **kw)
        env.reset(seed=4)
        envs.append(env)
    rng = np.random.default_rng(2)
    saw_terminal = 0
    for s in range(3 * (T + 1)):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        p = rng.standard_normal((N, K if task == "cls" else 2)).astype(np.float32)
        od, rd, ted, _, infd = envs[0].step({"action": a, "prediction": p})
        os_, rs, tes, _, infs = envs[1].step({"action": a, "prediction": p})
        for k in od:
            assert np.array_equal(od[k], os_[k])
        assert np.array_equal(ted, tes)
        w = tes.astype(np.float32)
        assert np.array_equal(infs["prediction"]["target"]["weight"], w)
        assert np.array_equal(infs["prediction"]["target"]["target"], infd["prediction"]["target"])
        assert np.array_equal(infs["prediction"]["loss"], infd["prediction"]["loss"])
        base, loss = infd["base_reward"], infd["prediction"]["loss"]
        if task == "cls":  # float32 base - float64(loss) * float32 weight -> float64 (numpy promotion), emitted as f32
            want = (base.astype(np.float64) - loss.astype(np.float64) * w).astype(np.float32)
            assert np.allclose(rs, want, rtol=1e-6, atol=1e-7)  # loss itself is only available rounded to float32
        else:
            assert np.array_equal(rs, base - loss * w)
        assert np.array_equal(rs[tes], rd[tes])  # the terminating step charges the full loss
        assert np.array_equal(rs[~tes], base[~tes])
        saw_terminal += int(tes.any())
    assert saw_terminal == 3
    for e in envs:
        e.close()


def test_sparse_lidar_env_matches_dense_run():
    import ap_gym_b200 as apg
    bank = apg.FloorMapBank.from_maps(load_bank("rooms32", 64))
    N = 300
    envs = [apg.LIDARLocalization2DVectorEnv(N, bank, lidar_beam_count=8, max_episode_steps=7, output="numpy",
                                             sparse=sp) for sp in (False, True)]
    for e in envs:
        e.reset(seed=11)
    rng = np.random.default_rng(5)
    n_term = 0
    for s in range(30):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        od, rd, ted, _, infd = envs[0].step({"action": a, "prediction": p})
        os_, rs, tes, _, infs = envs[1].step({"action": a, "prediction": p})
        for k in od:
            assert np.array_equal(od[k], os_[k])
        assert np.array_equal(ted, tes)
        stepped = infd["_base_reward"]
        w = tes.astype(np.float32)
        assert np.array_equal(infs["prediction"]["target"]["weight"], w)
        base, loss = infd["base_reward"], infd["prediction"]["loss"]
        assert np.array_equal(rs[stepped], (base - loss * w)[stepped])
        assert np.array_equal(rs[~stepped], rd[~stepped])  # autoreset steps: reward 0 in both
        n_term += int(tes.sum())
    assert n_term > N
    for e in envs:
        e.close()
