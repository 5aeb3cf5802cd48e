"""CPU, needs /root/reference: pins tests/stats_formulas.py to the live reference LogWrappers."""
import os
import sys

import numpy as np
import pytest

from stats_formulas import classification_episode_stats, regression_episode_stats

pytestmark = pytest.mark.reference


def _ref():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "oracle", "shim"), "/root/reference"):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ap_gym

    return ap_gym


def test_classification_formula_matches_vector_log_wrapper():
    ap_gym = _ref()
    env = ap_gym.make_vec("CircleSquare-v0", num_envs=6)
    env.reset(seed=4)
    rng = np.random.default_rng(0)
    preds, targets = [], []
    for s in range(16):
        p = (rng.standard_normal((6, 2)) * 2).astype(np.float32)
        _, _, te, _, info = env.step({"action": rng.uniform(-1, 1, (6, 2)).astype(np.float32), "prediction": p})
        preds.append(p)
        targets.append(info["prediction"]["target"].copy())
    assert te.all()
    want = classification_episode_stats(np.stack(preds), np.stack(targets), 2)
    sc = info["stats"]["scalar"]
    for k in ("avg_correct_label_prob", "final_correct_label_prob", "avg_accuracy", "final_accuracy"):
        np.testing.assert_allclose(sc[k], want[k], rtol=1e-6, atol=1e-7)
    for k in ("first_correct", "last_incorrect"):
        assert np.array_equal(sc["_" + k], want[k] >= 0)
        assert np.array_equal(sc[k][sc["_" + k]], want[k][want[k] >= 0])
    env.close()


def test_regression_formula_matches_single_env_log_wrapper():
    ap_gym = _ref()
    env = ap_gym.make_vec("LIDARLocRooms-v0", num_envs=2)
    env.reset(seed=1)
    rng = np.random.default_rng(1)
    preds, targets = [], []
    for s in range(100):
        p = rng.uniform(-1, 1, (2, 2)).astype(np.float32)
        _, _, te, _, info = env.step({"action": rng.uniform(-1, 1, (2, 2)).astype(np.float32), "prediction": p})
        preds.append(p)
        targets.append(info["prediction"]["target"].copy())
    assert te.all()
    want = regression_episode_stats(np.stack(preds), np.stack(targets))
    sc = info["stats"]["scalar"]
    for k, v in want.items():
        np.testing.assert_allclose(np.asarray(sc[k], dtype=np.float64), v, rtol=1e-5, atol=1e-6)
    env.close()


def _sparse_formula(base, loss, terminated):
    """The rule the engine implements for `<id>-sparse` and tests/test_stats_gpu.py checks on the GPU."""
    return base - loss * terminated.astype(np.float32)


def test_reference_lidar_sparse_ids_cannot_reset():
    """The reference's own LIDAR `-sparse` registrations are unusable: SparsifyWrapper.reset adds the weight to
    info["prediction"], which LIDARLocalization2DEnv.reset does not emit (sparsify_wrapper.py:126-135, 152-160).  The
    engine's `LIDARLocalization2DVectorEnv(sparse=True)` therefore follows the wrapper's STEP rule (the same one the
    image ids pin below) and has no reference run to be compared with."""
    ap_gym = _ref()
    env = ap_gym.make_vec("LIDARLocRooms-sparse-v0", num_envs=2)
    with pytest.raises(KeyError):
        env.reset(seed=0)
    env.close()


@pytest.mark.parametrize("env_id,pw", [("CircleSquare-v0", 2), ("CircleSquareInverted-v0", 2)])
def test_sparse_reward_formula_matches_the_reference_sparse_ids(env_id, pw):
    """Dense and `-sparse` registrations of the same id, same seed, same inputs: identical observations, flags, losses
    and base rewards; sparse reward == base_reward - loss * float32(terminated); target becomes {target, weight}."""
    ap_gym = _ref()
    n = 5
    dense = ap_gym.make_vec(env_id, num_envs=n)
    name, ver = env_id.rsplit("-", 1)
    sparse = ap_gym.make_vec(f"{name}-sparse-{ver}", num_envs=n)
    od, _ = dense.reset(seed=7)
    os_, infs = sparse.reset(seed=7)
    for k in od:
        assert np.array_equal(od[k], os_[k])
    rng = np.random.default_rng(3)
    terminals = 0
    for s in range(40 if "Circle" in env_id else 210):
        act = {"action": rng.uniform(-1, 1, (n, 2)).astype(np.float32),
               "prediction": rng.uniform(-1, 1, (n, pw)).astype(np.float32)}
        od, rd, ted, trd, infd = dense.step(act)
        os_, rs, tes, trs, infs = sparse.step(act)
        for k in od:
            assert np.array_equal(od[k], os_[k]), k
        assert np.array_equal(ted, tes) and np.array_equal(trd, trs)
        stepped = infd.get("_base_reward", np.ones(n, bool))  # SyncVectorEnv autoreset steps carry no step info
        if not stepped.any():
            continue
        base = np.asarray(infd["base_reward"])
        loss = np.asarray(infd["prediction"]["loss"])
        tgt = infs["prediction"]["target"]
        assert np.array_equal(np.asarray(tgt["weight"], dtype=np.float32)[stepped], tes.astype(np.float32)[stepped])
        assert np.array_equal(np.asarray(tgt["target"])[stepped], np.asarray(infd["prediction"]["target"])[stepped])
        assert np.array_equal(np.asarray(infs["prediction"]["loss"])[stepped], loss[stepped])
        want = _sparse_formula(base, loss, tes)
        np.testing.assert_allclose(np.asarray(rs, dtype=np.float64)[stepped], np.asarray(want, dtype=np.float64)[stepped],
                                   rtol=1e-6, atol=1e-7)
        assert np.array_equal(np.asarray(rs)[stepped & tes], np.asarray(rd)[stepped & tes])
        terminals += int(tes.sum())
    assert terminals >= n
    dense.close()
    sparse.close()
