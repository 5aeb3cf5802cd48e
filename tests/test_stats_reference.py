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
