"""CPU: the image oracle (oracle/image_oracle.py) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py), i.e. the pin that lets the oracle stand in for the reference."""
import numpy as np
import pytest

from helpers import IMAGE_CASES, assert_close, image_oracle_for, load_golden
from oracle import image_oracle as io
from oracle.datasets import sha256_of


@pytest.mark.parametrize("name", list(IMAGE_CASES))
def test_image_oracle_matches_reference_golden(name):
    case = IMAGE_CASES[name]
    g = load_golden("image", name)
    env, imgs, _ = image_oracle_for(case)
    assert sha256_of(imgs) == str(g["bank_sha256"])
    explicit = None
    if case["task"] == "loc" and case["N"] > 4:
        explicit = g["reset_target"]  # skip the O(P^2) unique sampling for the bigger case
    obs, info = env.reset(seed=case["seed"], explicit_targets=explicit)
    for k in obs:
        assert np.array_equal(obs[k], g["reset_" + k]), f"reset obs {k}"
    assert np.array_equal(info["index"], g["reset_index"])
    if case["task"] == "loc":
        assert np.array_equal(env.target, g["reset_target"]), "unique glimpse targets"
    Hr, Wr = g["counts"].shape[-2:]
    for s in range(case["steps"]):
        o, r, te, tr, inf = env.step(g["actions"][s], g["preds"][s])
        for k in o:
            assert o[k].dtype == g["obs_" + k].dtype or k == "inverted_label"
            assert np.array_equal(o[k], g["obs_" + k][s]), f"step {s} obs {k}"
        assert np.array_equal(env.pos, g["pos64"][s]), f"step {s} float64 sensor position"
        assert np.array_equal(te, g["terminated"][s]) and np.array_equal(tr, g["truncated"][s])
        assert np.array_equal(inf["index"], g["index"][s])
        assert np.array_equal(inf["prediction"]["target"], g["target"][s])
        assert np.array_equal(np.asarray(inf["base_reward"], np.float64), g["base_reward"][s])
        # exp/log are SIMD-path dependent in numpy: tolerance instead of bit equality
        assert_close(inf["prediction"]["loss"], g["loss"][s], rtol=1e-6, atol=1e-7, what=f"step {s} loss")
        assert_close(r, g["reward"][s], rtol=1e-6, atol=1e-7, what=f"step {s} reward")
        assert np.array_equal(env.counts[:2].astype(np.uint8), g["counts"][s]), f"step {s} visitation counts"
        assert_close(env.quality[:2], g["quality"][s], rtol=1e-6, atol=1e-7, what=f"step {s} quality map")


def test_ce_matches_reference_unit_test_tolerance():
    """The one numeric test the reference ships (test/test_active_classification_env.py:17-50): CE numpy vs
    torch at rtol 1e-4 on 20 random shapes, rng default_rng(0)."""
    import torch

    rng = np.random.default_rng(0)
    for _ in range(20):
        batch_shape = tuple(rng.integers(1, 5, size=rng.integers(0, 3)))
        k = int(rng.integers(2, 10))
        logits = rng.standard_normal(batch_shape + (k,)).astype(np.float32)
        target = rng.integers(0, k, size=batch_shape)
        ce = -np.take_along_axis(io.log_softmax_f32(logits), target[..., None], axis=-1)[..., 0]
        ce_t = -torch.take_along_dim(
            torch.nn.functional.log_softmax(torch.from_numpy(logits), dim=-1), torch.from_numpy(target)[..., None], dim=-1
        )[..., 0]
        np.testing.assert_allclose(ce, ce_t.numpy(), rtol=1e-4)


def test_tap_indices_bisect_semantics():
    # a coordinate a hair below a grid value must fall in the lower interval with weight ~1
    H = 28
    off = (H - 1) / 2
    y = np.array([np.nextafter(3 - off, -np.inf), 3 - off, -off, off])
    k, w = io.tap_indices(y, H)
    assert k.tolist() == [2, 3, 0, H - 2]
    assert w[1] == 0.0 and w[2] == 0.0 and w[3] == 1.0 and 0.0 < w[0] <= 1.0
