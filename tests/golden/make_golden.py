"""Generates tests/golden/image_*.npz by running the UNMODIFIED reference (/root/reference) through the
test-only gymnasium shim (oracle/shim).  Run here (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py image

Every case stores its inputs (seed, actions, predictions, config) and the reference's outputs per step.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))


def _ref_dataset_cls():
    from ap_gym.envs.image import ImageClassificationDataset

    class ArrayDataset(ImageClassificationDataset):
        """10-line finite bank through the reference's own dataset interface."""

        def __init__(self, imgs, labels, num_classes):
            self._imgs, self._labels, self._k = imgs, labels, num_classes

        def _get_num_classes(self):
            return self._k

        def _get_num_channels(self):
            return self._imgs.shape[-1]

        def _get_length(self):
            return len(self._imgs)

        def _get_data_point_batch(self, idx):
            return self._imgs[idx], self._labels[idx]

    return ArrayDataset


IMAGE_CASES = {
    # name: dict(task, bank spec, N, steps, config kwargs)
    "circlesquare_n64": dict(task="cls", bank=("circle_square", dict(image_shape=(28, 28), show_gradient=True)),
                             K=2, N=64, steps=53, seed=0, cfg={}),
    "mnistlike_cls_n32": dict(task="cls", bank=("u8", dict(seed=0, M=512, H=28, W=28, C=1, K=10)),
                              K=10, N=32, steps=36, seed=1, cfg={}),
    "cifarlike_cls_n16": dict(task="cls", bank=("u8", dict(seed=2, M=128, H=32, W=32, C=3, K=10)),
                              K=10, N=16, steps=36, seed=2, cfg={}),
    "inverted_cls_n16": dict(task="cls", bank=("circle_square", dict(image_shape=(20, 20), show_gradient=False)),
                             K=2, N=16, steps=36, seed=3, cfg=dict(randomly_invert_labels=True)),
    "scaled_cls_n8": dict(task="cls", bank=("u8", dict(seed=4, M=64, H=40, W=36, C=3, K=7)),
                          K=7, N=8, steps=20, seed=4,
                          cfg=dict(sensor_size=(6, 6), sensor_scale=1.5, step_limit=8, max_step_length=0.35)),
    "tinyloc_n4": dict(task="loc", bank=("u8", dict(seed=5, M=32, H=64, W=64, C=3, K=200)),
                       K=200, N=4, steps=36, seed=5, cfg=dict(sensor_size=(10, 10))),
    "mnistloc_n6": dict(task="loc", bank=("u8", dict(seed=6, M=64, H=28, W=28, C=1, K=10)),
                        K=10, N=6, steps=36, seed=6, cfg={}),
}


def make_bank(spec):
    from oracle.datasets import circle_square_bank, synthetic_u8_bank

    kind, kw = spec
    if kind == "circle_square":
        return circle_square_bank(**kw)
    return synthetic_u8_bank(**kw)


def run_image_case(name, case):
    import ap_gym  # noqa: F401  (registers; slow import)
    from ap_gym.envs.image import ImagePerceptionConfig, CircleSquareDataset
    from ap_gym.envs.image_classification import ImageClassificationVectorEnv
    from ap_gym.envs.image_localization import ImageLocalizationVectorEnv
    from oracle.datasets import sha256_of

    kind, kw = case["bank"]
    imgs, labels = make_bank(case["bank"])
    if kind == "circle_square":
        ds = CircleSquareDataset(**kw)  # the reference's own generator
        ref_imgs, ref_labels = ds.get_data_point_batch(np.arange(len(ds)))
        assert sha256_of(ref_imgs) == sha256_of(imgs), "oracle circle_square_bank != reference dataset"
        assert np.array_equal(ref_labels, labels)
    else:
        ds = _ref_dataset_cls()(imgs, labels, case["K"])
    cfg = ImagePerceptionConfig(dataset=ds, **case["cfg"])
    N = case["N"]
    cls = ImageClassificationVectorEnv if case["task"] == "cls" else ImageLocalizationVectorEnv
    env = cls(N, cfg)
    rng = np.random.default_rng(1000 + case["seed"])
    obs, info = env.reset(seed=case["seed"])
    mod = env._ImageClassificationVectorEnv__image_perception_module if case["task"] == "cls" else \
        env._ImageLocalizationVectorEnv__image_perception_module
    rec = {"reset_" + k: v for k, v in obs.items()}
    rec["reset_index"] = info["index"]
    if case["task"] == "loc":
        rec["reset_target"] = env._ImageLocalizationVectorEnv__current_prediction_target.copy()
    steps = case["steps"]
    P = case["K"] if case["task"] == "cls" else 2
    actions = np.empty((steps, N, 2), np.float32)
    preds = np.empty((steps, N, P), np.float32)
    out = {k: [] for k in ["reward", "terminated", "truncated", "base_reward", "loss", "target", "index",
                           "pos64", "counts", "quality"]}
    obs_rec = {k: [] for k in obs}
    for s in range(steps):
        # mix of in-disc, out-of-disc and zero actions
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        if s % 7 == 3:
            a[: N // 2] = 0
        if s % 5 == 1:
            a *= np.float32(1.4)
        if case["task"] == "cls":
            p = (rng.standard_normal((N, P)) * (1 + 3 * (s % 3))).astype(np.float32)
        else:
            p = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        actions[s], preds[s] = a, p
        o, r, te, tr, inf = env.step({"action": a, "prediction": p})
        for k in obs_rec:
            obs_rec[k].append(np.asarray(o[k]).copy())
        out["reward"].append(np.asarray(r, dtype=np.float64))
        out["terminated"].append(te.copy())
        out["truncated"].append(tr.copy())
        out["base_reward"].append(np.asarray(inf["base_reward"], dtype=np.float64))
        out["loss"].append(np.asarray(inf["prediction"]["loss"], dtype=np.float64))
        out["target"].append(np.asarray(inf["prediction"]["target"]).copy())
        out["index"].append(np.asarray(inf["index"]).copy())
        out["pos64"].append(mod._ImagePerceptionModule__current_sensor_pos_norm.copy())
        out["counts"].append(mod._ImagePerceptionModule__visitation_counts[:2].astype(np.uint8))
        out["quality"].append(mod._ImagePerceptionModule__prediction_quality_map[:2].copy())
    env.close()
    rec.update({"actions": actions, "preds": preds})
    rec.update({"obs_" + k: np.stack(v) for k, v in obs_rec.items()})
    rec.update({k: np.stack(v) for k, v in out.items()})
    rec["bank_sha256"] = np.array(sha256_of(imgs))
    np.savez_compressed(os.path.join(OUT, f"image_{name}.npz"), **rec)
    print(name, "ok", {k: v.shape for k, v in rec.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "image"
    if what == "image":
        for name, case in IMAGE_CASES.items():
            if len(sys.argv) > 2 and name not in sys.argv[2:]:
                continue
            run_image_case(name, case)
    elif what == "lidar":
        from make_golden_lidar import main as lidar_main  # type: ignore

        lidar_main(sys.argv[2:])
