"""Times the UNMODIFIED reference (/root/reference, through the test-only gymnasium shim of oracle/shim) on one host
core for the image workloads of bench.py -- SURVEY.md section 8(d): "run the reference's own code ... single process".

    python tools/time_real_reference.py mnist|tinyloc      # prints one JSON line

Only runs where /root/reference exists (the build container); bench.py reports the number recorded in
profiles/r02_real_reference.json elsewhere.  The LIDAR reference needs shapely/GEOS, which this image does not have.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)


def main(workload):
    import ap_gym  # noqa: F401
    from ap_gym.envs.image import ImageClassificationDataset, ImagePerceptionConfig
    from ap_gym.envs.image_classification import ImageClassificationVectorEnv
    from ap_gym.envs.image_localization import ImageLocalizationVectorEnv
    from oracle.datasets import synthetic_u8_bank

    class ArrayDataset(ImageClassificationDataset):
        def __init__(self, imgs, labels, k):
            self._imgs, self._labels, self._k = imgs, labels, k

        def _get_num_classes(self):
            return self._k

        def _get_num_channels(self):
            return self._imgs.shape[-1]

        def _get_length(self):
            return len(self._imgs)

        def _get_data_point_batch(self, idx):
            return self._imgs[idx], self._labels[idx]

    rng = np.random.default_rng(0)
    if workload == "mnist":
        imgs, labels = synthetic_u8_bank(seed=0, M=4096, H=28, W=28, C=1, K=10)
        n, steps, pw = 1024, 51, 10
        env = ImageClassificationVectorEnv(n, ImagePerceptionConfig(dataset=ArrayDataset(imgs.astype(np.float32) / 255, labels, 10)))
    elif workload == "tinyloc":
        imgs, labels = synthetic_u8_bank(seed=0, M=512, H=64, W=64, C=3, K=200)
        n, steps, pw = 64, 51, 2
        env = ImageLocalizationVectorEnv(n, ImagePerceptionConfig(dataset=ArrayDataset(imgs.astype(np.float32) / 255, labels, 200),
                                                                  sensor_size=(10, 10)))
        # explicit reset() would run sample_unique_glimpse_positions (seconds per env, not on the step path): start from
        # an autoreset instead
    else:
        raise SystemExit("no real reference for " + workload)
    t_reset = time.perf_counter()
    env.reset(seed=0)
    t_reset = time.perf_counter() - t_reset
    a = rng.uniform(-1, 1, (steps + 17, n, 2)).astype(np.float32)
    p = (rng.standard_normal((steps + 17, n, pw)) if pw > 2 else rng.uniform(-1, 1, (steps + 17, n, pw))).astype(np.float32)
    for s in range(17):  # one warm-up episode
        env.step({"action": a[s], "prediction": p[s]})
    t0 = time.perf_counter()
    for s in range(17, 17 + steps):
        env.step({"action": a[s], "prediction": p[s]})
    dt = time.perf_counter() - t0
    print(json.dumps({"value": n * steps / dt, "unit": "env-steps/s", "cores": 1, "kind": "reference",
                      "sample": f"unmodified ap_gym 0.5.0, N={n} x {steps} steps, {dt:.1f} s (explicit reset {t_reset:.1f} s)"}))


if __name__ == "__main__":
    main(sys.argv[1])
