import sys, numpy as np
sys.path.insert(0, "active-perception-gym_b200"); sys.path.insert(0, ".")
import ap_gym_b200 as apg
from oracle.datasets import synthetic_u8_bank
rng = np.random.default_rng(0)
for (H, W, C, S, K, task) in [(28, 28, 1, 5, 10, "cls"), (64, 64, 3, 10, 7, "loc"), (33, 35, 1, 5, 4, "cls"), (70, 72, 3, 10, 3, "loc"), (64, 64, 1, 36, 3, "cls")]:
    imgs, labels = synthetic_u8_bank(seed=1, M=40, H=H, W=W, C=C, K=K)
    cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, K), sensor_size=(S, S), step_limit=5)
    cls = apg.ImageClassificationVectorEnv if task == "cls" else apg.ImageLocalizationVectorEnv
    N = 131
    env = cls(N, cfg, output="numpy")
    env.reset(seed=1)
    for s in range(13):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32) * 1.5
        p = rng.standard_normal((N, K if task == "cls" else 2)).astype(np.float32)
        env.step({"action": a, "prediction": p})
    print("image ok", H, W, C, S, task, env.episode_stats(reduce=False)["episodes"])
    env.close()
z = np.load("tests/golden/maps_maze21.npz")
bank = apg.FloorMapBank(z["bits"][:16], int(z["width"]), int(z["height"]))
for B in (8, 12, 64, 256):
    env = apg.LIDARLocalization2DVectorEnv(77, bank, lidar_beam_count=B, emit_cells=True, output="numpy")
    env.reset(seed=3)
    for s in range(103):
        env.step({"action": rng.uniform(-1, 1, (77, 2)).astype(np.float32), "prediction": rng.uniform(-1, 1, (77, 2)).astype(np.float32)})
    print("lidar ok", B, env.episode_stats(reduce=False)["episodes"])
    env.close()
