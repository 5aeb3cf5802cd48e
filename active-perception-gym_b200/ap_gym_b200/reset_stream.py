"""Host-side replay of the reference's reset-path random streams.

All random draws of the image environments happen at (auto)reset and do not depend on the agent, so they are
produced here with numpy's own PCG64 Generators -- in exactly the reference's order -- and handed to the
engine as reset payloads.  Shards of a multi-GPU run draw the GLOBAL batch and slice their env range, so
results do not depend on the number of GPUs.

Reference chain:
  g0 = Generator(PCG64(SeedSequence(seed)))                    gymnasium.utils.seeding.np_random
  s1 = g0.integers(0, 2**32-1, endpoint=True) -> module.seed   ap_gym/envs/image_classification.py:144-149
  g1 = default_rng(s1); s2 = g1.integers(0, 2**32-1, endpoint=True)
  g2 = default_rng(s2)  (DatasetBatchIterator)                 ap_gym/envs/image/image_perception_module.py:105-118
  every (auto)reset: idx = g2.integers(0, len(dataset), N)     ap_gym/envs/dataset/dataset_iterator.py:57
                     [inv = g1.integers(0, 2, N) == 1]         ap_gym/envs/image/image_perception_module.py:131-133
                     pos = g1.uniform(-1, 1, (N, 2))           ap_gym/envs/image/image_perception_module.py:159-161
  localization autoreset: target = g0.uniform(-1, 1, (N, 2)).astype(f32), BEFORE the module reset
                                                               ap_gym/envs/image_localization.py:153-156
"""
from __future__ import annotations

import numpy as np


class ImageResetStream:
    def __init__(self, seed: int | None, global_num_envs: int, bank_len: int, env_slice: slice | None = None):
        self.g0 = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        self.g1 = np.random.default_rng(self.g0.integers(0, 2**32 - 1, endpoint=True))
        self.g2 = np.random.default_rng(self.g1.integers(0, 2**32 - 1, endpoint=True))
        self.n = int(global_num_envs)
        self.bank_len = int(bank_len)
        self.sl = env_slice if env_slice is not None else slice(0, self.n)

    def draw_episode(self, randomly_invert_labels: bool):
        """(idx int64 (n,), invert uint8 (n,) | None, pos float64 (n,2)) for this shard's env range."""
        idx = self.g2.integers(0, self.bank_len, self.n)
        inv = None
        if randomly_invert_labels:
            inv = (self.g1.integers(0, 2, size=self.n) == 1).astype(np.uint8)[self.sl]
        pos = self.g1.uniform(-1, 1, size=(self.n, 2))
        return np.ascontiguousarray(idx[self.sl]), inv, np.ascontiguousarray(pos[self.sl])

    def draw_autoreset_targets(self):
        return np.ascontiguousarray(self.g0.uniform(-1, 1, (self.n, 2)).astype(np.float32)[self.sl])

    def draw_unique_selection(self, top_k: int, max_cell: np.ndarray):
        """The two draws that end sample_unique_glimpse_positions (image_perception_module.py:278-289)."""
        sel = self.g1.integers(0, top_k, size=self.n)
        jitter = self.g1.uniform(-max_cell, max_cell, (self.n, 2))
        return sel[self.sl], jitter[self.sl]
