"""CPU (gloo, world_size 2): the host-side multi-GPU logic -- env-index sharding, shard-invariant reset payloads,
and the final statistics reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ap_gym_b200.distributed import reduce_stats, shard_range
from ap_gym_b200.reset_stream import ImageResetStream


def test_shard_ranges_partition_the_env_index():
    for n, w in [(65536, 8), (10, 3), (7, 8), (262144, 4)]:
        shards = [shard_range(n, r, w) for r in range(w)]
        assert shards[0].env_offset == 0 and sum(s.num_envs for s in shards) == n
        for a, b in zip(shards, shards[1:]):
            assert a.env_offset + a.num_envs == b.env_offset
        assert max(s.num_envs for s in shards) - min(s.num_envs for s in shards) <= 1


def test_reset_payloads_do_not_depend_on_the_number_of_shards():
    n, bank = 1000, 777
    full = ImageResetStream(5, n, bank)
    parts = [ImageResetStream(5, n, bank, shard_range(n, r, 4).slice) for r in range(4)]
    for _ in range(3):
        ft = full.draw_autoreset_targets()
        fi, finv, fp = full.draw_episode(True)
        pt = np.concatenate([p.draw_autoreset_targets() for p in parts])
        drawn = [p.draw_episode(True) for p in parts]
        assert np.array_equal(ft, pt)
        assert np.array_equal(fi, np.concatenate([d[0] for d in drawn]))
        assert np.array_equal(finv, np.concatenate([d[1] for d in drawn]))
        assert np.array_equal(fp, np.concatenate([d[2] for d in drawn]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = shard_range(1001, rank, world)
    # per-shard "episode statistics": count, sum of env indices, sum of squares
    idx = torch.arange(sh.env_offset, sh.env_offset + sh.num_envs, dtype=torch.float64)
    stats = torch.stack([torch.tensor(float(sh.num_envs), dtype=torch.float64), idx.sum(), (idx * idx).sum()])
    reduce_stats(stats)
    dist.barrier()
    if rank == 0:
        torch.save(stats, out)
    dist.destroy_process_group()


def test_stats_reduce_world_size_2_gloo(tmp_path):
    out = str(tmp_path / "stats.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    stats = torch.load(out)
    idx = np.arange(1001, dtype=np.float64)
    assert stats.tolist() == [1001.0, idx.sum(), (idx * idx).sum()]
