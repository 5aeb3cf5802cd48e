"""Multi-GPU plumbing: environments shard by env index, one process per GPU, no per-step collective.

The only exchange of the data path is ONE reduction of run/episode statistics at the end of a run
(`reduce_stats`, NCCL over NVLink on GPUs, gloo on CPU tensors in the tests).  Everything random is derived from
the GLOBAL env index (image envs: the global batch stream is drawn and sliced, reset_stream.py; LIDAR: env i is
seeded with seed + env_offset + i), so results do not depend on the number of GPUs.
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world_size: int
    global_num_envs: int
    env_offset: int
    num_envs: int

    @property
    def slice(self) -> slice:
        return slice(self.env_offset, self.env_offset + self.num_envs)


def shard_range(global_num_envs: int, rank: int, world_size: int) -> Shard:
    """Contiguous block [g*N/G, (g+1)*N/G) of env indices for rank g (remainder spread over the first ranks)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(global_num_envs), int(world_size))
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return Shard(rank, world_size, int(global_num_envs), off, n)


def env_from_torchrun() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def make_sharded(env_cls, global_num_envs: int, *args, rank: int | None = None, world_size: int | None = None,
                 device: int | None = None, **kwargs):
    """Constructs this rank's shard of a `global_num_envs`-wide vector env (image or LIDAR class)."""
    r, lr, w = env_from_torchrun()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    device = lr if device is None else device
    sh = shard_range(global_num_envs, rank, world_size)
    if "global_num_envs" in env_cls.__init__.__code__.co_varnames or hasattr(env_cls, "_TASK"):
        kwargs.setdefault("global_num_envs", sh.global_num_envs)
    return env_cls(sh.num_envs, *args, device=device, env_offset=sh.env_offset, **kwargs), sh


def reduce_stats(stats, group=None):
    """Sum a small float64 tensor of statistics over all ranks (in place) and return it.  No-op without an
    initialised process group.  Tensors live on the GPU for NCCL and on the CPU for gloo."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
