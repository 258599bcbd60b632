"""Multi-GPU plumbing: envs are independent, so the env index range is split
into contiguous shards, one process per GPU, and the only collective on the step path is one
all-gather of the observation tensor for the policy.  The reference has no distributed code;
there is nothing to be compatible with.  Backend-agnostic (nccl on GPUs, gloo in CPU tests)."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(offset, count) of the contiguous shard of ``rank``; the remainder goes to the first ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(num_envs_global, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def env_from_torchrun():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ObsAllGather:
    """Persistent buffers for the per-step observation all-gather.

    Equal shards use ``all_gather_into_tensor`` (one NCCL kernel into a contiguous [N,D] buffer);
    ragged shards (N not divisible by the world size) gather max-count padded blocks and compact."""

    def __init__(self, local_count: int, feat: int, num_envs_global: int, device, dtype=torch.float32,
                 group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.out = torch.empty(num_envs_global, feat, device=device, dtype=dtype)
        counts = [shard_range(num_envs_global, r, self.world)[1] for r in range(self.world)]
        if counts[self.rank] != local_count:
            raise ValueError(f"rank {self.rank}: local shard has {local_count} envs, expected {counts[self.rank]}")
        self.equal = len(set(counts)) == 1
        self.offs = [shard_range(num_envs_global, r, self.world)[0] for r in range(self.world)]
        self.counts = counts
        if not self.equal:
            self.pad_in = torch.zeros(max(counts), feat, device=device, dtype=dtype)
            self.pad_out = torch.empty(self.world * max(counts), feat, device=device, dtype=dtype)

    def __call__(self, obs_local: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            self.out.copy_(obs_local)
        elif self.equal:
            dist.all_gather_into_tensor(self.out, obs_local.contiguous(), group=self.group)
        else:
            mc = self.pad_in.shape[0]
            self.pad_in[: obs_local.shape[0]].copy_(obs_local)
            dist.all_gather_into_tensor(self.pad_out, self.pad_in, group=self.group)
            for r, (o, c) in enumerate(zip(self.offs, self.counts)):
                self.out[o:o + c].copy_(self.pad_out[r * mc: r * mc + c])
        return self.out
