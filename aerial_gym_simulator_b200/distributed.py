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


class P2PObsAllGather:
    """Same contract as ObsAllGather (equal shards only), but the collective is ONE hand-written
    kernel of NVLink peer stores + a flag handshake (csrc/p2p_allgather.cu) over buffers obtained
    from a torch symmetric-memory rendezvous -- no NCCL call on the step path."""

    def __init__(self, local_count: int, feat: int, device, group=None, dtype=torch.float32, num_buffers: int = 2):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        self._C, self._lib = C, _lib.load()
        self._check = _lib.check
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if dtype != torch.float32:
            raise ValueError("float32 only")
        self.bytes = local_count * feat * 4
        if self.bytes % 16:
            raise ValueError("shard bytes must be a multiple of 16")
        self.device = torch.device(device)
        # two gathered buffers, alternated by epoch parity: a peer can run at most one epoch ahead of
        # the slowest reader (everybody waits for everybody's flag), so epoch e+1 never overwrites
        # data of epoch e that a slower rank may still be consuming
        # (fused mode with lag 1 needs 4: a rank may push epoch e+4 once its peers finished epoch e+2,
        # i.e. after their consumers of epoch e -- stream-ordered before their step e+2 -- are done)
        if num_buffers not in (2, 4):
            raise ValueError("num_buffers must be 2 or 4")
        self.num_buffers = num_buffers
        self.outs, self.buf_ptrs = [], []
        for _ in range(num_buffers):
            o = symm_mem.empty(self.world * local_count, feat, dtype=dtype, device=self.device)
            h = symm_mem.rendezvous(o, group)
            self.outs.append(o)
            self.buf_ptrs.append(torch.tensor(list(h.buffer_ptrs), dtype=torch.int64, device=self.device))
            setattr(self, f"_h{len(self.outs)}", h)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=self.device)
        self.flags.zero_()
        self.h_flags = symm_mem.rendezvous(self.flags, group)
        self.flag_ptrs = torch.tensor(list(self.h_flags.buffer_ptrs), dtype=torch.int64, device=self.device)
        self.scratch = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group)  # every rank's flags are zeroed before anybody publishes epoch 1

    def next_epoch(self):
        """(epoch, buffer parity) of the next collective -- used by Hp1Engine when the all-gather is
        fused into the step kernel (Hp1Engine.attach_obs_gather)."""
        self.epoch += 1
        return self.epoch, self.epoch % self.num_buffers

    def __call__(self, obs_local: torch.Tensor) -> torch.Tensor:
        C = self._C
        if obs_local.numel() * 4 != self.bytes or not obs_local.is_contiguous():
            raise ValueError("obs_local must be the contiguous local shard")
        _, b = self.next_epoch()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self._check(self._lib.agx_p2p_allgather(
            C.c_void_p(obs_local.data_ptr()), C.c_void_p(self.buf_ptrs[b].data_ptr()), C.c_void_p(self.flag_ptrs.data_ptr()),
            self.world, self.rank, C.c_uint64(self.bytes), C.c_uint32(self.epoch), C.c_void_p(self.scratch.data_ptr()), stream),
            "agx_p2p_allgather")
        return self.outs[b]
