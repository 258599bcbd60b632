"""Common surface of every task (reference: aerial_gym/task/base_task.py): the gym-style attribute slots and the seeding rule
(-1 / None / negative -> a clock-derived seed; one seed for python, numpy and torch)."""
import abc
import os
import random
import time

import numpy as np
import torch

_GYM_SLOTS = ("action_space", "observation_space", "reward_range", "metadata", "spec")


def _clock_seed() -> int:
    return time.time_ns() % (2**32)


class BaseTask(abc.ABC):
    def __init__(self, task_config):
        self.task_config = task_config
        for slot in _GYM_SLOTS:
            setattr(self, slot, None)
        self.seed(_clock_seed() if task_config.seed == -1 else task_config.seed)

    def seed(self, seed):
        """Seed python / numpy / torch (all CUDA devices) with one value; remembered as self._seed."""
        seed = _clock_seed() if (seed is None or seed < 0) else int(seed)
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
        os.environ["PYTHONHASHSEED"] = str(seed)
        self._seed = seed

    # ---- multi-GPU: env-sharded task, one process per GPU ----------------------------------------------------------------
    # task_config.args = {"world_size": W, "rank": r} (or {"shard": "torchrun"}: RANK / WORLD_SIZE from the environment): this process
    # simulates num_envs envs -- global env ids rank * num_envs .. -- and the ONE collective of the step path is the all-gather of the
    # observation tensor handed to the policy: after step() / reset(), task_obs["observations"] is the GLOBAL [W * num_envs, D] tensor
    # (rank-major, one of a ring of symmetric-memory buffers, so a different tensor object each step) and task_obs["observations_local"]
    # this rank's rows inside it; rewards / terminations / truncations / infos stay local.  {"world_size": W, "loopback": True} emulates W
    # ranks inside one process (tests on one GPU: the peers' rows stay zero).
    def shard_spec(self):
        """(world, rank, loopback) from task_config.args"""
        a = dict(getattr(self.task_config, "args", None) or {})
        if a.get("shard") == "torchrun":
            return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), False
        return int(a.get("world_size", 1)), int(a.get("rank", 0)), bool(a.get("loopback", False))

    def init_sharding(self, num_envs, obs_dim, device):
        from ..distributed import PipelinedObsGather

        world, rank, loopback = self.shard_spec()
        self.shard_world, self.shard_rank, self.obs_gather = world, rank, None
        if world <= 1:
            return None
        if torch.device(device).type != "cuda":
            raise ValueError("a sharded task (args['world_size'] > 1) needs a CUDA device")
        self.obs_gather = PipelinedObsGather(num_envs, obs_dim, device, num_buffers=4, loopback_world=world if loopback else 0)
        if not loopback and self.obs_gather.rank != rank:
            raise ValueError(f"args['rank'] = {rank} but the process group says {self.obs_gather.rank}")
        return self.obs_gather

    def gather_observations(self, local_obs):
        """stream-ordered form for tasks whose observation is written by several kernels: push `local_obs` [N, D] (this rank's rows)
        into every rank's buffer, wait for everybody's rows, return (global [W*N, D], local view)"""
        g = self.obs_gather
        epoch, slot = g.next_epoch()
        cur = torch.cuda.current_stream(local_obs.device).cuda_stream
        g.push(local_obs.data_ptr(), epoch, slot, stream=cur)
        if getattr(g, "peer_outs", None) is not None:
            g.loopback_complete(epoch)
        out = g.wait(epoch)
        return out, g.own_slot[slot]

    # the five entry points every task implements
    @abc.abstractmethod
    def reset(self): ...

    @abc.abstractmethod
    def reset_idx(self, env_ids): ...

    @abc.abstractmethod
    def step(self, action): ...

    @abc.abstractmethod
    def render(self, mode="human"): ...

    @abc.abstractmethod
    def close(self): ...
