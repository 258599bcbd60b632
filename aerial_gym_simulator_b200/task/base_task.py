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
