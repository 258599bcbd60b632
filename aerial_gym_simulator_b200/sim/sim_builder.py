"""SimBuilder: the factory the reference's examples and tasks go through (sim/sim_builder.py:22-48 there).  It remembers what it was
last asked for and the EnvManager it made; delete_env() releases that env's device allocations."""
import torch

from ..env_manager import EnvManager


class SimBuilder:
    _REMEMBERED = ("sim_name", "env_name", "robot_name")

    def __init__(self):
        self.env = None
        self._request = {}

    def __getattr__(self, name):  # builder.sim_name / .env_name / .robot_name: the names of the last build_env call (None before)
        if name in SimBuilder._REMEMBERED:
            return self.__dict__.get("_request", {}).get(name)
        raise AttributeError(name)

    def build_env(self, sim_name, env_name, robot_name, controller_name, device, args=None, num_envs=None, use_warp=None, headless=None):
        request = {k: v for k, v in locals().items() if k != "self"}
        self._request = request
        self.env = EnvManager(**request)
        return self.env

    def delete_env(self):
        env, self.env = self.env, None
        if env is not None:
            env.delete_env()  # frees the engine's and the ray-caster's buffers
        if torch.cuda.is_available():
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
