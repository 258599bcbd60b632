"""PositionSetpointTaskSim2Real / PositionSetpointTaskAccelerationSim2Real
(task/position_setpoint_task_sim2real/position_setpoint_task_sim2real.py and
task/position_setpoint_task_acceleration_sim2real/position_setpoint_task_acceleration_sim2real.py): the lmf2 position tasks trained
for deployment, velocity or acceleration setpoints, 17-D noisy observation.  Same surface and step order as the reference --

    prev_actions / prev_dist bookkeeping -> sim_env.step -> reward + crashes (agx_s2r_reward) -> truncations ->
    post_reward_calculation_step (reset) -> observation (agx_s2r_obs; normalises the sign of robot_orientation IN PLACE like the
    reference, :204-207)

-- with the two epilogues as C-ABI kernel launches.  Quirks kept: ``self.actions`` becomes the caller's tensor (:160), which the
acceleration task then scales in place (:170); the four torch.randn_like draws keep the reference's order."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..sim import SimBuilder
from ..utils.math import quat_rotate
from .base_task import BaseTask
from .spaces import Box, Dict


class PositionSetpointTaskSim2Real(BaseTask):
    VARIANT = 0  # agx_s2r_reward: 0 = velocity commands, 1 = acceleration commands

    def __init__(self, task_config, seed=None, num_envs=None, headless=None, device=None, use_warp=None):
        for k, v in (("seed", seed), ("num_envs", num_envs), ("headless", headless), ("device", device), ("use_warp", use_warp)):
            if v is not None:
                setattr(task_config, k, v)
        super().__init__(task_config)
        self.device = torch.device(self.task_config.device)
        self.lib = _lib.load()
        args = dict(self.task_config.args or {})
        args.setdefault("seed", self._seed)
        self.sim_env = SimBuilder().build_env(
            sim_name=self.task_config.sim_name, env_name=self.task_config.env_name, robot_name=self.task_config.robot_name,
            controller_name=self.task_config.controller_name, args=args, device=self.device, num_envs=self.task_config.num_envs,
            use_warp=self.task_config.use_warp, headless=self.task_config.headless)
        N, dev, A = self.sim_env.num_envs, self.device, self.task_config.action_space_dim
        self.num_envs = N
        self.actions = torch.zeros((N, A), device=dev)
        self.prev_actions = torch.zeros_like(self.actions)
        self.prev_actions_vehicle_frame = torch.zeros_like(self.actions)
        self.actions_vehicle_frame = torch.zeros_like(self.actions)
        self.prev_dist = torch.zeros(N, device=dev)
        self.target_position = torch.zeros((N, 3), device=dev)
        self.obs_dict = self.sim_env.get_obs()
        self.obs_dict["num_obstacles_in_env"] = 1
        self.terminations = self.obs_dict["crashes"]
        self.truncations = self.obs_dict["truncations"]
        self.rewards = torch.zeros(N, device=dev)
        self.observation_space = Dict({"observations": Box(low=-1.0, high=1.0, shape=(self.task_config.observation_space_dim,), dtype=np.float32)})
        self.action_space = Box(low=-1.0, high=1.0, shape=(A,), dtype=np.float32)
        self.counter = 0
        self.task_obs = {
            "observations": torch.zeros((N, self.task_config.observation_space_dim), device=dev),
            "priviliged_obs": torch.zeros((N, self.task_config.privileged_observation_space_dim), device=dev),
            "collisions": torch.zeros((N, 1), device=dev),
            "rewards": torch.zeros((N, 1), device=dev),
        }
        self.infos = {}

    def close(self):
        self.sim_env.delete_env()

    def reset(self):
        self.target_position[:, 0:3] = 0.0
        self.infos = {}
        self.sim_env.reset()
        return self.get_return_tuple()

    def reset_idx(self, env_ids):
        self.target_position[:, 0:3] = 0.0
        self.infos = {}
        self.sim_env.reset_idx(env_ids)

    def render(self):
        return None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _pre_step(self, actions):
        self.prev_actions[:] = self.actions
        self.prev_dist[:] = torch.norm(self.target_position - self.obs_dict["robot_position"], dim=1)
        self.actions = actions  # (the caller's tensor from here on, like the reference)

    def compute_rewards_and_crashes(self, obs_dict):
        st = obs_dict["robot_state_tensor"]
        p = lambda t: C.c_void_p(t.data_ptr())
        prev = self.prev_actions_vehicle_frame if self.VARIANT else self.prev_actions
        act = self.actions if (self.actions.dtype == torch.float32 and self.actions.is_contiguous()) else self.actions.float().contiguous()
        _lib.check(self.lib.agx_s2r_reward(self.num_envs, self.VARIANT, p(st), st.stride(0), p(obs_dict["robot_vehicle_orientation"]),
                                           p(obs_dict["robot_body_linvel"]), p(self.target_position), p(self.prev_dist), p(act), p(prev),
                                           p(self.actions_vehicle_frame), p(obs_dict["crashes"]), p(self.rewards), self._stream()),
                   "agx_s2r_reward")
        return self.rewards, obs_dict["crashes"]

    def step(self, actions):
        self.counter += 1
        self._pre_step(actions)
        self.sim_env.step(actions=self.actions)
        self.compute_rewards_and_crashes(self.obs_dict)
        if self.task_config.return_state_before_reset:
            return_tuple = self.get_return_tuple()
        self.truncations[:] = self.sim_env.sim_steps > self.task_config.episode_len_steps
        self.sim_env.post_reward_calculation_step()
        self.infos = {}
        if not self.task_config.return_state_before_reset:
            return_tuple = self.get_return_tuple()
        return return_tuple

    def get_return_tuple(self):
        self.process_obs_for_task()
        return (self.task_obs, self.rewards, self.terminations, self.truncations, self.infos)

    def process_obs_for_task(self, noise=None):
        """:202-228; noise [N,12] = the four torch.randn_like draws (euler, position, body linvel, body angvel) unless given."""
        od, N, dev = self.obs_dict, self.num_envs, self.device
        if noise is None:
            noise = torch.cat([torch.randn((N, 3), device=dev) for _ in range(4)], dim=1)
        st, obs = od["robot_state_tensor"], self.task_obs["observations"]
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.agx_s2r_obs(N, p(st), st.stride(0), p(od["robot_body_linvel"]), p(od["robot_body_angvel"]), p(od["robot_actions"]),
                                        p(self.target_position), p(noise), p(obs), obs.stride(0), self._stream()), "agx_s2r_obs")
        self.task_obs["rewards"] = self.rewards
        self.task_obs["terminations"] = self.terminations
        self.task_obs["truncations"] = self.truncations


class PositionSetpointTaskAccelerationSim2Real(PositionSetpointTaskSim2Real):
    VARIANT = 1

    def _pre_step(self, actions):
        self.prev_actions[:] = self.actions
        self.prev_actions_vehicle_frame[:, 0:3] = quat_rotate(self.obs_dict["robot_orientation"], self.prev_actions[:, 0:3])  # :162-165
        self.prev_actions_vehicle_frame[:, 3] = self.prev_actions[:, 3]
        self.prev_dist[:] = torch.norm(self.target_position - self.obs_dict["robot_position"], dim=1)
        self.actions = actions
        self.actions[:, 0:3] = 2.0 * self.actions[:, 0:3]  # :170, in place on the caller's tensor
