"""PositionSetpointTaskSim2RealEndToEnd / PositionSetpointTaskSim2RealPX4
(task/position_setpoint_task_sim2real_end_to_end/position_setpoint_task_sim2real_end_to_end.py:20-311 and
task/position_setpoint_task_sim2real_px4/position_setpoint_task_sim2real_px4.py): the policy commands the four motor thrusts
directly (robot tinyprop / x500, controller no_control).  Same surface and step order as the reference --

    motor commands (process_actions_for_task) -> sim_env.step (n physics sub-steps fused in one launch) -> reward + crashes
    (agx_e2e_reward) -> truncations -> post_reward_calculation_step (reset) -> reset_idx (the reference resets those envs a SECOND
    time, :146-156: kept) -> observation (agx_e2e_obs: noisy position error, rotation-6D of the noisy ZYX Euler angles, noisy world
    velocity, noisy body rates) -> prev_actions / prev_pos_error bookkeeping

-- with the two epilogues as C-ABI kernel launches.  The four torch.normal draws of process_obs_for_task keep the reference's order."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..sim import SimBuilder
from .base_task import BaseTask
from .spaces import Box, Dict


class PositionSetpointTaskSim2RealEndToEnd(BaseTask):
    # AgxE2ERewardParams: the constants of compute_reward (:267-311)
    REWARD_CONSTANTS = dict(z_error_scale=11.0, upright_gain2=0.0, upright_exp2=0.0, align_gain1=6.0, align_exp1=5.0, align_gain2=0.0,
                            align_exp2=0.0, angvel_gain=0.3, hover_thrust=9.81 * 0.372 / 4, towards_gain_pos=10.0, towards_gain_neg=15.0,
                            action_diff_gain=1.3)
    RESET_PREV_POS_ERROR = False

    def __init__(self, task_config, seed=None, num_envs=None, headless=None, device=None, use_warp=None):
        for k, v in (("seed", seed), ("num_envs", num_envs), ("headless", headless), ("device", device), ("use_warp", use_warp)):
            if v is not None:
                setattr(task_config, k, v)
        super().__init__(task_config)
        self.device = torch.device(self.task_config.device)
        self.lib = _lib.load()
        self._params = _lib.AgxE2ERewardParams()
        for k, v in self.REWARD_CONSTANTS.items():
            setattr(self._params, k, float(v))
        self._params.crash_dist = float(self.task_config.crash_dist)
        args = dict(self.task_config.args or {})
        args.setdefault("seed", self._seed)
        self.sim_env = SimBuilder().build_env(
            sim_name=self.task_config.sim_name, env_name=self.task_config.env_name, robot_name=self.task_config.robot_name,
            controller_name=self.task_config.controller_name, args=args, device=self.device, num_envs=self.task_config.num_envs,
            use_warp=self.task_config.use_warp, headless=self.task_config.headless)
        N, dev, A = self.sim_env.num_envs, self.device, self.task_config.action_space_dim
        self.num_envs = N
        T = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev)
        self.action_limit_min, self.action_limit_max = T(self.task_config.action_limit_min), T(self.task_config.action_limit_max)
        self.actions = torch.zeros((N, A), device=dev)
        self.prev_actions = torch.zeros_like(self.actions)
        self.action_history = torch.zeros((N, A * 10), device=dev)
        self.counter = 0
        self.target_position = torch.zeros((N, 3), device=dev)
        self.obs_dict = self.sim_env.get_obs()
        self.obs_dict["num_obstacles_in_env"] = 1
        self.terminations = self.obs_dict["crashes"]
        self.truncations = self.obs_dict["truncations"]
        self.rewards = torch.zeros(N, device=dev)
        self.prev_position = torch.zeros((N, 3), device=dev)
        self.prev_pos_error = torch.zeros((N, 3), device=dev)
        self.observation_space = Dict({"observations": Box(low=-1.0, high=1.0, shape=(self.task_config.observation_space_dim,), dtype=np.float32)})
        self.action_space = Box(low=-1.0, high=1.0, shape=(A,), dtype=np.float32)
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
        self.action_history[env_ids] = 0.0
        if self.RESET_PREV_POS_ERROR:  # position_setpoint_task_sim2real_px4.py:123
            self.prev_pos_error[env_ids] = 0.0

    def render(self):
        return None

    def handle_action_history(self, actions):
        A = self.task_config.action_space_dim
        old = self.action_history.clone()
        self.action_history[:, A:] = old[:, :-A]
        self.action_history[:, :A] = actions

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def compute_rewards_and_crashes(self, obs_dict):
        """:232-251 + compute_reward :267-311, one launch; crashes are OR-ed in place."""
        st = obs_dict["robot_state_tensor"]
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.agx_e2e_reward(self.num_envs, p(st), st.stride(0), p(obs_dict["robot_body_angvel"]), p(self.target_position),
                                           p(self.actions), p(self.prev_actions), p(self.prev_pos_error), C.byref(self._params),
                                           p(obs_dict["crashes"]), p(self.rewards), self._stream()), "agx_e2e_reward")
        return self.rewards, obs_dict["crashes"]

    def step(self, actions):
        self.counter += 1
        self.actions[:] = self.task_config.process_actions_for_task(actions, self.action_limit_min, self.action_limit_max)
        self.prev_position[:] = self.obs_dict["robot_position"]
        self.sim_env.step(actions=self.actions)
        self.compute_rewards_and_crashes(self.obs_dict)
        if self.task_config.return_state_before_reset:
            return_tuple = self.get_return_tuple()
        self.truncations[:] = self.sim_env.sim_steps > self.task_config.episode_len_steps
        reset_envs = self.sim_env.post_reward_calculation_step()
        if len(reset_envs) > 0:
            self.reset_idx(reset_envs)
        self.infos = {}
        if not self.task_config.return_state_before_reset:
            return_tuple = self.get_return_tuple()
        self.prev_actions[:] = self.actions
        torch.sub(self.target_position, self.obs_dict["robot_position"], out=self.prev_pos_error)
        return return_tuple

    def get_return_tuple(self):
        self.process_obs_for_task()
        return (self.task_obs, self.rewards, self.terminations, self.truncations, self.infos)

    def process_obs_for_task(self, noise=None):
        """:204-229; noise [N,12] = the four torch.normal draws (position, orientation, linear velocity, body rates) unless given."""
        od, N, dev = self.obs_dict, self.num_envs, self.device
        if noise is None:
            z = torch.zeros((N, 3), device=dev)
            noise = torch.cat([torch.normal(mean=z, std=0.001), torch.normal(mean=z, std=torch.pi / 1032), torch.normal(mean=z, std=0.002),
                               torch.normal(mean=z, std=0.001)], dim=1)
        st, obs = od["robot_state_tensor"], self.task_obs["observations"]
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.agx_e2e_obs(N, p(st), st.stride(0), p(od["robot_body_angvel"]), p(self.target_position), p(noise), p(obs),
                                        obs.stride(0), self._stream()), "agx_e2e_obs")
        self.task_obs["rewards"] = self.rewards
        self.task_obs["terminations"] = self.terminations
        self.task_obs["truncations"] = self.truncations


class PositionSetpointTaskSim2RealPX4(PositionSetpointTaskSim2RealEndToEnd):
    """task/position_setpoint_task_sim2real_px4/position_setpoint_task_sim2real_px4.py: same task on the x500, other reward constants
    (:268-312) and prev_pos_error zeroed on reset (:123)."""
    REWARD_CONSTANTS = dict(z_error_scale=13.0, upright_gain2=2.5, upright_exp2=2.0, align_gain1=4.0, align_exp1=5.0, align_gain2=2.0,
                            align_exp2=2.0, angvel_gain=0.75, hover_thrust=9.81 * 1.6559999883174896 / 4, towards_gain_pos=50.0,
                            towards_gain_neg=100.0, action_diff_gain=0.5)
    RESET_PREV_POS_ERROR = True
