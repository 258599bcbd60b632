"""LiDARNavigationTask (task/lidar_navigation_task/lidar_navigation_task.py:21-499): same surface and step order as
the reference --

    prev/current action bookkeeping -> transformed action -> sim_env.step (n physics sub-steps, collision check after
    each) -> reward + crashes (agx_lidar_nav_reward, uses LAST step's time to collision) -> truncations, successes /
    timeouts, curriculum -> post_reward_calculation_step (reset + LiDAR render) -> process_image_observation
    (agx_lidar_nav_pool: point cloud -> clipped ranges -> 3x6 min-pool + time to collision; the task's own noise in
    torch; 1/x) -> observation assembly (agx_lidar_nav_obs)

-- with the three per-env / per-pixel epilogues as C-ABI kernel launches instead of ~90 torch ops and five
full-image temporaries.  Random numbers (target ratios, target yaw, observation perturbations, lidar noise) are drawn with
torch in the reference's call order."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..sim import SimBuilder
from ..utils.logging import CustomLogger
from ..utils.math import torch_interpolate_ratio, torch_rand_float_tensor
from .base_task import BaseTask
from .spaces import Box, Dict

logger = CustomLogger("lidar_navigation_task")

_PARAM_ORDER = (
    "pos_reward_magnitude", "pos_reward_exponent", "very_close_to_goal_reward_magnitude", "very_close_to_goal_reward_exponent",
    "vel_direction_component_reward_magnitude",
    "x_action_diff_penalty_magnitude", "x_action_diff_penalty_exponent", "y_action_diff_penalty_magnitude", "y_action_diff_penalty_exponent",
    "z_action_diff_penalty_magnitude", "z_action_diff_penalty_exponent", "yawrate_action_diff_penalty_magnitude",
    "yawrate_action_diff_penalty_exponent",
    "x_absolute_action_penalty_magnitude", "x_absolute_action_penalty_exponent", "y_absolute_action_penalty_magnitude",
    "y_absolute_action_penalty_exponent", "z_absolute_action_penalty_magnitude", "z_absolute_action_penalty_exponent",
    "yawrate_absolute_action_penalty_magnitude", "yawrate_absolute_action_penalty_exponent", "collision_penalty",
)  # AgxLidarNavRewardParams.v (include/aerial_gym_b200.h)


def add_noise_to_downsampled_lidar_data(ds):
    """lidar_navigation_task.py:286-310, same torch calls in the same order (so a seeded run reproduces the reference's
    draws): 3 % of the pixels get +U(0.2, 10), 2 % become max range, 2 % of columns 10.. become U(0.2, 1).  In place."""
    noise_mask = torch.bernoulli(0.03 * torch.ones_like(ds))
    sel = noise_mask == 1
    ds[sel] += torch_rand_float_tensor(0.2 * torch.ones_like(noise_mask[sel]), 10.0 * torch.ones_like(noise_mask[sel]))
    max_range_mask = torch.bernoulli(0.02 * torch.ones_like(ds))
    ds[max_range_mask == 1] = 10.0
    low_mask = torch.bernoulli(0.02 * torch.ones_like(ds[:, 10:]))
    low = torch_rand_float_tensor(0.2 * torch.ones_like(low_mask), 1.0 * torch.ones_like(low_mask))
    ds[:, 10:][low_mask == 1] = low[low_mask == 1]
    return ds


class LiDARNavigationTask(BaseTask):
    def __init__(self, task_config, seed=None, num_envs=None, headless=None, device=None, use_warp=None):
        for k, v in (("seed", seed), ("num_envs", num_envs), ("headless", headless), ("device", device), ("use_warp", use_warp)):
            if v is not None:
                setattr(task_config, k, v)
        super().__init__(task_config)
        self.device = torch.device(self.task_config.device)
        self.lib = _lib.load()
        self._params = _lib.AgxLidarNavRewardParams()
        for i, name in enumerate(_PARAM_ORDER):
            self._params.v[i] = float(self.task_config.reward_parameters[name])
        self.sim_env = SimBuilder().build_env(
            sim_name=self.task_config.sim_name, env_name=self.task_config.env_name, robot_name=self.task_config.robot_name,
            controller_name=self.task_config.controller_name, args=self.task_config.args, device=self.device,
            num_envs=self.task_config.num_envs, use_warp=self.task_config.use_warp, headless=self.task_config.headless)
        N, dev = self.sim_env.num_envs, self.device
        self.num_envs = N
        self.target_position = torch.zeros((N, 3), device=dev)
        T = lambda v: torch.tensor(v, dtype=torch.float32, device=dev).expand(N, -1)
        self.target_min_ratio, self.target_max_ratio = T(self.task_config.target_min_ratio), T(self.task_config.target_max_ratio)
        self.success_aggregate = self.crashes_aggregate = self.timeouts_aggregate = 0
        self.pos_error_vehicle_frame_prev = torch.zeros_like(self.target_position)
        self.pos_error_vehicle_frame = torch.zeros_like(self.target_position)
        self.current_action = torch.zeros((N, 4), device=dev)
        self.prev_action = torch.zeros((N, 4), device=dev)
        self.time_to_collision = torch.zeros(N, device=dev)
        self.target_yaw = torch.zeros(N, device=dev)
        self.obs_dict = self.sim_env.get_obs()
        cur = self.task_config.curriculum
        self.curriculum_level = self.obs_dict.get("curriculum_level", cur.min_level)
        self.obs_dict["curriculum_level"] = self.curriculum_level
        self.obs_dict["num_obstacles_in_env"] = self.curriculum_level
        self.curriculum_progress_fraction = (self.curriculum_level - cur.min_level) / (cur.max_level - cur.min_level)
        self.terminations = self.obs_dict["crashes"]
        self.truncations = self.obs_dict["truncations"]
        self.rewards = torch.zeros(N, device=dev)
        # LiDAR image geometry: the reference hard-codes a 48 x 120 cloud and a 16 x 20 pooled image (:83-85, :123-125)
        px = self.obs_dict.get("depth_range_pixels")
        if px is None or px.dim() != 5 or px.shape[1] != 1 or px.shape[-1] != 3:
            raise ValueError("LiDARNavigationTask needs a robot with one LiDAR that returns a point cloud "
                             "(depth_range_pixels [N,1,H,W,3]; e.g. magpie + RSLidar_Airy_Config)")
        self._H, self._W = int(px.shape[2]), int(px.shape[3])
        self._ph, self._pw = (int(v) for v in getattr(self.task_config, "lidar_pool_window", (3, 6)))
        self._OH, self._OW = self._H // self._ph, self._W // self._pw
        L = self._OH * self._OW
        if self.task_config.observation_space_dim != 17 + L:
            raise ValueError(f"observation_space_dim = {self.task_config.observation_space_dim}, but 13 + 4 + pooled LiDAR image "
                             f"({self._OH} x {self._OW}) = {17 + L}")
        self.downsampled_lidar_data = torch.zeros((N, L), device=dev)
        self._image_ds = torch.zeros((N, self._OH, self._OW), device=dev)
        self.observation_space = Dict({"observations": Box(low=-np.inf, high=np.inf, shape=(self.task_config.observation_space_dim,),
                                                           dtype=np.float32)})
        self.action_space = Box(low=-1.0, high=1.0, shape=(4,), dtype=np.float32)
        self.action_transformation_function = self.task_config.action_transformation_function
        self.task_obs = {"observations": torch.zeros((N, self.task_config.observation_space_dim), device=dev)}
        self.infos = {}
        self.num_task_steps = 0

    def close(self):
        self.sim_env.delete_env()

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        return self.get_return_tuple()

    def reset_idx(self, env_ids):
        """:164-181 -- full-N target-ratio draw, then a len(env_ids) target-yaw draw."""
        ratio = torch_rand_float_tensor(self.target_min_ratio, self.target_max_ratio)
        self.target_position[env_ids] = torch_interpolate_ratio(
            min=self.obs_dict["env_bounds_min"][env_ids], max=self.obs_dict["env_bounds_max"][env_ids], ratio=ratio[env_ids])
        self.obs_dict["robot_prev_actions"][env_ids] = 0.0
        k = len(env_ids)
        self.target_yaw[env_ids] = torch_rand_float_tensor(-torch.pi * torch.ones(k, device=self.device),
                                                           torch.pi * torch.ones(k, device=self.device))
        self.infos = {}

    def render(self):
        return self.sim_env.render()

    def check_and_update_curriculum_level(self, successes, crashes, timeouts):
        """:243-284."""
        cur = self.task_config.curriculum
        self.success_aggregate += torch.sum(successes)
        self.crashes_aggregate += torch.sum(crashes)
        self.timeouts_aggregate += torch.sum(timeouts)
        instances = self.success_aggregate + self.crashes_aggregate + self.timeouts_aggregate
        if instances >= cur.check_after_log_instances:
            success_rate = self.success_aggregate / instances
            if success_rate > cur.success_rate_for_increase:
                self.curriculum_level += cur.increase_step
            elif success_rate < cur.success_rate_for_decrease:
                self.curriculum_level -= cur.decrease_step
            self.curriculum_level = min(max(self.curriculum_level, cur.min_level), cur.max_level)
            self.obs_dict["curriculum_level"] = self.curriculum_level
            self.obs_dict["num_obstacles_in_env"] = self.curriculum_level
            self.curriculum_progress_fraction = (self.curriculum_level - cur.min_level) / (cur.max_level - cur.min_level)
            logger.warning(f"Curriculum Level: {self.curriculum_level}, Curriculum progress fraction: {self.curriculum_progress_fraction}")
            self.success_aggregate = self.crashes_aggregate = self.timeouts_aggregate = 0

    def add_noise_to_downsampled_lidar_data(self, ds_lidar_data):
        return add_noise_to_downsampled_lidar_data(ds_lidar_data)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def process_image_observation(self):
        """:313-362: one launch for everything up to the min-pooling, then the task's noise (torch RNG) and 1/x."""
        tc, od = self.task_config, self.obs_dict
        pc, st = od["depth_range_pixels"], od["robot_state_tensor"]
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.agx_lidar_nav_pool(
            self.num_envs, self._H, self._W, self._ph, self._pw, p(pc), p(st), st.stride(0),
            float(getattr(tc, "lidar_max_range", 10.0)), float(getattr(tc, "lidar_min_range", 0.2)),
            float(getattr(tc, "lidar_invalid_value", 10.0)), float(getattr(tc, "time_to_collision_max", 10.0)),
            p(self._image_ds), p(self.time_to_collision), self._stream()), "agx_lidar_nav_pool")
        noisy = self.add_noise_to_downsampled_lidar_data(self._image_ds)
        torch.reciprocal(noisy.view(self.num_envs, -1), out=self.downsampled_lidar_data)

    def compute_rewards_and_crashes(self, obs_dict):
        """:471-499 + compute_reward :554-720, one launch."""
        st = obs_dict["robot_state_tensor"]
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.agx_lidar_nav_reward(
            self.num_envs, p(st), st.stride(0), p(obs_dict["robot_vehicle_orientation"]), p(self.target_position),
            p(obs_dict["robot_euler_angles"]), p(self.target_yaw), p(obs_dict["robot_vehicle_linvel"]), p(obs_dict["robot_body_angvel"]),
            p(obs_dict["crashes"]), p(self.current_action), p(self.prev_action), p(self.time_to_collision),
            float(self.curriculum_progress_fraction), C.byref(self._params), p(self.pos_error_vehicle_frame),
            p(self.pos_error_vehicle_frame_prev), p(self.rewards), self._stream()), "agx_lidar_nav_reward")
        return self.rewards, obs_dict["crashes"]

    def step(self, actions):
        self.prev_action[:] = self.current_action
        transformed_action = self.action_transformation_function(actions)
        self.current_action[:] = transformed_action
        self.sim_env.step(actions=transformed_action)
        self.compute_rewards_and_crashes(self.obs_dict)  # rewards / terminations are written in place
        if self.task_config.return_state_before_reset:
            return_tuple = self.get_return_tuple()
        self.truncations[:] = self.sim_env.sim_steps > self.task_config.episode_len_steps
        successes = self.truncations * (torch.norm(self.target_position - self.obs_dict["robot_position"], dim=1) < 1.0)
        successes = torch.where(self.terminations > 0, torch.zeros_like(successes), successes)
        timeouts = torch.where(self.truncations > 0, torch.logical_not(successes), torch.zeros_like(successes))
        timeouts = torch.where(self.terminations > 0, torch.zeros_like(timeouts), timeouts)
        self.infos["successes"], self.infos["timeouts"], self.infos["crashes"] = successes, timeouts, self.terminations
        self.check_and_update_curriculum_level(successes, self.terminations, timeouts)
        reset_envs = self.sim_env.post_reward_calculation_step()
        if len(reset_envs) > 0:
            self.reset_idx(reset_envs)
        self.num_task_steps += 1
        self.process_image_observation()
        if not self.task_config.return_state_before_reset:
            return_tuple = self.get_return_tuple()
        return return_tuple

    def get_return_tuple(self):
        self.process_obs_for_task()
        return (self.task_obs, self.rewards, self.terminations, self.truncations, self.infos)

    def process_obs_for_task(self, u_vec=None, u_euler=None):
        """:440-469; the two rand_like draws in the reference's order unless given."""
        od, N, dev = self.obs_dict, self.num_envs, self.device
        if u_vec is None:
            u_vec = torch.rand((N, 3), device=dev)
        if u_euler is None:
            u_euler = torch.rand((N, 3), device=dev)
        st, obs = od["robot_state_tensor"], self.task_obs["observations"]
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.agx_lidar_nav_obs(
            N, p(st), st.stride(0), p(od["robot_vehicle_orientation"]), p(od["robot_euler_angles"]), p(od["robot_body_linvel"]),
            p(od["robot_body_angvel"]), p(od["robot_actions"]), p(self.target_position), p(self.target_yaw), p(u_vec), p(u_euler),
            p(self.downsampled_lidar_data), self.downsampled_lidar_data.shape[1], p(obs), obs.stride(0), self._stream()),
            "agx_lidar_nav_obs")
