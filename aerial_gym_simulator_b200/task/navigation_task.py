"""NavigationTask (task/navigation_task/navigation_task.py:21-521): same surface and step order as the
reference --

    transformed action -> sim_env.step (n physics sub-steps, collision check after each) ->
    reward + crashes (agx_nav_reward) -> truncations, successes / timeouts, curriculum ->
    post_reward_calculation_step (reset + sensor render) -> VAE latents (torch) ->
    observation assembly (agx_nav_obs)

-- with the two per-env epilogues as C-ABI kernel launches instead of ~60 torch ops.  Random numbers
(target ratios, observation perturbations) are drawn with torch in the reference's call order."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..sim import SimBuilder
from ..utils.logging import CustomLogger
from ..utils.math import torch_interpolate_ratio, torch_rand_float_tensor
from ..utils.vae_encoder import VAEImageEncoder
from .base_task import BaseTask
from .spaces import Box, Dict

logger = CustomLogger("navigation_task")

_PARAM_ORDER = (
    "pos_reward_magnitude", "pos_reward_exponent", "very_close_to_goal_reward_magnitude", "very_close_to_goal_reward_exponent",
    "getting_closer_reward_multiplier", "x_action_diff_penalty_magnitude", "x_action_diff_penalty_exponent",
    "z_action_diff_penalty_magnitude", "z_action_diff_penalty_exponent", "yawrate_action_diff_penalty_magnitude",
    "yawrate_action_diff_penalty_exponent", "x_absolute_action_penalty_magnitude", "x_absolute_action_penalty_exponent",
    "z_absolute_action_penalty_magnitude", "z_absolute_action_penalty_exponent", "yawrate_absolute_action_penalty_magnitude",
    "yawrate_absolute_action_penalty_exponent", "collision_penalty",
)


class NavigationTask(BaseTask):
    def __init__(self, task_config, seed=None, num_envs=None, headless=None, device=None, use_warp=None):
        for k, v in (("seed", seed), ("num_envs", num_envs), ("headless", headless), ("device", device), ("use_warp", use_warp)):
            if v is not None:
                setattr(task_config, k, v)
        super().__init__(task_config)
        self.device = torch.device(self.task_config.device)
        self.lib = _lib.load()
        self._params = _lib.AgxNavRewardParams()
        for i, name in enumerate(_PARAM_ORDER):
            self._params.v[i] = float(self.task_config.reward_parameters[name])
        args = dict(self.task_config.args or {})
        world, rank, _ = self.shard_spec()
        if world > 1:  # env-sharded: global env ids rank * num_envs ... (keys the device RNG)
            args.setdefault("env_id_offset", rank * int(self.task_config.num_envs))
        self.sim_env = SimBuilder().build_env(
            sim_name=self.task_config.sim_name, env_name=self.task_config.env_name, robot_name=self.task_config.robot_name,
            controller_name=self.task_config.controller_name, args=args, device=self.device,
            num_envs=self.task_config.num_envs, use_warp=self.task_config.use_warp, headless=self.task_config.headless)
        N, dev = self.sim_env.num_envs, self.device
        self.num_envs = N
        self.target_position = torch.zeros((N, 3), device=dev)
        T = lambda v: torch.tensor(v, dtype=torch.float32, device=dev).expand(N, -1)
        self.target_min_ratio, self.target_max_ratio = T(self.task_config.target_min_ratio), T(self.task_config.target_max_ratio)
        self.success_aggregate = self.crashes_aggregate = self.timeouts_aggregate = 0
        self.pos_error_vehicle_frame_prev = torch.zeros_like(self.target_position)
        self.pos_error_vehicle_frame = torch.zeros_like(self.target_position)
        vc = self.task_config.vae_config
        if vc.use_vae:
            self.vae_model = VAEImageEncoder(config=vc, device=dev)
            self.image_latents = torch.zeros((N, vc.latent_dims), device=dev)
        else:
            self.vae_model = lambda x: x
        self.obs_dict = self.sim_env.get_obs()
        cur = self.task_config.curriculum
        self.curriculum_level = self.obs_dict.get("curriculum_level", cur.min_level)
        self.obs_dict["curriculum_level"] = self.curriculum_level
        self.obs_dict["num_obstacles_in_env"] = self.curriculum_level
        self.curriculum_progress_fraction = (self.curriculum_level - cur.min_level) / (cur.max_level - cur.min_level)
        self.terminations = self.obs_dict["crashes"]
        self.truncations = self.obs_dict["truncations"]
        self.rewards = torch.zeros(N, device=dev)
        self.observation_space = Dict({"observations": Box(low=-1.0, high=1.0, shape=(self.task_config.observation_space_dim,), dtype=np.float32)})
        self.action_space = Box(low=-1.0, high=1.0, shape=(4,), dtype=np.float32)
        self.action_transformation_function = self.task_config.action_transformation_function
        self.task_obs = {"observations": torch.zeros((N, self.task_config.observation_space_dim), device=dev)}
        self.infos = {}
        self.num_task_steps = 0
        if self.init_sharding(N, self.task_config.observation_space_dim, dev) is not None:  # BaseTask: multi-GPU, observation all-gather
            self.task_obs["observations_local"] = self.task_obs["observations"]

    def close(self):
        self.sim_env.delete_env()

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        return self.get_return_tuple()

    def reset_idx(self, env_ids):
        ratio = torch_rand_float_tensor(self.target_min_ratio, self.target_max_ratio)  # full-N draw, navigation_task.py:167
        self.target_position[env_ids] = torch_interpolate_ratio(
            min=self.obs_dict["env_bounds_min"][env_ids], max=self.obs_dict["env_bounds_max"][env_ids], ratio=ratio[env_ids])
        self.infos = {}

    def render(self):
        return self.sim_env.render()

    def check_and_update_curriculum_level(self, successes, crashes, timeouts):
        """navigation_task.py:234-273."""
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

    def process_image_observation(self):
        if self.task_config.vae_config.use_vae and "depth_range_pixels" in self.obs_dict:
            self.image_latents[:] = self.vae_model.encode(self.obs_dict["depth_range_pixels"].squeeze(1))

    def post_image_reward_addition(self):
        """navigation_task.py:351-357.  The mask `terminations < 0` on a bool tensor is never true, so the
        reference adds nothing here; only min_pixel_dist is kept for inspection."""
        if "depth_range_pixels" not in self.obs_dict:
            return
        image_obs = 10.0 * self.obs_dict["depth_range_pixels"].squeeze(1)
        image_obs[image_obs < 0] = 10.0
        self.min_pixel_dist = torch.amin(image_obs, dim=(1, 2))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def compute_rewards_and_crashes(self, obs_dict):
        """navigation_task.py:397-418 + compute_reward :436-521, one launch."""
        st = obs_dict["robot_state_tensor"]
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.agx_nav_reward(
            self.num_envs, p(st), st.stride(0), p(obs_dict["robot_vehicle_orientation"]), p(self.target_position), p(obs_dict["crashes"]),
            p(obs_dict["robot_actions"]), p(obs_dict["robot_prev_actions"]), float(self.curriculum_progress_fraction),
            C.byref(self._params), p(self.pos_error_vehicle_frame), p(self.pos_error_vehicle_frame_prev), p(self.rewards),
            self._stream()), "agx_nav_reward")
        return self.rewards, obs_dict["crashes"]

    def step(self, actions):
        transformed_action = self.action_transformation_function(actions)
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
        self.post_image_reward_addition()
        if not self.task_config.return_state_before_reset:
            return_tuple = self.get_return_tuple()
        return return_tuple

    def get_return_tuple(self):
        self.process_obs_for_task()
        if self.obs_gather is not None:  # sharded: the policy gets the global [W * N, D] observation tensor
            self.task_obs["observations"], _ = self.gather_observations(self.task_obs["observations_local"])
        return (self.task_obs, self.rewards, self.terminations, self.truncations, self.infos)

    def process_obs_for_task(self, u_vec=None, u_euler=None):
        """navigation_task.py:369-395; the two rand_like draws in the reference's order unless given."""
        od, N, dev = self.obs_dict, self.num_envs, self.device
        if u_vec is None:
            u_vec = torch.rand((N, 3), device=dev)
        if u_euler is None:
            u_euler = torch.rand((N, 3), device=dev)
        st, obs = od["robot_state_tensor"], (self.task_obs["observations_local"] if self.obs_gather is not None else self.task_obs["observations"])
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.agx_nav_obs(
            N, p(st), st.stride(0), p(od["robot_vehicle_orientation"]), p(od["robot_euler_angles"]), p(od["robot_body_linvel"]),
            p(od["robot_body_angvel"]), p(od["robot_actions"]), p(self.target_position), p(u_vec), p(u_euler), p(obs), obs.stride(0),
            self._stream()), "agx_nav_obs")
        if self.task_config.vae_config.use_vae:
            obs[:, 17:] = self.image_latents
