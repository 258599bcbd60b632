"""config/task_config/position_setpoint_task_config.py"""


class position_setpoint_task_config:
    seed = 1
    sim_name = "base_sim"
    env_name = "empty_env"
    robot_name = "base_quadrotor"
    controller_name = "lee_attitude_control"
    args = {}
    num_envs = 4096
    use_warp = False
    headless = False
    device = "cuda:0"
    observation_space_dim = 13
    privileged_observation_space_dim = 0
    action_space_dim = 4
    episode_len_steps = 500
    return_state_before_reset = False
    reward_parameters = {
        "pos_error_gain1": [2.0, 2.0, 2.0],
        "pos_error_exp1": [1 / 3.5, 1 / 3.5, 1 / 3.5],
        "pos_error_gain2": [2.0, 2.0, 2.0],
        "pos_error_exp2": [2.0, 2.0, 2.0],
        "dist_reward_coefficient": 7.5,
        "max_dist": 15.0,
        "action_diff_penalty_gain": [1.0, 1.0, 1.0],
        "absolute_action_reward_gain": [2.0, 2.0, 2.0],
        "crash_penalty": -100,
    }


class position_setpoint_task_lmf2_config(position_setpoint_task_config):
    """config/task_config/position_setpoint_task_lmf2_config.py (not registered as a task by the reference either)"""
    robot_name, controller_name, num_envs = "lmf2", "lmf2_velocity_control", 16


class navigation_task_config:
    """config/task_config/navigation_task_config.py"""
    seed = -1
    sim_name = "base_sim"
    env_name = "env_with_obstacles"
    robot_name = "lmf2"
    controller_name = "lmf2_velocity_control"
    args = {}
    num_envs = 1024
    use_warp = True
    headless = True
    device = "cuda:0"
    observation_space_dim = 13 + 4 + 64  # root_state + action_dim + latent_dims
    privileged_observation_space_dim = 0
    action_space_dim = 4
    episode_len_steps = 100
    return_state_before_reset = False
    target_min_ratio = [0.90, 0.1, 0.1]
    target_max_ratio = [0.94, 0.90, 0.90]
    reward_parameters = {
        "pos_reward_magnitude": 5.0,
        "pos_reward_exponent": 1.0 / 3.5,
        "very_close_to_goal_reward_magnitude": 5.0,
        "very_close_to_goal_reward_exponent": 2.0,
        "getting_closer_reward_multiplier": 10.0,
        "x_action_diff_penalty_magnitude": 0.8,
        "x_action_diff_penalty_exponent": 3.333,
        "z_action_diff_penalty_magnitude": 0.8,
        "z_action_diff_penalty_exponent": 5.0,
        "yawrate_action_diff_penalty_magnitude": 0.8,
        "yawrate_action_diff_penalty_exponent": 3.33,
        "x_absolute_action_penalty_magnitude": 0.1,
        "x_absolute_action_penalty_exponent": 0.3,
        "z_absolute_action_penalty_magnitude": 1.5,
        "z_absolute_action_penalty_exponent": 1.0,
        "yawrate_absolute_action_penalty_magnitude": 1.5,
        "yawrate_absolute_action_penalty_exponent": 2.0,
        "collision_penalty": -100.0,
    }

    class vae_config:
        use_vae = True
        latent_dims = 64
        # the reference ships its weights inside its own tree (utils/vae/weights/, 44 MB, not redistributed here):
        # point model_file at that checkpoint; without it the encoder runs with its initial weights
        model_file = "ICRA_test_set_more_sim_data_kld_beta_3_LD_64_epoch_49.pth"
        model_folder = ""
        image_res = (270, 480)
        interpolation_mode = "nearest"
        return_sampled_latent = True

    class curriculum:
        min_level = 15
        max_level = 50
        check_after_log_instances = 2048
        increase_step = 2
        decrease_step = 1
        success_rate_for_increase = 0.7
        success_rate_for_decrease = 0.6

    @staticmethod
    def action_transformation_function(action):
        """3-D policy action -> [vx, 0, vz, yaw_rate] (navigation_task_config.py:87-117)."""
        import math

        import torch

        clamped = torch.clamp(action, -1.0, 1.0)
        max_speed, max_yawrate, max_incl = 2.0, math.pi / 3, math.pi / 4
        clamped[:, 0] += 1.0
        out = torch.zeros((clamped.shape[0], 4), device=action.device)
        out[:, 0] = clamped[:, 0] * torch.cos(max_incl * clamped[:, 1]) * max_speed / 2.0
        out[:, 2] = clamped[:, 0] * torch.sin(max_incl * clamped[:, 1]) * max_speed / 2.0
        out[:, 3] = clamped[:, 2] * max_yawrate
        return out


class lidar_navigation_task_config:
    """config/task_config/lidar_navigation_task_config.py"""
    seed = -1
    sim_name = "base_sim"
    env_name = "env_with_lidar_nav_obstacles"
    robot_name = "magpie"
    controller_name = "magpie_acceleration_control"
    args = {}
    num_envs = 512
    use_warp = True
    headless = False
    device = "cuda:0"
    observation_space_dim = 13 + 4 + 16 * 20  # root_state + action_dim + min-pooled 48x120 LiDAR image
    privileged_observation_space_dim = 0
    action_space_dim = 4
    episode_len_steps = 110
    return_state_before_reset = False
    target_min_ratio = [0.90, 0.15, 0.15]
    target_max_ratio = [0.92, 0.80, 0.80]
    # the reference task hard-codes these in process_image_observation (lidar_navigation_task.py:320-347)
    lidar_pool_window = (3, 6)
    lidar_max_range, lidar_min_range, lidar_invalid_value, time_to_collision_max = 10.0, 0.2, 10.0, 10.0
    reward_parameters = {
        "pos_reward_magnitude": 3.0,
        "pos_reward_exponent": 1.0,
        "very_close_to_goal_reward_magnitude": 5.0,
        "very_close_to_goal_reward_exponent": 8.0,
        "vel_direction_component_reward_magnitude": 1.0,
        "x_action_diff_penalty_magnitude": 0.3,
        "x_action_diff_penalty_exponent": 5.0,
        "y_action_diff_penalty_magnitude": 0.3,
        "y_action_diff_penalty_exponent": 5.0,
        "z_action_diff_penalty_magnitude": 0.3,
        "z_action_diff_penalty_exponent": 5.0,
        "yawrate_action_diff_penalty_magnitude": 0.3,
        "yawrate_action_diff_penalty_exponent": 5.0,
        "x_absolute_action_penalty_magnitude": 0.1,
        "x_absolute_action_penalty_exponent": 0.3,
        "y_absolute_action_penalty_magnitude": 0.1,
        "y_absolute_action_penalty_exponent": 0.3,
        "z_absolute_action_penalty_magnitude": 0.15,
        "z_absolute_action_penalty_exponent": 1.0,
        "yawrate_absolute_action_penalty_magnitude": 0.15,
        "yawrate_absolute_action_penalty_exponent": 2.0,
        "collision_penalty": -10.0,
    }

    class vae_config:
        use_vae = False
        latent_dims = 64
        model_file = "ICRA_test_set_more_sim_data_kld_beta_3_LD_64_epoch_49.pth"
        model_folder = ""
        image_res = (270, 480)
        interpolation_mode = "nearest"
        return_sampled_latent = True

    class curriculum:
        min_level = 25
        max_level = 70
        check_after_log_instances = 2048
        increase_step = 2
        decrease_step = 1
        success_rate_for_increase = 0.7
        success_rate_for_decrease = 0.6

    @staticmethod
    def action_transformation_function(action):
        """4-D policy action -> [2 ax, 2 ay, 2 az, yaw_rate] (lidar_navigation_task_config.py:105-115)."""
        import math

        import torch

        clamped = torch.clamp(action, -1.0, 1.0)
        out = torch.zeros((clamped.shape[0], 4), device=action.device)
        out[:, 0:3] = 2 * clamped[:, 0:3]
        out[:, 3] = clamped[:, 3] * (math.pi / 3)
        return out


def _rescale_motor_commands(actions, min_limit, max_limit):
    """process_actions_for_task of the motor-command tasks (position_setpoint_task_sim2real_end_to_end_config.py:27-32)."""
    import torch

    actions_clipped = torch.clamp(actions, -1, 1)
    return actions_clipped * (max_limit - min_limit) / 2 + (max_limit + min_limit) / 2


class position_setpoint_task_sim2real_end_to_end_config:
    """config/task_config/position_setpoint_task_sim2real_end_to_end_config.py (EVAL = False).  The action limits are plain lists
    here (the reference builds CUDA tensors at import time); the task turns them into tensors on its device."""
    seed = 56
    sim_name = "base_sim"
    env_name = "empty_env"
    robot_name = "tinyprop"
    controller_name = "no_control"
    args = {}
    num_envs = 4096
    use_warp = False
    headless = True
    device = "cuda:0"
    privileged_observation_space_dim = 0
    action_space_dim = 4
    observation_space_dim = 15
    episode_len_steps = 600
    return_state_before_reset = False
    reward_parameters = {}
    crash_dist = 1.5
    action_limit_max = [1.2] * 4
    action_limit_min = [0.2] * 4
    process_actions_for_task = staticmethod(_rescale_motor_commands)


class position_setpoint_task_sim2real_px4_config(position_setpoint_task_sim2real_end_to_end_config):
    """config/task_config/position_setpoint_task_sim2real_px4_config.py (EVAL = False)"""
    robot_name = "x500"
    num_envs = 24
    episode_len_steps = 500
    crash_dist = 6.5
    action_limit_max = [8.0] * 4
    action_limit_min = [0.0] * 4


class radar_navigation_task_config(lidar_navigation_task_config):
    """config/task_config/radar_navigation_task_config.py: the LiDAR-navigation config on lmf2_radar in env_with_obstacles"""
    env_name = "env_with_obstacles"
    robot_name = "lmf2_radar"
    controller_name = "lmf2_acceleration_control"


class position_setpoint_task_sim2real_config:
    """config/task_config/position_setpoint_task_sim2real_config.py"""
    seed = 1
    sim_name = "base_sim"
    env_name = "empty_env"
    robot_name = "lmf2"
    controller_name = "lmf2_velocity_control"
    args = {}
    num_envs = 16
    use_warp = False
    headless = False
    device = "cuda:0"
    observation_space_dim = 17
    privileged_observation_space_dim = 0
    action_space_dim = 4
    episode_len_steps = 800
    return_state_before_reset = False
    reward_parameters = {}


class position_setpoint_task_acceleration_sim2real_config(position_setpoint_task_sim2real_config):
    """config/task_config/position_setpoint_task_acceleration_sim2real_config.py"""
    controller_name = "lmf2_acceleration_control"
