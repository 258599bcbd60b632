"""CPU restatement (torch) of two SURVEY 8(f) rows: the NavigationTask epilogue (f3) and the IMU (f4).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product path is
aerial_gym_simulator_b200/csrc/hp1_aux.cu behind the C ABI.  Pinned against fixtures produced by running
the reference's own code (tests/golden/make_golden_aux.py -> nav_task_epilogue.npz, imu_sensor.npz).
Random draws are inputs (the reference draws them with torch; call order documented per function)."""
import math

import torch

from . import hp1_oracle as O

NAV_PARAM_NAMES = (
    "pos_reward_magnitude", "pos_reward_exponent", "very_close_to_goal_reward_magnitude", "very_close_to_goal_reward_exponent",
    "getting_closer_reward_multiplier", "x_action_diff_penalty_magnitude", "x_action_diff_penalty_exponent",
    "z_action_diff_penalty_magnitude", "z_action_diff_penalty_exponent", "yawrate_action_diff_penalty_magnitude",
    "yawrate_action_diff_penalty_exponent", "x_absolute_action_penalty_magnitude", "x_absolute_action_penalty_exponent",
    "z_absolute_action_penalty_magnitude", "z_absolute_action_penalty_exponent", "yawrate_absolute_action_penalty_magnitude",
    "yawrate_absolute_action_penalty_exponent", "collision_penalty",
)  # config/task_config/navigation_task_config.py:30-49 (order fixed by the C ABI: AgxNavRewardParams)


def _exp_reward(mag, exponent, value):  # navigation_task.py:420-425
    return mag * torch.exp(-(value * value) * exponent)


def _exp_penalty(mag, exponent, value):  # navigation_task.py:428-433
    return mag * (torch.exp(-(value * value) * exponent) - 1.0)


def nav_compute_reward(pos_error, prev_pos_error, crashes, action, prev_action, curriculum_progress_fraction, p):
    """navigation_task.py:436-521.  p: dict name -> float (NAV_PARAM_NAMES)."""
    mult = 1.0 + 2.0 * curriculum_progress_fraction
    dist = torch.norm(pos_error, dim=1)
    prev_dist = torch.norm(prev_pos_error, dim=1)
    pos_reward = _exp_reward(p["pos_reward_magnitude"], p["pos_reward_exponent"], dist)
    close_reward = _exp_reward(p["very_close_to_goal_reward_magnitude"], p["very_close_to_goal_reward_exponent"], dist)
    closer = prev_dist - dist
    k = p["getting_closer_reward_multiplier"]
    closer_reward = torch.where(closer > 0, k * closer, 2.0 * k * closer)
    dist_reward = (20.0 - dist) / 20.0
    diff = action - prev_action
    diff_pen = (_exp_penalty(p["x_action_diff_penalty_magnitude"], p["x_action_diff_penalty_exponent"], diff[:, 0])
                + _exp_penalty(p["z_action_diff_penalty_magnitude"], p["z_action_diff_penalty_exponent"], diff[:, 2])
                + _exp_penalty(p["yawrate_action_diff_penalty_magnitude"], p["yawrate_action_diff_penalty_exponent"], diff[:, 3]))
    f = curriculum_progress_fraction
    abs_pen = (f * _exp_penalty(p["x_absolute_action_penalty_magnitude"], p["x_absolute_action_penalty_exponent"], action[:, 0])
               + f * _exp_penalty(p["z_absolute_action_penalty_magnitude"], p["z_absolute_action_penalty_exponent"], action[:, 2])
               + f * _exp_penalty(p["yawrate_absolute_action_penalty_magnitude"], p["yawrate_absolute_action_penalty_exponent"], action[:, 3]))
    reward = mult * (pos_reward + close_reward + closer_reward + dist_reward) + (diff_pen + abs_pen)
    return torch.where(crashes > 0, torch.full_like(reward, p["collision_penalty"]), reward)


def nav_pos_error(vehicle_orientation, target, position):
    """navigation_task.py:405-408."""
    return O.quat_rotate_inverse(vehicle_orientation, target - position)


def nav_process_obs(vehicle_orientation, position, target, euler, body_linvel, body_angvel, robot_actions, u_vec, u_euler):
    """navigation_task.py:369-395, columns 0..16 (the VAE latents 17.. stay in torch).
    u_vec, u_euler: the two torch.rand_like draws in call order.  Note :374: the -0.5 sits INSIDE rand_like,
    so the position perturbation is 0.2 * U[0,1), not centred."""
    vec = O.quat_rotate_inverse(vehicle_orientation, target - position)
    pert = vec + 0.1 * 2 * u_vec
    dist = torch.norm(vec, dim=-1)
    obs = torch.zeros(position.shape[0], 17, dtype=position.dtype)
    obs[:, 0:3] = pert / dist.unsqueeze(1)
    obs[:, 3] = dist
    e = O.ssa(euler) + 0.1 * (u_euler - 0.5)
    obs[:, 4], obs[:, 5], obs[:, 6] = e[:, 0], e[:, 1], 0.0
    obs[:, 7:10], obs[:, 10:13], obs[:, 13:17] = body_linvel, body_angvel, robot_actions
    return obs


def imu_update(force, mass, robot_orientation, body_angvel, sensor_quats, gravity, world_frame, gravity_compensation, bias, n_noise,
               n_bias, imu_noise_std, bias_std, max_meas, dt, enable_noise=True, enable_bias=True):
    """sensors/imu_sensor.py:85-131.  n_noise, n_bias: standard-normal draws [N,6] in the reference's call order
    (sample_noise :74-77 first, update_bias :79-83 second).  Returns (imu_meas [N,6], new bias)."""
    sqrt_dt = math.sqrt(dt)
    g_world = gravity * (1 - int(gravity_compensation))
    accel_t = force[:, 0:3] / mass.unsqueeze(1)
    q_ws = O.quat_mul(robot_orientation, sensor_quats)
    if world_frame:
        acc = O.quat_rotate_inverse(q_ws, accel_t - g_world)
        rate = O.quat_rotate_inverse(q_ws, body_angvel)
    else:
        acc = O.quat_rotate_inverse(sensor_quats, accel_t) - O.quat_rotate_inverse(q_ws, g_world.expand_as(accel_t))
        rate = O.quat_rotate_inverse(sensor_quats, body_angvel)
    noise = n_noise * imu_noise_std / sqrt_dt
    bias = bias + n_bias * bias_std * sqrt_dt
    a = acc + int(enable_bias) * bias[:, :3] + int(enable_noise) * noise[:, :3]
    w = rate + int(enable_bias) * bias[:, 3:] + int(enable_noise) * noise[:, 3:]
    a = torch.max(torch.min(a, max_meas[0:3]), -max_meas[0:3])  # tensor_clamp, utils/math.py
    w = torch.max(torch.min(w, max_meas[3:6]), -max_meas[3:6])
    return torch.cat([a, w], dim=1), bias
