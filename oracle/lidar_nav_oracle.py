"""CPU restatement (torch) of the LiDARNavigationTask epilogue
(task/lidar_navigation_task/lidar_navigation_task.py of the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product path is
aerial_gym_simulator_b200/csrc/lidar_nav.cu behind the C ABI.  Pinned against fixtures produced by running the
reference's own code (tests/golden/make_golden_lidar_nav.py -> lidar_nav_task_epilogue.npz).  Random draws are
inputs (the reference draws them with torch; call order documented per function)."""
import torch

from . import hp1_oracle as O

LIDAR_NAV_PARAM_NAMES = (
    "pos_reward_magnitude", "pos_reward_exponent", "very_close_to_goal_reward_magnitude", "very_close_to_goal_reward_exponent",
    "vel_direction_component_reward_magnitude",
    "x_action_diff_penalty_magnitude", "x_action_diff_penalty_exponent", "y_action_diff_penalty_magnitude", "y_action_diff_penalty_exponent",
    "z_action_diff_penalty_magnitude", "z_action_diff_penalty_exponent", "yawrate_action_diff_penalty_magnitude",
    "yawrate_action_diff_penalty_exponent",
    "x_absolute_action_penalty_magnitude", "x_absolute_action_penalty_exponent", "y_absolute_action_penalty_magnitude",
    "y_absolute_action_penalty_exponent", "z_absolute_action_penalty_magnitude", "z_absolute_action_penalty_exponent",
    "yawrate_absolute_action_penalty_magnitude", "yawrate_absolute_action_penalty_exponent", "collision_penalty",
)  # config/task_config/lidar_navigation_task_config.py:28-51 (order fixed by the C ABI: AgxLidarNavRewardParams)


def _erf(mag, exponent, value):  # lidar_navigation_task.py:503-507
    return mag * torch.exp(-(value * value) * exponent)


def _epf(mag, exponent, value):  # lidar_navigation_task.py:511-515
    return mag * (torch.exp(-(value * value) * exponent) - 1.0)


def pool(pointcloud, robot_position, robot_linvel, pool_hw=(3, 6), max_range=10.0, min_range=0.2, invalid_value=10.0, ttc_max=10.0):
    """process_image_observation up to the min-pooling, lidar_navigation_task.py:313-347.
    pointcloud [N,H,W,3] world frame.  Returns (image_ds [N,H//ph,W//pw], time_to_collision [N])."""
    N = pointcloud.shape[0]
    dirs = pointcloud - robot_position.unsqueeze(1).unsqueeze(1)                    # :315
    rng = torch.norm(dirs, dim=-1)                                                  # :316
    flat = rng.view(N, -1)                                                          # :317 (a VIEW: sees the clipping below)
    unit = dirs.view(N, -1, 3) / (flat.unsqueeze(-1) + 1e-6)                        # :318 (before the clipping)
    rng[rng > max_range] = invalid_value                                            # :320
    rng[rng < min_range] = invalid_value                                            # :321
    vc = torch.sum(robot_linvel.unsqueeze(1) * unit, dim=-1)                        # :327-329
    ttc = torch.where(vc > 0, flat / (vc + 1e-6), ttc_max * torch.ones_like(flat))  # :331-335 (clipped ranges)
    ttc_env = torch.clamp(torch.min(ttc, dim=-1).values, 0.0, ttc_max)              # :339
    ds = -torch.nn.functional.max_pool2d(-rng.unsqueeze(1), pool_hw).squeeze(1)     # :346-347
    return ds, ttc_env


def add_noise(ds, generator=None):
    """add_noise_to_downsampled_lidar_data, lidar_navigation_task.py:286-310, with the reference's draw order
    (bernoulli, rand of the masked count, bernoulli, bernoulli, full-size rand).  In place, returns ds."""
    g = generator
    noise_mask = torch.bernoulli(0.03 * torch.ones_like(ds), generator=g)                     # :288-289
    n = int((noise_mask == 1).sum())
    ds[noise_mask == 1] += (10.0 - 0.2) * torch.rand(n, generator=g) + 0.2                    # :290-293
    max_mask = torch.bernoulli(0.02 * torch.ones_like(ds), generator=g)                       # :296-297
    ds[max_mask == 1] = 10.0                                                                  # :298
    low_mask = torch.bernoulli(0.02 * torch.ones_like(ds[:, 10:]), generator=g)               # :303-304
    low = (1.0 - 0.2) * torch.rand(low_mask.shape, generator=g) + 0.2                         # :305-308
    ds[:, 10:][low_mask == 1] = low[low_mask == 1]                                            # :309
    return ds


def add_noise_radar(ds, generator=None):
    """RadarNavigationTask.add_noise_to_downsampled_lidar_data (radar_navigation_task.py:7-21): 3 % of the pixels get +U(0.2, 10),
    then 80 % are invalidated (-1).  In place, returns ds."""
    g = generator
    noise_mask = torch.bernoulli(0.03 * torch.ones_like(ds), generator=g)
    n = int((noise_mask == 1).sum())
    ds[noise_mask == 1] += (10.0 - 0.2) * torch.rand(n, generator=g) + 0.2
    invalid = torch.bernoulli(0.8 * torch.ones_like(ds), generator=g)
    ds[invalid == 1] = -1.0
    return ds


def compute_reward(pos_error, vehicle_linvel, body_angvel, yaw_error, crashes, action, prev_action, time_to_collision,
                   curriculum_progress_fraction, p, radar_variant=False):
    """compute_reward, lidar_navigation_task.py:554-720 (prev_pos_error is an argument there but is not used).
    p: dict name -> float (LIDAR_NAV_PARAM_NAMES)."""
    f = curriculum_progress_fraction
    mult = 1.0 + 2.0 * f                                                            # :568
    dist = torch.norm(pos_error, dim=1)                                             # :569
    pos_reward = _erf(p["pos_reward_magnitude"], p["pos_reward_exponent"], dist)    # :571-575
    very_close = _erf(p["very_close_to_goal_reward_magnitude"], p["very_close_to_goal_reward_exponent"], dist)  # :576-580
    vn = torch.norm(vehicle_linvel, dim=1)                                          # :582
    vdir = vehicle_linvel / (vn.unsqueeze(1) + 1e-6)                                # :583-584
    ug = pos_error / (dist.unsqueeze(1) + 1e-6)                                     # :585
    reasonable_vel = _erf(2.0, 2.0, vn - 2.0)                                       # :587-591
    vdc = torch.sum(vdir * ug, dim=1)                                               # :594
    vdc_reward = torch.where(vdc > 0, p["vel_direction_component_reward_magnitude"] * vdc * reasonable_vel,
                             -0.2 * torch.ones_like(vdc)) * torch.min(dist / 3.0, torch.ones_like(dist))  # :596-599
    vel_mag_pen = _epf(2.0, 2.0, torch.clamp(vn - 3.0, min=0.0))                    # :603-607
    close_to_goal = 1.0 - _erf(1.0, 2.0, dist)                                      # :609-613
    vx = torch.clamp(vehicle_linvel[:, 0], max=0.0) if radar_variant else torch.clamp(vehicle_linvel[:, 0], min=0.0)
    neg_x_pen = _epf(2.0, 8.0, vx) * close_to_goal  # :616-620 (radar_navigation_task.py:251: max=0.0)
    vel_pen = vel_mag_pen + neg_x_pen                                               # :622
    low_vel = _erf(1.5, 10.0, vn) + _erf(1.5, 0.5, vn)                              # :625
    correct_yaw = _erf(2.0, 0.2, yaw_error) + _erf(4.0, 15.0, yaw_error)            # :630
    alignment = _erf(1.0, 2.0, yaw_error)                                           # :635
    low_angvel = _erf(1.5, 5.0, body_angvel[:, 2]) * alignment                      # :636
    stable = torch.where(dist < 1.0, low_vel + correct_yaw + low_angvel, torch.zeros_like(low_vel))  # :638-642
    dist_reward = (20.0 - dist) / 20.0                                              # :645
    d = action - prev_action                                                        # :646
    diff_pen = (_epf(p["x_action_diff_penalty_magnitude"], p["x_action_diff_penalty_exponent"], d[:, 0])
                + _epf(p["y_action_diff_penalty_magnitude"], p["y_action_diff_penalty_exponent"], d[:, 1])
                + _epf(p["z_action_diff_penalty_magnitude"], p["z_action_diff_penalty_exponent"], d[:, 2])
                + _epf(p["yawrate_action_diff_penalty_magnitude"], p["yawrate_action_diff_penalty_exponent"], d[:, 3]))  # :647-667
    x_abs = f * _epf(p["x_absolute_action_penalty_magnitude"], p["x_absolute_action_penalty_exponent"], action[:, 0])
    z_abs = f * _epf(p["z_absolute_action_penalty_magnitude"], p["z_absolute_action_penalty_exponent"], action[:, 2])
    w_abs = f * _epf(p["yawrate_absolute_action_penalty_magnitude"], p["yawrate_absolute_action_penalty_exponent"], action[:, 3])
    y_abs = f * _epf(p["y_absolute_action_penalty_magnitude"], p["y_absolute_action_penalty_exponent"], action[:, 1])
    abs_pen = x_abs + z_abs + w_abs + y_abs                                         # :669-690
    action_pen = diff_pen + abs_pen                                                 # :691
    ttc_pen = _erf(-3.0, 2.0, time_to_collision ** 2)                               # :693-697
    reward = mult * (pos_reward + very_close * alignment + vdc_reward + dist_reward + stable + vel_pen + action_pen + ttc_pen)  # :700-712
    return torch.where(crashes > 0, torch.full_like(reward, p["collision_penalty"]), reward)  # :714-718


def rewards_and_errors(vehicle_orientation, position, target, euler, target_yaw, vehicle_linvel, body_angvel, crashes, action,
                       prev_action, time_to_collision, curriculum_progress_fraction, p, radar_variant=False):
    """compute_rewards_and_crashes, lidar_navigation_task.py:471-499.  Returns (reward, pos_error)."""
    err = O.quat_rotate_inverse(vehicle_orientation, target - position)            # :480-482
    yaw_error = O.ssa(target_yaw - O.ssa(euler)[:, 2])                              # :483-484
    return compute_reward(err, vehicle_linvel, body_angvel, yaw_error, crashes, action, prev_action, time_to_collision,
                          curriculum_progress_fraction, p, radar_variant), err


def process_obs(vehicle_orientation, position, target, euler, target_yaw, body_linvel, body_angvel, robot_actions, lidar_obs,
                u_vec, u_euler):
    """process_obs_for_task, lidar_navigation_task.py:440-469.  u_vec, u_euler: the two torch.rand_like draws in
    call order (here the -0.5 is OUTSIDE rand_like, unlike navigation_task.py:374)."""
    vec = O.quat_rotate_inverse(vehicle_orientation, target - position)            # :441-444
    pert = vec + 0.1 * 2 * (u_vec - 0.5)                                            # :445-446
    dist = torch.norm(vec, dim=-1)                                                  # :447
    L = lidar_obs.shape[1]
    obs = torch.zeros(position.shape[0], 17 + L, dtype=position.dtype)
    obs[:, 0:3] = pert / dist.unsqueeze(1)                                          # :448-450
    obs[:, 3] = dist                                                                # :451
    e = O.ssa(euler)                                                                # :454
    pe = e + 0.1 * (u_euler - 0.5)                                                  # :455-456
    obs[:, 4], obs[:, 5] = pe[:, 0], pe[:, 1]                                       # :457-458
    obs[:, 6] = O.ssa(target_yaw - e[:, 2])                                         # :459-461
    obs[:, 7:10], obs[:, 10:13], obs[:, 13:17] = body_linvel, body_angvel, robot_actions  # :463-468
    obs[:, 17:] = lidar_obs                                                         # :469
    return obs
