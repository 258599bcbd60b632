"""CPU restatement (torch) of the reward / observation epilogue of the reference's two setpoint-command sim2real position tasks
(task/position_setpoint_task_sim2real/...py = variant 0, task/position_setpoint_task_acceleration_sim2real/...py = variant 1).

TEST INFRASTRUCTURE ONLY; product path: csrc/sim2real.cu behind agx_s2r_reward / agx_s2r_obs.  Pinned against fixtures produced by the
reference's own functions (tests/golden/make_golden_sim2real.py)."""
import torch

from . import hp1_oracle as O


def _exp(x, gain, exp):
    return gain * torch.exp(-exp * x * x)


def _abs_exp(x, gain, exp):
    return gain * torch.exp(-exp * torch.abs(x))


def _abs_exp_pen(x, gain, exp):
    return gain * (torch.exp(-exp * torch.abs(x)) - 1)


def _quat_apply_inverse(q, v):  # utils/math.py:314-325: quat_apply(quat_inverse(q), v)
    xyz = -q[:, :3]
    t = torch.cross(xyz, v, dim=-1) * 2
    return v + q[:, 3:] * t + torch.cross(xyz, t, dim=-1)


def reward(variant, position, orientation, vehicle_orientation, body_linvel, target, prev_dist, actions, prev_actions, crashes):
    """compute_rewards_and_crashes + compute_reward.  variant 1: `actions` are rotated by the vehicle orientation first (:254-257) and
    `prev_actions` must already be prev_actions_vehicle_frame.  Returns (reward, crashes, actions used by the reward)."""
    pe = _quat_apply_inverse(orientation if variant else vehicle_orientation, target - position)
    yaw_error = 0 - O.ssa(O.euler_xyz_from_quat(orientation))[:, 2]
    act = actions.clone()
    if variant:
        act[:, 0:3] = O.quat_rotate(vehicle_orientation, actions[:, 0:3])
    dist = torch.norm(pe, dim=1)
    pos_reward = _exp(dist, 2.0, 1.0) + _exp(dist, 3.0, 10.0) + _abs_exp(dist, 3.0, 50.0)
    speed = torch.norm(body_linvel, dim=1)
    diff = act - prev_actions
    if not variant:
        speed_reward, dist_reward = _exp(speed, 1.0, 3.0), (20 - dist) / 40.0
        ap, adp = torch.sum(_abs_exp_pen(act, 0.2, 4.0), dim=1), torch.sum(_abs_exp_pen(diff, 0.3, 6.0), dim=1)
        closer = 400.0 * (prev_dist - dist)
        total = (pos_reward + dist_reward + pos_reward * (speed_reward + ap + closer / 10.0)) + ap + adp + closer + _abs_exp(yaw_error, 2.0, 3.0)
    else:
        close_pos, speed_reward = _exp(dist, 2.0, 1.0), _exp(speed, 2.0, 2.5)
        ap, adp = torch.sum(_abs_exp_pen(act, 0.3, 4.0), dim=1), torch.sum(_abs_exp_pen(diff, 0.4, 6.0), dim=1)
        closer = torch.where(dist < prev_dist, 400.0 * (prev_dist - dist), 1200 * (prev_dist - dist))
        total = ((pos_reward + pos_reward * (closer / 9.0 + ap / 3.0 + speed_reward / 1.5)) + ap + adp + closer + _abs_exp(yaw_error, 3.0, 5.0)
                 + close_pos + speed_reward * 0.2)
    crashes = crashes | (dist > 10.0)
    return torch.where(crashes, torch.full_like(total, -50.0), total), crashes, act


def process_obs(position, orientation, body_linvel, body_angvel, robot_actions, target, noise):
    """process_obs_for_task.  noise [N,12]: euler, position, body linvel, body angvel draws (unscaled).  Returns (obs [N,17], the
    sign-normalised orientation the method leaves in the state)."""
    q = torch.sign(orientation[:, 3]).unsqueeze(1) * orientation
    e = O.ssa(O.euler_xyz_from_quat(q)) + noise[:, 0:3] * 0.02
    obs = torch.zeros(position.shape[0], 17, dtype=position.dtype)
    obs[:, 0:3] = (target - position) + noise[:, 3:6] * 0.03
    obs[:, 3:7] = O.quat_from_euler_xyz(e[:, 0], e[:, 1], e[:, 2])
    obs[:, 7:10] = body_linvel + noise[:, 6:9] * 0.02
    obs[:, 10:13] = body_angvel + noise[:, 9:12] * 0.02
    obs[:, 13:17] = robot_actions
    return obs, q
