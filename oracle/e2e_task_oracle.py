"""CPU restatement (torch) of the reward / observation epilogue of the reference's two motor-command position tasks
(task/position_setpoint_task_sim2real_end_to_end/...py and task/position_setpoint_task_sim2real_px4/...py).

TEST INFRASTRUCTURE ONLY; product path: csrc/e2e_task.cu behind agx_e2e_reward / agx_e2e_obs.  compute_reward is pinned against
fixtures produced by the reference's own functions (tests/golden/make_golden_e2e.py).  process_obs_for_task calls four pytorch3d
functions that are NOT in the reference tree (pytorch3d is an unpinned dependency, setup.py:15): they are restated below from
pytorch3d's published algorithms (pytorch3d/transforms/rotation_conversions.py), the fixture generator runs the reference method on
top of the same restatements -- parity against pytorch3d itself is UNPINNED."""
import torch

from . import hp1_oracle as O

E2E_PARAMS = {  # AgxE2ERewardParams, in struct order
    "end_to_end": dict(z_error_scale=11.0, upright_gain2=0.0, upright_exp2=0.0, align_gain1=6.0, align_exp1=5.0, align_gain2=0.0, align_exp2=0.0,
                       angvel_gain=0.3, hover_thrust=9.81 * 0.372 / 4, towards_gain_pos=10.0, towards_gain_neg=15.0, action_diff_gain=1.3),
    "px4": dict(z_error_scale=13.0, upright_gain2=2.5, upright_exp2=2.0, align_gain1=4.0, align_exp1=5.0, align_gain2=2.0, align_exp2=2.0,
                angvel_gain=0.75, hover_thrust=9.81 * 1.6559999883174896 / 4, towards_gain_pos=50.0, towards_gain_neg=100.0, action_diff_gain=0.5),
}


# ---- pytorch3d.transforms, published algorithms ----------------------------------------------------------------------------
def quaternion_to_matrix(q):  # (w, x, y, z)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _index_from_letter(letter):
    return {"X": 0, "Y": 1, "Z": 2}[letter]


def _angle_from_tan(axis, other_axis, data, horizontal, tait_bryan):
    i1, i2 = {"X": (2, 1), "Y": (0, 2), "Z": (1, 0)}[axis]
    if horizontal:
        i2, i1 = i1, i2
    even = (axis + other_axis) in ["XY", "YZ", "ZX"]
    if horizontal == even:
        return torch.atan2(data[..., i1], data[..., i2])
    if tait_bryan:
        return torch.atan2(-data[..., i2], data[..., i1])
    return torch.atan2(data[..., i2], -data[..., i1])


def matrix_to_euler_angles(matrix, convention):
    i0, i2 = _index_from_letter(convention[0]), _index_from_letter(convention[2])
    tait_bryan = i0 != i2
    if tait_bryan:
        central_angle = torch.asin(matrix[..., i0, i2] * (-1.0 if i0 - i2 in [-1, 2] else 1.0))
    else:
        central_angle = torch.acos(matrix[..., i0, i0])
    o = (_angle_from_tan(convention[0], convention[1], matrix[..., i2], False, tait_bryan), central_angle,
         _angle_from_tan(convention[2], convention[1], matrix[..., i0, :], True, tait_bryan))
    return torch.stack(o, -1)


def _axis_angle_rotation(axis, angle):
    cos, sin, one, zero = torch.cos(angle), torch.sin(angle), torch.ones_like(angle), torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, cos, -sin, zero, sin, cos)
    elif axis == "Y":
        flat = (cos, zero, sin, zero, one, zero, -sin, zero, cos)
    else:
        flat = (cos, -sin, zero, sin, cos, zero, zero, zero, one)
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles, convention):
    m = [_axis_angle_rotation(c, e) for c, e in zip(convention, torch.unbind(euler_angles, -1))]
    return torch.matmul(torch.matmul(m[0], m[1]), m[2])


def matrix_to_rotation_6d(matrix):
    return matrix[..., :2, :].clone().reshape(matrix.size()[:-2] + (6,))


# ---- the task ------------------------------------------------------------------------------------------------------------------
def _exp_func(x, gain, exp):  # :254-257
    return gain * torch.exp(-exp * x * x)


def _exp_penalty(x, gain, exp):  # :260-263
    return gain * (torch.exp(-exp * x * x) - 1)


def compute_reward(pos_error, quats, linvel, body_angvel, crashes, action, prev_action, prev_pos_error, crash_dist, p):
    """compute_reward (:267-311 of the end-to-end task; p = E2E_PARAMS[...]).  Returns (reward, crashes)."""
    pos_error = pos_error.clone()
    dist = torch.norm(pos_error[:, :3], dim=1)
    prev_dist = torch.norm(prev_pos_error, dim=1)
    pos_error[:, 2] = pos_error[:, 2] * p["z_error_scale"]
    pos_reward = torch.sum(_exp_func(pos_error, 10.0, 10.0), dim=1) + torch.sum(_exp_func(pos_error, 2.0, 2.0), dim=1)
    ez, ex = torch.zeros_like(linvel), torch.zeros_like(linvel)
    ez[:, 2], ex[:, 0] = 1.0, 1.0
    tilt = 1 - O.quat_rotate(quats, ez)[:, 2]
    upright = _exp_func(tilt, 2.5, 5.0) + _exp_func(tilt, p["upright_gain2"], p["upright_exp2"])
    al = 1 - O.quat_rotate(quats, ex)[:, 0]
    align = _exp_func(al, p["align_gain1"], p["align_exp1"]) + _exp_func(al, p["align_gain2"], p["align_exp2"])
    angvel_r = torch.sum(_exp_func(body_angvel, p["angvel_gain"], 10.0), dim=1)
    vel_r = torch.sum(_exp_func(linvel, 1.0, 5.0), dim=1)
    action_cost = torch.sum(_exp_penalty(action - p["hover_thrust"], 0.01, 10.0), dim=1)
    closer = prev_dist - dist
    towards = torch.where(closer >= 0, p["towards_gain_pos"] * closer, p["towards_gain_neg"] * closer)
    diff_pen = torch.sum(_exp_penalty(action - prev_action, p["action_diff_gain"], 6.0), dim=1)
    reward = towards + (pos_reward * (align + vel_r + angvel_r + diff_pen) + (angvel_r + vel_r + upright + pos_reward + action_cost)) / 100.0
    return reward, crashes | (dist > crash_dist)


def process_obs(position, orientation_xyzw, linvel, body_angvel, target, noise):
    """process_obs_for_task (:204-229).  noise [N,12]: position, orientation, linear velocity, body-rate draws side by side."""
    obs = torch.zeros(position.shape[0], 15, dtype=position.dtype)
    obs[:, 0:3] = (target - position) + noise[:, 0:3]
    or_euler = matrix_to_euler_angles(quaternion_to_matrix(orientation_xyzw[:, [3, 0, 1, 2]]), "ZYX")[:, [2, 1, 0]]
    noisy = or_euler + noise[:, 3:6]
    obs[:, 3:9] = matrix_to_rotation_6d(euler_angles_to_matrix(noisy[:, [2, 1, 0]], "ZYX"))
    obs[:, 9:12] = linvel + noise[:, 6:9]
    obs[:, 12:15] = body_angvel + noise[:, 9:12]
    return obs
