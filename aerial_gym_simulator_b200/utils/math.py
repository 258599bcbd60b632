"""Torch helpers with the names user task code imports from the reference's utils/math.py
(xyzw quaternions, batched [..., k] tensors).  Host-side conveniences for rewards/observations
written in torch; the hot paths use the device versions in csrc/agx_math.cuh."""
import math

import torch


def quat_conjugate(a):
    return torch.cat((-a[..., :3], a[..., 3:4]), dim=-1)


quat_inverse = quat_conjugate


def quat_mul(a, b):
    x1, y1, z1, w1 = a.unbind(-1)
    x2, y2, z2, w2 = b.unbind(-1)
    return torch.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], dim=-1)


def quat_apply(a, b):
    xyz = a[..., :3]
    t = torch.cross(xyz, b, dim=-1) * 2
    return b + a[..., 3:4] * t + torch.cross(xyz, t, dim=-1)


def quat_apply_inverse(a, b):
    return quat_apply(quat_conjugate(a), b)


def quat_rotate(q, v):
    qw, qv = q[..., 3:4], q[..., :3]
    return v * (2.0 * qw * qw - 1.0) + torch.cross(qv, v, dim=-1) * qw * 2.0 + qv * (qv * v).sum(-1, keepdim=True) * 2.0


def quat_rotate_inverse(q, v):
    qw, qv = q[..., 3:4], q[..., :3]
    return v * (2.0 * qw * qw - 1.0) - torch.cross(qv, v, dim=-1) * qw * 2.0 + qv * (qv * v).sum(-1, keepdim=True) * 2.0


def quat_axis(q, axis=0):
    basis = torch.zeros(q.shape[:-1] + (3,), device=q.device, dtype=q.dtype)
    basis[..., axis] = 1
    return quat_rotate(q, basis)


def quat_from_euler_xyz(roll, pitch, yaw):
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    return torch.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp,
                        cy * cr * cp + sy * sr * sp], dim=-1)


def quat_from_euler_xyz_tensor(e):
    return quat_from_euler_xyz(e[..., 0], e[..., 1], e[..., 2])


def get_euler_xyz_tensor(q):
    qx, qy, qz, qw = q.unbind(-1)
    roll = torch.atan2(2.0 * (qw * qx + qy * qz), qw * qw - qx * qx - qy * qy + qz * qz)
    sinp = 2.0 * (qw * qy - qz * qx)
    pitch = torch.where(torch.abs(sinp) >= 1, torch.sign(sinp) * (math.pi / 2.0), torch.asin(sinp))
    yaw = torch.atan2(2.0 * (qw * qz + qx * qy), qw * qw + qx * qx - qy * qy - qz * qz)
    two_pi = 2 * math.pi
    return torch.stack([roll % two_pi, pitch % two_pi, yaw % two_pi], dim=-1)


def get_euler_xyz(q):
    e = get_euler_xyz_tensor(q)
    return e[..., 0], e[..., 1], e[..., 2]


def ssa(a):
    return torch.remainder(a + math.pi, 2 * math.pi) - math.pi


def vehicle_frame_quat_from_quat(q):
    e = get_euler_xyz_tensor(q)
    z = torch.zeros_like(e[..., 0])
    return quat_from_euler_xyz(z, z, e[..., 2])


def quat_to_rotation_matrix(a):
    x, y, z, w = a.unbind(-1)
    m = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1)
    return m.view(a.shape[:-1] + (3, 3))


def normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def tf_apply(q, t, v):
    return quat_apply(q, v) + t


def torch_rand_float_tensor(lower, upper):
    return (upper - lower) * torch.rand_like(upper) + lower


def torch_interpolate_ratio(min, max, ratio):
    return min + (max - min) * ratio


def tensor_clamp(t, min_t, max_t):
    return torch.max(torch.min(t, max_t), min_t)


def exponential_reward_function(magnitude: float, base_width: float, value):
    return magnitude * torch.exp(-(value * value) / base_width)


def exponential_penalty_function(magnitude: float, base_width: float, value):
    return magnitude * (torch.exp(-(value * value) / base_width) - 1.0)


# ---- the rest of the reference's utils/math.py surface (names user task code may import) --------------------------------------
def compute_vee_map(skew_matrix):
    """vee of a batch of skew-symmetric 3x3 matrices (utils/math.py:34-42)."""
    return torch.stack([-skew_matrix[..., 1, 2], skew_matrix[..., 0, 2], -skew_matrix[..., 0, 1]], dim=-1)


def torch_rand_float(lower, upper, shape, device):
    return (upper - lower) * torch.rand(*shape, device=device) + lower


def torch_rand_float_vec(lower, upper, shape, device):
    return torch.rand(*shape, device=device) * (upper - lower) + lower


def torch_random_dir_2(shape, device):
    angle = torch_rand_float(-math.pi, math.pi, shape, device).squeeze(-1)
    return torch.stack([torch.cos(angle), torch.sin(angle)], dim=-1)


def copysign(a, b):
    """scalar magnitude a with the sign of every element of the 1-D tensor b"""
    return torch.full((b.shape[0],), abs(float(a)), device=b.device, dtype=torch.float) * torch.sign(b)


def scale(x, lower, upper):
    return 0.5 * (x + 1.0) * (upper - lower) + lower


def unscale(x, lower, upper):
    return (2.0 * x - upper - lower) / (upper - lower)


unscale_np = unscale


def to_torch(x, dtype=torch.float, device="cuda:0", requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


def quat_unit(a):
    return normalize(a)


def quat_from_angle_axis(angle, axis):
    half = (angle / 2).unsqueeze(-1)
    return quat_unit(torch.cat([normalize(axis) * half.sin(), half.cos()], dim=-1))


def normalize_angle(x):
    return torch.atan2(torch.sin(x), torch.cos(x))


def tf_inverse(q, t):
    q_inv = quat_conjugate(q)
    return q_inv, -quat_apply(q_inv, t)


def tf_vector(q, v):
    return quat_apply(q, v)


def tf_combine(q1, t1, q2, t2):
    return quat_mul(q1, q2), quat_apply(q1, t2) + t1


def get_basis_vector(q, v):
    return quat_rotate(q, v)


def pd_control(pos_error, vel_error, stiffness, damping):
    return stiffness * pos_error + damping * vel_error
