"""Reward / observation epilogue of the two motor-command position tasks (sim2real_end_to_end, sim2real_px4) on CPU: the oracle and
the device code (csrc/e2e_task_core.cuh, compiled for the host) against the fixtures produced by the reference's own functions."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from aerial_gym_simulator_b200 import _lib
from oracle import e2e_task_oracle as E

from . import _shadow

G = os.path.join(os.path.dirname(__file__), "golden", "e2e_task_epilogue.npz")


def _t(x):
    return torch.tensor(np.asarray(x))


def _params(tag, crash_dist):
    p = _lib.AgxE2ERewardParams()
    for k, v in E.E2E_PARAMS[tag].items():
        setattr(p, k, float(v))
    p.crash_dist = float(crash_dist)
    return p


@pytest.mark.parametrize("tag", ["end_to_end", "px4"])
def test_oracle_matches_reference_fixture(tag):
    d = np.load(G)
    rew, cr = E.compute_reward(-_t(d["pos"]), _t(d["quat"]), _t(d["linvel"]), _t(d["body_angvel"]), _t(d["crashes_in"]), _t(d[f"{tag}_actions"]),
                               _t(d[f"{tag}_prev_actions"]), _t(d["prev_pos_error"]), float(d[f"{tag}_crash_dist"]), E.E2E_PARAMS[tag])
    assert torch.equal(cr, _t(d[f"{tag}_crashes_out"])) and cr.any() and not cr.all()
    assert torch.allclose(rew, _t(d[f"{tag}_reward"]), rtol=1e-6, atol=1e-6)
    obs = E.process_obs(_t(d["pos"]), _t(d["quat"]), _t(d["linvel"]), _t(d["body_angvel"]), torch.zeros(d["pos"].shape), _t(d[f"{tag}_noise"]))
    assert torch.allclose(obs, _t(d[f"{tag}_obs"]), rtol=1e-6, atol=1e-6)


def test_pytorch3d_restatements_are_consistent():
    """ZYX Euler round trip through the restated pytorch3d functions, and the closed form the kernel uses"""
    g = torch.Generator().manual_seed(1)
    e = (torch.rand(500, 3, generator=g) * 2 - 1) * torch.tensor([3.1, 1.5, 3.1])  # yaw, pitch, roll
    R = E.euler_angles_to_matrix(e, "ZYX")
    assert torch.allclose(E.matrix_to_euler_angles(R, "ZYX"), e, atol=2e-5)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(500, 3, 3), atol=1e-5)
    y, p, r = e[:, 0], e[:, 1], e[:, 2]
    want = torch.stack([torch.cos(y) * torch.cos(p), torch.cos(y) * torch.sin(p) * torch.sin(r) - torch.sin(y) * torch.cos(r),
                        torch.cos(y) * torch.sin(p) * torch.cos(r) + torch.sin(y) * torch.sin(r), torch.sin(y) * torch.cos(p),
                        torch.sin(y) * torch.sin(p) * torch.sin(r) + torch.cos(y) * torch.cos(r),
                        torch.sin(y) * torch.sin(p) * torch.cos(r) - torch.cos(y) * torch.sin(r)], dim=1)
    assert torch.allclose(E.matrix_to_rotation_6d(R), want, atol=1e-6)


def _c(a, dtype=np.float32):
    a = np.ascontiguousarray(np.asarray(a), dtype)
    return a, a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("tag", ["end_to_end", "px4"])
def test_shadow_matches_reference_fixture(tag):
    d = np.load(G)
    lib, n = _shadow.load(), d["pos"].shape[0]
    st = np.zeros((n, 13), np.float32)
    st[:, 0:3], st[:, 3:7], st[:, 7:10] = d["pos"], d["quat"], d["linvel"]
    keep = [_c(st), _c(d["body_angvel"]), _c(d[f"{tag}_actions"]), _c(d[f"{tag}_prev_actions"]), _c(d["prev_pos_error"])]
    p = _params(tag, d[f"{tag}_crash_dist"])
    cr, rew = np.array(d["crashes_in"], np.uint8), np.zeros(n, np.float32)
    lib.shadow_e2e_reward(n, keep[0][1], 13, keep[1][1], None, keep[2][1], keep[3][1], keep[4][1], C.cast(C.byref(p), C.c_void_p),
                          cr.ctypes.data_as(C.c_void_p), rew.ctypes.data_as(C.c_void_p))
    assert np.array_equal(cr.astype(bool), d[f"{tag}_crashes_out"])  # bit-exact flags
    # towards_goal = gain * (|prev error| - |error|): a difference of two nearly equal fp32 norms times 10..15 (end_to_end) or 50..100
    # (px4); one ulp in either norm (1.2e-7 * |error|, |error| up to 10 m here) is already 1e-4 in the px4 reward
    np.testing.assert_allclose(rew, d[f"{tag}_reward"], rtol=1e-5, atol=3e-4 if tag == "px4" else 5e-5)
    for stride in (15, 24):
        nz = _c(d[f"{tag}_noise"])
        obs = np.full((n, stride), 7.0, np.float32)
        lib.shadow_e2e_obs(n, keep[0][1], 13, keep[1][1], None, nz[1], obs.ctypes.data_as(C.c_void_p), stride)
        np.testing.assert_allclose(obs[:, :15], d[f"{tag}_obs"], rtol=1e-5, atol=1e-5)
        assert (obs[:, 15:] == 7.0).all()
