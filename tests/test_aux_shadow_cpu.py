"""NavigationTask epilogue and IMU on CPU: the device code of hp1_aux.cu (csrc/aux_core.cuh), compiled for the host, against
the fixtures produced by the reference's own code -- the CPU twins of tests/test_aux_gpu.py."""
import ctypes as C
import os

import numpy as np
import pytest

from aerial_gym_simulator_b200 import _lib
from oracle import aux_oracle as A

from . import _shadow

G = os.path.join(os.path.dirname(__file__), "golden")


def _c(a, dtype=np.float32):
    a = np.ascontiguousarray(np.asarray(a), dtype)
    assert a.ctypes.data % 16 == 0 or a.size < 4  # the [N,4] arrays are read as float4
    return a, a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("tag", ["c0", "c1"])
def test_nav_reward_matches_reference_fixture(tag):
    d = np.load(os.path.join(G, "nav_task_epilogue.npz"))
    assert tuple(d["param_names"]) == A.NAV_PARAM_NAMES
    p = _lib.AgxNavRewardParams()
    for i, v in enumerate(d["param_values"]):
        p.v[i] = float(v)
    n = d["pos"].shape[0]
    st = np.zeros((n, 13), np.float32)
    st[:, 0:3] = d["pos"]
    keep = [_c(st), _c(d["vehicle_orientation"]), _c(d["target"]), _c(d["crashes"], np.uint8), _c(d["actions"]), _c(d["prev_actions"])]
    pe, pp, rew = np.array(d["prev_pos_error"], np.float32, copy=True), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
    ptr = [k[1] for k in keep]
    _shadow.load().shadow_nav_reward(n, ptr[0], 13, *ptr[1:], float(d[f"frac_{tag}"]), C.cast(C.byref(p), C.c_void_p),
                                     pe.ctypes.data_as(C.c_void_p), pp.ctypes.data_as(C.c_void_p), rew.ctypes.data_as(C.c_void_p))
    np.testing.assert_allclose(rew, d[f"reward_{tag}"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(pe, d["pos_error"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(pp, d["prev_pos_error"])


def test_nav_obs_matches_reference_fixture():
    d = np.load(os.path.join(G, "nav_task_epilogue.npz"))
    n = d["pos"].shape[0]
    st = np.zeros((n, 13), np.float32)
    st[:, 0:3] = d["pos"]
    keep = [_c(st)] + [_c(d[k]) for k in ("vehicle_orientation", "euler", "body_linvel", "body_angvel", "robot_actions", "target", "obs_draw_vec",
                                          "obs_draw_euler")]
    obs = np.full((n, 81), 7.0, np.float32)
    ptr = [k[1] for k in keep]
    _shadow.load().shadow_nav_obs(n, ptr[0], 13, *ptr[1:], obs.ctypes.data_as(C.c_void_p), 81)
    np.testing.assert_allclose(obs, d["obs"], rtol=1e-5, atol=1e-5)  # incl. the untouched latent columns


@pytest.mark.parametrize("tag", ["body", "world", "gcomp"])
def test_imu_matches_reference_fixture(tag):
    """three IMUSensor.update calls with the reference's recorded draws, bias random walk included"""
    d = np.load(os.path.join(G, "imu_sensor.npz"))
    world, gcomp = bool(d[f"{tag}_cfg"][0]), bool(d[f"{tag}_cfg"][1])
    n = d[f"{tag}_robot_mass"].shape[0]
    c = _lib.AgxImuConfig()
    c.world_frame, c.enable_noise, c.enable_bias, c.sqrt_dt = int(world), 1, 1, float(np.sqrt(np.float32(0.01)))
    g = np.array([0.0, 0.0, -9.81], np.float32) * (0.0 if gcomp else 1.0)
    for i in range(3):
        c.g_world[i] = float(g[i])
    for i in range(6):
        c.bias_std[i], c.noise_std[i], c.max_meas[i] = float(d["bias_std"][i]), float(d["imu_noise_std"][i]), float(d["max_measurement_value"][i])
    st = np.zeros((n, 13), np.float32)
    st[:, 3:7] = d[f"{tag}_robot_orientation"]
    keep = {k: _c(v) for k, v in (("force", d[f"{tag}_force_sensor_tensor"]), ("mass", d[f"{tag}_robot_mass"]), ("state", st),
                                  ("bav", d[f"{tag}_robot_body_angvel"]), ("sq", d[f"{tag}_sensor_quats"]))}
    bias, meas = np.array(d[f"{tag}_bias0"], np.float32, copy=True), np.zeros((n, 6), np.float32)
    draws = d[f"{tag}_draws"]
    for k in range(3):
        nn, nb = _c(draws[2 * k]), _c(draws[2 * k + 1])
        _shadow.load().shadow_imu_update(n, C.cast(C.byref(c), C.c_void_p), keep["force"][1], 6, keep["mass"][1], keep["state"][1], 13,
                                         keep["bav"][1], keep["sq"][1], nn[1], nb[1], bias.ctypes.data_as(C.c_void_p),
                                         meas.ctypes.data_as(C.c_void_p))
        np.testing.assert_allclose(meas, d[f"{tag}_meas"][k], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(bias, d[f"{tag}_bias_end"], rtol=1e-6, atol=1e-9)
