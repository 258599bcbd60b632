"""LiDARNavigationTask epilogue on CPU:
  (a) the oracle (oracle/lidar_nav_oracle.py) against the fixtures produced by the reference's own code;
  (b) the CPU shadow of the device code (the very functions lidar_nav.cu runs, compiled for the host, tests/csrc/host_shadow.cu)
      against the same fixtures and against the oracle on bigger random inputs, ragged shapes included.
The -m gpu counterpart (tests/test_zz_lidar_nav_gpu.py) runs the real kernels through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import hp1_oracle as O
from oracle import lidar_nav_oracle as L

from . import _shadow

G = os.path.join(os.path.dirname(__file__), "golden", "lidar_nav_task_epilogue.npz")


def _t(x):
    return torch.tensor(np.asarray(x))


def _params_dict(d):
    assert tuple(d["param_names"]) == L.LIDAR_NAV_PARAM_NAMES
    return {k: float(v) for k, v in zip(d["param_names"], d["param_values"])}


def _c(a):
    a = np.ascontiguousarray(a)
    return a, a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------------ oracle vs reference fixtures
def test_oracle_pool_matches_reference_fixture():
    d = np.load(G)
    ds, ttc = L.pool(_t(d["pool_pointcloud"]), _t(d["pool_robot_position"]), _t(d["pool_robot_linvel"]))
    assert torch.equal(ds, _t(d["pool_image_ds"]))  # same torch ops in the same order: bit-identical
    assert torch.equal(ttc, _t(d["pool_time_to_collision"]))
    assert float(ttc[1]) == 10.0  # the hovering env
    assert (d["pool_image_ds"] == 10.0).any() and (d["pool_image_ds"] < 1.0).any()  # both clipping outcomes are in the fixture


def test_oracle_noise_matches_reference_fixture():
    """The task's own lidar noise (torch RNG): same seed -> the reference's draws, bit for bit; then 1/x."""
    d = np.load(G)
    torch.manual_seed(int(d["pool_noise_seed"]))
    noisy = L.add_noise(_t(d["pool_image_ds"]).clone())
    assert torch.equal(noisy, _t(d["pool_image_noisy"]))
    assert torch.equal((1 / noisy).reshape(noisy.shape[0], -1), _t(d["pool_downsampled_lidar_data"]))


@pytest.mark.parametrize("tag", ["c0", "c1"])
def test_oracle_reward_matches_reference_fixture(tag):
    d = np.load(G)
    rew, err = L.rewards_and_errors(_t(d["vehicle_orientation"]), _t(d["pos"]), _t(d["target"]), _t(d["euler"]), _t(d["target_yaw"]),
                                    _t(d["vehicle_linvel"]), _t(d["body_angvel"]), _t(d["crashes"]), _t(d["actions"]), _t(d["prev_actions"]),
                                    _t(d["time_to_collision"]), float(d[f"frac_{tag}"]), _params_dict(d))
    assert torch.allclose(rew, _t(d[f"reward_{tag}"]), rtol=1e-6, atol=1e-6), (rew - _t(d[f"reward_{tag}"])).abs().max()
    assert torch.equal(err, _t(d["pos_error"]))


def test_oracle_obs_matches_reference_fixture():
    d = np.load(G)
    ref = _t(d["obs"])
    obs = L.process_obs(_t(d["vehicle_orientation"]), _t(d["pos"]), _t(d["target"]), _t(d["euler"]), _t(d["target_yaw"]), _t(d["body_linvel"]),
                        _t(d["body_angvel"]), _t(d["robot_actions"]), ref[:, 17:], _t(d["obs_draw_vec"]), _t(d["obs_draw_euler"]))
    assert torch.allclose(obs, ref, rtol=1e-6, atol=1e-6), (obs - ref).abs().max()


# ------------------------------------------------------------------------------------------------ device code, compiled for the host
def _shadow_pool(pc, state, ph=3, pw=6, max_range=10.0, min_range=0.2, invalid=10.0, ttc_max=10.0, force_scalar=False):
    lib = _shadow.load()
    n, H, W, _ = pc.shape
    pc, ppc = _c(pc.astype(np.float32))
    state, pst = _c(state.astype(np.float32))
    ds = np.full((n, H // ph, W // pw), -7.0, np.float32)
    ttc = np.full((n,), -7.0, np.float32)
    vec4 = lib.shadow_lidar_nav_pool(n, H, W, ph, pw, ppc, pst, state.shape[1], max_range, min_range, invalid, ttc_max,
                                     ds.ctypes.data_as(C.c_void_p), ttc.ctypes.data_as(C.c_void_p), int(force_scalar))
    assert vec4 >= 0
    return ds, ttc, bool(vec4)


def _state(pos, linvel):
    st = np.zeros((pos.shape[0], 13), np.float32)
    st[:, 0:3], st[:, 7:10], st[:, 6] = pos, linvel, 1.0
    return st


@pytest.mark.parametrize("force_scalar", [False, True])
def test_shadow_pool_matches_reference_fixture(force_scalar):
    d = np.load(G)
    ds, ttc, vec4 = _shadow_pool(d["pool_pointcloud"], _state(d["pool_robot_position"], d["pool_robot_linvel"]), force_scalar=force_scalar)
    assert vec4 == (not force_scalar)  # 48 x 120 x 3 floats per env, 3 x 120 x 3 per band: the 16-byte path applies
    np.testing.assert_allclose(ds, d["pool_image_ds"], rtol=1e-6, atol=0)
    assert np.array_equal(ds == 10.0, d["pool_image_ds"] == 10.0)  # clipping decisions agree exactly
    np.testing.assert_allclose(ttc, d["pool_time_to_collision"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("H,W,ph,pw", [(48, 120, 3, 6), (7, 13, 3, 6), (5, 9, 2, 2), (1, 40, 1, 7), (64, 33, 4, 3), (3, 6, 3, 6)])
def test_shadow_pool_matches_oracle_any_shape(H, W, ph, pw):
    """ragged shapes: trailing rows / columns are dropped by the pooling (max_pool2d floors) but still count for the
    time to collision; more bands than warps; bands that are not a multiple of 16 bytes (scalar staging path)."""
    g = torch.Generator().manual_seed(H * 1000 + W)
    n = 5
    pos = torch.randn(n, 3, generator=g)
    vel = torch.randn(n, 3, generator=g) * 2
    dirs = torch.randn(n, H, W, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    rng = torch.rand(n, H, W, 1, generator=g) * 13 + 0.05
    # keep every return clear of the two clipping thresholds: 1 ulp must not decide a pixel
    rng = torch.where((rng - 10.0).abs() < 1e-3, rng + 0.01, rng)
    rng = torch.where((rng - 0.2).abs() < 1e-3, rng + 0.01, rng)
    pc = pos.view(n, 1, 1, 3) + dirs * rng
    want_ds, want_ttc = L.pool(pc, pos, vel, (ph, pw))
    ds, ttc, _ = _shadow_pool(pc.numpy(), _state(pos.numpy(), vel.numpy()), ph, pw)
    np.testing.assert_allclose(ds, want_ds.numpy(), rtol=1e-6, atol=0)
    np.testing.assert_allclose(ttc, want_ttc.numpy(), rtol=2e-5, atol=1e-6)


def _shadow_reward(d_or_arrays, frac, params, radar=False):
    lib = _shadow.load()
    a = d_or_arrays
    n = a["pos"].shape[0]
    st = np.zeros((n, 13), np.float32)
    st[:, 0:3] = a["pos"]
    keep = [_c(st)] + [_c(np.asarray(a[k], np.float32)) for k in ("vehicle_orientation", "target", "euler", "target_yaw", "vehicle_linvel",
                                                                  "body_angvel")]
    keep.append(_c(np.asarray(a["crashes"]).astype(np.uint8)))
    keep += [_c(np.asarray(a[k], np.float32)) for k in ("actions", "prev_actions", "time_to_collision")]
    from aerial_gym_simulator_b200 import _lib
    p = _lib.AgxLidarNavRewardParams()
    for i, k in enumerate(L.LIDAR_NAV_PARAM_NAMES):
        p.v[i] = float(params[k])
    p.radar_variant = int(radar)
    pe = np.array(a["prev_pos_error"], np.float32, copy=True)
    pp, rew = np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
    ptr = [k[1] for k in keep]
    lib.shadow_lidar_nav_reward(n, ptr[0], 13, *ptr[1:], float(frac), C.cast(C.byref(p), C.c_void_p), pe.ctypes.data_as(C.c_void_p),
                                pp.ctypes.data_as(C.c_void_p), rew.ctypes.data_as(C.c_void_p))
    return rew, pe, pp


@pytest.mark.parametrize("tag", ["c0", "c1"])
def test_shadow_reward_matches_reference_fixture(tag):
    d = np.load(G)
    rew, pe, pp = _shadow_reward(d, d[f"frac_{tag}"], _params_dict(d))
    np.testing.assert_allclose(rew, d[f"reward_{tag}"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(pe, d["pos_error"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(pp, d["prev_pos_error"])  # previous error handed over bit for bit
    assert np.array_equal(rew[::11], np.full_like(rew[::11], -10.0))  # crashed envs: the collision penalty, exactly


def test_shadow_reward_matches_oracle_on_random_inputs():
    g = torch.Generator().manual_seed(77)
    n = 5000
    r = lambda *s: torch.randn(*s, generator=g)
    q = r(n, 4)
    q = q / q.norm(dim=1, keepdim=True)
    a = {"pos": r(n, 3) * 4, "target": r(n, 3) * 4, "vehicle_orientation": O.vehicle_frame_quat_from_quat(q),
         "euler": torch.rand(n, 3, generator=g) * 2 * np.pi, "target_yaw": (torch.rand(n, generator=g) * 2 - 1) * np.pi,
         "vehicle_linvel": r(n, 3) * 2, "body_angvel": r(n, 3), "crashes": torch.rand(n, generator=g) < 0.05,
         "actions": torch.rand(n, 4, generator=g) * 2 - 1, "prev_actions": torch.rand(n, 4, generator=g) * 2 - 1,
         "time_to_collision": torch.rand(n, generator=g) * 10, "prev_pos_error": r(n, 3)}
    a["target"][:200] = a["pos"][:200] + 0.4 * r(200, 3)
    d = np.load(G)
    p = _params_dict(d)
    want, err = L.rewards_and_errors(a["vehicle_orientation"], a["pos"], a["target"], a["euler"], a["target_yaw"], a["vehicle_linvel"],
                                     a["body_angvel"], a["crashes"], a["actions"], a["prev_actions"], a["time_to_collision"], 0.6, p)
    rew, pe, _ = _shadow_reward({k: v.numpy() for k, v in a.items()}, 0.6, p)
    np.testing.assert_allclose(rew, want.numpy(), rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(pe, err.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("obs_stride,num_lidar", [(337, 320), (400, 320), (17, 0)])
def test_shadow_obs_matches_reference_fixture(obs_stride, num_lidar):
    d = np.load(G)
    lib = _shadow.load()
    n = d["pos"].shape[0]
    st = np.zeros((n, 13), np.float32)
    st[:, 0:3] = d["pos"]
    lidar = np.ascontiguousarray(d["obs"][:, 17:17 + num_lidar]) if num_lidar else None
    keep = [_c(st)] + [_c(d[k]) for k in ("vehicle_orientation", "euler", "body_linvel", "body_angvel", "robot_actions", "target", "target_yaw",
                                          "obs_draw_vec", "obs_draw_euler")]
    obs = np.full((n, obs_stride), 7.0, np.float32)
    ptr = [k[1] for k in keep]
    lib.shadow_lidar_nav_obs(n, ptr[0], 13, *ptr[1:], lidar.ctypes.data_as(C.c_void_p) if num_lidar else None, num_lidar,
                             obs.ctypes.data_as(C.c_void_p), obs_stride)
    np.testing.assert_allclose(obs[:, :17 + num_lidar], d["obs"][:, :17 + num_lidar], rtol=1e-5, atol=1e-5)
    assert (obs[:, 17 + num_lidar:] == 7.0).all()  # nothing beyond the requested columns is touched


# ------------------------------------------------------------------------------------------------ RadarNavigationTask variant
@pytest.mark.parametrize("tag", ["c0", "c1"])
def test_radar_variant_matches_reference_fixture(tag):
    """RadarNavigationTask's own compute_reward (radar_navigation_task.py: the x-velocity penalty on clamp(vx, max=0)): oracle and
    device code against the fixture produced by that file's function"""
    d = np.load(G)
    ref = d[f"radar_reward_{tag}"]
    assert not np.array_equal(ref, d[f"reward_{tag}"])  # the variant matters on this input
    rew, _ = L.rewards_and_errors(_t(d["vehicle_orientation"]), _t(d["pos"]), _t(d["target"]), _t(d["euler"]), _t(d["target_yaw"]),
                                  _t(d["vehicle_linvel"]), _t(d["body_angvel"]), _t(d["crashes"]), _t(d["actions"]), _t(d["prev_actions"]),
                                  _t(d["time_to_collision"]), float(d[f"frac_{tag}"]), _params_dict(d), radar_variant=True)
    assert torch.allclose(rew, _t(ref), rtol=1e-6, atol=1e-6)
    got, _, _ = _shadow_reward(d, d[f"frac_{tag}"], _params_dict(d), radar=True)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-4)


def test_radar_noise_matches_reference_fixture():
    from aerial_gym_simulator_b200.task.radar_navigation_task import add_noise_to_downsampled_radar_data

    d = np.load(G)
    torch.manual_seed(int(d["pool_noise_seed"]))
    assert torch.equal(L.add_noise_radar(_t(d["radar_noise_in"]).clone()), _t(d["radar_noise_out"]))
    torch.manual_seed(int(d["pool_noise_seed"]))
    out = add_noise_to_downsampled_radar_data(_t(d["radar_noise_in"]).clone())
    assert torch.equal(out, _t(d["radar_noise_out"])) and 0.7 < float((out == -1.0).float().mean()) < 0.9
