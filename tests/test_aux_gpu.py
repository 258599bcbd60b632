"""SURVEY 8(f) rows 3 and 4 on the GPU, through the C ABI: navigation-task reward / observation epilogue and
the IMU against (a) the fixtures produced by the reference's own code and (b) the oracle on bigger random
inputs; plus NavigationTask end to end (env_with_obstacles, lmf2, depth camera, VAE encoder in torch)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from aerial_gym_simulator_b200 import _lib
from oracle import aux_oracle as A

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _d(x, dtype=torch.float32):
    return torch.tensor(np.asarray(x), dtype=dtype, device=DEV).contiguous()


def _p(t):
    return C.c_void_p(t.data_ptr())


def _params(names, values):
    assert tuple(names) == A.NAV_PARAM_NAMES
    p = _lib.AgxNavRewardParams()
    for i, v in enumerate(values):
        p.v[i] = float(v)
    return p


def _nav_reward_gpu(pos, veh_q, target, crashes, act, prev_act, frac, params, pos_err_in):
    lib, n = _lib.load(), pos.shape[0]
    state = torch.zeros(n, 13, device=DEV)
    state[:, 0:3] = pos
    pe, pp, rew = pos_err_in.clone(), torch.zeros(n, 3, device=DEV), torch.zeros(n, device=DEV)
    _lib.check(lib.agx_nav_reward(n, _p(state), 13, _p(veh_q), _p(target), _p(crashes), _p(act), _p(prev_act), float(frac), C.byref(params),
                                  _p(pe), _p(pp), _p(rew), None), "agx_nav_reward")
    torch.cuda.synchronize()
    return rew.cpu(), pe.cpu(), pp.cpu()


def test_nav_reward_kernel_matches_reference_fixture():
    d = np.load(os.path.join(G, "nav_task_epilogue.npz"))
    params = _params(d["param_names"], d["param_values"])
    for tag in ("c0", "c1"):
        rew, pe, pp = _nav_reward_gpu(_d(d["pos"]), _d(d["vehicle_orientation"]), _d(d["target"]), _d(d["crashes"], torch.uint8),
                                      _d(d["actions"]), _d(d["prev_actions"]), d[f"frac_{tag}"], params, _d(d["prev_pos_error"]))
        ref = torch.tensor(d[f"reward_{tag}"])
        assert torch.allclose(rew, ref, rtol=1e-5, atol=1e-4), (rew - ref).abs().max()
        assert torch.allclose(pe, torch.tensor(d["pos_error"]), rtol=1e-5, atol=1e-5)
        assert torch.equal(pp, torch.tensor(d["prev_pos_error"]))  # previous error handed over bit for bit


def test_nav_obs_kernel_matches_reference_fixture():
    d = np.load(os.path.join(G, "nav_task_epilogue.npz"))
    lib, n = _lib.load(), d["pos"].shape[0]
    state = torch.zeros(n, 13, device=DEV)
    state[:, 0:3] = _d(d["pos"])
    obs = torch.full((n, 81), 7.0, device=DEV)
    # (the tensors must outlive the call: keep them in a list, a temporary would be freed before the launch)
    t = [_d(d[k]) for k in ("vehicle_orientation", "euler", "body_linvel", "body_angvel", "robot_actions", "target", "obs_draw_vec",
                            "obs_draw_euler")]
    _lib.check(lib.agx_nav_obs(n, _p(state), 13, *[_p(x) for x in t], _p(obs), 81, None), "agx_nav_obs")
    ref = torch.tensor(d["obs"])
    assert torch.allclose(obs.cpu(), ref, rtol=1e-5, atol=1e-5), (obs.cpu() - ref).abs().max()  # incl. untouched latent columns


def test_nav_kernels_match_oracle_on_random_inputs():
    g = torch.Generator().manual_seed(5)
    n = 20000
    r = lambda *s: torch.randn(*s, generator=g)
    q = r(n, 4)
    q = q / q.norm(dim=1, keepdim=True)
    from oracle import hp1_oracle as O
    veh = O.vehicle_frame_quat_from_quat(q)
    pos, target, prev_err = r(n, 3) * 4, r(n, 3) * 4, r(n, 3) * 4
    act, prev_act = torch.rand(n, 4, generator=g) * 2 - 1, torch.rand(n, 4, generator=g) * 2 - 1
    crashes = torch.rand(n, generator=g) < 0.05
    d = np.load(os.path.join(G, "nav_task_epilogue.npz"))
    p = {k: float(v) for k, v in zip(d["param_names"], d["param_values"])}
    params = _params(d["param_names"], d["param_values"])
    rew, pe, _ = _nav_reward_gpu(pos.to(DEV), veh.to(DEV).contiguous(), target.to(DEV), crashes.to(DEV).to(torch.uint8), act.to(DEV),
                                 prev_act.to(DEV), 0.6, params, prev_err.to(DEV))
    err = A.nav_pos_error(veh, target, pos)
    want = A.nav_compute_reward(err, prev_err, crashes, act, prev_act, 0.6, p)
    assert torch.allclose(rew, want, rtol=1e-5, atol=2e-4), (rew - want).abs().max()
    assert torch.allclose(pe, err, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ["body", "world", "gcomp"])
def test_imu_kernel_matches_reference_fixture(tag):
    """IMUSensor mirror driven with the reference's recorded draws: three updates, bias random walk included."""
    from aerial_gym_simulator_b200.config.sensor_config import BaseImuConfig
    from aerial_gym_simulator_b200.sensors import IMUSensor

    d = np.load(os.path.join(G, "imu_sensor.npz"))

    class Cfg(BaseImuConfig):
        pass
    Cfg.world_frame, Cfg.gravity_compensation = bool(d[f"{tag}_cfg"][0]), bool(d[f"{tag}_cfg"][1])
    n = d[f"{tag}_robot_mass"].shape[0]
    state = torch.zeros(n, 13, device=DEV)
    state[:, 3:7] = _d(d[f"{tag}_robot_orientation"])
    gtd = {"robot_state_tensor": state, "robot_body_angvel": _d(d[f"{tag}_robot_body_angvel"]), "robot_mass": _d(d[f"{tag}_robot_mass"]),
           "force_sensor_tensor": _d(d[f"{tag}_force_sensor_tensor"]), "dt": 0.01, "gravity": torch.tensor([0.0, 0.0, -9.81]).expand(n, -1)}
    imu = IMUSensor(Cfg, n, DEV)
    imu.init_tensors(gtd)
    imu.sensor_quats.copy_(_d(d[f"{tag}_sensor_quats"]))
    imu.bias.copy_(_d(d[f"{tag}_bias0"]))
    draws = d[f"{tag}_draws"]
    for k in range(3):
        imu.update(n_noise=_d(draws[2 * k]), n_bias=_d(draws[2 * k + 1]))
        ref = torch.tensor(d[f"{tag}_meas"][k])
        got = gtd["imu_measurement"].cpu()
        assert torch.allclose(got, ref, rtol=1e-5, atol=2e-5), (tag, k, (got - ref).abs().max())
    assert torch.allclose(imu.bias.cpu(), torch.tensor(d[f"{tag}_bias_end"]), rtol=1e-6, atol=1e-9)


def test_navigation_task_end_to_end():
    """The reference's navigation_task configuration, shrunk: 16 envs of env_with_obstacles (44 boxes), lmf2 with the
    velocity controller, 135x240 depth camera, VAE encoder (initial weights: the checkpoint does not travel)."""
    import aerial_gym_simulator_b200.task  # noqa: F401
    from aerial_gym_simulator_b200.registry.task_registry import task_registry

    task = task_registry.make_task("navigation_task", seed=3, num_envs=16, headless=True)
    obs, rew, term, trunc, info = task.reset()
    assert obs["observations"].shape == (16, 81)
    n_resets = 0
    for step in range(25):
        a = torch.rand(16, 4, device=DEV) * 2 - 1
        obs, rew, term, trunc, info = task.step(a)
        o = obs["observations"]
        assert torch.isfinite(o).all() and torch.isfinite(rew).all()
        assert (o[:, 6] == 0).all()
        # column 3 is the distance to the target in the vehicle frame = |target - position|
        dist = torch.norm(task.target_position - task.obs_dict["robot_position"], dim=1)
        assert torch.allclose(o[:, 3], dist, rtol=1e-5, atol=1e-5)
        assert torch.equal(o[:, 13:17], task.obs_dict["robot_actions"])
        assert (rew[term] == -100.0).all() or not term.any()
        n_resets += int((term | trunc).sum())
    assert task.obs_dict["depth_range_pixels"].shape == (16, 1, 135, 240)
    assert task.image_latents.abs().sum() > 0
    task.close()


def test_imu_in_env_manager():
    """base_quadrotor_with_imu in an empty env: a hovering robot's accelerometer reads +g along body z (specific force),
    the gyro reads the body rates; both within the sensor's noise."""
    from aerial_gym_simulator_b200.sim import SimBuilder

    env = SimBuilder().build_env(sim_name="base_sim", env_name="empty_env", robot_name="base_quadrotor_with_imu",
                                 controller_name="lee_position_control", args={"seed": 1}, device=DEV, num_envs=64, headless=True)
    env.reset()
    gtd = env.get_obs()
    hold = gtd["robot_position"].clone()
    act = torch.cat([hold, torch.zeros(64, 1, device=DEV)], dim=1)  # hold position, yaw 0
    for _ in range(300):
        env.step(act)
        env.render()
    imu = gtd["imu_measurement"]
    assert imu.shape == (64, 6) and torch.isfinite(imu).all()
    up = torch.nn.functional.normalize(imu[:, 0:3], dim=1)
    assert (imu[:, 0:3].norm(dim=1) - 9.81).abs().max() < 1.0       # |specific force| ~ g once settled
    assert (imu[:, 3:6] - gtd["robot_body_angvel"]).abs().max() < 0.2  # gyro = body rates + noise + bias
    assert up[:, 2].min() > 0.9
    env.delete_env()
