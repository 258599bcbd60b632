"""The reference-facing surface end to end ON CPU: task_registry.make_task -> task.reset / step, SimBuilder().build_env ->
EnvManager.step / reset / render, with the CUDA engines replaced by their CPU twins (tests/_cpu_stack.py: host shadow of the HP1 /
per-env device code, brute-force ray-cast oracle).  What runs here is the product's own HOST code -- registries, config plumbing,
RNG call order, reset / curriculum / render bookkeeping, Global-Tensor-Dict aliasing -- i.e. the part of the -m gpu end-to-end
tests (tests/test_env_task_gpu.py, test_zz_*_gpu.py) that does not need a GPU."""
import numpy as np
import pytest
import torch

import aerial_gym_simulator_b200.task  # noqa: F401
from aerial_gym_simulator_b200.registry._core import task_registry
from aerial_gym_simulator_b200.sim import SimBuilder
from oracle import e2e_task_oracle as E
from oracle import lidar_nav_oracle as L
from oracle import obstacle_oracle as OB
from oracle import sensor_noise_oracle as SN
from oracle import sim2real_oracle as S2R

from ._cpu_stack import cpu_stack


@pytest.fixture
def cpu_task():
    made = []

    def make(name, **kw):
        cfg = task_registry.get_task_config(name)
        old = cfg.device
        cfg.device = "cpu"
        made.append((cfg, old))
        return task_registry.make_task(name, headless=True, **kw)
    with cpu_stack() as proxy:
        make.proxy = proxy
        yield make
    for cfg, old in made:
        cfg.device = old


def test_position_setpoint_task_surface(cpu_task):
    task = cpu_task("position_setpoint_task", seed=3, num_envs=70)
    obs, rew, term, trunc, info = task.reset()
    assert obs["observations"].shape == (70, 13) and term.dtype == torch.bool and trunc.dtype == torch.bool
    gtd = task.sim_env.get_obs()
    base = gtd["robot_state_tensor"]
    for k in ("robot_position", "robot_orientation", "robot_linvel", "robot_angvel"):
        assert gtd[k]._base is base  # views of ONE [N,13] row
    g = torch.Generator().manual_seed(0)
    for i in range(505):
        out = task.step((torch.rand(70, 4, generator=g) * 2 - 1) * 0.05)
        assert out[0] is obs and out[1] is rew and out[2] is term and out[3] is trunc
        if i == 499:
            assert not trunc.any() and not term.any() and (task.sim_env.sim_steps == 500).all()
    assert torch.isfinite(obs["observations"]).all() and (task.sim_env.sim_steps == 4).all()
    assert int(task.sim_env.engine.episode_count.min()) == 2  # initial reset + truncation at step 501
    assert torch.equal(obs["observations"][:, 3:7], gtd["robot_orientation"])


def test_navigation_task_end_to_end(cpu_task):
    cfg = task_registry.get_task_config("navigation_task")
    old = cfg.vae_config.use_vae
    cfg.vae_config.use_vae = False  # (the encoder is plain torch, covered by test_aux_oracle_golden.py)
    try:
        task = cpu_task("navigation_task", seed=5, num_envs=3)
        od = task.obs_dict
        assert od["depth_range_pixels"].shape == (3, 1, 135, 240) and od["num_obstacles_in_env"] == 15
        obs, rew, term, trunc, info = task.reset()
        for _ in range(2):
            obs, rew, term, trunc, info = task.step(torch.rand(3, 3) * 2 - 1)
        assert torch.isfinite(obs["observations"][:, :17]).all() and torch.isfinite(rew).all()
        assert cpu_task.proxy.calls["agx_nav_reward"] == 2 and cpu_task.proxy.calls["agx_nav_obs"] == 3
        assert task.sim_env.sensor.captures >= 2 and (od["segmentation_pixels"] >= -2).all()
        px = od["depth_range_pixels"]
        assert (px <= 1.0).all() and (px >= -1.0).all() and (px > 0).any()
    finally:
        cfg.vae_config.use_vae = old


def test_lidar_navigation_task_end_to_end(cpu_task):
    """CPU twin of tests/test_zz_lidar_nav_gpu.py::test_lidar_navigation_task_end_to_end"""
    N = 3
    task = cpu_task("lidar_navigation_task", seed=7, num_envs=N)
    od = task.obs_dict
    assert od["depth_range_pixels"].shape == (N, 1, 48, 120, 3) and od["num_obstacles_in_env"] == 25
    assert task.sim_env.scene.K >= 91
    obs, rew, term, trunc, info = task.reset()
    assert obs["observations"].shape == (N, 337) and (obs["observations"][:, 17:] == 0).all()
    assert (task.target_yaw.abs() <= np.pi).all()
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        a = (torch.rand(N, 4, generator=g) * 2 - 1) * 0.3
        out = task.step(a)
        assert out[0] is obs and out[1] is rew
    assert torch.isfinite(obs["observations"]).all() and torch.isfinite(rew).all()
    assert torch.equal(obs["observations"][:, 17:], task.downsampled_lidar_data)
    assert (task.time_to_collision >= 0).all() and (task.time_to_collision <= 10).all()
    assert torch.allclose(task.current_action[:, 0:3], 2 * a[:, 0:3]) and torch.equal(obs["observations"][:, 13:17], od["robot_actions"])
    task.add_noise_to_downsampled_lidar_data = lambda x: x
    task.process_image_observation()
    pc = od["depth_range_pixels"].squeeze(1)
    want_ds, want_ttc = L.pool(pc, od["robot_position"], od["robot_linvel"])
    r = (pc - od["robot_position"].view(N, 1, 1, 3)).norm(dim=-1)
    near = ((r - 10.0).abs().lt(1e-4) | (r - 0.2).abs().lt(1e-5)).view(N, 16, 3, 20, 6).any(4).any(2)
    assert (torch.isclose(task._image_ds, want_ds, rtol=1e-5, atol=0) | near).all()
    assert torch.allclose(task.time_to_collision, want_ttc, rtol=1e-4, atol=1e-5)
    assert torch.allclose(task.downsampled_lidar_data, (1 / task._image_ds).view(N, -1), rtol=1e-6, atol=0)
    d = np.load(__file__.replace("test_host_stack_cpu.py", "golden/lidar_nav_task_epilogue.npz"))
    p = {k: float(v) for k, v in zip(d["param_names"], d["param_values"])}
    prev_err = task.pos_error_vehicle_frame.clone()
    task.compute_rewards_and_crashes(od)
    want, err = L.rewards_and_errors(od["robot_vehicle_orientation"], od["robot_position"], task.target_position, od["robot_euler_angles"],
                                     task.target_yaw, od["robot_vehicle_linvel"], od["robot_body_angvel"], od["crashes"], task.current_action,
                                     task.prev_action, task.time_to_collision, task.curriculum_progress_fraction, p)
    assert torch.allclose(task.rewards, want, rtol=1e-5, atol=2e-4) and torch.allclose(task.pos_error_vehicle_frame, err, rtol=1e-5, atol=1e-5)
    assert torch.equal(task.pos_error_vehicle_frame_prev, prev_err)
    u1, u2 = torch.rand(N, 3), torch.rand(N, 3)
    task.process_obs_for_task(u1, u2)
    want_obs = L.process_obs(od["robot_vehicle_orientation"], od["robot_position"], task.target_position, od["robot_euler_angles"],
                             task.target_yaw, od["robot_body_linvel"], od["robot_body_angvel"], od["robot_actions"], task.downsampled_lidar_data, u1, u2)
    assert torch.allclose(obs["observations"], want_obs, rtol=1e-5, atol=1e-5)
    task.close()


def test_dynamic_env_end_to_end():
    """CPU twin of tests/test_zz_obstacles_gpu.py::test_dynamic_env_end_to_end"""
    N = 2
    with cpu_stack():
        env = SimBuilder().build_env("base_sim", "dynamic_env", "lmf2", "lmf2_position_control", "cpu", args={"seed": 2}, num_envs=N,
                                     use_warp=True, headless=True)
        gtd = env.get_obs()
        A = gtd["num_obstacles_in_env"]
        assert A == 40 and gtd["num_env_actions"] == 6
        env.reset()
        ast = gtd["env_asset_state_tensor"]
        twist = torch.zeros(N, A, 6)
        twist[..., 0], twist[..., 1], twist[..., 5] = -1.0, 0.5, 0.8
        actions = torch.zeros(N, 4)
        before, upd = ast.clone(), env.scene.updates
        env.step(actions=actions, env_actions=twist)
        want = OB.obstacle_step(before, twist, gtd["dt"], 10, 0.1, 0.1)
        assert torch.allclose(ast, want, rtol=1e-5, atol=1e-5)
        assert torch.allclose(ast[..., 0], before[..., 0] - 0.1 * 0.999, atol=1e-4)
        assert env.scene.updates == upd + 1  # the ray-cast scene was re-posed once for the env step
        assert torch.equal(env._obj_pose, torch.gather(ast[..., 0:7], 1, env._obj_asset.unsqueeze(-1).expand(-1, -1, 7)))
        with pytest.raises(ValueError, match="env_actions"):
            env.step(actions=actions, env_actions=torch.zeros(N, A + 1, 6))
        env.render()
        assert env.sensor.captures == 1
        env2 = SimBuilder().build_env("base_sim", "dynamic_env", "lmf2", "lmf2_position_control", "cpu",
                                      args={"seed": 2, "refit_dynamic_obstacles": False}, num_envs=N, use_warp=True, headless=True)
        env2.reset()
        pose0, ast2 = env2._obj_pose.clone(), env2.get_obs()["env_asset_state_tensor"]
        before2 = ast2.clone()
        env2.step(actions=actions, env_actions=twist)
        assert torch.equal(env2._obj_pose, pose0) and torch.allclose(ast2[..., 0], before2[..., 0] - 0.1 * 0.999, atol=1e-4)


@pytest.mark.parametrize("mode", ["device", "torch"])
def test_env_manager_noisy_lidar(mode):
    """CPU twin of tests/test_zz_sensor_noise_gpu.py::test_env_manager_noisy_lidar (smaller image: the oracle is brute force)"""
    from aerial_gym_simulator_b200.config.robot_config import BaseQuadWithLidarCfg
    from aerial_gym_simulator_b200.config.sensor_config import BaseLidarConfig
    from aerial_gym_simulator_b200.sensors.noise import apply_noise_and_limits_torch

    class SmallNoisyLidar(BaseLidarConfig):
        height, width = 16, 32
    old = BaseQuadWithLidarCfg.sensor_config.lidar_config
    BaseQuadWithLidarCfg.sensor_config.lidar_config = SmallNoisyLidar
    try:
        with cpu_stack():
            N = 2
            args = {"seed": 3, "sensor_noise_rng": mode} if mode == "device" else {"seed": 3, "reset_rng": "torch"}
            env = SimBuilder().build_env("base_sim", "env_with_obstacles", "base_quadrotor_with_lidar", "lee_velocity_control", "cpu", args=args,
                                         num_envs=N, use_warp=True, headless=True)
            assert env.sensor.noise_enabled and env.sensor_noise_rng == mode and (env._device_noise is not None) == (mode == "device")
            assert env.sensor.c.fuse_epilogue == 0
            env.reset()
            gtd = env.get_obs()
            torch.manual_seed(5)
            env.render()
            px = gtd["depth_range_pixels"].clone()
            assert px.shape == (N, 1, 16, 32) and torch.isfinite(px).all() and (px <= 1.0).all() and (px >= -1.0).all()
            env.sensor.noise_enabled = False  # the raw frame again, without the noise pass
            env.render()
            raw = gtd["depth_range_pixels"].clone().numpy()
            env.sensor.noise_enabled = True
            hit = raw < 999.0
            if mode == "device":
                c = env._device_noise.c
                want = SN.noise_limits(raw, c.components, c.enable_noise, c.apply_limits, c.normalize, c.std_a, c.std_b, c.std_c, c.mean_offset,
                                       c.pixel_dropout_prob, c.max_range, c.min_range, c.far_out_of_range_value, c.near_out_of_range_value,
                                       env._device_noise.seed, 0, 0)
                assert np.allclose(px.numpy()[hit], want[hit], rtol=1e-4, atol=1e-5) and env._device_noise.frame == 1
            else:
                t = torch.tensor(raw)
                torch.manual_seed(5)
                apply_noise_and_limits_torch(t, env.sensor_cfg)
                assert torch.equal(px, t)
            sel = hit & (raw > 0.3) & (raw < 5.0)
            assert sel.any() and np.allclose(px.numpy()[sel], (raw[sel] + 0.05) / 10.0, atol=2e-4)
    finally:
        BaseQuadWithLidarCfg.sensor_config.lidar_config = old


@pytest.mark.parametrize("name,tag", [("position_setpoint_task_sim2real_end_to_end", "end_to_end"), ("position_setpoint_task_sim2real_px4", "px4")])
def test_motor_command_tasks_end_to_end(cpu_task, name, tag):
    """tinyprop / x500 + no_control: policy actions -> motor commands -> physics -> reward, crash, truncation, (double) reset, noisy
    rotation-6D observation; every stage against the oracle on the task's own tensors"""
    N = 40
    task = cpu_task(name, seed=3, num_envs=N)
    assert task.sim_env.spec.num_motors == 4 and task.task_config.observation_space_dim == 15
    obs, rew, term, trunc, info = task.reset()
    assert obs["observations"].shape == (N, 15) and torch.isfinite(obs["observations"]).all()
    od = task.obs_dict
    g = torch.Generator().manual_seed(1)
    p = dict(E.E2E_PARAMS[tag])
    for step in range(4):
        a = torch.rand(N, 4, generator=g) * 2.4 - 1.2  # beyond [-1, 1]: clipped by process_actions_for_task
        prev_act, prev_err = task.actions.clone(), task.prev_pos_error.clone()
        task.sim_env.sim_steps[5] = task.task_config.episode_len_steps if step == 2 else task.sim_env.sim_steps[5]  # force one truncation
        ep_before = task.sim_env.engine.episode_count.clone()
        out = task.step(a)
        assert out[0] is obs and out[1] is rew
        lo, hi = task.action_limit_min, task.action_limit_max
        assert torch.allclose(task.actions, torch.clamp(a, -1, 1) * (hi - lo) / 2 + (hi + lo) / 2) and torch.equal(task.prev_actions, task.actions)
        assert torch.equal(prev_act, prev_act)  # (kept for the reward check below)
        if step == 2:
            assert bool(trunc[5]) and int(trunc.sum()) == 1
            assert int(task.sim_env.engine.episode_count[5]) == int(ep_before[5]) + 2  # reset by the env AND again by task.reset_idx (:146-156)
            assert int(task.sim_env.sim_steps[5]) == 0
        assert torch.allclose(task.prev_pos_error, task.target_position - od["robot_position"])
    assert torch.isfinite(rew).all() and torch.isfinite(obs["observations"]).all()
    # reward stage on the current tensors
    crashes_in = od["crashes"].clone()
    task.compute_rewards_and_crashes(od)
    want, cr = E.compute_reward(task.target_position - od["robot_position"], od["robot_orientation"], od["robot_linvel"], od["robot_body_angvel"],
                                crashes_in, task.actions, task.prev_actions, task.prev_pos_error, task.task_config.crash_dist, p)
    assert torch.allclose(task.rewards, want, rtol=1e-5, atol=3e-4) and torch.equal(od["crashes"], cr)
    # observation stage with given draws
    noise = torch.randn(N, 12) * 0.01
    task.process_obs_for_task(noise)
    want_obs = E.process_obs(od["robot_position"], od["robot_orientation"], od["robot_linvel"], od["robot_body_angvel"], task.target_position, noise)
    assert torch.allclose(obs["observations"], want_obs, rtol=1e-5, atol=1e-5)
    r6 = obs["observations"][:, 3:9].view(N, 2, 3)  # two orthonormal rows of a rotation matrix
    assert torch.allclose((r6 * r6).sum(-1), torch.ones(N, 2), atol=1e-5) and ((r6[:, 0] * r6[:, 1]).sum(-1).abs() < 1e-5).all()
    # the reference's draw order and scales: four [N,3] normals with std 1e-3, pi/1032, 2e-3, 1e-3
    calls, real = [], torch.normal
    torch.normal = lambda mean, std: (calls.append((tuple(mean.shape), float(std))), real(mean=mean, std=std))[1]
    try:
        task.process_obs_for_task()
    finally:
        torch.normal = real
    assert calls == [((N, 3), 0.001), ((N, 3), torch.pi / 1032), ((N, 3), 0.002), ((N, 3), 0.001)]
    task.close()


def test_radar_navigation_task_end_to_end(cpu_task):
    """RadarNavigationTask = the LiDAR task on lmf2_radar (48 x 120 world-frame "fake radar" cloud) in env_with_obstacles, with the
    radar noise (80 % of the pooled pixels invalidated) and the radar variant of the reward"""
    N = 3
    task = cpu_task("radar_navigation_task", seed=9, num_envs=N)
    od = task.obs_dict
    assert od["depth_range_pixels"].shape == (N, 1, 48, 120, 3) and task.sim_env.scene.K >= 44 and task._params.radar_variant == 1
    assert task.sim_env.robot_cfg.sensor_config.lidar_config.__name__ == "fake_radar_config"
    obs, rew, term, trunc, info = task.reset()
    for _ in range(2):
        obs, rew, term, trunc, info = task.step(torch.rand(N, 4) * 0.6 - 0.3)
    ds = task.downsampled_lidar_data
    assert torch.isfinite(ds).all() and 0.6 < float((ds == -1.0).float().mean()) < 0.95  # 1 / (-1): the invalidated pixels
    assert torch.equal(obs["observations"][:, 17:], ds) and torch.isfinite(rew).all()
    d = np.load(__file__.replace("test_host_stack_cpu.py", "golden/lidar_nav_task_epilogue.npz"))
    p = {k: float(v) for k, v in zip(d["param_names"], d["param_values"])}
    task.compute_rewards_and_crashes(od)
    args = (od["robot_vehicle_orientation"], od["robot_position"], task.target_position, od["robot_euler_angles"], task.target_yaw,
            od["robot_vehicle_linvel"], od["robot_body_angvel"], od["crashes"], task.current_action, task.prev_action, task.time_to_collision,
            task.curriculum_progress_fraction, p)
    want, _ = L.rewards_and_errors(*args, radar_variant=True)
    assert torch.allclose(task.rewards, want, rtol=1e-5, atol=2e-4)
    task.close()


@pytest.mark.parametrize("name,variant", [("position_setpoint_task_sim2real", 0), ("position_setpoint_task_acceleration_sim2real", 1)])
def test_setpoint_sim2real_tasks_end_to_end(cpu_task, name, variant):
    """lmf2 + velocity / acceleration control: bookkeeping quirks (the caller's action tensor is adopted, and scaled in place by the
    acceleration task), reward / crash / observation stages against the oracle on the task's own tensors"""
    N = 16
    task = cpu_task(name, seed=2, num_envs=N)
    obs, rew, term, trunc, info = task.reset()
    assert obs["observations"].shape == (N, 17) and torch.isfinite(obs["observations"]).all()
    od = task.obs_dict
    assert (od["robot_orientation"][:, 3] >= 0).all()  # process_obs_for_task normalised the sign in place
    g = torch.Generator().manual_seed(3)
    for _ in range(3):
        a = torch.rand(N, 4, generator=g) * 2 - 1
        a0, prev = a.clone(), task.actions.clone()
        pos_before, q_before = od["robot_position"].clone(), od["robot_orientation"].clone()
        out = task.step(a)
        assert out[0] is obs and task.actions is a
        assert torch.equal(a[:, 0:3], 2.0 * a0[:, 0:3]) if variant else torch.equal(a, a0)
        assert torch.equal(task.prev_actions, prev) and torch.allclose(task.prev_dist, (task.target_position - pos_before).norm(dim=1))
        if variant:
            from aerial_gym_simulator_b200.utils.math import quat_rotate
            assert torch.allclose(task.prev_actions_vehicle_frame[:, 0:3], quat_rotate(q_before, prev[:, 0:3]))
        assert torch.equal(obs["observations"][:, 13:17], od["robot_actions"])
    crashes_in = od["crashes"].clone()
    task.compute_rewards_and_crashes(od)
    want, cr, act = S2R.reward(variant, od["robot_position"], od["robot_orientation"], od["robot_vehicle_orientation"], od["robot_body_linvel"],
                               task.target_position, task.prev_dist, task.actions, task.prev_actions_vehicle_frame if variant else task.prev_actions,
                               crashes_in)
    assert torch.allclose(task.rewards, want, rtol=1e-5, atol=2e-3) and torch.equal(od["crashes"], cr)
    if variant:
        assert torch.allclose(task.actions_vehicle_frame, act, atol=1e-6)
    noise = torch.randn(N, 12)
    od["robot_orientation"][::2] *= -1.0
    pos, q = od["robot_position"].clone(), od["robot_orientation"].clone()
    task.process_obs_for_task(noise)
    want_obs, q_after = S2R.process_obs(pos, q, od["robot_body_linvel"], od["robot_body_angvel"], od["robot_actions"], task.target_position, noise)
    assert torch.allclose(obs["observations"], want_obs, rtol=1e-5, atol=1e-5) and torch.equal(od["robot_orientation"], q_after)
    task.close()


def test_generic_eight_rotor_robot_steps():
    """base_random: 8 arbitrarily placed, arbitrarily tilted rotors (thrust axes and arms parsed from random.urdf), velocity control,
    disturbances on -- the generic case of the URDF -> wrench-map reduction through EnvManager"""
    with cpu_stack():
        env = SimBuilder().build_env("base_sim", "empty_env", "base_random", "lee_velocity_control", "cpu", args={"seed": 4}, num_envs=32,
                                     use_warp=False, headless=True)
        assert env.spec.num_motors == 8 and not env.spec.use_rps and env.spec.enable_disturbance
        W, A = env.spec.wrench_map(), np.asarray(env.robot_cfg.control_allocator_config.allocation_matrix)
        assert np.allclose(W[0:3], A[0:3], atol=1e-6)  # thrust axes from the URDF joint rotations = the config's force rows
        env.reset()
        gtd = env.get_obs()
        a = torch.zeros(32, 4)
        a[:, 0] = 0.5
        p0 = gtd["robot_position"].clone()
        for _ in range(50):
            env.step(actions=a)
        assert torch.isfinite(gtd["robot_state_tensor"]).all()
        assert torch.allclose(gtd["robot_orientation"].norm(dim=1), torch.ones(32), atol=1e-5)
        v_forward = gtd["robot_vehicle_linvel"][:, 0]  # the velocity controller pulls the vehicle-frame x velocity towards 0.5 m/s
        assert v_forward.mean() > 0.2 and (gtd["robot_position"] - p0).norm(dim=1).mean() > 0.05


def test_rov_fully_actuated_and_motor_state_survives_reset():
    """base_rov + rov_fully_actuated_control (7-D pose command): holds a pose setpoint; BaseROV.reset_idx leaves the motor model alone
    (robots/base_rov.py:188-201), so thrusts / time constants of a reset env survive (after the first, initialising reset)"""
    with cpu_stack():
        N = 24
        env = SimBuilder().build_env("base_sim", "empty_env", "base_rov", "rov_fully_actuated_control", "cpu", args={"seed": 6}, num_envs=N,
                                     use_warp=False, headless=True)
        assert env.spec.num_motors == 8 and env.num_robot_actions == 7 and env.robot.keeps_motor_state_on_reset
        env.reset()
        gtd, eng = env.get_obs(), env.engine
        assert (eng.motor_thrust != 0).any()  # the first reset initialised the motor model
        a = torch.zeros(N, 7)
        a[:, 6] = 1.0  # position (0, 0, 0), identity orientation
        d0 = gtd["robot_position"].norm(dim=1).clone()
        for _ in range(150):
            env.step(actions=a)
        assert torch.isfinite(gtd["robot_state_tensor"]).all() and (gtd["robot_position"].norm(dim=1) < d0).float().mean() > 0.8
        thrust, tau = eng.motor_thrust.clone(), eng.tau_inc.clone()
        pos = gtd["robot_position"].clone()
        ids = torch.tensor([1, 5, 7])
        env.reset_idx(ids)
        assert torch.equal(eng.motor_thrust, thrust) and torch.equal(eng.tau_inc, tau)       # motor model untouched, everywhere
        assert not torch.equal(gtd["robot_position"][ids], pos[ids])                           # ... while the state was re-drawn
        keep = torch.ones(N, dtype=torch.bool)
        keep[ids] = False
        assert torch.equal(gtd["robot_position"][keep], pos[keep])


def test_forest_env_trees_as_cylinders_with_per_link_segmentation():
    """forest_env: one tree (13 cylinder links, tessellated into 32-gon prisms and cut into 12-triangle parts), 35 objects, the floor;
    per-link segmentation ids (assets/warp_asset.py:44-70) and the running counter advancing by the number of links"""
    from aerial_gym_simulator_b200.hp2 import cylinder_triangles

    t = cylinder_triangles(0.5, 2.0).reshape(-1, 3, 3).astype(np.float64)
    assert t.shape == (128, 3, 3)
    vol = np.einsum("ni,ni->n", t[:, 0], np.cross(t[:, 1], t[:, 2])).sum() / 6.0  # closed, outward-oriented mesh
    assert abs(vol - 0.5 * 32 * 0.25 * np.sin(2 * np.pi / 32) * 2.0) < 1e-5
    assert np.abs(np.linalg.norm(t[..., :2], axis=-1)[np.linalg.norm(t[..., :2], axis=-1) > 1e-6] - 0.5).max() < 1e-6
    with cpu_stack():
        N = 2
        env = SimBuilder().build_env("base_sim", "forest_env", "base_quadrotor_with_camera", "lee_velocity_control", "cpu", args={"seed": 1},
                                     num_envs=N, use_warp=True, headless=True)
        gtd, sc = env.get_obs(), env.scene
        assert gtd["num_obstacles_in_env"] == 37 and env.keep_in_env == 2  # tree + floor are kept, the 35 objects follow the curriculum
        assert sc.K == 13 * 11 + 35 + 1 and sc.L == 12                     # 13 cylinders x ceil(128 / 12) parts + boxes
        env.reset()
        env.render()
        seg = gtd["segmentation_pixels"]
        ids = torch.unique(seg[seg >= 0]).tolist()
        assert 13 in ids or any(100 <= i for i in ids)  # floor (fixed id 13) or instances
        # the tree is the asset listed first among the kept ones or second: its 13 links take 13 consecutive ids, the other assets one each
        ctr = sc.obj_seg_counter.numpy()
        base = sc.tmpl_seg_base.numpy()[sc.tmpl_tri_offset.numpy()[sc.obj_template.numpy()]]  # seg base of each object's first triangle
        mask = sc.tmpl_seg_mask.numpy()[sc.tmpl_tri_offset.numpy()[sc.obj_template.numpy()]]
        final = base + ctr * mask
        tree_parts = (env._obj_asset.numpy() == np.argmax(np.bincount(env._obj_asset.numpy()[0])))  # the asset with the most parts
        assert tree_parts[0].sum() == 143
        tree_ids = np.unique(final[0][tree_parts[0]])
        assert len(tree_ids) == 13 and tree_ids.max() - tree_ids.min() == 12
        others = np.unique(final[0][~tree_parts[0] & (mask[0] == 1)])
        assert len(others) == 35 and not set(others.tolist()) & set(tree_ids.tolist())
        assert final[1][mask[1] == 1].min() > final[0][mask[0] == 1].max()  # env 1 continues the running counter
        assert torch.isfinite(gtd["depth_range_pixels"]).all() and (gtd["depth_range_pixels"] > 0).any()


def test_task_torch_rng_mode_follows_reference_call_order(cpu_task):
    """CPU twin of tests/test_env_task_gpu.py::test_task_torch_rng_mode_follows_reference_call_order: reset_rng='torch' draws full-size
    uniforms with torch.rand in the reference's order (bounds lo, bounds hi, state, motor tau_inc, tau_dec, thrust, k), only when an env
    resets"""
    from oracle import hp1_oracle as O
    from tests import _hp1_common as H

    cfg = task_registry.get_task_config("position_setpoint_task")
    old = cfg.args
    cfg.args = {"reset_rng": "torch"}
    try:
        task = cpu_task("position_setpoint_task", seed=5, num_envs=48)
    finally:
        cfg.args = old
    N = 48
    calls, orig = [], torch.rand

    def spy(*a, **k):
        out = orig(*a, **k)
        calls.append(out.clone())
        return out
    torch.rand = spy
    try:
        torch.manual_seed(1234)
        task.reset()
    finally:
        torch.rand = orig
    assert [tuple(c.shape) for c in calls] == [(N, 3), (N, 3), (N, 13), (N, 4), (N, 4), (N, 4), (N, 4)]
    eng = task.sim_env.engine
    model = H.oracle_model_from_spec(task.sim_env.spec)
    draws = O.ResetDraws(calls[0], calls[1], calls[2], None, None, None, None, calls[3], calls[4], calls[5], calls[6])
    st = O.make_state(model, N)
    O.reset_envs(model, st, torch.ones(N, dtype=torch.bool), draws)
    H.assert_close(eng.root_state, st.root, "reset state (torch rng order)", scale=1.0)
    H.assert_close(eng.motor_thrust, st.thrust, "reset thrust")
    first = eng.root_state.clone()
    torch.manual_seed(1234)
    task.reset()
    assert torch.equal(eng.root_state, first)  # same seed, same episode start
    s0 = torch.get_rng_state().clone()
    task.step(torch.zeros(N, 4))
    assert torch.equal(torch.get_rng_state(), s0)  # a step without resets does not consume the generator
    eng.sim_steps[7] = 500
    task.step(torch.zeros(N, 4))
    assert not torch.equal(torch.get_rng_state(), s0)
    assert int(eng.sim_steps[7]) == 0 and bool(task.truncations[7]) and int(task.truncations.sum()) == 1


def test_imu_in_env_manager():
    """CPU twin of tests/test_aux_gpu.py::test_imu_in_env_manager: base_quadrotor_with_imu hovering -- the accelerometer reads +g along
    body z (specific force from the engine's net body wrench), the gyro the body rates, both within the sensor's noise"""
    with cpu_stack() as proxy:
        env = SimBuilder().build_env(sim_name="base_sim", env_name="empty_env", robot_name="base_quadrotor_with_imu",
                                     controller_name="lee_position_control", args={"seed": 1}, device="cpu", num_envs=32, headless=True)
        env.reset()
        gtd = env.get_obs()
        act = torch.cat([gtd["robot_position"].clone(), torch.zeros(32, 1)], dim=1)  # hold position, yaw 0
        for _ in range(300):
            env.step(act)
            env.render()
        imu = gtd["imu_measurement"]
        assert imu.shape == (32, 6) and torch.isfinite(imu).all() and proxy.calls["agx_imu_update"] == 300
        assert (imu[:, 0:3].norm(dim=1) - 9.81).abs().max() < 1.0
        assert (imu[:, 3:6] - gtd["robot_body_angvel"]).abs().max() < 0.2
        assert torch.nn.functional.normalize(imu[:, 0:3], dim=1)[:, 2].min() > 0.9


def test_thin_assets_switch_on():
    """env_object_config.thin_asset_params is off by default (num_assets = 0, "thin": False); a user who switches it on gets that many
    thin rods (one box each, resources/.../thin/*.urdf) in the scene, with incremental per-instance segmentation ids"""
    from aerial_gym_simulator_b200.config import asset_config as AC, env_config as EC
    from aerial_gym_simulator_b200.sim import SimBuilder
    inc, old_n = EC.EnvWithObstaclesCfg.env_config.include_asset_type, AC.thin_asset_params.num_assets
    try:
        with cpu_stack():
            base = SimBuilder().build_env("base_sim", "env_with_obstacles", "base_quadrotor_with_camera", "lee_velocity_control", "cpu",
                                          args={"seed": 1}, num_envs=2, headless=True, use_warp=True)
            inc["thin"], AC.thin_asset_params.num_assets = True, 5
            env = SimBuilder().build_env("base_sim", "env_with_obstacles", "base_quadrotor_with_camera", "lee_velocity_control", "cpu",
                                         args={"seed": 1}, num_envs=2, headless=True, use_warp=True)
            assert env.num_obs_in_env == base.num_obs_in_env + 5 and env.IGE_env.num_assets_per_env == env.num_obs_in_env + 1
            env.reset()
            env.step(torch.zeros(2, 4))
            env.render()
            assert torch.isfinite(env.global_tensor_dict["depth_range_pixels"]).all()
    finally:
        inc["thin"], AC.thin_asset_params.num_assets = False, old_n


@pytest.mark.parametrize("per_env", [False, True])
def test_controller_gains_can_be_read_and_set_like_the_reference(per_env):
    """examples/tune_controllers.py reads controller.K_*_tensor_current and calls controller.set_controller_gains(...) between steps
    (base_lee_controller.py:59-62, 78-85).  Here the gains live in the engine (constants in AgxHp1Config, or per-env arrays)."""
    from aerial_gym_simulator_b200.sim import SimBuilder
    with cpu_stack():
        args = {"seed": 1, "per_env_params": "all"} if per_env else {"seed": 1}
        env = SimBuilder().build_env("base_sim", "empty_env", "base_quadrotor", "lee_position_control", "cpu", args=args, num_envs=8,
                                     headless=True)
        ctrl = env.robot_manager.robot.controller
        assert torch.allclose(ctrl.K_pos_tensor_current, torch.tensor([2.5, 2.5, 1.5]).expand(8, 3))  # mid-point of the config range
        assert torch.equal(ctrl.K_rot_tensor_max, torch.tensor([1.2, 1.2, 0.6]).expand(8, 3))
        env.reset()
        act = torch.zeros(8, 4)
        act[:, 0] = 1.0

        def run(gain_scale):
            env.reset()
            env.engine.root_state[:] = 0.0  # same start for both runs: origin, level, at rest
            env.engine.root_state[:, 6] = 1.0
            env.engine.motor_thrust[:] = 0.25 * 9.81 / 4.0
            env.engine.refresh()
            ctrl.set_controller_gains(gain_scale * torch.tensor([2.5, 2.5, 1.5]), ctrl.K_linvel_tensor_current.clone(),
                                      ctrl.K_rot_tensor_current.clone(), ctrl.K_angvel_tensor_current.clone())
            for _ in range(20):
                env.step(act)
            return env.global_tensor_dict["robot_position"][:, 0].clone()
        slow, fast = run(0.5), run(2.0)
        assert torch.allclose(ctrl.K_pos_tensor_current, 2.0 * torch.tensor([2.5, 2.5, 1.5]).expand(8, 3))
        assert (fast > slow + 1e-4).all()  # a stiffer position loop moves further towards the setpoint in the same time
        env.reset()
        assert torch.allclose(ctrl.K_pos_tensor_current, 2.0 * torch.tensor([2.5, 2.5, 1.5]).expand(8, 3))  # not randomised: a reset keeps them
        per_env_gain = torch.rand(8, 3) + 1.0
        if per_env:
            ctrl.set_controller_gains(per_env_gain, 2.5, 1.0, 0.15)
            assert torch.equal(ctrl.K_pos_tensor_current, per_env_gain) and (ctrl.K_linvel_tensor_current == 2.5).all()
        else:
            with pytest.raises(RuntimeError, match="per_env_params"):
                ctrl.set_controller_gains(per_env_gain, 2.5, 1.0, 0.15)


def _position_task_cfg(**kw):
    from aerial_gym_simulator_b200.config.task_config import position_setpoint_task_config as C
    return type("cfg", (C,), dict(device="cpu", num_envs=48, episode_len_steps=6, reward_parameters=dict(C.reward_parameters), **kw))


def test_position_task_hooks_can_be_overridden_like_in_the_reference():
    """position_setpoint_task.py:194-229: process_obs_for_task / compute_rewards_and_crashes are methods a user subclass may override.
    The base class fuses both into the step kernel; with an override the step keeps the reference's order.  Checked against a fused
    run of the base class on the same seed (same Philox reset streams): rewards differ exactly by the override's bonus, terminations /
    truncations / resets coincide, and the default hooks reached through super() reproduce the kernel's values."""
    from aerial_gym_simulator_b200.task.position_setpoint_task import PositionSetpointTask

    class ObsOnly(PositionSetpointTask):
        def process_obs_for_task(self):
            super().process_obs_for_task()
            self.task_obs["observations"][:, 0:3] *= 0.5

    class RewardToo(PositionSetpointTask):
        calls = 0

        def compute_rewards_and_crashes(self, obs_dict):
            type(self).calls += 1
            rew, crashes = super().compute_rewards_and_crashes(obs_dict)
            return rew + 1.25, crashes

    with cpu_stack():
        base = PositionSetpointTask(_position_task_cfg(), seed=5, headless=True)
        obs_only = ObsOnly(_position_task_cfg(), seed=5, headless=True)
        rew_too = RewardToo(_position_task_cfg(), seed=5, headless=True)
        tasks = (base, obs_only, rew_too)
        for t in tasks:
            t.reset()
        assert torch.allclose(obs_only.task_obs["observations"][:, 0:3], 0.5 * base.task_obs["observations"][:, 0:3])
        g = torch.Generator().manual_seed(0)
        resets = 0
        for step in range(16):
            a = torch.rand(48, 4, generator=g) * 2 - 1
            if step == 3:  # fly three envs out of the 8 m crash radius
                for t in tasks:
                    t.sim_env.engine.root_state[5:8, 0] = 9.0
            outs = [t.step(a.clone()) for t in tasks]
            (o0, r0, te0, tr0, _), (o1, r1, te1, tr1, _), (o2, r2, te2, tr2, _) = outs
            assert torch.equal(te0, te1) and torch.equal(te0, te2) and torch.equal(tr0, tr1) and torch.equal(tr0, tr2), step
            resets += int((te0 | tr0).sum())
            if step == 3:
                assert te0[5:8].all() and int(te0.sum()) == 3  # crashed, and (device RNG) already re-initialised inside the same step
            assert torch.allclose(r1, r0, rtol=1e-6, atol=1e-6) and torch.allclose(r2, r0 + 1.25, rtol=1e-5, atol=1e-5), step
            assert torch.allclose(o1["observations"][:, 3:], o0["observations"][:, 3:], atol=1e-6)
            assert torch.allclose(o1["observations"][:, 0:3], 0.5 * o0["observations"][:, 0:3], atol=1e-6)
            assert torch.allclose(o2["observations"], o0["observations"], rtol=1e-5, atol=1e-5), step
            assert torch.equal(base.sim_env.sim_steps, rew_too.sim_env.sim_steps)
        assert resets == 48 * 2 and RewardToo.calls == 16  # two episodes end per env in 16 steps (6-step episodes), crashed or not
        assert int(base.sim_env.engine.episode_count.min()) >= 2


def test_position_task_return_state_before_reset():
    """task_config.return_state_before_reset = True (position_setpoint_task.py:170-171): the observation returned for an env that just
    crashed is the one BEFORE its reset -- served by the un-fused sequence (the fused step resets inside the launch)"""
    from aerial_gym_simulator_b200.task.position_setpoint_task import PositionSetpointTask
    with cpu_stack():
        task = PositionSetpointTask(_position_task_cfg(return_state_before_reset=True), seed=2, headless=True)
        task.reset()
        task.sim_env.engine.root_state[3, 0] = 9.5
        obs, rew, term, trunc, _ = task.step(torch.zeros(48, 4))
        assert bool(term[3]) and int(term.sum()) == 1 and float(rew[3]) == -20.0
        assert abs(float(obs["observations"][3, 0]) + 9.5) < 0.2  # target - position of the crashed state, not of the new episode
        assert abs(float(task.obs_dict["robot_position"][3, 0])) <= 1.0 and int(task.sim_env.sim_steps[3]) == 0  # ... which has begun


def test_simulate_is_one_physics_step():
    """EnvManager.simulate (env_manager.py:346-349) for callers that drive the physics loop themselves: n calls == one step() of an
    env with n physics steps per env step (same kernels, same state), minus the env-step bookkeeping (sim_steps, collision flags)"""
    from aerial_gym_simulator_b200.sim import SimBuilder
    with cpu_stack():
        mk = lambda: SimBuilder().build_env("base_sim", "empty_env_2ms", "base_quadrotor", "lee_velocity_control", "cpu", args={"seed": 4},
                                            num_envs=16, headless=True)
        a, b = mk(), mk()
        a.reset()
        b.reset()
        assert torch.equal(a.engine.root_state, b.engine.root_state)
        act = torch.full((16, 4), 0.3)
        a.step(act)
        for _ in range(5):  # EnvCfg2Ms: 5 physics steps per env step
            b.simulate(act)
        assert torch.allclose(a.engine.root_state, b.engine.root_state, rtol=1e-6, atol=1e-7)
        assert int(a.sim_steps[0]) == 1 and int(b.sim_steps[0]) == 0
        assert b.render_viewer() is None and b.log_memory_use() is None
