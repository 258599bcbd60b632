"""CPU tests of the host-side mirror of the reference surface: registries, configs, URDF-derived
robot constants, the MultirotorSpec -> AgxHp1Config packing.  No GPU, no compute calls."""
import os

import numpy as np
import pytest

from aerial_gym_simulator_b200 import _lib, urdf
from aerial_gym_simulator_b200.config import RESOURCES_DIRECTORY
from aerial_gym_simulator_b200.config.env_config import EmptyEnvCfg, EnvWithObstaclesCfg
from aerial_gym_simulator_b200.config.sim_config import BaseSimConfig
from aerial_gym_simulator_b200.hp1 import build_config
from aerial_gym_simulator_b200.registry._core import (controller_registry, env_config_registry, robot_registry,
                                                sim_config_registry, task_registry)
import aerial_gym_simulator_b200.task  # noqa: F401  (registers everything)


def test_registry_surface_matches_reference_names():
    # control/__init__.py:42-99 of the reference registers these 22 names
    want = {"no_control", "rov_fully_actuated_control"}
    for fam in ("lee", "magpie", "lmf2", "octarotor"):
        want |= {f"{fam}_{k}_control" for k in ("position", "velocity", "attitude", "rates", "acceleration")}
    assert want <= set(controller_registry.get_controller_names())
    assert {"base_quadrotor", "base_octarotor", "base_quad_root_link_control", "lmf2", "x500"} <= set(robot_registry.get_robot_names())
    assert "position_setpoint_task" in task_registry.get_task_names()
    assert {"empty_env", "env_with_obstacles"} <= set(env_config_registry.get_env_names())
    assert "base_sim" in sim_config_registry.get_sim_names()
    with pytest.raises(ValueError):
        controller_registry.make_controller("nope", 4, "cpu")
    with pytest.raises(ValueError):
        robot_registry.make_robot("nope", "lee_attitude_control", EmptyEnvCfg, "cpu")


def test_urdf_composite_inertia_quad():
    """SURVEY 8c hand derivation: m = 0.25 kg, J = diag(8.45e-4, 8.45e-4, 1.69e-3) about a centred COM."""
    m = urdf.parse_urdf(os.path.join(RESOURCES_DIRECTORY, "robots/quad/quad.urdf"))
    mass, com, J = m.composite_inertia()
    assert abs(mass - 0.25) < 1e-12 and np.abs(com).max() < 1e-12
    assert np.allclose(np.diag(J), [8.45e-4, 8.45e-4, 1.69e-3], rtol=1e-9) and np.abs(J - np.diag(np.diag(J))).max() < 1e-15
    order = m.body_order()
    assert order[0] == "base_link" and order[5:9] == ["motor_0", "motor_1", "motor_2", "motor_3"]  # application_mask 5..8


def test_wrench_map_reproduces_quad_allocation_matrix():
    """Appendix B: for base_quadrotor the per-link force/torque reduction equals the config's
    allocation matrix (base_quad_config.py:166-173)."""
    robot, cfg = robot_registry.make_robot("base_quadrotor", "lee_attitude_control", EmptyEnvCfg, "cpu")
    spec = robot.make_spec(BaseSimConfig, EmptyEnvCfg)
    assert np.allclose(spec.wrench_map(), np.asarray(cfg.control_allocator_config.allocation_matrix), atol=1e-12)
    # tilted octarotor: thrust axes come from the URDF joint rotations
    robot8, cfg8 = robot_registry.make_robot("base_octarotor", "octarotor_velocity_control", EmptyEnvCfg, "cpu")
    spec8 = robot8.make_spec(BaseSimConfig, EmptyEnvCfg)
    W = spec8.wrench_map()
    assert np.allclose(W[0:3], np.asarray(cfg8.control_allocator_config.allocation_matrix)[0:3], atol=1e-6)
    assert spec8.randomize_params and spec8.num_actions == 4 and not spec8.use_rps


def test_spec_to_abi_config_packing():
    robot, _ = robot_registry.make_robot("base_quadrotor", "lee_position_control", EnvWithObstaclesCfg, "cpu")
    spec = robot.make_spec(BaseSimConfig, EnvWithObstaclesCfg)
    c = build_config(spec, 128, physics_steps=10, seed=0x1_0000_0002, env_id_offset=256)
    assert (c.num_envs, c.num_motors, c.controller, c.num_actions, c.physics_steps) == (128, 4, _lib.CTRL_POSITION, 4, 10)
    assert c.seed == 0x1_0000_0002 and c.env_id_offset == 256
    assert c.flags & _lib.F_USE_RPS and c.flags & _lib.F_MOTOR_RK4 and c.flags & _lib.F_DISCRETE_MIX
    assert abs(c.mass - 0.25) < 1e-7 and abs(c.inertia[8] - 1.69e-3) < 1e-9 and abs(c.inertia_inv[0] - 1 / 8.45e-4) < 1e-2
    # pinv of the rank-4 quad allocation matrix: f = pinv(A) w reproduces collective thrust
    P = np.array(list(c.alloc_pinv)[:24]).reshape(4, 6)
    assert np.allclose(P @ np.array([0, 0, 4.0, 0, 0, 0]), [1, 1, 1, 1], atol=1e-5)
    assert list(c.bounds_lo_min) == [-2.0, -4.0, -3.0] and list(c.bounds_hi_max) == [10.0, 4.0, 3.0]
    assert abs(c.K_rot[2] - 0.5) < 1e-7  # mid-point of [0.4, 0.6]


def test_fused_controllers_refuse_direct_calls():
    ctrl, cfg = controller_registry.make_controller("lee_velocity_control", 8, "cpu")
    assert ctrl.CONTROLLER_ID == _lib.CTRL_VELOCITY and cfg.num_actions == 4
    with pytest.raises(RuntimeError, match="fused"):
        ctrl(None)


def test_root_link_no_control_is_rejected():
    with pytest.raises(ValueError, match="root_link"):
        robot_registry.make_robot("base_quad_root_link_control", "no_control", EmptyEnvCfg, "cpu")


def test_compat_import_paths_of_the_reference():
    """rl_training/* and examples/* import these paths; compat.install()
    makes them resolve to this package, including the isaacgym shim."""
    import aerial_gym_simulator_b200.compat as compat

    compat.install()
    import isaacgym  # noqa: F401
    from isaacgym import gymutil
    from aerial_gym.registry.task_registry import task_registry as tr
    from aerial_gym.registry.robot_registry import robot_registry as rr
    from aerial_gym.sim.sim_builder import SimBuilder  # noqa: F401
    from aerial_gym.utils.helpers import parse_arguments  # noqa: F401
    from aerial_gym.utils.logging import CustomLogger
    from aerial_gym.config.task_config.position_setpoint_task_config import task_config

    assert tr is task_registry and rr is robot_registry
    assert gymutil.parse_device_str("cuda:3") == ("cuda", 3)
    assert task_config.controller_name == "lee_attitude_control" and task_config.episode_len_steps == 500
    CustomLogger("t").setLoggerLevel("INFO")
    # SURVEY 8(f) rows built in round 1: navigation task, IMU, VAE encoder wrapper
    from aerial_gym.config.task_config.navigation_task_config import task_config as nav_cfg
    from aerial_gym.sensors.imu_sensor import IMUSensor  # noqa: F401
    from aerial_gym.task.navigation_task.navigation_task import NavigationTask  # noqa: F401
    from aerial_gym.utils.vae.vae_image_encoder import VAEImageEncoder  # noqa: F401

    assert tr.get_task_config("navigation_task") is nav_cfg and nav_cfg.robot_name == "lmf2"
    # the one-class-per-file sensor catalogue
    from aerial_gym.config.sensor_config.camera_config.d455_depth_config import RsD455Config
    from aerial_gym.config.sensor_config.imu_config.vn100_config import VN100Config  # noqa: F401
    from aerial_gym.config.sensor_config.lidar_config.os1_64_config import OS_1_64_Config
    from aerial_gym.config.env_config.dynamic_environment import DynamicEnvironmentCfg  # noqa: F401

    assert (RsD455Config.height, RsD455Config.far_out_of_range_value) == (270, 15.0)
    assert (OS_1_64_Config.max_range, OS_1_64_Config.far_out_of_range_value) == (90.0, 35.0)  # the reference's inherited far value


def test_navigation_task_config_matches_reference_values():
    """Reward parameters / curriculum / action transform of config/task_config/navigation_task_config.py."""
    import math

    import torch

    from aerial_gym_simulator_b200.config.task_config import navigation_task_config as C

    assert C.observation_space_dim == 81 and C.episode_len_steps == 100 and C.env_name == "env_with_obstacles"
    assert C.reward_parameters["collision_penalty"] == -100.0 and C.reward_parameters["pos_reward_exponent"] == 1.0 / 3.5
    assert (C.curriculum.min_level, C.curriculum.max_level, C.curriculum.check_after_log_instances) == (15, 50, 2048)
    a = torch.tensor([[1.0, 0.5, -1.0, 0.3], [-1.0, 0.0, 0.5, 0.0], [3.0, -2.0, 2.0, 0.0]])
    out = C.action_transformation_function(a.clone())
    c = torch.clamp(a, -1, 1)
    want_x = (c[:, 0] + 1) * torch.cos(math.pi / 4 * c[:, 1]) * 2.0 / 2.0
    want_z = (c[:, 0] + 1) * torch.sin(math.pi / 4 * c[:, 1]) * 2.0 / 2.0
    assert torch.allclose(out[:, 0], want_x) and torch.allclose(out[:, 2], want_z)
    assert torch.allclose(out[:, 3], c[:, 2] * math.pi / 3) and (out[:, 1] == 0).all()


def test_lidar_navigation_task_surface_and_config():
    """LiDARNavigationTask is registered with the reference's names and values (config/task_config/
    lidar_navigation_task_config.py, env_with_lidar_nav_obstacles.py, rslidar_airy_config.py, magpie_config.py:52-53)."""
    import math

    import torch

    import aerial_gym_simulator_b200.compat as compat
    from aerial_gym_simulator_b200.config import asset_config as ac
    from aerial_gym_simulator_b200.config.robot_config import MagpieCfg
    from aerial_gym_simulator_b200.config.sensor_config import RSLidar_Airy_Config
    from aerial_gym_simulator_b200.config.task_config import lidar_navigation_task_config as C

    compat.install()
    from aerial_gym.config.task_config.lidar_navigation_task_config import task_config
    from aerial_gym.task.lidar_navigation_task.lidar_navigation_task import LiDARNavigationTask

    assert task_registry.get_task_config("lidar_navigation_task") is task_config is C
    assert task_registry.get_task_class("lidar_navigation_task") is LiDARNavigationTask
    assert (C.robot_name, C.controller_name, C.env_name) == ("magpie", "magpie_acceleration_control", "env_with_lidar_nav_obstacles")
    assert C.observation_space_dim == 337 and C.episode_len_steps == 110 and C.reward_parameters["collision_penalty"] == -10.0
    assert (C.curriculum.min_level, C.curriculum.max_level) == (25, 70)
    assert "magpie_acceleration_control" in controller_registry.get_controller_names()
    assert MagpieCfg.sensor_config.enable_lidar and MagpieCfg.sensor_config.lidar_config is RSLidar_Airy_Config
    L = RSLidar_Airy_Config
    assert (L.height, L.width, L.return_pointcloud, L.pointcloud_in_world_frame, L.normalize_range) == (48, 120, True, True, False)
    env = env_config_registry.make_env("env_with_lidar_nav_obstacles")
    m = env.env_config.asset_type_to_dict_map
    on = [p for k, p in m.items() if env.env_config.include_asset_type.get(k, True)]  # (thin / trees / tiles are in the map, switched off)
    assert sum(p.num_assets for p in on) == 15 + 70 + 6 and not any(p.keep_in_env for p in m.values())
    assert m["panels"].min_state_ratio[0] == 0.35 and m["objects"].max_state_ratio[0:3] == [1.0, 1.0, 1.0]
    assert env.env.lower_bound_min == [-7.5, -7.5, -5.0] and env.env.upper_bound_max == [7.5, 7.5, 5.0]
    assert ac.panel_asset_params.num_assets == 3 and ac.left_wall.keep_in_env  # the navigation-task scene is untouched
    a = torch.tensor([[0.5, -2.0, 0.25, 1.0], [3.0, 0.1, -0.3, -0.5]])
    out = C.action_transformation_function(a)
    assert torch.allclose(out[:, 0:3], 2 * torch.clamp(a[:, 0:3], -1, 1)) and torch.allclose(out[:, 3], torch.clamp(a[:, 3], -1, 1) * math.pi / 3)


def test_lidar_task_noise_reproduces_the_reference_draws():
    """The host-side noise of LiDARNavigationTask (torch RNG, CPU here) against the fixture recorded from the reference's own
    add_noise_to_downsampled_lidar_data under the same seed."""
    import torch

    from aerial_gym_simulator_b200.task.lidar_navigation_task import add_noise_to_downsampled_lidar_data

    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "lidar_nav_task_epilogue.npz"))
    torch.manual_seed(int(d["pool_noise_seed"]))
    noisy = add_noise_to_downsampled_lidar_data(torch.tensor(d["pool_image_ds"]))
    assert torch.equal(noisy, torch.tensor(d["pool_image_noisy"]))


def test_dynamic_env_config():
    """config/env_config/dynamic_environment.py + dynamic_env_object_config.py"""
    env = env_config_registry.make_env("dynamic_env")
    m = env.env_config.asset_type_to_dict_map
    on = [k for k in m if env.env_config.include_asset_type.get(k, True)]
    assert on == ["objects"] and m["objects"].num_assets == 40 and not m["objects"].fix_base_link and m["objects"].disable_gravity
    assert all(p.disable_gravity and not p.fix_base_link for p in m.values())  # every class of that file floats
    assert env.env.num_env_actions == 6 and env.env.write_to_sim_at_every_timestep and env.env.lower_bound_min[2] == 0.0
    assert env.env.num_physics_steps_per_env_step_mean == 10


def test_position_task_default_reward_hook_matches_reference_fixture():
    """PositionSetpointTask.compute_rewards_and_crashes (the torch body a subclass reaches through super(); the base class itself uses
    the fused kernel) against the fixture recorded from the reference's compute_reward, including pre-set collision flags"""
    import torch

    from aerial_gym_simulator_b200.task.position_setpoint_task import PositionSetpointTask
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hp1_position_reward.npz"))
    n = z["pos"].shape[0]
    fake = type("T", (), {"target_position": torch.zeros(n, 3)})()
    od = {"robot_position": torch.tensor(z["pos"]), "robot_orientation": torch.tensor(z["quat"]),
          "robot_body_angvel": torch.tensor(z["body_angvel"]), "crashes": torch.tensor(z["crashes_in"]).clone()}
    rew, crashes = PositionSetpointTask.compute_rewards_and_crashes(fake, od)
    assert torch.equal(crashes, torch.tensor(z["crashes_out"])) and crashes is od["crashes"]
    assert torch.allclose(rew, torch.tensor(z["reward"]), rtol=1e-5, atol=1e-5)
