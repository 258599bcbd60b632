"""EnvManager -- reference surface (env_manager/env_manager.py:37-436) over Hp1Engine / RayScene /
RaySensor.  Host logic only: RNG draws (in the reference's call order), bookkeeping, the Global
Tensor Dict.  All per-env arithmetic is in the CUDA library.

Differences that are deliberate and documented:
  * no Isaac Gym: ``robot_state_tensor`` is its own contiguous [N,13] allocation (not a stride-13*A
    view of ``vec_root_tensor``); obstacles live in ``env_asset_state_tensor`` [N,K,13];
  * construction is array-native (no O(N) actor loop), except asset-file selection which keeps the
    reference's per-env ``random.choices`` / ``random.shuffle`` order;
  * collision flags (a14) come from a geometric sphere-vs-mesh overlap test, not PhysX contact
    forces; ``robot_contact_force_tensor`` stays 0.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import random
from collections import deque

import numpy as np
import torch

from .. import _lib
from .. import robots  # noqa: F401  (registers robots + controllers)
from .. import urdf
from ..hp1 import Hp1Engine
from ..hp2 import RayScene, RaySensor
from ..registry._core import env_config_registry, robot_registry, sim_config_registry


def _lerp(lo, hi, u):  # utils/math.py:51-54 torch_rand_float_tensor arithmetic
    return (hi - lo) * u + lo


def _quat_from_euler(e):  # utils/math.py:155-172
    roll, pitch, yaw = e[..., 0], e[..., 1], e[..., 2]
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    return torch.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp,
                        cy * cr * cp + sy * sr * sp], dim=-1)


def _quat_rotate_inverse(q, v):
    qw, qv = q[:, 3:4], q[:, 0:3]
    return v * (2.0 * qw * qw - 1.0) - torch.cross(qv, v, dim=-1) * qw * 2.0 + qv * (qv * v).sum(-1, keepdim=True) * 2.0


class _RobotManagerView:
    """robots/robot_manager.py: RobotManagerIGE as callers see it (the arithmetic it hosts in the reference runs in the HP1 kernel)"""

    def __init__(self, env):
        self._env, self.robot, self.cfg = env, env.robot, env.robot_cfg
        self.num_envs, self.device = env.num_envs, env.device
        self.robot_mass, self.robot_inertia = env.robot.robot_mass, env.robot.robot_inertia  # :45-46, scalars / 3x3 of the URDF
        self.dof_control_mode = "none"

    robot_masses = property(lambda self: self._env.global_tensor_dict["robot_mass"])      # :47  [N]
    robot_inertias = property(lambda self: self._env.global_tensor_dict["robot_inertia"])  # :48  [N,3,3]
    imu_sensor = property(lambda self: self._env.imu)
    warp_sensor = lidar_sensor = property(lambda self: self._env.sensor)
    camera_sensor = property(lambda self: None)  # Isaac Gym's rasterised cameras do not exist here


class _SimView:
    """env_manager/IGE_env_manager.py: the counts callers read off IsaacGymEnv"""

    def __init__(self, env):
        self._env, self.sim_config, self.num_envs = env, env.sim_config, env.num_envs
        self.sim_has_dof, self.dof_control_mode, self.viewer, self.has_viewer = False, "none", None, False

    num_assets_per_env = property(lambda self: self._env.num_obs_in_env + 1)  # :269-276, the robot is an actor too
    num_rigid_bodies_robot = property(lambda self: self._env.robot.num_bodies)
    global_tensor_dict = property(lambda self: self._env.global_tensor_dict)


class EnvManager:
    def __init__(self, sim_name, env_name, robot_name, controller_name, device, args=None, num_envs=None,
                 use_warp=None, headless=None):
        self.robot_name, self.controller_name = robot_name, controller_name
        self.sim_config = sim_config_registry.make_sim(sim_name)
        self.cfg = env_config_registry.make_env(env_name)
        self.device = torch.device(device)
        if num_envs is not None:
            self.cfg.env.num_envs = num_envs
        if use_warp is not None:
            self.cfg.env.use_warp = use_warp
        if headless is not None:
            self.sim_config.viewer.headless = headless
        self.num_envs = self.cfg.env.num_envs
        self.use_warp = self.cfg.env.use_warp
        self.env_args = dict(args or {})
        # "device": in-kernel Philox resets, no host round trip (default).
        # "torch" : uniforms drawn with torch.rand in the reference's call order, host sync per step
        #           exactly where the reference has one (env_manager.py:364-375).
        self.reset_rng = self.env_args.get("reset_rng", "device")
        if self.reset_rng not in ("device", "torch"):
            raise ValueError("args['reset_rng'] must be 'device' or 'torch'")
        self.global_tensor_dict = {}
        self.step_counter = 0
        self._graphs, self._dist_ctr_dev = {}, None
        self._populate()
        # "graph" (default where it applies): the per-sub-step launches of an env step with obstacles / disturbances replay as one CUDA
        # graph; "launch": one C-ABI call per kernel, as in round 1.  The torch-order RNG mode draws with torch inside the loop: launches.
        want = self.env_args.get("step_mode", "graph")
        if want not in ("graph", "launch"):
            raise ValueError("args['step_mode'] must be 'graph' or 'launch'")
        self.step_mode = want if (self.reset_rng == "device" and self.device.type == "cuda") else "launch"
        self.sim_steps = self.engine.sim_steps  # int32 [N] (env_manager.py:78-80)
        # the reference's object graph, as far as its examples / tasks / trainers walk it: env.robot_manager.robot.{cfg, controller,
        # controller_config}, env.IGE_env.num_assets_per_env (robot included, IGE_env_manager.py:269-276)
        self.robot_manager = _RobotManagerView(self)
        self.IGE_env = _SimView(self)

    # ------------------------------------------------------------------------------------------
    # construction (populate_env + prepare_sim, env_manager.py:127-271)
    # ------------------------------------------------------------------------------------------
    def _populate(self):
        N, dev = self.num_envs, self.device
        self.robot, self.robot_cfg = robot_registry.make_robot(self.robot_name, self.controller_name, self.cfg, dev)
        self.num_robot_actions = self.robot.num_actions
        spec = self.robot.make_spec(self.sim_config, self.cfg)
        self.spec = spec
        self.engine = Hp1Engine(
            spec, N, dev, physics_steps=1, seed=int(self.env_args.get("seed", 0)),
            env_id_offset=int(self.env_args.get("env_id_offset", 0)), device_rng_reset=(self.reset_rng == "device"),
            strict_stale_obs=bool(self.env_args.get("strict_stale_obs", True)), materialize_derived=True,
            per_env_params=self.env_args.get("per_env_params", "auto"), host_io=bool(self.env_args.get("host_io", False)),
            debug_wrench=bool(getattr(self.robot_cfg.sensor_config, "enable_imu", False)))  # the IMU reads the net body force
        eng, gtd = self.engine, self.global_tensor_dict
        self.robot.controller.bind_engine(eng)  # controller.K_*_tensor_current / set_controller_gains act on the engine's gains
        gtd["crashes"], gtd["truncations"] = eng.terminations, eng.truncations
        self.collision_tensor, self.truncation_tensor = eng.terminations, eng.truncations
        self.num_env_actions = self.cfg.env.num_env_actions
        gtd["num_env_actions"], gtd["env_actions"], gtd["prev_env_actions"] = self.num_env_actions, None, None
        # robot tensors (IGE_env_manager.py:301-311, base_multirotor.py:85-91, robot_manager.py:112-126)
        rs = eng.root_state
        gtd["robot_state_tensor"] = rs
        gtd["robot_position"], gtd["robot_orientation"] = rs[:, 0:3], rs[:, 3:7]
        gtd["robot_linvel"], gtd["robot_angvel"] = rs[:, 7:10], rs[:, 10:13]
        gtd["robot_euler_angles"], gtd["robot_vehicle_orientation"] = eng.euler, eng.vehicle_orientation
        gtd["robot_vehicle_linvel"], gtd["robot_body_linvel"], gtd["robot_body_angvel"] = (
            eng.vehicle_linvel, eng.body_linvel, eng.body_angvel)
        gtd["robot_actions"] = torch.zeros(N, self.num_robot_actions, device=dev)
        gtd["robot_prev_actions"] = torch.zeros_like(gtd["robot_actions"])
        gtd["num_robot_actions"] = self.num_robot_actions
        gtd["robot_mass"] = torch.full((N,), float(self.robot.robot_mass), device=dev)
        gtd["robot_inertia"] = torch.tensor(self.robot.robot_inertia, dtype=torch.float32, device=dev).expand(N, -1, -1)
        B = self.robot.num_bodies
        gtd["robot_force_tensor"] = torch.zeros(N, B, 3, device=dev)   # not materialised by the fused step
        gtd["robot_torque_tensor"] = torch.zeros(N, B, 3, device=dev)
        gtd["robot_contact_force_tensor"] = torch.zeros(N, 3, device=dev)
        gtd["env_bounds_min"], gtd["env_bounds_max"] = eng.bounds_min, eng.bounds_max
        gtd["gravity"] = torch.tensor(self.sim_config.sim.gravity, device=dev).expand(N, -1)
        gtd["dt"] = self.sim_config.sim.dt
        gtd["dof_control_mode"] = "none"
        self._bounds_rng = tuple(torch.tensor(v, dtype=torch.float32, device=dev) for v in (
            self.cfg.env.lower_bound_min, self.cfg.env.lower_bound_max, self.cfg.env.upper_bound_min, self.cfg.env.upper_bound_max))
        # IsaacGymEnv.__init__ draws the initial bounds (IGE_env_manager.py:59-64)
        eng.bounds_min.copy_(_lerp(self._bounds_rng[0], self._bounds_rng[1], torch.rand(N, 3, device=dev)))
        eng.bounds_max.copy_(_lerp(self._bounds_rng[2], self._bounds_rng[3], torch.rand(N, 3, device=dev)))
        self._build_obstacles()
        self._build_sensors()
        if eng.host_io and (self.scene is not None or self.reset_rng != "device"):
            raise NotImplementedError("args['host_io'] needs an obstacle-free env and reset_rng='device': the flags "
                                      "live in host memory and are written only by the fused task step")

    # ---- obstacles: AssetLoader.select_assets_for_sim + WarpEnv (asset_loader.py:148-194, warp_env_manager.py)
    def _select_assets(self):
        ec = self.cfg.env_config
        ordered, keep = deque(), 0
        for asset_type, params in ec.asset_type_to_dict_map.items():
            if asset_type in ec.include_asset_type and ec.include_asset_type[asset_type] is False:
                continue
            n = params.num_assets
            if n <= 0:
                continue
            if params.file is None:
                files = sorted(f for f in os.listdir(params.asset_folder) if f.endswith(".urdf"))
                chosen = random.choices(files, k=n)
            else:
                chosen = [params.file] * n
            for f in chosen:
                item = (params, os.path.join(params.asset_folder, f))
                if params.keep_in_env:
                    ordered.appendleft(item)
                    keep += 1
                else:
                    ordered.append(item)
        ordered = list(ordered)
        tail = ordered[keep:]
        random.shuffle(tail)
        ordered[keep:] = tail
        return ordered, keep

    def _build_obstacles(self):
        N, dev, gtd = self.num_envs, self.device, self.global_tensor_dict
        self.scene = None
        per_env, self.keep_in_env = [], 0
        if self.cfg.env_config.asset_type_to_dict_map:
            for _ in range(N):
                sel, keep = self._select_assets()
                per_env.append(sel)
                self.keep_in_env = keep
        A = len(per_env[0]) if per_env else 0
        if any(len(s) != A for s in per_env):
            raise ValueError("All environments should have the same number of assets")
        # dynamic obstacles (agx_obstacle_step): one pair of damping coefficients for the whole scene -- every shipped asset
        # class has 0.1 / 0.1 (config/asset_config/*_config.py: asset_state_params)
        p0 = per_env[0][0][0] if A else None
        self._asset_damping = (float(getattr(p0, "linear_damping", 0.0)), float(getattr(p0, "angular_damping", 0.0)))
        self.num_obs_in_env = A
        gtd["num_obstacles_in_env"] = A
        gtd["env_asset_state_tensor"] = torch.zeros(N, A, 13, device=dev)
        if A:
            gtd["env_asset_state_tensor"][..., 6] = 1.0
        gtd["asset_min_state_ratio"] = torch.zeros(N, A, 13, device=dev)
        gtd["asset_max_state_ratio"] = torch.zeros(N, A, 13, device=dev)
        if not A:
            return
        ast = gtd["env_asset_state_tensor"]
        gtd["obstacle_position"], gtd["obstacle_orientation"] = ast[..., 0:3], ast[..., 3:7]
        gtd["obstacle_linvel"], gtd["obstacle_angvel"] = ast[..., 7:10], ast[..., 10:13]
        lo = np.zeros((N, A, 13), np.float32)
        hi = np.zeros((N, A, 13), np.float32)
        # templates: every <box> / <cylinder> visual of every distinct URDF; an asset with b parts = b objects sharing its pose
        # (a box is one part, a tessellated cylinder eleven: urdf.visual_parts)
        tmpl_index, templates, tmpl_obbs, boxes_of, tmpl_link = {}, [], [], {}, []
        for e in range(N):
            for a, (params, path) in enumerate(per_env[e]):
                lo[e, a], hi[e, a] = params.min_state_ratio, params.max_state_ratio
                if path not in boxes_of:
                    model = urdf.parse_urdf(path)
                    ids = []
                    for _, link_index, tris, obb in urdf.visual_parts(model, params.use_collision_mesh_instead_of_visual):
                        tmpl_index[(path, len(ids))] = len(templates)
                        ids.append(len(templates))
                        templates.append(tris)
                        tmpl_obbs.append(obb)
                        tmpl_link.append(link_index)
                    boxes_of[path] = ids
        gtd["asset_min_state_ratio"].copy_(torch.from_numpy(lo))
        gtd["asset_max_state_ratio"].copy_(torch.from_numpy(hi))
        if not self.use_warp:
            return
        nb = [sum(len(boxes_of[p]) for _, p in per_env[e]) for e in range(N)]
        K = max(nb)
        if K == 0:
            return
        obj_t = np.zeros((N, K), np.int32)
        obj_c = np.zeros((N, K), np.int32)
        obj_asset = np.zeros((N, K), np.int64)
        seg_base, seg_mask = [0] * len(templates), [0] * len(templates)
        seg_ctr = 100  # env_manager.py:147; global across envs like the reference
        for e in range(N):
            k = 0
            for a, (params, path) in enumerate(per_env[e]):
                per_link = bool(getattr(params, "per_link_semantic", False))
                for t in boxes_of[path]:
                    obj_t[e, k], obj_asset[e, k] = t, a
                    if per_link:  # one id per link: link counter + the env's running counter (warp_asset.py:44-70, e.g. the trees)
                        seg_base[t], seg_mask[t], obj_c[e, k] = tmpl_link[t], 1, seg_ctr
                    elif params.semantic_id < 0:  # per-instance id = counter (warp_asset.py:100-104, warp_env_manager.py:76-80)
                        seg_base[t], seg_mask[t], obj_c[e, k] = 0, 1, seg_ctr
                    else:
                        seg_base[t], seg_mask[t], obj_c[e, k] = params.semantic_id, 0, 0
                    k += 1
                # the running counter advances by the number of distinct variable ids the asset used (warp_env_manager.py:90-95)
                seg_ctr += (max(tmpl_link[t] for t in boxes_of[path]) + 1) if (per_link and boxes_of[path]) else 1
            while k < K:  # pad ragged envs with a repeat of the last object (same pose, same id)
                obj_t[e, k], obj_c[e, k], obj_asset[e, k] = obj_t[e, k - 1], obj_c[e, k - 1], obj_asset[e, k - 1]
                k += 1
        # object poses: one row per object, gathered from the asset rows after every asset reset
        self._obj_asset = torch.from_numpy(obj_asset).to(dev)
        self._obj_pose = torch.zeros(N, K, 7, device=dev)
        self.scene = RayScene(templates, seg_base, seg_mask, obj_t, obj_c, self._obj_pose, dev,
                              bounds_min=self.engine.bounds_min, bounds_max=self.engine.bounds_max,
                              tmpl_obb=np.stack(tmpl_obbs))
        # Until the first reset_idx every actor sits where it was created: at the env origin, identity orientation
        # (IGE_env_manager.py: create_actor with the default start pose), the robot among them.  Tasks whose reset() does not reset the
        # sim (NavigationTask.reset, navigation_task.py:162-175) therefore begin in contact: the first step flags a crash and the
        # first real episode starts with the reset that follows.  Build the scene for that configuration.
        self._obj_pose[..., 6] = 1.0
        self.scene.update()

    def _build_sensors(self):
        gtd, N, dev = self.global_tensor_dict, self.num_envs, self.device
        sc = self.robot_cfg.sensor_config
        self.sensor, self.sensor_cfg = None, None
        self.imu = None
        if getattr(sc, "enable_imu", False):  # robot_manager.py:91-96, 246-262
            from ..sensors import IMUSensor
            # PhysX's force sensor reports the TOTAL force on the link, gravity included (base_imu_config.py:48), in the link
            # frame: here = applied base-frame force of the last physics step + m R(q)^T g  (refreshed in render_sensors)
            gtd["force_sensor_tensor"] = torch.zeros(N, 6, device=dev)
            self.imu = IMUSensor(sc.imu_config, N, dev)
            self.imu.init_tensors(gtd)
        if not self.use_warp:
            return
        if sc.enable_camera and sc.enable_lidar:
            raise ValueError("Both camera and lidar are enabled: they share depth_range_pixels (robot_manager.py:69-89).")
        cfg = sc.camera_config if sc.enable_camera else (sc.lidar_config if sc.enable_lidar else None)
        if cfg is None:
            return
        if self.scene is None:
            return  # "Warp camera is enabled but there is nothing in the environment" (robot_manager.py:189-192)
        S, H, W = cfg.num_sensors, cfg.height, cfg.width
        pc = bool(cfg.return_pointcloud)
        gtd["depth_range_pixels"] = torch.zeros((N, S, H, W, 3) if pc else (N, S, H, W), device=dev)
        seg = None
        if cfg.segmentation_camera:
            seg = gtd["segmentation_pixels"] = torch.zeros(N, S, H, W, dtype=torch.int32, device=dev)
        self.sensor_cfg = cfg
        T = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
        self._mount_rng = (T(cfg.min_translation), T(cfg.max_translation),
                           torch.deg2rad(T(cfg.min_euler_rotation_deg)), torch.deg2rad(T(cfg.max_euler_rotation_deg)))
        self.sensor_mount = torch.zeros(N, S, 7, device=dev)
        mean_e = (self._mount_rng[2] + self._mount_rng[3]) / 2.0  # warp_sensor.py:118-123
        self.sensor_mount[..., 3:7] = _quat_from_euler(mean_e.expand(N, S, 3))
        self.sensor = RaySensor(cfg, self.scene, self.engine.root_state, gtd["depth_range_pixels"], seg, self.sensor_mount)
        self.sensor.mount = self.sensor_mount  # keep the live tensor (RaySensor made it contiguous already)
        self.sensor.c.mount = self.sensor_mount.data_ptr()
        # sensor noise (WarpSensor.apply_noise, warp_sensor.py:227-250): "device" = one in-place pass with a counter-based
        # device RNG (agx_hp2_noise_limits; default next to the in-kernel Philox resets), "torch" = the reference's torch calls
        # in the reference's order (default when reset_rng == "torch", i.e. when every random number comes from torch)
        self.sensor_noise_rng = self.env_args.get("sensor_noise_rng", "device" if self.reset_rng == "device" else "torch")
        if self.sensor_noise_rng not in ("device", "torch"):
            raise ValueError("args['sensor_noise_rng'] must be 'device' or 'torch'")
        self._device_noise = None
        if self.sensor.noise_enabled and self.sensor_noise_rng == "device":
            from ..sensors.noise import DeviceSensorNoise
            px = gtd["depth_range_pixels"]
            per_env = px[0].numel() // (3 if pc else 1)
            self._device_noise = DeviceSensorNoise(cfg, px, seed=int(self.env_args.get("seed", 0)) ^ 0x5E4503_4E015E,
                                                   first_pixel=int(self.env_args.get("env_id_offset", 0)) * per_env)

    # ------------------------------------------------------------------------------------------
    # reset (env_manager.py:273-301)
    # ------------------------------------------------------------------------------------------
    def _reset_assets(self, env_ids, num_obstacles):
        """AssetManager.reset_idx (asset_manager.py:51-71): full-size draw, gather, park the rest."""
        gtd = self.global_tensor_dict
        ast = gtd["env_asset_state_tensor"]
        n_keep = max(num_obstacles, self._num_keep)
        lo, hi = gtd["asset_min_state_ratio"], gtd["asset_max_state_ratio"]
        r = _lerp(lo, hi, torch.rand_like(hi))
        bmin = self.engine.bounds_min.unsqueeze(1)
        bmax = self.engine.bounds_max.unsqueeze(1)
        pos = bmin + (bmax - bmin) * r[..., 0:3]
        ast[env_ids, :, 0:3] = pos[env_ids]
        ast[env_ids, :, 3:7] = _quat_from_euler(r[env_ids][..., 3:6])
        ast[env_ids, n_keep:, 0:3] = -1000.0

    def reset_idx(self, env_ids=None):
        eng, gtd, dev, N = self.engine, self.global_tensor_dict, self.device, self.num_envs
        if env_ids is None:
            env_ids = torch.arange(N, device=dev)
        env_ids = env_ids.to(dev)
        mask = torch.zeros(N, dtype=torch.bool, device=dev)
        mask[env_ids] = True
        A = gtd["num_obstacles_in_env"]
        # robots whose reset leaves the motor model alone (BaseROV, robots/base_rov.py:188-201): put it back after the engine's reset
        motor_names = ("motor_thrust", "tau_inc", "tau_dec", "k_thrust")
        saved_motors = None
        if getattr(self.robot, "keeps_motor_state_on_reset", False) and getattr(self, "_motors_initialised", False):
            saved_motors = {k: getattr(eng, k).clone() for k in motor_names if getattr(eng, k, None) is not None}
        if self.reset_rng == "device":
            # Philox reset of bounds + robot + motor/controller params in one launch; the asset
            # sampling below only needs the new bounds (reference order: bounds -> assets -> robot,
            # and the robot sampling does not depend on the assets)
            eng.reset(mask, None)
            if A > 0:
                self._sample_assets_and_update_scene(env_ids, mask, A)
        else:
            M = self.spec.num_motors
            r = lambda *s: torch.rand(*s, device=dev)
            # IsaacGymEnv.reset_idx: two full-size draws (IGE_env_manager.py:513-519)
            draws = {"bounds_lo": r(N, 3), "bounds_hi": r(N, 3)}
            if A > 0:
                lo0, lo1, hi0, hi1 = self._bounds_rng
                eng.bounds_min[env_ids] = _lerp(lo0, lo1, draws["bounds_lo"])[env_ids]
                eng.bounds_max[env_ids] = _lerp(hi0, hi1, draws["bounds_hi"])[env_ids]
                self._sample_assets_and_update_scene(env_ids, mask, A)
            draws["state"] = r(N, 13)  # base_multirotor.py:182
            if self.spec.randomize_params:  # base_lee_controller.py:105-118 draws [k,3] per gain
                k = len(env_ids)
                for name in ("K_pos", "K_vel", "K_rot", "K_angvel"):
                    t = torch.zeros(N, 3, device=dev)
                    t[env_ids] = r(k, 3)
                    draws[name] = t
            draws["tau_inc"], draws["tau_dec"], draws["thrust"] = r(N, M), r(N, M), r(N, M)  # motor_model.py:140-150
            if self.spec.use_rps:
                draws["k_thrust"] = r(N, M)
            eng.reset(mask, draws)
        if saved_motors is not None:
            for k, t in saved_motors.items():
                getattr(eng, k)[env_ids] = t[env_ids]
        self._motors_initialised = True
        if self.sensor is not None and self.sensor_cfg.randomize_placement:  # warp_sensor.py:153-172
            k, S = len(env_ids), self.sensor_cfg.num_sensors
            t0, t1, e0, e1 = self._mount_rng
            self.sensor_mount[env_ids, :, 0:3] = _lerp(t0, t1, torch.rand(k, S, 3, device=dev))
            self.sensor_mount[env_ids, :, 3:7] = _quat_from_euler(_lerp(e0, e1, torch.rand(k, S, 3, device=dev)))
        if self.imu is not None:
            self.imu.reset_idx(env_ids)
        eng.refresh()  # update_states of ALL envs (base_multirotor.py:204-205); sim_steps zeroed by the kernel

    def _sample_assets_and_update_scene(self, env_ids, mask, A):
        gtd, dev = self.global_tensor_dict, self.device
        self._num_keep = self.keep_in_env
        self._reset_assets(env_ids, A)
        self._num_keep = self.keep_in_env // 2  # env_manager.py:285-295 (second, thinner sampling)
        samples = torch.bernoulli(0.15 * torch.ones(len(env_ids), device=dev))
        sel = torch.nonzero(samples).squeeze(-1)
        if len(sel) > 0:
            self._reset_assets(env_ids[sel], A // 2)
        self._num_keep = self.keep_in_env
        if self.scene is not None:  # WarpEnv.reset_idx (warp_env_manager.py:40-54)
            idx = self._obj_asset.unsqueeze(-1).expand(-1, -1, 7)
            self._obj_pose.copy_(torch.gather(gtd["env_asset_state_tensor"][..., 0:7], 1, idx))
            self.scene.update(mask)

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))

    # ------------------------------------------------------------------------------------------
    # stepping (env_manager.py:328-432)
    # ------------------------------------------------------------------------------------------
    def reset_tensors(self):
        self.collision_tensor[:] = 0
        self.truncation_tensor[:] = 0

    def _draw_disturbance(self):
        """apply_disturbance (base_multirotor.py:213-234).  reset_rng = "torch": bernoulli, rand, rand in the reference's order.
        reset_rng = "device" (default): one launch of agx_disturbance_draw (Philox per env, keyed by the global env id and a draw
        counter) into a persistent [N,6] buffer -- same distributions, no torch RNG launches inside the physics loop."""
        sp, N, dev = self.spec, self.num_envs, self.device
        if not sp.enable_disturbance:
            return None
        if self.reset_rng == "device":
            if getattr(self, "_dist_buf", None) is None:
                self._dist_buf = torch.zeros(N, 6, device=dev)
                self._dist_max = (C.c_float * 6)(*[float(v) for v in sp.max_disturbance])
                self._dist_counter = 0
                self._dist_seed = (int(self.env_args.get("seed", 0)) ^ 0xD157_0000_D157) & (2**64 - 1)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().agx_disturbance_draw(N, int(self.env_args.get("env_id_offset", 0)), float(sp.prob_apply_disturbance),
                                                        self._dist_max, self._dist_seed, self._dist_counter & 0xFFFFFFFF,
                                                        C.c_void_p(self._dist_buf.data_ptr()), stream), "agx_disturbance_draw")
            self._dist_counter += 1
            if self._dist_ctr_dev is not None:  # the graph path's device-resident copy of the counter follows
                _lib.check(_lib.load().agx_counter_add(C.c_void_p(self._dist_ctr_dev.data_ptr()), 1, stream), "agx_counter_add")
            return self._dist_buf
        occ = torch.bernoulli(sp.prob_apply_disturbance * torch.ones(N, device=dev))
        mx = torch.tensor(sp.max_disturbance, dtype=torch.float32, device=dev).expand(N, -1)
        f = _lerp(-mx[:, 0:3], mx[:, 0:3], torch.rand(N, 3, device=dev)) * occ.unsqueeze(1)
        t = _lerp(-mx[:, 3:6], mx[:, 3:6], torch.rand(N, 3, device=dev)) * occ.unsqueeze(1)
        return torch.cat([f, t], dim=1).contiguous()

    def _graph_for(self, n):
        """CUDA graph of the n physics sub-steps of one env step (captured once per n): per sub-step ONE physics launch that draws its
        disturbance in the kernel (AgxHp1Buffers.dist_counter: the draw counter lives in device memory, so a replay draws afresh) and
        ONE collision launch, then one single-thread kernel that advances the draw counter.  Same kernels, same arithmetic, same
        draws as the launch-by-launch loop (tests: bit-identical trajectories); every C-ABI call is capture-safe (no allocation, no
        synchronisation).  Navigation-type envs run 10 sub-steps per env step (config/env_config/env_with_obstacles.py:29-30):
        ~30 launches become one."""
        g = self._graphs.get(n)
        if g is not None:
            return g
        eng, a = self.engine, self.global_tensor_dict["robot_actions"]
        dist = self.spec.enable_disturbance
        if dist and self._dist_ctr_dev is None:
            self._draw_disturbance()  # creates the stream state (seed, counters); the draw itself is discarded ...
            self._dist_counter -= 1   # ... and not counted
            self._dist_ctr_dev = torch.tensor([self._dist_counter & 0xFFFFFFFF], dtype=torch.int32, device=self.device)
        lib = _lib.load()

        def body():
            for i in range(n):
                eng.physics_step(a, physics_steps=1, dist_counter=self._dist_ctr_dev if dist else None, dist_offset=i)
                self.compute_observations()
            if dist and n:
                _lib.check(lib.agx_counter_add(C.c_void_p(self._dist_ctr_dev.data_ptr()), n,
                                               C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "agx_counter_add")

        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up outside the capture (first launches may load modules), on throw-away copies of the state
            keep = (eng.root_state.clone(), eng.motor_thrust.clone(), self.collision_tensor.clone(),
                    None if not dist else self._dist_ctr_dev.clone())
            body()
            eng.root_state.copy_(keep[0]); eng.motor_thrust.copy_(keep[1]); self.collision_tensor.copy_(keep[2])
            if dist:
                self._dist_ctr_dev.copy_(keep[3])
        cur.wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            body()
        # capturing does not execute: the state is untouched, the caller replays
        self._graphs[n] = g
        return g

    def sample_physics_steps(self):
        e = self.cfg.env
        return max(math.floor(random.gauss(e.num_physics_steps_per_env_step_mean, e.num_physics_steps_per_env_step_std)), 0)

    def step(self, actions, env_actions=None):
        gtd = self.global_tensor_dict
        self.reset_tensors()
        if env_actions is not None:
            if self.num_obs_in_env > 1:
                want = (self.num_envs, self.num_obs_in_env, 6)
                if tuple(env_actions.shape) != want or env_actions.dtype != torch.float32 or env_actions.device.type != self.device.type:
                    raise ValueError(f"env_actions must be a float32 {list(want)} tensor on {self.device} (linear + angular velocity per obstacle)")
            if gtd["env_actions"] is None:
                gtd["env_actions"], gtd["prev_env_actions"] = env_actions, env_actions.clone()
            gtd["prev_env_actions"][:] = gtd["env_actions"]
            gtd["env_actions"][:] = env_actions
        n = self.sample_physics_steps()
        gtd["robot_prev_actions"][:] = gtd["robot_actions"]  # robot_manager.py:486-488
        gtd["robot_actions"][:] = actions
        a = gtd["robot_actions"]
        if self.spec.enable_disturbance or self.scene is not None:
            # the reference loop (env_manager.py:426-428): fresh disturbance draws and a collision check after EVERY physics step
            if self.step_mode == "graph":
                self._graph_for(n).replay()  # the n x (physics launch with in-kernel draw, collision launch) of this env step: one graph launch
                if self.spec.enable_disturbance:
                    self._dist_counter += n
            else:
                for i in range(n):
                    self.engine.physics_step(a, disturbance=self._draw_disturbance(), physics_steps=1)
                    self.compute_observations()
        elif n > 0:
            self.engine.physics_step(a, physics_steps=n)  # n sub-steps fused in one launch
        if env_actions is not None and self.num_obs_in_env > 1 and n > 0:  # ObstacleManager.pre_physics_step, obstacle_manager.py:40-44
            self._step_obstacles(gtd["env_actions"], n)
        self.engine.sim_steps += 1
        self.step_counter += 1

    def simulate(self, actions, env_actions=None):
        """ONE physics step (env_manager.py:346-349: pre_physics_step -> physics -> post_physics_step), for callers that drive the loop
        themselves; step() batches the n physics steps of an env step into as few launches as the env allows"""
        gtd = self.global_tensor_dict
        gtd["robot_prev_actions"][:] = gtd["robot_actions"]
        gtd["robot_actions"][:] = actions
        self.engine.physics_step(gtd["robot_actions"], disturbance=self._draw_disturbance(), physics_steps=1)
        if env_actions is not None and self.num_obs_in_env > 1:
            self._step_obstacles(env_actions, 1)

    def render_viewer(self):
        return None  # headless by construction: there is no Isaac Gym viewer (env_manager.py:389-391)

    def log_memory_use(self):
        if self.device.type == "cuda":  # env_manager.py:303-326
            gb = 1024.0 ** 3
            print(f"torch.cuda.memory_allocated: {torch.cuda.memory_allocated(self.device) / gb:.3f} GB, "
                  f"reserved: {torch.cuda.memory_reserved(self.device) / gb:.3f} GB, max reserved: {torch.cuda.max_memory_reserved(self.device) / gb:.3f} GB")

    def _step_obstacles(self, twist, n):
        """dynamic_env: the obstacles' twist is overwritten with env_actions [N,A,6] before each of the n physics steps and
        the obstacles move (PhysX in the reference; agx_obstacle_step's kinematic advance here, one launch for the n steps).
        The reference leaves its Warp meshes stale until the next reset ("refit() ... expensive", env_manager.py:340-342);
        here re-posing + rebuilding the BVHs of all envs is one more launch, done by default (args['refit_dynamic_obstacles'])
        so the ray-cast sensors and the collision test of the next env step see the obstacles where they are."""
        gtd, N, A = self.global_tensor_dict, self.num_envs, self.num_obs_in_env
        ast = gtd["env_asset_state_tensor"]
        twist = twist.contiguous()  # shape / dtype / device were checked in step()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.load().agx_obstacle_step(N, A, C.c_void_p(ast.data_ptr()), ast.stride(1), C.c_void_p(twist.data_ptr()),
                                                 float(self.sim_config.sim.dt), int(n), self._asset_damping[0], self._asset_damping[1],
                                                 stream), "agx_obstacle_step")
        if self.scene is not None and bool(self.env_args.get("refit_dynamic_obstacles", True)):
            idx = self._obj_asset.unsqueeze(-1).expand(-1, -1, 7)
            self._obj_pose.copy_(torch.gather(ast[..., 0:7], 1, idx))
            self.scene.update()

    def compute_observations(self):
        """crashes += contact (env_manager.py:358-362).  PhysX contact forces are replaced by a
        geometric test: the base-link collision sphere overlaps the env mesh (agx_hp2_collide)."""
        if self.scene is not None:
            self.scene.collide(self.engine.root_state, self.robot.collision_radius, self.collision_tensor)

    def reset_terminated_and_truncated_envs(self):
        flags = self.collision_tensor * int(self.cfg.env.reset_on_collision) + self.truncation_tensor
        envs_to_reset = flags.nonzero(as_tuple=False).squeeze(-1)
        if len(envs_to_reset) > 0:
            self.reset_idx(envs_to_reset)
        return envs_to_reset

    def render(self, render_components="sensors"):
        if render_components == "sensors":
            self.render_sensors()

    def render_sensors(self):
        if self.imu is not None:
            gtd = self.global_tensor_dict
            fs, w = gtd["force_sensor_tensor"], self.engine.body_wrench
            fs[:, 0:3] = w[:, 0:3] + gtd["robot_mass"].unsqueeze(1) * _quat_rotate_inverse(gtd["robot_orientation"], gtd["gravity"])
            fs[:, 3:6] = w[:, 3:6]
            self.imu.update()
        if self.sensor is None:
            return
        self.sensor.capture()
        if self.sensor.noise_enabled:  # WarpSensor.apply_noise + apply_range_limits + normalize_observation (warp_sensor.py:197-250)
            if self._device_noise is not None:
                self._device_noise.apply()
            else:
                from ..sensors.noise import apply_noise_and_limits_torch
                apply_noise_and_limits_torch(self.global_tensor_dict["depth_range_pixels"], self.sensor_cfg)

    def post_reward_calculation_step(self):
        envs_to_reset = self.reset_terminated_and_truncated_envs()
        self.render(render_components="sensors")
        return envs_to_reset

    def get_obs(self):
        return self.global_tensor_dict

    def delete_env(self):
        if self.engine is not None:
            self.engine.close()
        self.engine = None
        self.scene = None
        self.sensor = None
