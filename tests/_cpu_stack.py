"""CPU twin of the whole stack, for testing the HOST logic (EnvManager, tasks, sensors wiring) without a GPU.

``with cpu_stack():`` swaps, for the duration of the block,
  * ``Hp1Engine``           -> the host shadow of the HP1 device code (tests/_shadow_hp1.py: the kernels' own arithmetic),
  * ``RayScene / RaySensor``-> the brute-force ray-cast ORACLE (oracle/hp2_oracle.c) behind the same attributes and methods,
  * ``_lib.load()``         -> a proxy whose agx_* entry points for the small per-env kernels call their shadow_* twins
                               (same argument lists minus the stream),
  * ``DeviceSensorNoise``   -> the shadow of noise_limits_kernel,
  * ``IMUSensor``           -> the product class with its CUDA-device check skipped (agx_imu_update goes through the proxy),
  * ``torch.cuda.current_stream`` -> a dummy (the tasks read ``.cuda_stream`` to pass it along).
Everything above the C ABI -- registries, config plumbing, RNG call order, reset / curriculum / render bookkeeping, tensor
aliasing in the Global Tensor Dict -- is the PRODUCT's code, running on "cpu" tensors.  TEST INFRASTRUCTURE ONLY: the product
has no CPU path (Hp1Engine / RayScene refuse non-CUDA devices), and nothing here is importable from the package."""
import contextlib
import ctypes as C
import types

import numpy as np
import torch

from aerial_gym_simulator_b200 import _lib
from oracle import hp2_oracle as RO

from . import _hp2_common as H2
from . import _shadow
from ._shadow_hp1 import ShadowHp1Engine


class CpuHp1Engine(ShadowHp1Engine):
    DERIVED = ("euler", "vehicle_orientation", "vehicle_linvel", "body_linvel", "body_angvel")

    def __init__(self, spec, num_envs, device="cpu", *, physics_steps=1, episode_len_steps=500, seed=0, env_id_offset=0,
                 device_rng_reset=True, strict_stale_obs=True, materialize_derived=True, per_env_params="auto", debug_wrench=False,
                 host_io=False):
        if host_io:
            raise NotImplementedError("host_io is a CUDA feature")
        super().__init__(spec, num_envs, physics_steps=physics_steps, episode_len_steps=episode_len_steps, seed=seed,
                         env_id_offset=env_id_offset, device_rng_reset=device_rng_reset, strict_stale_obs=strict_stale_obs,
                         materialize_derived=materialize_derived, per_env_params=per_env_params, debug_wrench=debug_wrench,
                         coop_reset=True)
        self.device, self.host_io = torch.device("cpu"), False

    def position_task_step(self, actions, disturbance=None, physics_steps=None, mid_event=None):
        super().position_task_step(actions.contiguous(), disturbance, physics_steps)

    def physics_step(self, actions, disturbance=None, physics_steps=None):
        super().physics_step(actions.contiguous(), disturbance, physics_steps)

    def reset(self, mask, draws=None):
        if draws is not None:
            draws = {k: (v.contiguous() if v is not None else None) for k, v in draws.items()}
        super().reset(mask, draws)

    def refresh(self, only_if_flag=False):
        super().refresh()

    def close(self):
        pass


class CpuRayScene:
    def __init__(self, templates, tmpl_seg_base, tmpl_seg_mask, obj_template, obj_seg_counter, obj_pose, device="cpu", tris_per_object=None,
                 bounds_min=None, bounds_max=None, tmpl_obb=None):
        obj_template = np.asarray(obj_template, np.int32)
        self.E, self.K = obj_template.shape
        self.L = int(tris_per_object or max(len(t) for t in templates))
        self.device = torch.device("cpu")
        offs = np.zeros(len(templates) + 1, np.int32)
        offs[1:] = np.cumsum([len(t) for t in templates])
        self.tmpl_tri_offset = torch.tensor(offs)
        self.tmpl_tris = torch.tensor(np.concatenate([np.asarray(t, np.float32).reshape(-1, 9) for t in templates], 0))
        self.tmpl_seg_base = torch.tensor(np.concatenate([np.broadcast_to(np.asarray(b, np.int32), (len(t),)) for b, t in zip(tmpl_seg_base, templates)]))
        self.tmpl_seg_mask = torch.tensor(np.concatenate([np.broadcast_to(np.asarray(m, np.int32), (len(t),)) for m, t in zip(tmpl_seg_mask, templates)]))
        self.obj_template, self.obj_seg_counter = torch.tensor(obj_template), torch.tensor(np.asarray(obj_seg_counter, np.int32))
        self.obj_pose = obj_pose
        self._tris = np.zeros((self.E, self.K * self.L, 9), np.float32)
        self._segs = np.zeros((self.E, self.K * self.L), np.int32)
        self._cnt = np.zeros(self.E, np.int32)
        self.updates = 0

    def update(self, mask=None):
        if mask is not None and (mask.dtype != torch.bool or mask.shape != (self.E,)):
            raise ValueError("mask must be bool [E]")
        t, s, c = RO.build_world_tris(self.obj_pose[..., :7].numpy(), self.obj_template.numpy(), self.obj_seg_counter.numpy(),
                                      self.tmpl_tri_offset.numpy(), self.tmpl_tris.numpy(), self.tmpl_seg_base.numpy(),
                                      self.tmpl_seg_mask.numpy(), self.K * self.L)
        m = np.ones(self.E, bool) if mask is None else mask.numpy()
        self._tris[m], self._segs[m], self._cnt[m] = t[m], s[m], c[m]
        self.updates += 1

    def collide(self, robot_state, radius, crashes, min_dist=None):
        if crashes.dtype != torch.bool or crashes.shape != (self.E,):
            raise ValueError("crashes must be bool [E]")
        flags, _ = RO.collide(robot_state[:, :7].numpy(), float(radius), self._tris, self._cnt)
        crashes |= torch.from_numpy(flags)


class CpuRaySensor:
    def __init__(self, cfg, scene, robot_state, pixels, seg_pixels=None, mount=None):
        E, S, H, W = scene.E, cfg.num_sensors, cfg.height, cfg.width
        pc = bool(getattr(cfg, "return_pointcloud", False)) or cfg.sensor_type.startswith("normal_faceID")
        want = (E, S, H, W, 3) if pc else (E, S, H, W)
        if tuple(pixels.shape) != want or pixels.dtype != torch.float32 or not pixels.is_contiguous():
            raise ValueError(f"pixels must be contiguous float32 {want}")
        if seg_pixels is not None and (tuple(seg_pixels.shape) != (E, S, H, W) or seg_pixels.dtype != torch.int32):
            raise ValueError("seg_pixels must be int32 [E,S,H,W]")
        self.cfg, self.scene, self.robot_state, self.pixels, self.seg_pixels = cfg, scene, robot_state, pixels, seg_pixels
        if mount is None:
            mount = torch.zeros(E, S, 7)
            mount[..., 6] = 1.0
        self.mount = mount.contiguous()
        noise = getattr(cfg, "sensor_noise", None)
        self.noise_enabled = bool(noise is not None and getattr(noise, "enable_sensor_noise", False))
        self._so, table = H2.oracle_sensor(cfg, fuse=not self.noise_enabled)
        self._so.segmentation = int(seg_pixels is not None)
        self.ray_table = torch.tensor(table) if table is not None else None
        self.c = types.SimpleNamespace(mount=None, fuse_epilogue=int(not self.noise_enabled))
        self.captures = 0

    def capture(self):
        pix, seg = RO.cast(self._so, self.robot_state[:, :7].numpy(), self.mount.numpy(),
                           self.ray_table.numpy() if self.ray_table is not None else None, self.scene._tris, self.scene._segs, self.scene._cnt)
        self.pixels.copy_(torch.from_numpy(pix))
        if self.seg_pixels is not None:
            self.seg_pixels.copy_(torch.from_numpy(seg))
        self.captures += 1
        return self.pixels

    def rays_per_frame(self):
        return self.scene.E * self.cfg.num_sensors * self.cfg.height * self.cfg.width


class CpuDeviceSensorNoise:
    def __init__(self, cfg, pixels, seed, first_pixel=0):
        from aerial_gym_simulator_b200.sensors.noise import noise_struct
        self.pixels, self.seed, self.frame, self.first_pixel = pixels, int(seed) & (2**64 - 1), 0, int(first_pixel)
        self.c = noise_struct(cfg)
        self.num_pixels = pixels.numel() // self.c.components

    def apply(self):
        _shadow.load().shadow_noise_limits(C.c_void_p(self.pixels.data_ptr()), self.num_pixels, self.first_pixel,
                                           C.cast(C.byref(self.c), C.c_void_p), self.seed, self.frame & 0xFFFFFFFF)
        self.frame += 1
        return self.pixels


class CpuIMUSensor:
    """IMUSensor minus its CUDA-device check (the rest -- init_tensors / update / reset -- is the product class's own code)"""

    def __new__(cls, sensor_config, num_envs, device):
        from aerial_gym_simulator_b200.sensors.imu_sensor import IMUSensor
        self = object.__new__(IMUSensor)
        self.cfg, self.num_envs, self.device = sensor_config, int(num_envs), torch.device("cpu")
        self.lib = _lib.load()
        self.world_frame, self.gravity_compensation = sensor_config.world_frame, sensor_config.gravity_compensation
        return self


class _LibProxy:
    """agx_* of the small per-env kernels -> the real entry point for its argument validation, then shadow_* (same arguments, no
    stream) for the arithmetic; everything else -> the real library"""
    _MAP = {"agx_nav_reward": "shadow_nav_reward", "agx_nav_obs": "shadow_nav_obs", "agx_imu_update": "shadow_imu_update",
            "agx_lidar_nav_pool": "shadow_lidar_nav_pool", "agx_lidar_nav_reward": "shadow_lidar_nav_reward",
            "agx_lidar_nav_obs": "shadow_lidar_nav_obs", "agx_obstacle_step": "shadow_obstacle_step",
            "agx_e2e_reward": "shadow_e2e_reward", "agx_e2e_obs": "shadow_e2e_obs",
            "agx_s2r_reward": "shadow_s2r_reward", "agx_s2r_obs": "shadow_s2r_obs", "agx_disturbance_draw": "shadow_disturbance_draw",
            "agx_hp2_noise_limits": "shadow_noise_limits"}

    def __init__(self, real):
        self._real, self._sh = real, _shadow.load()
        self.calls = {}

    def __getattr__(self, name):
        if name not in self._MAP:
            return getattr(self._real, name)
        fn = getattr(self._sh, self._MAP[name])

        real_fn = getattr(self._real, name)

        def call(*args):
            self.calls[name] = self.calls.get(name, 0) + 1
            # the PRODUCT's host-side argument validation runs first: on a box without a GPU the real entry point either rejects
            # the arguments (AGX_E_INVALID / AGX_E_NULL, before any launch) or fails at the launch itself (AGX_E_CUDA) -- only
            # then does the shadow compute
            rc = real_fn(*args)
            if rc in (-1, -3):
                return rc
            args = list(args[:-1])  # drop the stream
            if name == "agx_lidar_nav_pool":
                args.append(0)  # force_scalar = 0
            args = [C.cast(a, C.c_void_p) if hasattr(a, "_obj") else a for a in args]  # byref(struct) -> void*
            rc = fn(*args)
            return 0 if (rc is None or name == "agx_lidar_nav_pool") else rc
        return call


@contextlib.contextmanager
def cpu_stack():
    import aerial_gym_simulator_b200.env_manager.env_manager as EM
    import aerial_gym_simulator_b200.sensors as S
    import aerial_gym_simulator_b200.sensors.noise as SN

    real_lib = _lib.load()
    proxy = _LibProxy(real_lib)
    saved = (EM.Hp1Engine, EM.RayScene, EM.RaySensor, _lib.load, SN.DeviceSensorNoise, torch.cuda.current_stream)
    saved_imu = S.IMUSensor
    S.IMUSensor = CpuIMUSensor
    EM.Hp1Engine, EM.RayScene, EM.RaySensor = CpuHp1Engine, CpuRayScene, CpuRaySensor
    _lib.load = lambda: proxy
    SN.DeviceSensorNoise = CpuDeviceSensorNoise
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0)
    try:
        yield proxy
    finally:
        EM.Hp1Engine, EM.RayScene, EM.RaySensor, _lib.load, SN.DeviceSensorNoise, torch.cuda.current_stream = saved
        S.IMUSensor = saved_imu
