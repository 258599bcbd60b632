"""Host side of HP2 (ray-cast sensors): device tensors + C-ABI calls, no hot-path arithmetic.

``RayScene``  replaces WarpEnv (env_manager/warp_env_manager.py): per-env triangle scene + BVH,
              rebuilt on reset from the asset root poses.
``RaySensor`` replaces WarpSensor/WarpCam/WarpLidar (sensors/warp/warp_sensor.py, warp_cam.py,
              warp_lidar.py): one camera or LiDAR type mounted on every robot.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import AgxHp2Scene, AgxHp2Sensor


def box_triangles(size: Sequence[float]) -> np.ndarray:
    """[12,9] object-frame triangles (v0,v1,v2) of a URDF <box size="x y z"/> centred at the origin:
    the mesh the reference obtains through urdfpy/trimesh for every shipped obstacle
    (assets/warp_asset.py:23-60; resources/models/environment_assets/{objects,panels,walls})."""
    hx, hy, hz = (float(s) / 2.0 for s in size)
    v = np.array([[-hx, -hy, -hz], [-hx, -hy, hz], [-hx, hy, -hz], [-hx, hy, hz],
                  [hx, -hy, -hz], [hx, -hy, hz], [hx, hy, -hz], [hx, hy, hz]], dtype=np.float32)
    f = np.array([[1, 3, 0], [4, 1, 0], [0, 3, 2], [2, 4, 0], [1, 7, 3], [5, 1, 4],
                  [5, 7, 1], [3, 7, 2], [6, 4, 2], [2, 7, 6], [6, 5, 4], [7, 5, 6]])
    return v[f].reshape(12, 9).astype(np.float32)


def box_obb(size: Sequence[float], R=None, p=None) -> np.ndarray:
    """[16] oriented-box record of a box template: centre, 3 axes (columns of R), half extents, kind.
    kind = 2: the template's triangles are box_triangles(size) in that order, transformed by (R, p) --
    the ray-caster may then test only the two triangles of the face a ray enters through; kind = 1: some
    other mesh with this bounding box (culling only); kind = 0: no box."""
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    p = np.zeros(3) if p is None else np.asarray(p, dtype=np.float64)
    return np.concatenate([p, R[:, 0], R[:, 1], R[:, 2], np.asarray(size, dtype=np.float64) / 2.0, [2.0]]).astype(np.float32)


def cylinder_triangles(radius: float, length: float, sections: int = 32) -> np.ndarray:
    """[4*sections,9] object-frame triangles of a URDF <cylinder radius length/> (axis = z, centred): a `sections`-gon prism with fan
    caps -- what the reference obtains through urdfpy -> trimesh.creation.cylinder (a revolved rectangle, 32 sections by default;
    assets/warp_asset.py:19-24).  trimesh is not installable here: vertex angles 2 pi k / sections are the published construction,
    parity against trimesh's own tessellation unpinned."""
    r, h = float(radius), float(length) / 2.0
    ang = 2.0 * np.pi * np.arange(sections) / sections
    ring = np.stack([r * np.cos(ang), r * np.sin(ang)], axis=1)
    tris = []
    for k in range(sections):
        a, b = ring[k], ring[(k + 1) % sections]
        lo_a, lo_b, hi_a, hi_b = (a[0], a[1], -h), (b[0], b[1], -h), (a[0], a[1], h), (b[0], b[1], h)
        tris.append(((0.0, 0.0, -h), lo_b, lo_a))   # bottom cap (outward normal -z)
        tris.append((lo_a, lo_b, hi_b))             # side
        tris.append((lo_a, hi_b, hi_a))
        tris.append(((0.0, 0.0, h), hi_a, hi_b))    # top cap (+z)
    return np.asarray(tris, dtype=np.float32).reshape(-1, 9)


def _next_pow2(v: int) -> int:
    p = 1
    while p < v:
        p <<= 1
    return p


class RayScene:
    """templates: list of [n_i,9] float32 triangle arrays (n_i <= tris_per_object; split bigger
    meshes into several templates sharing a pose).  obj_template / obj_seg_counter: int [E,K].
    obj_pose: float32 device tensor [E,K,>=7] (e.g. env_asset_state_tensor, a view is fine as long
    as the last dim is contiguous and rows are uniformly strided)."""

    def __init__(self, templates, tmpl_seg_base, tmpl_seg_mask, obj_template, obj_seg_counter, obj_pose: torch.Tensor,
                 device="cuda:0", tris_per_object: Optional[int] = None, bounds_min=None, bounds_max=None, tmpl_obb=None):
        """tmpl_obb: optional [T,16] oriented-box records (box_obb) of the templates that ARE boxes (valid=1);
        purely a culling aid -- results are identical with or without it."""
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.AgxError("RayScene needs a CUDA device: there is no CPU path")
        dev = self.device
        obj_template = np.asarray(obj_template, dtype=np.int32)
        E, K = obj_template.shape
        L = int(tris_per_object or max(len(t) for t in templates))
        if any(len(t) > L for t in templates):
            raise ValueError("template larger than tris_per_object: split it on the host")
        offs = np.zeros(len(templates) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(t) for t in templates])
        tris = np.concatenate([np.asarray(t, np.float32).reshape(-1, 9) for t in templates], axis=0)
        seg_base = np.concatenate([np.broadcast_to(np.asarray(b, np.int32), (len(t),)) for b, t in zip(tmpl_seg_base, templates)])
        seg_mask = np.concatenate([np.broadcast_to(np.asarray(m, np.int32), (len(t),)) for m, t in zip(tmpl_seg_mask, templates)])
        T = lambda a, dt: torch.tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
        self.E, self.K, self.L, self.P = E, K, L, _next_pow2(K)
        self.tmpl_tri_offset = T(offs, torch.int32)
        self.tmpl_tris = T(tris, torch.float32)
        self.tmpl_seg_base = T(seg_base, torch.int32)
        self.tmpl_seg_mask = T(seg_mask, torch.int32)
        self.obj_template = T(obj_template, torch.int32)
        self.obj_seg_counter = T(np.asarray(obj_seg_counter, np.int32), torch.int32)
        self.obj_pose = obj_pose
        if obj_pose.device != dev or obj_pose.dtype != torch.float32 or obj_pose.shape[:2] != (E, K) or obj_pose.shape[2] < 7:
            raise ValueError("obj_pose must be float32 [E,K,>=7] on the scene's device")
        if obj_pose.stride(2) != 1 or obj_pose.stride(0) != K * obj_pose.stride(1):
            raise ValueError("obj_pose rows must be uniformly strided with a contiguous last dim")
        self.bounds_min, self.bounds_max = bounds_min, bounds_max
        nb = lambda which: int(self.lib.agx_hp2_scene_bytes(K, L, which))
        self.tris = torch.zeros(E * nb(0) // 4, dtype=torch.float32, device=dev)
        self.nodes = torch.zeros(E * nb(1) // 4, dtype=torch.float32, device=dev)
        self.leaf_object = torch.full((E * nb(2) // 4,), -1, dtype=torch.int32, device=dev)
        self.face_offset = torch.zeros(E * K, dtype=torch.int32, device=dev)
        self.tmpl_obb = self.obb = None
        if tmpl_obb is not None:
            tmpl_obb = np.asarray(tmpl_obb, dtype=np.float32).reshape(len(templates), 16)
            self.tmpl_obb = T(tmpl_obb, torch.float32)
            self.obb = torch.zeros(E * K * 16, dtype=torch.float32, device=dev)
        s = AgxHp2Scene()
        s.num_envs, s.num_objects, s.leaves_pow2, s.tris_per_object = E, K, self.P, L
        s.num_templates, s.obj_pose_stride = len(templates), obj_pose.stride(1)
        for name in ("tmpl_tri_offset", "tmpl_tris", "tmpl_seg_base", "tmpl_seg_mask", "obj_pose", "obj_template",
                     "obj_seg_counter", "tris", "nodes", "leaf_object", "face_offset"):
            setattr(s, name, getattr(self, name).data_ptr())
        s.tmpl_obb = self.tmpl_obb.data_ptr() if self.tmpl_obb is not None else None
        s.obb = self.obb.data_ptr() if self.obb is not None else None
        s.bounds_min = bounds_min.data_ptr() if bounds_min is not None else None
        s.bounds_max = bounds_max.data_ptr() if bounds_max is not None else None
        self.c = s

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def update(self, mask: Optional[torch.Tensor] = None):
        """Re-transform + rebuild the BVH of the masked envs (all when mask is None)."""
        if mask is not None and (mask.dtype != torch.bool or mask.shape != (self.E,)):
            raise ValueError("mask must be bool [E]")
        m = C.c_void_p(mask.data_ptr()) if mask is not None else None
        _lib.check(self.lib.agx_hp2_update_scene(C.byref(self.c), m, self._stream()), "agx_hp2_update_scene")


    def collide(self, robot_state: torch.Tensor, radius: float, crashes: torch.Tensor, min_dist: Optional[torch.Tensor] = None):
        """crashes (bool [E]) |= robot collision sphere overlaps the env mesh (row a14)."""
        if crashes.dtype != torch.bool or crashes.shape != (self.E,):
            raise ValueError("crashes must be bool [E]")
        if robot_state.shape[0] != self.E or robot_state.stride(1) != 1:
            raise ValueError("robot_state must be [E,>=3] with contiguous rows")
        md = C.c_void_p(min_dist.data_ptr()) if min_dist is not None else None
        _lib.check(self.lib.agx_hp2_collide(C.byref(self.c), C.c_void_p(robot_state.data_ptr()), robot_state.stride(0),
                                            C.c_float(radius), C.c_void_p(crashes.data_ptr()), md, self._stream()),
                   "agx_hp2_collide")


def camera_intrinsics(width, height, horizontal_fov_deg):
    """WarpCam.initialize_camera_matrices (sensors/warp/warp_cam.py:31-64): fp32 K, its inverse."""
    hf = math.radians(horizontal_fov_deg)
    u0, v0 = width / 2, height / 2
    f = width / 2 * 1 / math.tan(hf / 2)
    vfov = 2 * math.atan(height / (2 * f))
    au, av = u0 / math.tan(hf / 2), v0 / math.tan(vfov / 2)
    K = np.array([[au, 0, u0, 0], [0, av, v0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    Kinv = np.linalg.inv(K.astype(np.float64)).astype(np.float32)
    return Kinv[:3, :3].copy(), int(u0), int(v0)


def lidar_ray_table(height, width, hmin_deg, hmax_deg, vmin_deg, vmax_deg):
    """WarpLidar.initialize_ray_vectors (sensors/warp/warp_lidar.py:40-64)."""
    hmin, hmax = math.radians(hmin_deg), math.radians(hmax_deg)
    vmin, vmax = math.radians(vmin_deg), math.radians(vmax_deg)
    i = np.arange(height, dtype=np.float64)[:, None]
    j = np.arange(width, dtype=np.float64)[None, :]
    az = hmax - (hmax - hmin) * (j / (width - 1))
    el = vmax - (vmax - vmin) * (i / (height - 1))
    t = np.stack([np.cos(az) * np.cos(el), np.sin(az) * np.cos(el), np.sin(el) * np.ones_like(az)], axis=-1).astype(np.float32)
    n = np.sqrt((t ** 2).sum(-1, keepdims=True, dtype=np.float32))
    return (t / n).astype(np.float32)


def _quat_from_euler_deg(e):
    r, p, y = (math.radians(float(x)) for x in e)
    cy, sy, cr, sr, cp, sp = math.cos(y / 2), math.sin(y / 2), math.cos(r / 2), math.sin(r / 2), math.cos(p / 2), math.sin(p / 2)
    return [cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp, cy * cr * cp + sy * sr * sp]


class RaySensor:
    """cfg: a reference-style sensor config class (config/sensor_config/**): reads sensor_type,
    width, height, num_sensors, horizontal_fov_deg | *_fov_deg_min/max, max_range, min_range,
    calculate_depth, return_pointcloud, pointcloud_in_world_frame, segmentation_camera,
    normalize_range, far/near_out_of_range_value, euler_frame_rot_deg, sensor_noise."""

    def __init__(self, cfg, scene: RayScene, robot_state: torch.Tensor, pixels: torch.Tensor,
                 seg_pixels: Optional[torch.Tensor] = None, mount: Optional[torch.Tensor] = None):
        self.lib = _lib.load()
        self.cfg, self.scene = cfg, scene
        dev = scene.device
        E, S, H, W = scene.E, cfg.num_sensors, cfg.height, cfg.width
        st = cfg.sensor_type
        kinds = {"camera": _lib.SENSOR_CAMERA, "lidar": _lib.SENSOR_LIDAR, "stereo_camera": _lib.SENSOR_STEREO_CAMERA,
                 "normal_faceID_camera": _lib.SENSOR_NORMAL_FACEID_CAMERA, "normal_faceID_lidar": _lib.SENSOR_NORMAL_FACEID_LIDAR}
        if st not in kinds:
            raise NotImplementedError(f"sensor_type {st}")
        is_normal = st.startswith("normal_faceID")
        pc = bool(getattr(cfg, "return_pointcloud", False)) or is_normal
        want = (E, S, H, W, 3) if pc else (E, S, H, W)
        if tuple(pixels.shape) != want or pixels.dtype != torch.float32 or not pixels.is_contiguous():
            raise ValueError(f"pixels must be contiguous float32 {want}")
        if seg_pixels is not None and (tuple(seg_pixels.shape) != (E, S, H, W) or seg_pixels.dtype != torch.int32):
            raise ValueError("seg_pixels must be int32 [E,S,H,W]")
        if robot_state.shape[0] != E or robot_state.stride(1) != 1 or robot_state.shape[1] < 7:
            raise ValueError("robot_state must be [E,>=7] with contiguous rows")
        self.robot_state, self.pixels, self.seg_pixels = robot_state, pixels, seg_pixels
        if mount is None:  # identity mount
            mount = torch.zeros(E, S, 7, device=dev)
            mount[..., 6] = 1.0
        self.mount = mount.contiguous()
        noise = getattr(cfg, "sensor_noise", None)
        self.noise_enabled = bool(noise is not None and getattr(noise, "enable_sensor_noise", False))
        s = AgxHp2Sensor()
        s.kind = kinds[st]
        s.baseline = float(getattr(cfg, "baseline", 0.0))
        s.normal_in_world_frame = int(bool(getattr(cfg, "normal_in_world_frame", getattr(cfg, "pointcloud_in_world_frame", True))))
        s.width, s.height, s.num_sensors = W, H, S
        s.calculate_depth = int(bool(getattr(cfg, "calculate_depth", False)))
        s.return_pointcloud = int(pc and not is_normal)
        s.pointcloud_in_world_frame = int(bool(getattr(cfg, "pointcloud_in_world_frame", False)))
        s.segmentation = int(seg_pixels is not None)
        s.fuse_epilogue = int(not self.noise_enabled and not is_normal)  # limits only for camera/lidar/stereo (warp_sensor.py:198-200)
        s.normalize_range = int(bool(cfg.normalize_range))
        self.ray_table = None
        if "camera" in st:
            kinv, cx, cy = camera_intrinsics(W, H, cfg.horizontal_fov_deg)
            for i, v in enumerate(kinv.reshape(-1)):
                s.kinv[i] = float(v)
            s.c_x, s.c_y = cx, cy
        else:
            self.ray_table = torch.tensor(lidar_ray_table(H, W, cfg.horizontal_fov_deg_min, cfg.horizontal_fov_deg_max,
                                                          cfg.vertical_fov_deg_min, cfg.vertical_fov_deg_max), device=dev)
            s.ray_table = self.ray_table.data_ptr()
        s.far_plane = s.max_range = float(cfg.max_range)
        s.min_range = float(cfg.min_range)
        s.far_out_of_range_value = float(cfg.far_out_of_range_value)
        s.near_out_of_range_value = float(cfg.near_out_of_range_value)
        fq = torch.tensor(_quat_from_euler_deg(cfg.euler_frame_rot_deg), dtype=torch.float64).float()
        for i in range(4):
            s.frame_quat[i] = float(fq[i])
        s.robot_pose_stride = robot_state.stride(0)
        s.robot_pose, s.mount = robot_state.data_ptr(), self.mount.data_ptr()
        s.pixels = pixels.data_ptr()
        s.seg_pixels = seg_pixels.data_ptr() if seg_pixels is not None else None
        self.c = s

    def capture(self):
        """One launch: pose compose + ray gen + traversal + (range limits, normalise)."""
        _lib.check(self.lib.agx_hp2_cast(C.byref(self.scene.c), C.byref(self.c), self.scene._stream()), "agx_hp2_cast")
        return self.pixels

    def rays_per_frame(self):
        return self.scene.E * self.cfg.num_sensors * self.cfg.height * self.cfg.width
