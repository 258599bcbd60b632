"""ctypes front-end of oracle/hp2_oracle.c (brute-force ray-cast oracle) + numpy helpers that
restate the host-side constants of the reference sensors.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libhp2_oracle.so")


class Hp2oSensor(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("num_sensors", C.c_int32),
        ("calculate_depth", C.c_int32), ("return_pointcloud", C.c_int32), ("pointcloud_in_world_frame", C.c_int32),
        ("segmentation", C.c_int32), ("fuse_epilogue", C.c_int32), ("normalize_range", C.c_int32),
        ("c_x", C.c_int32), ("c_y", C.c_int32), ("kinv", C.c_float * 9), ("far_plane", C.c_float),
        ("max_range", C.c_float), ("min_range", C.c_float), ("far_out_of_range_value", C.c_float),
        ("near_out_of_range_value", C.c_float), ("frame_quat", C.c_float * 4),
        ("baseline", C.c_float), ("normal_in_world_frame", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "hp2_oracle.c")):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(LIB)
        assert _lib.hp2o_sizeof_sensor() == C.sizeof(Hp2oSensor)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def camera_kinv(width, height, hfov_deg):
    """WarpCam.initialize_camera_matrices (sensors/warp/warp_cam.py:31-64): fp32 K, fp32 inverse."""
    hf = math.radians(hfov_deg)
    u0, v0 = width / 2, height / 2
    f = width / 2 * 1 / math.tan(hf / 2)
    vfov = 2 * math.atan(height / (2 * f))
    au = u0 / math.tan(hf / 2)
    av = v0 / math.tan(vfov / 2)
    K = np.array([[au, 0, u0, 0], [0, av, v0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    Kinv = np.linalg.inv(K.astype(np.float64)).astype(np.float32)
    return Kinv[:3, :3].copy(), int(u0), int(v0)


def lidar_ray_table(height, width, hmin_deg, hmax_deg, vmin_deg, vmax_deg):
    """WarpLidar.initialize_ray_vectors (sensors/warp/warp_lidar.py:40-64): fp32 table, normalised."""
    hmin, hmax = math.radians(hmin_deg), math.radians(hmax_deg)
    vmin, vmax = math.radians(vmin_deg), math.radians(vmax_deg)
    t = np.zeros((height, width, 3), dtype=np.float32)
    for i in range(height):
        el = vmax - (vmax - vmin) * (i / (height - 1))
        for j in range(width):
            az = hmax - (hmax - hmin) * (j / (width - 1))
            t[i, j] = (math.cos(az) * math.cos(el), math.sin(az) * math.cos(el), math.sin(el))
    n = np.sqrt((t.astype(np.float32) ** 2).sum(-1, keepdims=True, dtype=np.float32))
    return (t / n).astype(np.float32)


def quat_from_euler_deg(e):
    r, p, y = [math.radians(float(x)) for x in e]
    cy, sy, cr, sr, cp, sp = math.cos(y / 2), math.sin(y / 2), math.cos(r / 2), math.sin(r / 2), math.cos(p / 2), math.sin(p / 2)
    return np.array([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp,
                     cy * cr * cp + sy * sr * sp], dtype=np.float32)


def box_template(size):
    """12 triangles (9 floats each: v0,v1,v2) of an axis-aligned box centred at the origin --
    what a URDF <box> visual becomes in the reference (trimesh.creation.box, assets/warp_asset.py)."""
    hx, hy, hz = [s / 2.0 for s in size]
    v = np.array([[-hx, -hy, -hz], [-hx, -hy, hz], [-hx, hy, -hz], [-hx, hy, hz],
                  [hx, -hy, -hz], [hx, -hy, hz], [hx, hy, -hz], [hx, hy, hz]], dtype=np.float32)
    f = np.array([[1, 3, 0], [4, 1, 0], [0, 3, 2], [2, 4, 0], [1, 7, 3], [5, 1, 4],
                  [5, 7, 1], [3, 7, 2], [6, 4, 2], [2, 7, 6], [6, 5, 4], [7, 5, 6]])
    return v[f].reshape(12, 9).astype(np.float32)


def build_world_tris(obj_pose, obj_template, obj_seg_counter, tmpl_tri_offset, tmpl_tris, tmpl_seg_base,
                     tmpl_seg_mask, max_tris):
    E, K = obj_template.shape
    obj_pose = np.ascontiguousarray(obj_pose, np.float32)
    out_t = np.zeros((E, max_tris, 9), np.float32)
    out_s = np.zeros((E, max_tris), np.int32)
    out_c = np.zeros(E, np.int32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    a = [obj_pose, i32(obj_template), i32(obj_seg_counter), i32(tmpl_tri_offset),
         np.ascontiguousarray(tmpl_tris, np.float32), i32(tmpl_seg_base), i32(tmpl_seg_mask)]
    lib().hp2o_build_world_tris(
        E, K, _p(a[0], C.c_float), _p(a[1], C.c_int32), _p(a[2], C.c_int32), _p(a[3], C.c_int32),
        _p(a[4], C.c_float), _p(a[5], C.c_int32), _p(a[6], C.c_int32), max_tris,
        _p(out_t, C.c_float), _p(out_s, C.c_int32), _p(out_c, C.c_int32))
    return out_t, out_s, out_c


def cast(sensor: Hp2oSensor, robot_pose, mount, ray_table, tris, seg_ids, tri_count):
    E = robot_pose.shape[0]
    S, H, W = sensor.num_sensors, sensor.height, sensor.width
    shape = (E, S, H, W, 3) if (sensor.return_pointcloud or sensor.kind in (3, 4)) else (E, S, H, W)
    pix = np.zeros(shape, np.float32)
    seg = np.zeros((E, S, H, W), np.int32) if (sensor.segmentation or sensor.kind in (3, 4)) else None
    rp = np.ascontiguousarray(robot_pose, np.float32)
    mt = np.ascontiguousarray(mount, np.float32)
    rt = np.ascontiguousarray(ray_table, np.float32) if ray_table is not None else None
    tris = np.ascontiguousarray(tris, np.float32)
    seg_ids = np.ascontiguousarray(seg_ids, np.int32)
    tri_count = np.ascontiguousarray(tri_count, np.int32)
    lib().hp2o_cast(C.byref(sensor), E, _p(rp, C.c_float), _p(mt, C.c_float), _p(rt, C.c_float),
                    _p(tris, C.c_float), _p(seg_ids, C.c_int32), _p(tri_count, C.c_int32), tris.shape[1],
                    _p(pix, C.c_float), _p(seg, C.c_int32))
    return pix, seg


def collide(robot_pose, radius, tris, tri_count):
    E = robot_pose.shape[0]
    rp = np.ascontiguousarray(robot_pose, np.float32)
    tris = np.ascontiguousarray(tris, np.float32)
    tri_count = np.ascontiguousarray(tri_count, np.int32)
    flags = np.zeros(E, np.uint8)
    d2 = np.zeros(E, np.float32)
    lib().hp2o_collide(E, _p(rp, C.c_float), C.c_float(radius), _p(tris, C.c_float), _p(tri_count, C.c_int32),
                       tris.shape[1], _p(flags, C.c_uint8), _p(d2, C.c_float))
    return flags.astype(bool), d2
