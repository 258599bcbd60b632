"""Shared scene / sensor builders for the HP2 tests (oracle side in numpy, product side in torch)."""
import numpy as np
import torch

from oracle import hp2_oracle as RO


def rand_quat(g, *shape):
    q = torch.randn(*shape, 4, generator=g)
    return q / q.norm(dim=-1, keepdim=True)


def make_scene(E, K, seed, extent=5.0, parked=0, n_templates=5):
    """K randomly posed boxes per env inside [-extent, extent]^3; the last `parked` objects are
    parked at -1000 like the reference does with unused obstacles (asset_manager.py:71)."""
    g = torch.Generator().manual_seed(seed)
    sizes = torch.rand(n_templates, 3, generator=g) * 1.2 + 0.15
    templates = [RO.box_template(s.tolist()) for s in sizes]
    pose = torch.zeros(E, K, 13)
    pose[..., 0:3] = (torch.rand(E, K, 3, generator=g) * 2 - 1) * extent
    pose[..., 3:7] = rand_quat(g, E, K)
    if parked:
        pose[:, K - parked:, 0:3] = -1000.0
    tm = torch.randint(0, n_templates, (E, K), generator=g).numpy().astype(np.int32)
    ctr = (100 + torch.arange(E * K).reshape(E, K)).numpy().astype(np.int32)
    offs = np.arange(0, 12 * (n_templates + 1), 12, dtype=np.int32)
    from aerial_gym_simulator_b200.hp2 import box_obb
    obbs = np.stack([box_obb(s.tolist()) for s in sizes])
    return dict(templates=templates, obbs=obbs, tm=tm, ctr=ctr, pose=pose, offs=offs, tmpl_tris=np.concatenate(templates),
                seg_base=np.zeros(12 * n_templates, np.int32), seg_mask=np.ones(12 * n_templates, np.int32), E=E, K=K)


def oracle_tris(sc):
    return RO.build_world_tris(sc["pose"][..., :7].numpy(), sc["tm"], sc["ctr"], sc["offs"], sc["tmpl_tris"],
                               sc["seg_base"], sc["seg_mask"], sc["K"] * 12)


def robot_poses(E, seed, extent=4.0):
    g = torch.Generator().manual_seed(seed)
    r = torch.zeros(E, 13)
    r[:, 0:3] = (torch.rand(E, 3, generator=g) * 2 - 1) * extent
    r[:, 3:7] = rand_quat(g, E)
    return r


def mounts(E, S, seed):
    g = torch.Generator().manual_seed(seed)
    m = torch.zeros(E, S, 7)
    m[..., 0:3] = torch.rand(E, S, 3, generator=g) * 0.1
    e = (torch.rand(E, S, 3, generator=g) * 10 - 5) * np.pi / 180
    cy, sy = torch.cos(e[..., 2] / 2), torch.sin(e[..., 2] / 2)
    cr, sr = torch.cos(e[..., 0] / 2), torch.sin(e[..., 0] / 2)
    cp, sp = torch.cos(e[..., 1] / 2), torch.sin(e[..., 1] / 2)
    m[..., 3] = cy * sr * cp - sy * cr * sp
    m[..., 4] = cy * cr * sp + sy * sr * cp
    m[..., 5] = sy * cr * cp - cy * sr * sp
    m[..., 6] = cy * cr * cp + sy * sr * sp
    return m


class CamCfg:
    """BaseDepthCameraConfig values (config/sensor_config/camera_config/base_depth_camera_config.py)."""
    sensor_type, num_sensors, height, width = "camera", 1, 48, 64
    horizontal_fov_deg, max_range, min_range = 87.0, 10.0, 0.2
    calculate_depth, return_pointcloud, pointcloud_in_world_frame = True, False, False
    segmentation_camera, normalize_range = True, True
    far_out_of_range_value, near_out_of_range_value = 10.0, -10.0
    euler_frame_rot_deg = [-90.0, 0, -90.0]

    class sensor_noise:
        enable_sensor_noise = False


class LidarCfg:
    """OSDome_64_Config values (config/sensor_config/lidar_config/osdome_64_config.py), smaller grid."""
    sensor_type, num_sensors, height, width = "lidar", 1, 16, 64
    horizontal_fov_deg_min, horizontal_fov_deg_max = -180, 180
    vertical_fov_deg_min, vertical_fov_deg_max = 0, 90
    max_range, min_range = 20.0, 0.5
    return_pointcloud, pointcloud_in_world_frame = False, False
    segmentation_camera, normalize_range = True, True
    far_out_of_range_value, near_out_of_range_value = 20.0, -20.0
    euler_frame_rot_deg = [0.0, 0.0, 0.0]

    class sensor_noise:
        enable_sensor_noise = False


def cfg_variant(base, **kw):
    return type("Cfg", (base,), kw)


def oracle_sensor(cfg, fuse=True):
    s = RO.Hp2oSensor()
    s.kind = {"camera": 0, "lidar": 1, "stereo_camera": 2, "normal_faceID_camera": 3, "normal_faceID_lidar": 4}[cfg.sensor_type]
    s.baseline = float(getattr(cfg, "baseline", 0.0))
    s.normal_in_world_frame = int(getattr(cfg, "normal_in_world_frame", getattr(cfg, "pointcloud_in_world_frame", True)))
    s.width, s.height, s.num_sensors = cfg.width, cfg.height, cfg.num_sensors
    s.calculate_depth = int(getattr(cfg, "calculate_depth", False))
    s.return_pointcloud = int(cfg.return_pointcloud and s.kind < 3)
    s.pointcloud_in_world_frame = int(cfg.pointcloud_in_world_frame)
    s.segmentation = int(cfg.segmentation_camera)
    s.fuse_epilogue = int(fuse and s.kind < 3)
    s.normalize_range = int(cfg.normalize_range)
    table = None
    if s.kind in (0, 2, 3):
        kinv, cx, cy = RO.camera_kinv(cfg.width, cfg.height, cfg.horizontal_fov_deg)
        for i, v in enumerate(kinv.reshape(-1)):
            s.kinv[i] = float(v)
        s.c_x, s.c_y = cx, cy
    else:
        table = RO.lidar_ray_table(cfg.height, cfg.width, cfg.horizontal_fov_deg_min, cfg.horizontal_fov_deg_max,
                                   cfg.vertical_fov_deg_min, cfg.vertical_fov_deg_max)
    s.far_plane = s.max_range = cfg.max_range
    s.min_range = cfg.min_range
    s.far_out_of_range_value, s.near_out_of_range_value = cfg.far_out_of_range_value, cfg.near_out_of_range_value
    fq = RO.quat_from_euler_deg(cfg.euler_frame_rot_deg)
    for i in range(4):
        s.frame_quat[i] = float(fq[i])
    return s, table
