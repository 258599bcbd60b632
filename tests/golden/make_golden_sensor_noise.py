"""Golden vectors for the ray-cast sensors' noise / range-limit / normalisation epilogue, produced by RUNNING THE
REFERENCE'S OWN CODE on CPU (this container only):

    python tests/golden/make_golden_sensor_noise.py

sensors/warp/warp_sensor.py imports warp, so WarpSensor.apply_noise, apply_range_limits and normalize_observation are pulled
out of the file with ``ast`` and executed unchanged on a stand-in ``self`` (cfg + pixels), under torch.manual_seed, in the order
WarpSensor.update calls them (:197-200).  Cases: a LiDAR range image with the LiDAR configs' noise parameters, a
sensor-frame point cloud, a world-frame point cloud (noise only), and a depth image without noise (limits + normalisation)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_loader  # noqa: E402
from make_golden_aux import _funcs_from  # noqa: E402

PATH = os.path.join(_ref_loader.REF_ROOT, "aerial_gym/sensors/warp/warp_sensor.py")
SEED = 4321

CASES = {
    # name: (sensor_type, pointcloud, world, normalize, noise on, dropout)
    "lidar_range": ("lidar", False, False, True, True, 0.02),
    "lidar_pc_sensor": ("lidar", True, False, True, True, 0.01),
    "lidar_pc_world": ("lidar", True, True, False, True, 0.0),
    "camera_depth_nonoise": ("camera", False, False, True, False, 0.0),
}


def make_cfg(sensor_type, pc, world, normalize, noise, dropout):
    nz = types.SimpleNamespace(enable_sensor_noise=noise, std_a=0.00038089, std_b=-0.00343351, std_c=0.01553284, mean_offset=-0.025,
                               pixel_dropout_prob=dropout)
    return types.SimpleNamespace(sensor_type=sensor_type, return_pointcloud=pc, pointcloud_in_world_frame=world, normalize_range=normalize,
                                 max_range=10.0, min_range=0.2, far_out_of_range_value=10.0, near_out_of_range_value=-10.0, sensor_noise=nz)


def main():
    ns = {"torch": torch}
    _funcs_from(PATH, {"apply_noise", "apply_range_limits", "normalize_observation"}, ns, in_class="WarpSensor")
    out = {"seed": np.int64(SEED)}
    g = torch.Generator().manual_seed(9)
    for name, spec in CASES.items():
        cfg = make_cfg(*spec)
        shape = (3, 1, 12, 20, 3) if spec[1] else (3, 1, 12, 20)
        px = torch.rand(shape, generator=g) * 12.0 + 0.05  # some beyond max range, some below min range
        if spec[1]:
            px = (torch.rand(shape, generator=g) * 2 - 1) * 8.0
            px[0, 0, :2] *= 0.01  # points closer than min_range
        px.view(-1)[::37] = 1000.0  # rays that hit nothing (NO_HIT_RAY_VAL)
        out[f"{name}_in"] = px.numpy().copy()
        me = types.SimpleNamespace(cfg=cfg, pixels=px)
        torch.manual_seed(SEED)
        ns["apply_noise"](me)
        if cfg.sensor_type in ["camera", "lidar", "stereo_camera"]:  # warp_sensor.py:198-200
            ns["apply_range_limits"](me)
            ns["normalize_observation"](me)
        out[f"{name}_out"] = me.pixels.numpy().copy()
        out[f"{name}_spec"] = np.array([spec[0], *[str(int(x)) if isinstance(x, bool) else str(x) for x in spec[1:]]])
    np.savez_compressed(os.path.join(HERE, "sensor_noise.npz"), **out)
    print("wrote sensor_noise.npz")


if __name__ == "__main__":
    main()
