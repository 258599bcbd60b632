"""Ray-cast sensor noise, range limits and normalisation (sensors/warp/warp_sensor.py:202-247), two modes:

* ``apply_noise_and_limits_torch``: the reference's torch calls in the reference's order (torch.normal, then
  torch.bernoulli), so a seeded run reproduces its draws -- pinned bit for bit against the reference's own functions by
  tests/test_sensor_noise_cpu.py.  Used when ``args['reset_rng'] == 'torch'`` (every random number from torch).
* ``DeviceSensorNoise``: one in-place CUDA pass with a counter-based device RNG (agx_hp2_noise_limits, csrc/noise_core.cuh):
  same distributions, its own stream, no host RNG state, ~10x less HBM traffic.  Default (like the in-kernel Philox resets)."""
import ctypes as C

import torch

from .. import _lib


def apply_noise_and_limits_torch(px, cfg):
    """WarpSensor.apply_noise (:227-250) -> apply_range_limits (:202-220) -> normalize_observation (:222-225), in place."""
    nz = cfg.sensor_noise
    if nz.enable_sensor_noise:
        std_val = nz.std_a * px**2 + nz.std_b * px + nz.std_c
        px[:] = torch.normal(mean=(px - nz.mean_offset), std=std_val)
        px[torch.bernoulli(torch.ones_like(px) * nz.pixel_dropout_prob) > 0] = cfg.near_out_of_range_value
    if cfg.sensor_type in ("camera", "lidar", "stereo_camera"):
        if cfg.return_pointcloud:
            if not cfg.pointcloud_in_world_frame:
                px[px.norm(dim=4, keepdim=True).expand(-1, -1, -1, -1, 3) > cfg.max_range] = cfg.far_out_of_range_value
                px[px.norm(dim=4, keepdim=True).expand(-1, -1, -1, -1, 3) < cfg.min_range] = cfg.near_out_of_range_value
        else:
            px[px > cfg.max_range] = cfg.far_out_of_range_value
            px[px < cfg.min_range] = cfg.near_out_of_range_value
        if cfg.normalize_range and not cfg.pointcloud_in_world_frame:
            px[:] = px / cfg.max_range
    return px


def noise_struct(cfg):
    """cfg (a reference-style sensor config class) -> AgxHp2Noise."""
    n, nz = _lib.AgxHp2Noise(), cfg.sensor_noise
    pc = bool(getattr(cfg, "return_pointcloud", False))
    world = bool(getattr(cfg, "pointcloud_in_world_frame", False))
    ranged = cfg.sensor_type in ("camera", "lidar", "stereo_camera")
    n.components = 3 if pc else 1
    n.enable_noise = int(bool(nz.enable_sensor_noise))
    n.apply_limits = int(ranged and not (pc and world))
    n.normalize = int(ranged and bool(cfg.normalize_range) and not world)
    n.std_a, n.std_b, n.std_c = float(nz.std_a), float(nz.std_b), float(nz.std_c)
    n.mean_offset, n.pixel_dropout_prob = float(nz.mean_offset), float(nz.pixel_dropout_prob)
    n.max_range, n.min_range = float(cfg.max_range), float(cfg.min_range)
    n.far_out_of_range_value, n.near_out_of_range_value = float(cfg.far_out_of_range_value), float(cfg.near_out_of_range_value)
    return n


class DeviceSensorNoise:
    """In-place noise + limits + normalisation of a sensor's pixel tensor; ``frame`` advances once per call."""

    def __init__(self, cfg, pixels: torch.Tensor, seed: int, first_pixel: int = 0):
        if not pixels.is_cuda or pixels.dtype != torch.float32 or not pixels.is_contiguous():
            raise ValueError("pixels must be a contiguous float32 CUDA tensor")
        self.lib, self.pixels, self.seed, self.frame = _lib.load(), pixels, int(seed) & (2**64 - 1), 0
        self.c = noise_struct(cfg)
        self.num_pixels = pixels.numel() // self.c.components
        self.first_pixel = int(first_pixel)  # global index of pixels[0]: env_id_offset x pixels per env on a sharded run

    def apply(self):
        stream = C.c_void_p(torch.cuda.current_stream(self.pixels.device).cuda_stream)
        _lib.check(self.lib.agx_hp2_noise_limits(C.c_void_p(self.pixels.data_ptr()), self.num_pixels, self.first_pixel, C.byref(self.c), self.seed,
                                                 self.frame & 0xFFFFFFFF, stream), "agx_hp2_noise_limits")
        self.frame += 1
        return self.pixels
