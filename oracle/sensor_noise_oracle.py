"""CPU restatement (numpy) of the device-RNG sensor noise pass (csrc/noise_core.cuh, agx_hp2_noise_limits).

TEST INFRASTRUCTURE ONLY.  The ARITHMETIC follows WarpSensor.apply_noise / apply_range_limits / normalize_observation
(sensors/warp/warp_sensor.py:202-247); the RANDOM STREAM has no counterpart in the reference (it draws with torch.normal /
torch.bernoulli -- that path is aerial_gym_simulator_b200/sensors/noise.py: apply_noise_and_limits_torch, pinned against the
reference's own functions).  Here: Philox4x32-10 (oracle/philox.py, Random123 known answers), counter = (pixel lo, pixel hi,
frame, component), key = seed; normal = Box-Muller cosine branch of words 0, 1 with u1 = ((x >> 8) + 1) 2^-24; dropout
uniform = word 2."""
import numpy as np

from . import philox


def noise_limits(pixels, components, enable_noise, apply_limits, normalize, std_a, std_b, std_c, mean_offset, dropout, max_range, min_range,
                 far_value, near_value, seed, frame, first_pixel=0):
    """pixels: float32 [..., components] flattened pixel-major.  Returns a new float32 array of the same shape."""
    f = np.float32
    px = np.asarray(pixels, np.float32).reshape(-1, components).copy()
    P = px.shape[0]
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    idx = np.arange(P, dtype=np.uint64) + np.uint64(first_pixel)
    if enable_noise:
        for c in range(components):
            ctr = np.stack([(idx & np.uint64(0xFFFFFFFF)).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32),
                            np.full(P, frame, np.uint32), np.full(P, c, np.uint32)], axis=-1)
            r = philox.philox4x32_10(ctr, key)
            u1 = ((r[:, 0] >> np.uint32(8)).astype(np.float32) + f(1.0)) * f(1.0 / 16777216.0)
            u2 = philox.u01(r[:, 1])
            z = np.sqrt(f(-2.0) * np.log(u1)) * np.cos(f(6.283185307179586) * u2)
            p = px[:, c]
            std = f(std_a) * (p * p) + f(std_b) * p + f(std_c)
            v = (p - f(mean_offset)) + std * z.astype(np.float32)
            v = np.where(philox.u01(r[:, 2]) < f(dropout), f(near_value), v)
            px[:, c] = v
    if apply_limits:
        if components == 3:
            far = np.sqrt((px * px).sum(1, dtype=np.float32)) > f(max_range)
            px[far] = f(far_value)
            near = np.sqrt((px * px).sum(1, dtype=np.float32)) < f(min_range)
            px[near] = f(near_value)
        else:
            px[px > f(max_range)] = f(far_value)
            px[px < f(min_range)] = f(near_value)
    if normalize:
        px = px / f(max_range)
    return px.reshape(np.asarray(pixels).shape).astype(np.float32)
