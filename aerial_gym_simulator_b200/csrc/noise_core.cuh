// Ray-cast sensor noise + range limits + normalisation in ONE pass over the image, device RNG
// (WarpSensor.apply_noise / apply_range_limits / normalize_observation, sensors/warp/warp_sensor.py:202-247: in the reference
// ~10 full-image torch passes -- pixels**2, two scalings, torch.normal, ones_like, bernoulli, three masked assignments, a division).
// Same distributions, different stream: the reference draws with torch.normal / torch.bernoulli; here every value gets its own
// Philox4x32-10 block, counter = (pixel lo, pixel hi, frame, component), key = seed -- reproducible, independent of launch
// geometry and of sharding, no host RNG state.  The reference-order torch path stays available
// (aerial_gym_simulator_b200/sensors/noise.py, pinned bit for bit against the reference's own functions on CPU).
// AGX_DEV: device-only in the product, host+device in the CPU shadow build.  Oracle: oracle/sensor_noise_oracle.py.
#pragma once
#include "../../include/aerial_gym_b200.h"
#include "agx_math.cuh"

namespace agx {

// standard normal from two Philox words (Box-Muller, cosine branch); u1 in (0, 1] so the log is finite
AGX_DEV float normal_from_u32(uint32_t a, uint32_t b) {
    const float u1 = (float)((a >> 8) + 1u) * (1.0f / 16777216.0f);
    const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cosf(AGX_TWO_PI_F * u2);
}

// one value: torch.normal(mean = p - mean_offset, std = a p^2 + b p + c) (:239-243), then dropout -> near value (:244-250)
AGX_DEV float noisy_value(float p, const AgxHp2Noise& n, uint64_t pixel, uint32_t frame, uint32_t comp, uint32_t k0, uint32_t k1) {
    const U4 r = philox4x32_10(U4{(uint32_t)pixel, (uint32_t)(pixel >> 32), frame, comp}, k0, k1);
    const float std_val = n.std_a * (p * p) + n.std_b * p + n.std_c;  // :238
    float v = (p - n.mean_offset) + std_val * normal_from_u32(r.x, r.y);
    if (u01(r.z) < n.pixel_dropout_prob) v = n.near_out_of_range_value;
    return v;
}

// local pixel i (global index g = first_pixel + i keys the RNG) of an image with n.components (1 or 3) floats per pixel, in place
AGX_DEV void noise_limits_pixel(uint64_t i, uint64_t g, float* pixels, const AgxHp2Noise& n, uint32_t frame, uint32_t k0, uint32_t k1) {
    if (n.components == 3) {
        float* p = pixels + i * 3;
        float v0 = p[0], v1 = p[1], v2 = p[2];
        if (n.enable_noise) {
            v0 = noisy_value(v0, n, g, frame, 0u, k0, k1);
            v1 = noisy_value(v1, n, g, frame, 1u, k0, k1);
            v2 = noisy_value(v2, n, g, frame, 2u, k0, k1);
        }
        if (n.apply_limits) {  // sensor-frame point clouds: limits on the point's norm, all three components replaced (:204-215)
            if (sqrtf(v0 * v0 + v1 * v1 + v2 * v2) > n.max_range) v0 = v1 = v2 = n.far_out_of_range_value;
            if (sqrtf(v0 * v0 + v1 * v1 + v2 * v2) < n.min_range) v0 = v1 = v2 = n.near_out_of_range_value;
        }
        if (n.normalize) { v0 = v0 / n.max_range; v1 = v1 / n.max_range; v2 = v2 / n.max_range; }  // :222-225
        p[0] = v0; p[1] = v1; p[2] = v2;
    } else {
        float v = pixels[i];
        if (n.enable_noise) v = noisy_value(v, n, g, frame, 0u, k0, k1);
        if (n.apply_limits) {  // :217-219, two masked assignments one after the other
            if (v > n.max_range) v = n.far_out_of_range_value;
            if (v < n.min_range) v = n.near_out_of_range_value;
        }
        if (n.normalize) v = v / n.max_range;
        pixels[i] = v;
    }
}

}  // namespace agx
