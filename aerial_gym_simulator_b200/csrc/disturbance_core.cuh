// Random wrench disturbance on body 0 (BaseMultirotor.apply_disturbance, robots/base_multirotor.py:213-234): with probability p an
// env receives a force ~ U(-max, max)^3 and a torque ~ U(-max, max)^3, otherwise zero.  The reference draws with torch.bernoulli +
// two rand_like per physics step (three RNG launches + arithmetic); this is the device-RNG form: one Philox stream per env,
// counter = (global env id, draw counter, block, 'DIST'), key = seed -- same distributions, its own stream, used when the env runs with
// reset_rng = "device" (the torch-order path stays for reset_rng = "torch").  AGX_DEV: device-only in the product, host+device in
// the CPU shadow build.  Oracle: oracle/disturbance_oracle.py.
#pragma once
#include "agx_math.cuh"

namespace agx {

constexpr uint32_t kDisturbanceTag = 0x44495354u;  // "DIST"

// out6 = gate * lerp(-max, max, u) per component (utils/math.py:51-54: (upper - lower) * u + lower)
AGX_DEV void disturbance_env(uint32_t env_gid, uint32_t counter, float prob, const float* max6, uint32_t k0, uint32_t k1, float* out6) {
    const U4 a = philox4x32_10(U4{env_gid, counter, 0u, kDisturbanceTag}, k0, k1);
    const U4 b = philox4x32_10(U4{env_gid, counter, 1u, kDisturbanceTag}, k0, k1);
    const float gate = (u01(a.x) < prob) ? 1.0f : 0.0f;  // torch.bernoulli(p): 1 with probability p
    const float u[6] = {u01(a.y), u01(a.z), u01(a.w), u01(b.x), u01(b.y), u01(b.z)};
    for (int j = 0; j < 6; ++j) {
        const float lo = -max6[j], hi = max6[j];
        // explicit round-to-nearest multiply THEN add: nvcc would contract (hi-lo)*u+lo into one FMA (one rounding), the numpy oracle
        // and torch's own (upper - lower) * u + lower round twice -- the bit contract is the two-rounding form
#if defined(__CUDA_ARCH__)
        out6[j] = __fmul_rn(__fadd_rn(__fmul_rn(hi - lo, u[j]), lo), gate);
#else
        out6[j] = ((hi - lo) * u[j] + lo) * gate;  // host shadow: built with -ffp-contract=off
#endif
    }
}

}  // namespace agx
