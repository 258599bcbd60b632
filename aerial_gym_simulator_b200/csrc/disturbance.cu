// agx_disturbance_draw: the per-physics-step random wrench of BaseMultirotor.apply_disturbance as ONE launch with a device RNG
// (disturbance_core.cuh) instead of torch.bernoulli + 2 x rand_like + arithmetic; feeds AgxHp1Buffers.disturbance.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aerial_gym_b200.h"
#include "agx_common.cuh"
#include "disturbance_core.cuh"

namespace {
using namespace agx;
constexpr int kThreads = 128;
struct Max6 {
    float v[6];
};

__global__ void __launch_bounds__(kThreads)
disturbance_kernel(int N, uint32_t env_id_offset, uint32_t counter, float prob, const __grid_constant__ Max6 mx, uint32_t k0, uint32_t k1,
                   float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    disturbance_env(env_id_offset + (uint32_t)e, counter, prob, mx.v, k0, k1, out + (size_t)e * 6);
}
}  // namespace

extern "C" int agx_disturbance_draw(int num_envs, int env_id_offset, float prob, const float* max_force_and_torque, uint64_t seed,
                                    uint32_t counter, float* disturbance, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_envs == 0) return AGX_OK;
    if (!max_force_and_torque || !disturbance) return agx_set_error(AGX_E_NULL, "agx_disturbance_draw: NULL argument");
    if (!(prob >= 0.0f && prob <= 1.0f)) return agx_set_error(AGX_E_INVALID, "agx_disturbance_draw: prob must be in [0, 1]");
    Max6 mx;
    for (int j = 0; j < 6; ++j) mx.v[j] = max_force_and_torque[j];  // a HOST array of six floats (config values)
    disturbance_kernel<<<(num_envs + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(
        num_envs, (uint32_t)env_id_offset, counter, prob, mx, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), disturbance);
    return agx_check_launch("disturbance_kernel");
}
