// agx_obstacle_step: kinematic advance of the obstacles of the "dynamic_env" environment (env_manager/obstacle_manager.py:40-44 +
// the PhysX motion the reference relies on) -- one thread per obstacle row, arithmetic in obstacle_core.cuh (our spec, DESIGN 1 row f2').
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aerial_gym_b200.h"
#include "agx_common.cuh"
#include "obstacle_core.cuh"

namespace {
using namespace agx;
constexpr int kEnvThreads = 128;

__global__ void __launch_bounds__(kEnvThreads)
obstacle_step_kernel(long long total, float* __restrict__ state, int stride, const float* __restrict__ twist, float dt, int substeps,
                     float lin_damp, float ang_damp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    obstacle_step_item(i, state, stride, twist, dt, substeps, lin_damp, ang_damp);
}

inline int blocks_for(long long n) { return (int)((n + kEnvThreads - 1) / kEnvThreads); }
}  // namespace

extern "C" int agx_obstacle_step(int num_envs, int num_assets, float* asset_state, int asset_stride, const float* twist, float dt, int substeps,
                      float linear_damping, float angular_damping, void* stream) {
    if (num_envs < 0 || num_assets < 0 || substeps < 0) return agx_set_error(AGX_E_INVALID, "agx_obstacle_step: negative size");
    if (asset_stride < 13) return agx_set_error(AGX_E_INVALID, "agx_obstacle_step: asset_stride < 13");
    if (!(dt > 0.0f)) return agx_set_error(AGX_E_INVALID, "agx_obstacle_step: dt must be > 0");
    const long long total = (long long)num_envs * num_assets;
    if (total == 0 || substeps == 0) return AGX_OK;
    if (!asset_state) return agx_set_error(AGX_E_NULL, "agx_obstacle_step: asset_state is NULL");
    obstacle_step_kernel<<<blocks_for(total), kEnvThreads, 0, (cudaStream_t)stream>>>(total, asset_state, asset_stride, twist, dt, substeps,
                                                                                      linear_damping, angular_damping);
    return agx_check_launch("obstacle_step_kernel");
}
