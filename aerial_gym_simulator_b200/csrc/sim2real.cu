// agx_s2r_reward / agx_s2r_obs: reward and observation epilogues of the reference's two setpoint-command sim2real position tasks
// (velocity commands, acceleration commands; lmf2) -- one thread per env, arithmetic in sim2real_core.cuh.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aerial_gym_b200.h"
#include "agx_common.cuh"
#include "sim2real_core.cuh"

namespace {
using namespace agx;
constexpr int kEnvThreads = 128;

__global__ void __launch_bounds__(kEnvThreads)
s2r_reward_kernel(int N, int variant, const float* __restrict__ state, int stride, const float* __restrict__ veh_q,
                  const float* __restrict__ body_linvel, const float* __restrict__ target, const float* __restrict__ prev_dist,
                  const float* __restrict__ act, const float* __restrict__ prev_act, float* __restrict__ act_vehicle_out,
                  uint8_t* __restrict__ crashes, float* __restrict__ rewards) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    s2r_reward_env(e, variant, state, stride, veh_q, body_linvel, target, prev_dist, act, prev_act, act_vehicle_out, crashes, rewards);
}

__global__ void __launch_bounds__(kEnvThreads)
s2r_obs_kernel(int N, float* __restrict__ state, int stride, const float* __restrict__ body_linvel, const float* __restrict__ body_angvel,
               const float* __restrict__ robot_actions, const float* __restrict__ target, const float* __restrict__ noise,
               float* __restrict__ obs, int obs_stride) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    s2r_obs_env(e, state, stride, body_linvel, body_angvel, robot_actions, target, noise, obs, obs_stride);
}

inline int blocks_for(int n) { return (n + kEnvThreads - 1) / kEnvThreads; }
}  // namespace

extern "C" {

int agx_s2r_reward(int num_envs, int variant, const float* robot_state, int robot_state_stride, const float* vehicle_orientation,
                   const float* body_linvel, const float* target_position, const float* prev_dist, const float* actions,
                   const float* prev_actions, float* actions_vehicle_frame, uint8_t* crashes, float* rewards, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (variant != 0 && variant != 1) return agx_set_error(AGX_E_INVALID, "agx_s2r_reward: variant must be 0 (velocity) or 1 (acceleration)");
    if (num_envs == 0) return AGX_OK;
    if (!robot_state || !vehicle_orientation || !body_linvel || !prev_dist || !actions || !prev_actions || !crashes || !rewards)
        return agx_set_error(AGX_E_NULL, "agx_s2r_reward: NULL argument");
    if (robot_state_stride < 7) return agx_set_error(AGX_E_INVALID, "agx_s2r_reward: robot_state_stride < 7");
    s2r_reward_kernel<<<blocks_for(num_envs), kEnvThreads, 0, (cudaStream_t)stream>>>(num_envs, variant, robot_state, robot_state_stride,
                                                                                     vehicle_orientation, body_linvel, target_position, prev_dist,
                                                                                     actions, prev_actions, actions_vehicle_frame, crashes, rewards);
    return agx_check_launch("s2r_reward_kernel");
}

int agx_s2r_obs(int num_envs, float* robot_state, int robot_state_stride, const float* body_linvel, const float* body_angvel,
                const float* robot_actions, const float* target_position, const float* noise, float* obs, int obs_stride, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_envs == 0) return AGX_OK;
    if (!robot_state || !body_linvel || !body_angvel || !robot_actions || !noise || !obs) return agx_set_error(AGX_E_NULL, "agx_s2r_obs: NULL argument");
    if (robot_state_stride < 7 || obs_stride < 17) return agx_set_error(AGX_E_INVALID, "agx_s2r_obs: stride too small");
    s2r_obs_kernel<<<blocks_for(num_envs), kEnvThreads, 0, (cudaStream_t)stream>>>(num_envs, robot_state, robot_state_stride, body_linvel, body_angvel,
                                                                                  robot_actions, target_position, noise, obs, obs_stride);
    return agx_check_launch("s2r_obs_kernel");
}

}  // extern "C"
