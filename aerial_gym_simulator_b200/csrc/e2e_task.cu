// agx_e2e_reward / agx_e2e_obs: reward and observation epilogues of the reference's two motor-command position tasks
// (sim2real_end_to_end, sim2real_px4) -- one thread per env, arithmetic in e2e_task_core.cuh.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aerial_gym_b200.h"
#include "agx_common.cuh"
#include "e2e_task_core.cuh"

namespace {
using namespace agx;
constexpr int kEnvThreads = 128;

__global__ void __launch_bounds__(kEnvThreads)
e2e_reward_kernel(int N, const float* __restrict__ state, int stride, const float* __restrict__ body_angvel, const float* __restrict__ target,
                  const float* __restrict__ act, const float* __restrict__ prev_act, const float* __restrict__ prev_pos_err,
                  const __grid_constant__ AgxE2ERewardParams p, uint8_t* __restrict__ crashes, float* __restrict__ rewards) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    e2e_reward_env(e, state, stride, body_angvel, target, act, prev_act, prev_pos_err, p, crashes, rewards);
}

__global__ void __launch_bounds__(kEnvThreads)
e2e_obs_kernel(int N, const float* __restrict__ state, int stride, const float* __restrict__ body_angvel, const float* __restrict__ target,
               const float* __restrict__ noise, float* __restrict__ obs, int obs_stride) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    e2e_obs_env(e, state, stride, body_angvel, target, noise, obs, obs_stride);
}

inline int blocks_for(int n) { return (n + kEnvThreads - 1) / kEnvThreads; }
}  // namespace

extern "C" {

int agx_e2e_reward(int num_envs, const float* robot_state, int robot_state_stride, const float* body_angvel, const float* target_position,
                   const float* actions, const float* prev_actions, const float* prev_pos_error, const AgxE2ERewardParams* params,
                   uint8_t* crashes, float* rewards, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_envs == 0) return AGX_OK;
    if (!robot_state || !body_angvel || !actions || !prev_actions || !prev_pos_error || !params || !crashes || !rewards)
        return agx_set_error(AGX_E_NULL, "agx_e2e_reward: NULL argument");
    if (robot_state_stride < 10) return agx_set_error(AGX_E_INVALID, "agx_e2e_reward: robot_state_stride < 10");
    e2e_reward_kernel<<<blocks_for(num_envs), kEnvThreads, 0, (cudaStream_t)stream>>>(num_envs, robot_state, robot_state_stride, body_angvel,
                                                                                     target_position, actions, prev_actions, prev_pos_error,
                                                                                     *params, crashes, rewards);
    return agx_check_launch("e2e_reward_kernel");
}

int agx_e2e_obs(int num_envs, const float* robot_state, int robot_state_stride, const float* body_angvel, const float* target_position,
                const float* noise, float* obs, int obs_stride, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_envs == 0) return AGX_OK;
    if (!robot_state || !body_angvel || !noise || !obs) return agx_set_error(AGX_E_NULL, "agx_e2e_obs: NULL argument");
    if (robot_state_stride < 10 || obs_stride < 15) return agx_set_error(AGX_E_INVALID, "agx_e2e_obs: stride too small");
    e2e_obs_kernel<<<blocks_for(num_envs), kEnvThreads, 0, (cudaStream_t)stream>>>(num_envs, robot_state, robot_state_stride, body_angvel,
                                                                                  target_position, noise, obs, obs_stride);
    return agx_check_launch("e2e_obs_kernel");
}

}  // extern "C"
