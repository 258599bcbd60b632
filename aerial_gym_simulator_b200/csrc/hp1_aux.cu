// Small per-env kernels next to the HP1 step: the NavigationTask epilogue
// (reward + observation assembly; the VAE encoder stays in torch) and the IMU sensor.
// One thread per env, plain fp32, the reference's operation order; random draws are INPUTS (the
// reference draws them with torch, the host keeps that call order).  Oracle: oracle/aux_oracle.py.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aerial_gym_b200.h"
#include "agx_common.cuh"
#include "agx_math.cuh"
#include "aux_core.cuh"

namespace {
using namespace agx;

constexpr int kAuxThreads = 128;


// NavigationTask.compute_rewards_and_crashes (:397-418) + compute_reward (:436-521)
__global__ void __launch_bounds__(kAuxThreads)
nav_reward_kernel(int N, const float* __restrict__ state, int stride, const float* __restrict__ veh_q, const float* __restrict__ target,
                  const uint8_t* __restrict__ crashes, const float* __restrict__ act, const float* __restrict__ prev_act, float frac,
                  const __grid_constant__ AgxNavRewardParams p, float* __restrict__ pos_err, float* __restrict__ pos_err_prev,
                  float* __restrict__ rewards) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    nav_reward_env(e, state, stride, veh_q, target, crashes, act, prev_act, frac, p, pos_err, pos_err_prev, rewards);
}

// NavigationTask.process_obs_for_task (:369-395), columns 0..16
__global__ void __launch_bounds__(kAuxThreads)
nav_obs_kernel(int N, const float* __restrict__ state, int stride, const float* __restrict__ veh_q, const float* __restrict__ euler,
               const float* __restrict__ blv, const float* __restrict__ bav, const float* __restrict__ actions,
               const float* __restrict__ target, const float* __restrict__ u_vec, const float* __restrict__ u_euler,
               float* __restrict__ obs, int obs_stride) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    nav_obs_env(e, state, stride, veh_q, euler, blv, bav, actions, target, u_vec, u_euler, obs, obs_stride);
}

// IMUSensor.update (sensors/imu_sensor.py:85-131)
__global__ void __launch_bounds__(kAuxThreads)
imu_kernel(int N, const __grid_constant__ AgxImuConfig c, const float* __restrict__ force, int force_stride, const float* __restrict__ mass,
           const float* __restrict__ state, int stride, const float* __restrict__ bav, const float* __restrict__ sensor_q,
           const float* __restrict__ n_noise, const float* __restrict__ n_bias, float* __restrict__ bias, float* __restrict__ meas) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    imu_env(e, c, force, force_stride, mass, state, stride, bav, sensor_q, n_noise, n_bias, bias, meas);
}

inline int blocks_for(int n) { return (n + kAuxThreads - 1) / kAuxThreads; }

}  // namespace

extern "C" {

int agx_nav_reward(int num_envs, const float* robot_state, int robot_state_stride, const float* vehicle_orientation,
                   const float* target_position, const uint8_t* crashes, const float* actions, const float* prev_actions,
                   float curriculum_progress_fraction, const AgxNavRewardParams* params, float* pos_error, float* pos_error_prev,
                   float* rewards, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_envs == 0) return AGX_OK;
    if (!robot_state || !vehicle_orientation || !target_position || !crashes || !actions || !prev_actions || !params || !pos_error ||
        !pos_error_prev || !rewards)
        return agx_set_error(AGX_E_NULL, "agx_nav_reward: NULL argument");
    if (robot_state_stride < 3) return agx_set_error(AGX_E_INVALID, "agx_nav_reward: robot_state_stride < 3");
    if (((uintptr_t)vehicle_orientation | (uintptr_t)actions | (uintptr_t)prev_actions) & 15)
        return agx_set_error(AGX_E_INVALID, "agx_nav_reward: [N,4] arrays must be 16-byte aligned");
    nav_reward_kernel<<<blocks_for(num_envs), kAuxThreads, 0, (cudaStream_t)stream>>>(
        num_envs, robot_state, robot_state_stride, vehicle_orientation, target_position, crashes, actions, prev_actions,
        curriculum_progress_fraction, *params, pos_error, pos_error_prev, rewards);
    return agx_check_launch("nav_reward_kernel");
}

int agx_nav_obs(int num_envs, const float* robot_state, int robot_state_stride, const float* vehicle_orientation, const float* euler,
                const float* body_linvel, const float* body_angvel, const float* robot_actions, const float* target_position,
                const float* u_vec, const float* u_euler, float* obs, int obs_stride, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_envs == 0) return AGX_OK;
    if (!robot_state || !vehicle_orientation || !euler || !body_linvel || !body_angvel || !robot_actions || !target_position || !u_vec ||
        !u_euler || !obs)
        return agx_set_error(AGX_E_NULL, "agx_nav_obs: NULL argument");
    if (robot_state_stride < 3 || obs_stride < 17) return agx_set_error(AGX_E_INVALID, "agx_nav_obs: stride too small");
    if (((uintptr_t)vehicle_orientation | (uintptr_t)robot_actions) & 15)
        return agx_set_error(AGX_E_INVALID, "agx_nav_obs: [N,4] arrays must be 16-byte aligned");
    nav_obs_kernel<<<blocks_for(num_envs), kAuxThreads, 0, (cudaStream_t)stream>>>(
        num_envs, robot_state, robot_state_stride, vehicle_orientation, euler, body_linvel, body_angvel, robot_actions, target_position,
        u_vec, u_euler, obs, obs_stride);
    return agx_check_launch("nav_obs_kernel");
}

int agx_imu_update(int num_envs, const AgxImuConfig* cfg, const float* force, int force_stride, const float* mass,
                   const float* robot_state, int robot_state_stride, const float* body_angvel, const float* sensor_quats,
                   const float* n_noise, const float* n_bias, float* bias, float* imu_meas, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_envs == 0) return AGX_OK;
    if (!cfg || !force || !mass || !robot_state || !body_angvel || !sensor_quats || !n_noise || !n_bias || !bias || !imu_meas)
        return agx_set_error(AGX_E_NULL, "agx_imu_update: NULL argument");
    if (force_stride < 3 || robot_state_stride < 7) return agx_set_error(AGX_E_INVALID, "agx_imu_update: stride too small");
    if ((uintptr_t)sensor_quats & 15) return agx_set_error(AGX_E_INVALID, "agx_imu_update: sensor_quats must be 16-byte aligned");
    if (!(cfg->sqrt_dt > 0.0f)) return agx_set_error(AGX_E_INVALID, "agx_imu_update: sqrt_dt must be > 0");
    imu_kernel<<<blocks_for(num_envs), kAuxThreads, 0, (cudaStream_t)stream>>>(
        num_envs, *cfg, force, force_stride, mass, robot_state, robot_state_stride, body_angvel, sensor_quats, n_noise, n_bias, bias, imu_meas);
    return agx_check_launch("imu_kernel");
}

}  // extern "C"
