// Small per-env kernels next to the HP1 step: the NavigationTask epilogue
// (reward + observation assembly; the VAE encoder stays in torch) and the IMU sensor.
// One thread per env, plain fp32, the reference's operation order; random draws are INPUTS (the
// reference draws them with torch, the host keeps that call order).  Oracle: oracle/aux_oracle.py.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aerial_gym_b200.h"
#include "agx_common.cuh"
#include "agx_math.cuh"

namespace {
using namespace agx;

constexpr int kAuxThreads = 128;

__device__ __forceinline__ float exp_reward(float mag, float ex, float v) { return mag * expf(-(v * v) * ex); }           // navigation_task.py:420-425
__device__ __forceinline__ float exp_penalty(float mag, float ex, float v) { return mag * (expf(-(v * v) * ex) - 1.0f); }  // :428-433

// NavigationTask.compute_rewards_and_crashes (:397-418) + compute_reward (:436-521)
__global__ void __launch_bounds__(kAuxThreads)
nav_reward_kernel(int N, const float* __restrict__ state, int stride, const float* __restrict__ veh_q, const float* __restrict__ target,
                  const uint8_t* __restrict__ crashes, const float* __restrict__ act, const float* __restrict__ prev_act, float frac,
                  const __grid_constant__ AgxNavRewardParams p, float* __restrict__ pos_err, float* __restrict__ pos_err_prev,
                  float* __restrict__ rewards) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const V3 prev = ld3(pos_err + (size_t)e * 3);  // :404 prev[:] = current
    const float4 q4 = reinterpret_cast<const float4*>(veh_q)[e];
    const V3 x = ld3(state + (size_t)e * stride), tg = ld3(target + (size_t)e * 3);
    const V3 err = quat_rotate_inverse(Q4{q4.x, q4.y, q4.z, q4.w}, tg - x);  // :405-408
    float* pe = pos_err + (size_t)e * 3;
    float* pp = pos_err_prev + (size_t)e * 3;
    pp[0] = prev.x; pp[1] = prev.y; pp[2] = prev.z;
    pe[0] = err.x; pe[1] = err.y; pe[2] = err.z;
    const float mult = 1.0f + 2.0f * frac;
    const float dist = norm3(err), prev_dist = norm3(prev);
    const float pos_reward = exp_reward(p.v[0], p.v[1], dist);
    const float close_reward = exp_reward(p.v[2], p.v[3], dist);
    const float closer = prev_dist - dist;
    const float closer_reward = closer > 0.0f ? p.v[4] * closer : 2.0f * p.v[4] * closer;
    const float dist_reward = (20.0f - dist) / 20.0f;
    const float4 a = reinterpret_cast<const float4*>(act)[e], b = reinterpret_cast<const float4*>(prev_act)[e];
    const float dx = a.x - b.x, dz = a.z - b.z, dw = a.w - b.w;
    const float diff_pen = exp_penalty(p.v[5], p.v[6], dx) + exp_penalty(p.v[7], p.v[8], dz) + exp_penalty(p.v[9], p.v[10], dw);
    const float abs_pen = frac * exp_penalty(p.v[11], p.v[12], a.x) + frac * exp_penalty(p.v[13], p.v[14], a.z) +
                          frac * exp_penalty(p.v[15], p.v[16], a.w);
    float r = mult * (pos_reward + close_reward + closer_reward + dist_reward) + (diff_pen + abs_pen);
    if (crashes[e]) r = p.v[17];
    rewards[e] = r;
}

// NavigationTask.process_obs_for_task (:369-395), columns 0..16
__global__ void __launch_bounds__(kAuxThreads)
nav_obs_kernel(int N, const float* __restrict__ state, int stride, const float* __restrict__ veh_q, const float* __restrict__ euler,
               const float* __restrict__ blv, const float* __restrict__ bav, const float* __restrict__ actions,
               const float* __restrict__ target, const float* __restrict__ u_vec, const float* __restrict__ u_euler,
               float* __restrict__ obs, int obs_stride) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const float4 q4 = reinterpret_cast<const float4*>(veh_q)[e];
    const V3 x = ld3(state + (size_t)e * stride), tg = ld3(target + (size_t)e * 3);
    const V3 vec = quat_rotate_inverse(Q4{q4.x, q4.y, q4.z, q4.w}, tg - x);
    const V3 uv = ld3(u_vec + (size_t)e * 3);
    // :374  vec + 0.1 * 2 * rand_like(vec - 0.5): the -0.5 is inside rand_like, so the noise is 0.2 * U[0,1)
    const V3 pert{vec.x + 0.1f * 2.0f * uv.x, vec.y + 0.1f * 2.0f * uv.y, vec.z + 0.1f * 2.0f * uv.z};
    const float dist = norm3(vec);
    float* o = obs + (size_t)e * obs_stride;
    o[0] = pert.x / dist; o[1] = pert.y / dist; o[2] = pert.z / dist;
    o[3] = dist;
    const V3 eu = ld3(euler + (size_t)e * 3), ue = ld3(u_euler + (size_t)e * 3);
    // ssa (utils/math.py:150-152): remainder(a + pi, 2 pi) - pi, python-style remainder (result in [0, 2 pi))
    auto ssa = [](float a) {
        float t = a + AGX_PI_F;
        float r = fmodf(t, AGX_TWO_PI_F);
        if (r < 0.0f) r += AGX_TWO_PI_F;
        return r - AGX_PI_F;
    };
    o[4] = ssa(eu.x) + 0.1f * (ue.x - 0.5f);
    o[5] = ssa(eu.y) + 0.1f * (ue.y - 0.5f);
    o[6] = 0.0f;
    const V3 lv = ld3(blv + (size_t)e * 3), av = ld3(bav + (size_t)e * 3);
    o[7] = lv.x; o[8] = lv.y; o[9] = lv.z;
    o[10] = av.x; o[11] = av.y; o[12] = av.z;
    const float4 a = reinterpret_cast<const float4*>(actions)[e];
    o[13] = a.x; o[14] = a.y; o[15] = a.z; o[16] = a.w;
}

// IMUSensor.update (sensors/imu_sensor.py:85-131)
__global__ void __launch_bounds__(kAuxThreads)
imu_kernel(int N, const __grid_constant__ AgxImuConfig c, const float* __restrict__ force, int force_stride, const float* __restrict__ mass,
           const float* __restrict__ state, int stride, const float* __restrict__ bav, const float* __restrict__ sensor_q,
           const float* __restrict__ n_noise, const float* __restrict__ n_bias, float* __restrict__ bias, float* __restrict__ meas) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const float m = mass[e];
    const float* f = force + (size_t)e * force_stride;
    const V3 accel_t{f[0] / m, f[1] / m, f[2] / m};  // :89
    const float* st = state + (size_t)e * stride;
    const Q4 rq{st[3], st[4], st[5], st[6]};
    const float4 s4 = reinterpret_cast<const float4*>(sensor_q)[e];
    const Q4 sq{s4.x, s4.y, s4.z, s4.w};
    const Q4 q_ws = quat_mul(rq, sq);
    const V3 g{c.g_world[0], c.g_world[1], c.g_world[2]};
    const V3 w_body = ld3(bav + (size_t)e * 3);
    V3 acc, rate;
    if (c.world_frame) {  // :90-98
        acc = quat_rotate_inverse(q_ws, accel_t - g);
        rate = quat_rotate_inverse(q_ws, w_body);
    } else {  // :99-106
        acc = quat_rotate_inverse(sq, accel_t) - quat_rotate_inverse(q_ws, g);
        rate = quat_rotate_inverse(sq, w_body);
    }
    const float a6[6] = {acc.x, acc.y, acc.z, rate.x, rate.y, rate.z};
    const float eb = (float)c.enable_bias, en = (float)c.enable_noise;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float noise = n_noise[(size_t)e * 6 + k] * c.noise_std[k] / c.sqrt_dt;                  // :74-77
        const float b = bias[(size_t)e * 6 + k] + n_bias[(size_t)e * 6 + k] * c.bias_std[k] * c.sqrt_dt;  // :79-83
        bias[(size_t)e * 6 + k] = b;
        float v = a6[k] + eb * b + en * noise;                                                         // :110-117
        v = fmaxf(fminf(v, c.max_meas[k]), -c.max_meas[k]);                                            // tensor_clamp :119-128
        meas[(size_t)e * 6 + k] = v;
    }
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
