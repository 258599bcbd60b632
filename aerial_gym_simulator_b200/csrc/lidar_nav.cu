// LiDARNavigationTask epilogue behind the C ABI (SURVEY 8f "next"):
//   agx_lidar_nav_pool    point cloud -> clipped ranges -> min-pool + time to collision   (HBM-bound: 12 B read per return)
//   agx_lidar_nav_reward  compute_rewards_and_crashes + compute_reward, one thread per env
//   agx_lidar_nav_obs     process_obs_for_task, one thread per env
// All arithmetic lives in lidar_nav_core.cuh (shared, text for text, with the CPU shadow build of the test-suite);
// the kernels below only map threads to envs / pixels.  Oracle: oracle/lidar_nav_oracle.py.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aerial_gym_b200.h"
#include "agx_common.cuh"
#include "lidar_nav_core.cuh"

namespace {
using namespace agx;

constexpr int kEnvThreads = 128;
constexpr int kPoolWarps = kLnavPoolWarps;
constexpr int kPoolThreads = kPoolWarps * 32;

// One CTA per env, one warp per band of pool_h image rows.  A band (pool_h * W returns, 12 B each; 4320 B for the
// reference's 3 x 120) is staged into the warp's slice of shared memory with coalesced 16-byte loads (streaming: every
// byte is used once), each lane then turns returns lane, lane+32, ... into (clipped range, ttc) -- shared-memory reads at
// a 3-float stride are bank-conflict free -- stores the range back over the return's x slot, and the first W/pool_w
// lanes take the window minima.  The env's time to collision is a warp-shuffle + shared-memory minimum.
template <bool VEC4>
__global__ void __launch_bounds__(kPoolThreads)
lidar_nav_pool_kernel(int H, int W, int ph, int pw, const float* __restrict__ pc, const float* __restrict__ state, int stride,
                      float max_range, float min_range, float invalid_value, float ttc_max, float* __restrict__ image_ds,
                      float* __restrict__ ttc_out) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float warp_min[kPoolWarps];
    const int env = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int band_floats = ph * W * 3;
    const int band_slot = (band_floats + 3) & ~3;  // keep every warp's slice 16-byte aligned
    float* buf = smem + (size_t)warp * band_slot;
    const float* st = state + (size_t)env * stride;
    const V3 pos{st[0], st[1], st[2]}, vel{st[7], st[8], st[9]};
    const int OH = H / ph, OW = W / pw;       // max_pool2d floors: trailing rows / columns are not pooled ...
    const int nbands = (H + ph - 1) / ph;     // ... but every return counts for the time to collision
    const float* env_pc = pc + (size_t)env * H * W * 3;
    float tmin = ttc_max;                     // min(.., ttc_max) == the final clamp's upper bound
    for (int b = warp; b < nbands; b += kPoolWarps) {
        const int rows = min(ph, H - b * ph);
        const int nfl = rows * W * 3, npx = rows * W;
        const float* src = env_pc + (size_t)b * band_floats;
        lnav_band_stage(lane, VEC4, src, buf, nfl);
        __syncwarp();
        tmin = lnav_band_pixels(lane, buf, npx, pos, vel, max_range, min_range, invalid_value, ttc_max, tmin);
        __syncwarp();
        if (b < OH) lnav_band_pool(lane, buf, W, ph, pw, OW, image_ds + ((size_t)env * OH + b) * OW);
        __syncwarp();  // the next band overwrites buf
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, off));
    if (lane == 0) warp_min[warp] = tmin;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = warp_min[0];
#pragma unroll
        for (int w = 1; w < kPoolWarps; ++w) m = fminf(m, warp_min[w]);
        ttc_out[env] = fminf(fmaxf(m, 0.0f), ttc_max);  // :339 clamp(min, 0, 10)
    }
}

__global__ void __launch_bounds__(kEnvThreads)
lidar_nav_reward_kernel(int N, const float* __restrict__ state, int stride, const float* __restrict__ veh_q, const float* __restrict__ target,
                        const float* __restrict__ euler, const float* __restrict__ target_yaw, const float* __restrict__ veh_linvel,
                        const float* __restrict__ body_angvel, const uint8_t* __restrict__ crashes, const float* __restrict__ act,
                        const float* __restrict__ prev_act, const float* __restrict__ ttc, float frac,
                        const __grid_constant__ AgxLidarNavRewardParams p, float* __restrict__ pos_err, float* __restrict__ pos_err_prev,
                        float* __restrict__ rewards) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    lnav_reward_env(e, state, stride, veh_q, target, euler, target_yaw, veh_linvel, body_angvel, crashes, act, prev_act, ttc, frac, p,
                    pos_err, pos_err_prev, rewards);
}

__global__ void __launch_bounds__(kEnvThreads)
lidar_nav_obs_kernel(int N, const float* __restrict__ state, int stride, const float* __restrict__ veh_q, const float* __restrict__ euler,
                     const float* __restrict__ blv, const float* __restrict__ bav, const float* __restrict__ actions,
                     const float* __restrict__ target, const float* __restrict__ target_yaw, const float* __restrict__ u_vec,
                     const float* __restrict__ u_euler, float* __restrict__ obs, int obs_stride) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    lnav_obs_env(e, state, stride, veh_q, euler, blv, bav, actions, target, target_yaw, u_vec, u_euler, obs, obs_stride);
}

// obs[:, 17:17+L] = lidar_obs: one thread per element, coalesced on both sides (rows of the two arrays differ in stride only)
__global__ void __launch_bounds__(kEnvThreads)
lidar_nav_obs_copy_kernel(long long total, int L, const float* __restrict__ lidar_obs, float* __restrict__ obs, int obs_stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long e = i / L;
    const int k = (int)(i - e * L);
    obs[(size_t)e * obs_stride + 17 + k] = lidar_obs[i];
}

inline int blocks_for(long long n) { return (int)((n + kEnvThreads - 1) / kEnvThreads); }

}  // namespace

extern "C" {

int agx_lidar_nav_pool(int num_envs, int height, int width, int pool_h, int pool_w, const float* pointcloud, const float* robot_state,
                       int robot_state_stride, float max_range, float min_range, float invalid_value, float ttc_max, float* image_ds,
                       float* time_to_collision, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (height < 1 || width < 1 || pool_h < 1 || pool_w < 1 || pool_h > height || pool_w > width)
        return agx_set_error(AGX_E_INVALID, "agx_lidar_nav_pool: need 1 <= pool_h <= height and 1 <= pool_w <= width");
    if (robot_state_stride < 10) return agx_set_error(AGX_E_INVALID, "agx_lidar_nav_pool: robot_state_stride < 10 (position 0..2, linvel 7..9)");
    if (num_envs == 0) return AGX_OK;
    if (!pointcloud || !robot_state || !image_ds || !time_to_collision) return agx_set_error(AGX_E_NULL, "agx_lidar_nav_pool: NULL argument");
    const long long band_floats = (long long)pool_h * width * 3;
    const long long smem_bytes = (long long)kPoolWarps * ((band_floats + 3) & ~3LL) * (long long)sizeof(float);
    if (smem_bytes > 200 * 1024)
        return agx_set_error(AGX_E_INVALID, "agx_lidar_nav_pool: a band of pool_h x width returns (%lld B x %d warps) exceeds shared memory",
                             band_floats * 4, kPoolWarps);
    // 16-byte loads need every band to start on a 16-byte boundary
    const bool vec4 = (((uintptr_t)pointcloud & 15) == 0) && (band_floats % 4 == 0) && (((long long)height * width * 3) % 4 == 0);
    auto kern = vec4 ? lidar_nav_pool_kernel<true> : lidar_nav_pool_kernel<false>;
    if (smem_bytes > 48 * 1024) {
        int rc = agx_check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes), "cudaFuncSetAttribute");
        if (rc) return rc;
    }
    kern<<<num_envs, kPoolThreads, (size_t)smem_bytes, (cudaStream_t)stream>>>(height, width, pool_h, pool_w, pointcloud, robot_state,
                                                                             robot_state_stride, max_range, min_range, invalid_value,
                                                                             ttc_max, image_ds, time_to_collision);
    return agx_check_launch("lidar_nav_pool_kernel");
}

int agx_lidar_nav_reward(int num_envs, const float* robot_state, int robot_state_stride, const float* vehicle_orientation,
                         const float* target_position, const float* euler, const float* target_yaw, const float* vehicle_linvel,
                         const float* body_angvel, const uint8_t* crashes, const float* actions, const float* prev_actions,
                         const float* time_to_collision, float curriculum_progress_fraction, const AgxLidarNavRewardParams* params,
                         float* pos_error, float* pos_error_prev, float* rewards, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_envs == 0) return AGX_OK;
    if (!robot_state || !vehicle_orientation || !target_position || !euler || !target_yaw || !vehicle_linvel || !body_angvel || !crashes ||
        !actions || !prev_actions || !time_to_collision || !params || !pos_error || !pos_error_prev || !rewards)
        return agx_set_error(AGX_E_NULL, "agx_lidar_nav_reward: NULL argument");
    if (robot_state_stride < 3) return agx_set_error(AGX_E_INVALID, "agx_lidar_nav_reward: robot_state_stride < 3");
    lidar_nav_reward_kernel<<<blocks_for(num_envs), kEnvThreads, 0, (cudaStream_t)stream>>>(
        num_envs, robot_state, robot_state_stride, vehicle_orientation, target_position, euler, target_yaw, vehicle_linvel, body_angvel,
        crashes, actions, prev_actions, time_to_collision, curriculum_progress_fraction, *params, pos_error, pos_error_prev, rewards);
    return agx_check_launch("lidar_nav_reward_kernel");
}

int agx_lidar_nav_obs(int num_envs, const float* robot_state, int robot_state_stride, const float* vehicle_orientation, const float* euler,
                      const float* body_linvel, const float* body_angvel, const float* robot_actions, const float* target_position,
                      const float* target_yaw, const float* u_vec, const float* u_euler, const float* lidar_obs, int num_lidar, float* obs,
                      int obs_stride, void* stream) {
    if (num_envs < 0) return agx_set_error(AGX_E_INVALID, "num_envs < 0");
    if (num_lidar < 0) return agx_set_error(AGX_E_INVALID, "agx_lidar_nav_obs: num_lidar < 0");
    if (num_envs == 0) return AGX_OK;
    if (!robot_state || !vehicle_orientation || !euler || !body_linvel || !body_angvel || !robot_actions || !target_position ||
        !target_yaw || !u_vec || !u_euler || !obs)
        return agx_set_error(AGX_E_NULL, "agx_lidar_nav_obs: NULL argument");
    if (!lidar_obs) num_lidar = 0;
    if (robot_state_stride < 3 || obs_stride < 17 + num_lidar) return agx_set_error(AGX_E_INVALID, "agx_lidar_nav_obs: stride too small");
    lidar_nav_obs_kernel<<<blocks_for(num_envs), kEnvThreads, 0, (cudaStream_t)stream>>>(
        num_envs, robot_state, robot_state_stride, vehicle_orientation, euler, body_linvel, body_angvel, robot_actions, target_position,
        target_yaw, u_vec, u_euler, obs, obs_stride);
    int rc = agx_check_launch("lidar_nav_obs_kernel");
    if (rc || num_lidar == 0) return rc;
    const long long total = (long long)num_envs * num_lidar;
    lidar_nav_obs_copy_kernel<<<blocks_for(total), kEnvThreads, 0, (cudaStream_t)stream>>>(total, num_lidar, lidar_obs, obs, obs_stride);
    return agx_check_launch("lidar_nav_obs_copy_kernel");
}

}  // extern "C"
