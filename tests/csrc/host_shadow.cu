// TEST INFRASTRUCTURE ONLY -- CPU "shadow" of the per-env device code.
//
// The build container has no GPU.  The headers under aerial_gym_simulator_b200/csrc that hold per-env / per-lane arithmetic
// (AGX_DEV functions) are compiled here a second time as HOST code (-DAGX_HOST_SHADOW) and driven by plain loops that stand in
// for the kernels' thread mapping: envs one after the other, and for the pooling kernel warps and lanes one after the other with
// the kernel's phase order.  tests/test_lidar_nav_shadow.py checks the result against the reference-generated fixtures and the
// oracle, so arithmetic and indexing of a new kernel are verified before it ever reaches a GPU; the -m gpu tests then run the
// real kernels through the C ABI.  Nothing in the product loads this library.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../aerial_gym_simulator_b200/csrc/aux_core.cuh"
#include "../../aerial_gym_simulator_b200/csrc/disturbance_core.cuh"
#include "../../aerial_gym_simulator_b200/csrc/e2e_task_core.cuh"
#include "../../aerial_gym_simulator_b200/csrc/hp1_core.cuh"
#include "../../aerial_gym_simulator_b200/csrc/lidar_nav_core.cuh"
#include "../../aerial_gym_simulator_b200/csrc/noise_core.cuh"
#include "../../aerial_gym_simulator_b200/csrc/obstacle_core.cuh"
#include "../../aerial_gym_simulator_b200/csrc/sim2real_core.cuh"

using namespace agx;

extern "C" {

// stands in for lidar_nav_pool_kernel<VEC4> (same locals, same order of the three phases per band)
int shadow_lidar_nav_pool(int num_envs, int H, int W, int ph, int pw, const float* pc, const float* state, int stride, float max_range,
                          float min_range, float invalid_value, float ttc_max, float* image_ds, float* ttc_out, int force_scalar) {
    const long long band_floats = (long long)ph * W * 3;
    const bool vec4 = !force_scalar && (((uintptr_t)pc & 15) == 0) && (band_floats % 4 == 0) && (((long long)H * W * 3) % 4 == 0);
    const int band_slot = (int)((band_floats + 3) & ~3LL);
    float* smem = nullptr;
    if (posix_memalign((void**)&smem, 16, sizeof(float) * (size_t)kLnavPoolWarps * band_slot)) return -1;
    for (int env = 0; env < num_envs; ++env) {
        float warp_min[kLnavPoolWarps];
        const float* st = state + (size_t)env * stride;
        const V3 pos{st[0], st[1], st[2]}, vel{st[7], st[8], st[9]};
        const int OH = H / ph, OW = W / pw;
        const int nbands = (H + ph - 1) / ph;
        const float* env_pc = pc + (size_t)env * H * W * 3;
        for (int warp = 0; warp < kLnavPoolWarps; ++warp) {
            float* buf = smem + (size_t)warp * band_slot;
            float tmin_lane[32];
            for (int lane = 0; lane < 32; ++lane) tmin_lane[lane] = ttc_max;
            for (int b = warp; b < nbands; b += kLnavPoolWarps) {
                const int rows = (ph < H - b * ph) ? ph : (H - b * ph);
                const int nfl = rows * W * 3, npx = rows * W;
                const float* src = env_pc + (size_t)b * band_floats;
                for (int lane = 0; lane < 32; ++lane) lnav_band_stage(lane, vec4, src, buf, nfl);
                for (int lane = 0; lane < 32; ++lane)
                    tmin_lane[lane] = lnav_band_pixels(lane, buf, npx, pos, vel, max_range, min_range, invalid_value, ttc_max, tmin_lane[lane]);
                if (b < OH)
                    for (int lane = 0; lane < 32; ++lane) lnav_band_pool(lane, buf, W, ph, pw, OW, image_ds + ((size_t)env * OH + b) * OW);
            }
            float tmin = tmin_lane[0];
            for (int lane = 1; lane < 32; ++lane) tmin = fminf(tmin, tmin_lane[lane]);
            warp_min[warp] = tmin;
        }
        float m = warp_min[0];
        for (int w = 1; w < kLnavPoolWarps; ++w) m = fminf(m, warp_min[w]);
        ttc_out[env] = fminf(fmaxf(m, 0.0f), ttc_max);
    }
    free(smem);
    return vec4 ? 1 : 0;
}

void shadow_lidar_nav_reward(int num_envs, const float* state, int stride, const float* veh_q, const float* target, const float* euler,
                             const float* target_yaw, const float* veh_linvel, const float* body_angvel, const uint8_t* crashes,
                             const float* act, const float* prev_act, const float* ttc, float frac, const AgxLidarNavRewardParams* p,
                             float* pos_err, float* pos_err_prev, float* rewards) {
    for (int e = 0; e < num_envs; ++e)
        lnav_reward_env(e, state, stride, veh_q, target, euler, target_yaw, veh_linvel, body_angvel, crashes, act, prev_act, ttc, frac, *p,
                        pos_err, pos_err_prev, rewards);
}

void shadow_lidar_nav_obs(int num_envs, const float* state, int stride, const float* veh_q, const float* euler, const float* blv,
                          const float* bav, const float* actions, const float* target, const float* target_yaw, const float* u_vec,
                          const float* u_euler, const float* lidar_obs, int num_lidar, float* obs, int obs_stride) {
    for (int e = 0; e < num_envs; ++e)
        lnav_obs_env(e, state, stride, veh_q, euler, blv, bav, actions, target, target_yaw, u_vec, u_euler, obs, obs_stride);
    // lidar_nav_obs_copy_kernel: one "thread" per element
    if (lidar_obs && num_lidar > 0) {
        const long long total = (long long)num_envs * num_lidar;
        for (long long i = 0; i < total; ++i) {
            const long long e = i / num_lidar;
            const int k = (int)(i - e * num_lidar);
            obs[(size_t)e * obs_stride + 17 + k] = lidar_obs[i];
        }
    }
}

// stand in for nav_reward_kernel / nav_obs_kernel / imu_kernel (hp1_aux.cu): one "thread" per env
void shadow_nav_reward(int n, const float* state, int stride, const float* veh_q, const float* target, const uint8_t* crashes, const float* act,
                       const float* prev_act, float frac, const AgxNavRewardParams* p, float* pos_err, float* pos_err_prev, float* rewards) {
    for (int e = 0; e < n; ++e) nav_reward_env(e, state, stride, veh_q, target, crashes, act, prev_act, frac, *p, pos_err, pos_err_prev, rewards);
}
void shadow_nav_obs(int n, const float* state, int stride, const float* veh_q, const float* euler, const float* blv, const float* bav,
                    const float* actions, const float* target, const float* u_vec, const float* u_euler, float* obs, int obs_stride) {
    for (int e = 0; e < n; ++e) nav_obs_env(e, state, stride, veh_q, euler, blv, bav, actions, target, u_vec, u_euler, obs, obs_stride);
}
void shadow_imu_update(int n, const AgxImuConfig* c, const float* force, int force_stride, const float* mass, const float* state, int stride,
                       const float* bav, const float* sensor_q, const float* n_noise, const float* n_bias, float* bias, float* meas) {
    for (int e = 0; e < n; ++e) imu_env(e, *c, force, force_stride, mass, state, stride, bav, sensor_q, n_noise, n_bias, bias, meas);
}

// stand in for e2e_reward_kernel / e2e_obs_kernel (e2e_task.cu)
void shadow_e2e_reward(int n, const float* state, int stride, const float* body_angvel, const float* target, const float* act,
                       const float* prev_act, const float* prev_pos_err, const AgxE2ERewardParams* p, uint8_t* crashes, float* rewards) {
    for (int e = 0; e < n; ++e) e2e_reward_env(e, state, stride, body_angvel, target, act, prev_act, prev_pos_err, *p, crashes, rewards);
}
void shadow_e2e_obs(int n, const float* state, int stride, const float* body_angvel, const float* target, const float* noise, float* obs,
                    int obs_stride) {
    for (int e = 0; e < n; ++e) e2e_obs_env(e, state, stride, body_angvel, target, noise, obs, obs_stride);
}

// stand in for s2r_reward_kernel / s2r_obs_kernel (sim2real.cu)
void shadow_s2r_reward(int n, int variant, const float* state, int stride, const float* veh_q, const float* body_linvel, const float* target,
                       const float* prev_dist, const float* act, const float* prev_act, float* act_vehicle_out, uint8_t* crashes, float* rewards) {
    for (int e = 0; e < n; ++e) s2r_reward_env(e, variant, state, stride, veh_q, body_linvel, target, prev_dist, act, prev_act, act_vehicle_out, crashes, rewards);
}
void shadow_s2r_obs(int n, float* state, int stride, const float* body_linvel, const float* body_angvel, const float* robot_actions,
                    const float* target, const float* noise, float* obs, int obs_stride) {
    for (int e = 0; e < n; ++e) s2r_obs_env(e, state, stride, body_linvel, body_angvel, robot_actions, target, noise, obs, obs_stride);
}

// stands in for disturbance_kernel (disturbance.cu)
int shadow_disturbance_draw(int n, int env_id_offset, float prob, const float* max6, uint64_t seed, uint32_t counter, float* out) {
    for (int e = 0; e < n; ++e)
        disturbance_env((uint32_t)env_id_offset + (uint32_t)e, counter, prob, max6, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), out + (size_t)e * 6);
    return 0;
}

// stands in for obstacle_step_kernel: one "thread" per obstacle
void shadow_obstacle_step(int num_envs, int num_assets, float* state, int stride, const float* twist, float dt, int substeps,
                          float lin_damp, float ang_damp) {
    const long long total = (long long)num_envs * num_assets;
    for (long long i = 0; i < total; ++i) obstacle_step_item(i, state, stride, twist, dt, substeps, lin_damp, ang_damp);
}

// stands in for noise_limits_kernel: one "thread" per pixel
void shadow_noise_limits(float* pixels, uint64_t num_pixels, uint64_t first_pixel, const AgxHp2Noise* cfg, uint64_t seed, uint32_t frame) {
    for (uint64_t i = 0; i < num_pixels; ++i) noise_limits_pixel(i, first_pixel + i, pixels, *cfg, frame, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
}

// the device Philox4x32-10 itself (agx_math.cuh), for the Random123 known-answer vectors
void shadow_philox4x32_10(const uint32_t* ctr, uint32_t k0, uint32_t k1, uint32_t* out) {
    const U4 r = philox4x32_10(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, k0, k1);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"

#include "host_shadow_hp1.inc"
