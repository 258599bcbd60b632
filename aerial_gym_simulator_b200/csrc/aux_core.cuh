// Per-env arithmetic of the NavigationTask epilogue (reward, observation) and the IMU, moved out of
// hp1_aux.cu verbatim: AGX_DEV = device-only in the product (SASS of hp1_aux.cu unchanged by the move), host+device in the CPU
// shadow build of the test-suite (tests/csrc/host_shadow.cu).
#pragma once
#include "../../include/aerial_gym_b200.h"
#include "agx_math.cuh"

namespace agx {

AGX_DEV float exp_reward(float mag, float ex, float v) { return mag * expf(-(v * v) * ex); }           // navigation_task.py:420-425
AGX_DEV float exp_penalty(float mag, float ex, float v) { return mag * (expf(-(v * v) * ex) - 1.0f); }  // :428-433

AGX_DEV void nav_reward_env(int e, const float* state, int stride, const float* veh_q, const float* target, const uint8_t* crashes, const float* act, const float* prev_act, float frac, const AgxNavRewardParams& p, float* pos_err, float* pos_err_prev, float* rewards) {
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

AGX_DEV void nav_obs_env(int e, const float* state, int stride, const float* veh_q, const float* euler, const float* blv, const float* bav, const float* actions, const float* target, const float* u_vec, const float* u_euler, float* obs, int obs_stride) {
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

AGX_DEV void imu_env(int e, const AgxImuConfig& c, const float* force, int force_stride, const float* mass, const float* state, int stride, const float* bav, const float* sensor_q, const float* n_noise, const float* n_bias, float* bias, float* meas) {
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

}  // namespace agx
