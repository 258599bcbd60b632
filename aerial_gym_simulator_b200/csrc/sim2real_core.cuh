// Per-env arithmetic of the reference's two setpoint-command sim2real position tasks (lmf2, 16-env deployments),
//   task/position_setpoint_task_sim2real/position_setpoint_task_sim2real.py                          (velocity commands;  variant 0)
//   task/position_setpoint_task_acceleration_sim2real/position_setpoint_task_acceleration_sim2real.py (acceleration commands; variant 1)
// compute_rewards_and_crashes + compute_reward (:230-339 | :239-356) and process_obs_for_task (:202-228 | :211-237, identical text).
// AGX_DEV: device-only in the product, host+device in the CPU shadow build.  Oracle: oracle/sim2real_oracle.py.
#pragma once
#include "../../include/aerial_gym_b200.h"
#include "agx_math.cuh"

namespace agx {

AGX_DEV float s2r_exp(float x, float gain, float ex) { return gain * expf(((-ex) * x) * x); }                     // exp_func
AGX_DEV float s2r_abs_exp(float x, float gain, float ex) { return gain * expf((-ex) * fabsf(x)); }                // abs_exp_func
AGX_DEV float s2r_abs_exp_pen(float x, float gain, float ex) { return gain * (expf((-ex) * fabsf(x)) - 1.0f); }   // abs_exp_penalty_func

// reward + crash flag of env e.  act / prev_act [N,4]: variant 0 = the task's actions / prev_actions; variant 1 = the current actions
// (rotated HERE by the vehicle orientation into actions_vehicle_frame, :254-257, also stored) / prev_actions_vehicle_frame.
AGX_DEV void s2r_reward_env(int e, int variant, const float* state, int stride, const float* veh_q, const float* body_linvel,
                            const float* target, const float* prev_dist, const float* act, const float* prev_act, float* act_vehicle_out,
                            uint8_t* crashes, float* rewards) {
    const size_t e3 = (size_t)e * 3, e4 = (size_t)e * 4;
    const float* st = state + (size_t)e * stride;
    const Q4 q{st[3], st[4], st[5], st[6]}, qv{veh_q[e4], veh_q[e4 + 1], veh_q[e4 + 2], veh_q[e4 + 3]};
    const V3 tg = target ? ld3(target + e3) : V3{0.0f, 0.0f, 0.0f};
    const V3 d = tg - ld3(st);
    // quat_apply_inverse = quat_apply(conj(q), .): velocity task in the (stale) vehicle frame (:241-243), acceleration task in the body frame (:250-252)
    const V3 pe = quat_apply(quat_conj(variant ? q : qv), d);
    const float yaw_error = 0.0f - ssa_0_2pi(euler_xyz_0_2pi(q).z);     // :245-246 from the CURRENT orientation
    float a[4] = {act[e4], act[e4 + 1], act[e4 + 2], act[e4 + 3]};
    if (variant) {                                                     // :254-257
        const V3 r = quat_rotate(qv, V3{a[0], a[1], a[2]});
        a[0] = r.x; a[1] = r.y; a[2] = r.z;
        if (act_vehicle_out) { act_vehicle_out[e4] = a[0]; act_vehicle_out[e4 + 1] = a[1]; act_vehicle_out[e4 + 2] = a[2]; act_vehicle_out[e4 + 3] = a[3]; }
    }
    const float dist = norm3(pe);
    const float pos_reward = s2r_exp(dist, 2.0f, 1.0f) + s2r_exp(dist, 3.0f, 10.0f) + s2r_abs_exp(dist, 3.0f, 50.0f);
    const float speed = norm3(ld3(body_linvel + e3));
    const float pd = prev_dist[e];
    float ap = 0.0f, adp = 0.0f;
    for (int i = 0; i < 4; ++i) {
        ap += s2r_abs_exp_pen(a[i], variant ? 0.3f : 0.2f, 4.0f);
        adp += s2r_abs_exp_pen(a[i] - prev_act[e4 + i], variant ? 0.4f : 0.3f, 6.0f);
    }
    float total;
    if (!variant) {  // position_setpoint_task_sim2real.py:299-334
        const float speed_reward = s2r_exp(speed, 1.0f, 3.0f);
        const float dist_reward = (20.0f - dist) / 40.0f;
        const float closer = 400.0f * (pd - dist);
        const float yaw_r = s2r_abs_exp(yaw_error, 2.0f, 3.0f);
        total = (pos_reward + dist_reward + pos_reward * (speed_reward + ap + closer / 10.0f)) + ap + adp + closer + yaw_r;
    } else {         // position_setpoint_task_acceleration_sim2real.py:313-351
        const float close_pos = s2r_exp(dist, 2.0f, 1.0f);
        const float speed_reward = s2r_exp(speed, 2.0f, 2.5f);
        const float closer = (dist < pd) ? 400.0f * (pd - dist) : 1200.0f * (pd - dist);
        const float yaw_r = s2r_abs_exp(yaw_error, 3.0f, 5.0f);
        total = (pos_reward + pos_reward * (closer / 9.0f + ap / 3.0f + speed_reward / 1.5f)) + ap + adp + closer + yaw_r + close_pos + speed_reward * 0.2f;
    }
    total = 1.0f * total;                       // curriculum_level_multiplier = 1.0
    uint8_t c = crashes[e];
    if (dist > 10.0f) c = 1;
    if (c) total = -50.0f;
    crashes[e] = c;
    rewards[e] = total;
}

// process_obs_for_task for env e.  noise [N,12] = the four torch.randn_like draws (euler, position, body linvel, body angvel) in the
// method's order, UNSCALED.  Side effect kept: robot_orientation is multiplied by sign(qw) in place (:204-207).
AGX_DEV void s2r_obs_env(int e, float* state, int stride, const float* body_linvel, const float* body_angvel, const float* robot_actions,
                         const float* target, const float* noise, float* obs, int obs_stride) {
    const size_t e3 = (size_t)e * 3, e4 = (size_t)e * 4;
    float* st = state + (size_t)e * stride;
    const float* nz = noise + (size_t)e * 12;
    const V3 tg = target ? ld3(target + e3) : V3{0.0f, 0.0f, 0.0f};
    const V3 pe = tg - ld3(st);                                           // :203
    const float sgn = (st[6] > 0.0f) ? 1.0f : ((st[6] < 0.0f) ? -1.0f : 0.0f);  // torch.sign
    st[3] = sgn * st[3]; st[4] = sgn * st[4]; st[5] = sgn * st[5]; st[6] = sgn * st[6];
    const V3 eu = euler_xyz_0_2pi(Q4{st[3], st[4], st[5], st[6]});
    const float r = ssa_0_2pi(eu.x) + nz[0] * 0.02f, p = ssa_0_2pi(eu.y) + nz[1] * 0.02f, y = ssa_0_2pi(eu.z) + nz[2] * 0.02f;  // :209-210
    float* o = obs + (size_t)e * obs_stride;
    o[0] = pe.x + nz[3] * 0.03f; o[1] = pe.y + nz[4] * 0.03f; o[2] = pe.z + nz[5] * 0.03f;   // :211-213
    const Q4 qn = quat_from_euler(r, p, y);                                // :215
    o[3] = qn.x; o[4] = qn.y; o[5] = qn.z; o[6] = qn.w;
    o[7] = body_linvel[e3] + nz[6] * 0.02f; o[8] = body_linvel[e3 + 1] + nz[7] * 0.02f; o[9] = body_linvel[e3 + 2] + nz[8] * 0.02f;       // :216-219
    o[10] = body_angvel[e3] + nz[9] * 0.02f; o[11] = body_angvel[e3 + 1] + nz[10] * 0.02f; o[12] = body_angvel[e3 + 2] + nz[11] * 0.02f;  // :220-223
    o[13] = robot_actions[e4]; o[14] = robot_actions[e4 + 1]; o[15] = robot_actions[e4 + 2]; o[16] = robot_actions[e4 + 3];              // :224
}

}  // namespace agx
