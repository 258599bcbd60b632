// Per-env arithmetic of the two motor-command ("end to end") position tasks of the reference,
//   task/position_setpoint_task_sim2real_end_to_end/position_setpoint_task_sim2real_end_to_end.py   (tinyprop, 4096 envs)
//   task/position_setpoint_task_sim2real_px4/position_setpoint_task_sim2real_px4.py                 (x500)
// which differ only in reward constants (AgxE2ERewardParams): compute_rewards_and_crashes + compute_reward (:232-311) and
// process_obs_for_task (:204-229).  The observation goes through pytorch3d (quaternion_to_matrix, matrix_to_euler_angles "ZYX",
// euler_angles_to_matrix "ZYX", matrix_to_rotation_6d) -- not in the reference tree; restated from its published algorithms,
// parity against pytorch3d itself unpinned.  AGX_DEV: device-only in the product, host+device in the CPU shadow build.
// Oracle: oracle/e2e_task_oracle.py.
#pragma once
#include "../../include/aerial_gym_b200.h"
#include "agx_math.cuh"

namespace agx {

AGX_DEV float e2e_exp_func(float x, float gain, float ex) { return gain * expf(((-ex) * x) * x); }              // :254-257  gain * exp(-exp * x * x)
AGX_DEV float e2e_exp_penalty(float x, float gain, float ex) { return gain * (expf(((-ex) * x) * x) - 1.0f); }  // :260-263

// compute_rewards_and_crashes (:232-251) + compute_reward (:267-311) for env e
AGX_DEV void e2e_reward_env(int e, const float* state, int stride, const float* body_angvel, const float* target, const float* act,
                            const float* prev_act, const float* prev_pos_err, const AgxE2ERewardParams& p, uint8_t* crashes, float* rewards) {
    const size_t e3 = (size_t)e * 3, e4 = (size_t)e * 4;
    const float* st = state + (size_t)e * stride;
    const V3 x = ld3(st), v = ld3(st + 7), w = ld3(body_angvel + e3);
    const Q4 q{st[3], st[4], st[5], st[6]};
    const V3 tg = target ? ld3(target + e3) : V3{0.0f, 0.0f, 0.0f};
    V3 pe = tg - x;                                                    // :240
    const float dist = norm3(pe);                                      // :279
    const float prev_dist = norm3(ld3(prev_pos_err + e3));             // :281
    pe.z = pe.z * p.z_error_scale;                                     // :283 (after the distance was taken)
    const float pos_reward = (e2e_exp_func(pe.x, 10.0f, 10.0f) + e2e_exp_func(pe.y, 10.0f, 10.0f) + e2e_exp_func(pe.z, 10.0f, 10.0f)) +
                             (e2e_exp_func(pe.x, 2.0f, 2.0f) + e2e_exp_func(pe.y, 2.0f, 2.0f) + e2e_exp_func(pe.z, 2.0f, 2.0f));  // :284
    const V3 ups = quat_rotate(q, V3{0.0f, 0.0f, 1.0f});               // :286 quat_axis(q, 2)
    const float tilt = 1.0f - ups.z;
    const float upright = e2e_exp_func(tilt, 2.5f, 5.0f) + e2e_exp_func(tilt, p.upright_gain2, p.upright_exp2);  // :288 (px4: two terms)
    const V3 forw = quat_rotate(q, V3{1.0f, 0.0f, 0.0f});              // :290
    const float al = 1.0f - forw.x;
    const float align = e2e_exp_func(al, p.align_gain1, p.align_exp1) + e2e_exp_func(al, p.align_gain2, p.align_exp2);  // :292
    const float angvel_r = e2e_exp_func(w.x, p.angvel_gain, 10.0f) + e2e_exp_func(w.y, p.angvel_gain, 10.0f) + e2e_exp_func(w.z, p.angvel_gain, 10.0f);  // :294
    const float vel_r = e2e_exp_func(v.x, 1.0f, 5.0f) + e2e_exp_func(v.y, 1.0f, 5.0f) + e2e_exp_func(v.z, 1.0f, 5.0f);  // :295
    float action_cost = 0.0f, diff_pen = 0.0f;
    for (int i = 0; i < 4; ++i) {
        const float a = act[e4 + i];
        action_cost += e2e_exp_penalty(a - p.hover_thrust, 0.01f, 10.0f);           // :297-298
        diff_pen += e2e_exp_penalty(a - prev_act[e4 + i], p.action_diff_gain, 6.0f);  // :303-304
    }
    const float closer = prev_dist - dist;                                            // :300
    const float towards = (closer >= 0.0f) ? p.towards_gain_pos * closer : p.towards_gain_neg * closer;  // :301
    rewards[e] = towards + (pos_reward * (align + vel_r + angvel_r + diff_pen) + (angvel_r + vel_r + upright + pos_reward + action_cost)) / 100.0f;  // :306
    if (dist > p.crash_dist) crashes[e] = 1;                                          // :308 (no crash penalty on the reward)
}

// process_obs_for_task (:204-229) for env e.  noise [N,12]: the four torch.normal draws (position, orientation, linear velocity, body
// rates; 3 columns each) in that order.  Columns 0..14 of obs.
AGX_DEV void e2e_obs_env(int e, const float* state, int stride, const float* body_angvel, const float* target, const float* noise, float* obs,
                         int obs_stride) {
    const size_t e3 = (size_t)e * 3;
    const float* st = state + (size_t)e * stride;
    const float* nz = noise + (size_t)e * 12;
    const V3 tg = target ? ld3(target + e3) : V3{0.0f, 0.0f, 0.0f};
    float* o = obs + (size_t)e * obs_stride;
    o[0] = (tg.x - st[0]) + nz[0]; o[1] = (tg.y - st[1]) + nz[1]; o[2] = (tg.z - st[2]) + nz[2];  // :208
    // pytorch3d.quaternion_to_matrix on (w, x, y, z) = (st[6], st[3], st[4], st[5])  (:211-212)
    const float r = st[6], i = st[3], j = st[4], k = st[5];
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    const float R00 = 1.0f - two_s * (j * j + k * k), R10 = two_s * (i * j + k * r), R20 = two_s * (i * k - j * r);
    const float R21 = two_s * (j * k + i * r), R22 = 1.0f - two_s * (i * i + j * j);
    // pytorch3d.matrix_to_euler_angles(M, "ZYX") = (atan2(R10, R00), asin(-R20), atan2(R21, R22)); [:, [2, 1, 0]] -> (roll, pitch, yaw)
    const float roll = atan2f(R21, R22) + nz[3], pitch = asinf(-R20) + nz[4], yaw = atan2f(R10, R00) + nz[5];  // :213 (+ or_noise)
    // pytorch3d.euler_angles_to_matrix((yaw, pitch, roll), "ZYX") = Rz(yaw) Ry(pitch) Rx(roll); rotation_6d = its first two rows  (:223-224)
    float sy, cy, sp, cp, sr, cr;
    sincos_(yaw, &sy, &cy);
    sincos_(pitch, &sp, &cp);
    sincos_(roll, &sr, &cr);
    o[3] = cy * cp; o[4] = cy * sp * sr - sy * cr; o[5] = cy * sp * cr + sy * sr;
    o[6] = sy * cp; o[7] = sy * sp * sr + cy * cr; o[8] = sy * sp * cr - cy * sr;
    o[9] = st[7] + nz[6]; o[10] = st[8] + nz[7]; o[11] = st[9] + nz[8];                           // :216, :225 (WORLD linear velocity)
    o[12] = body_angvel[e3] + nz[9]; o[13] = body_angvel[e3 + 1] + nz[10]; o[14] = body_angvel[e3 + 2] + nz[11];  // :219, :226
}

}  // namespace agx
