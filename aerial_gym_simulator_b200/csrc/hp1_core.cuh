// Per-env arithmetic of the HP1 step: derived states, Lee controllers, motor model, rigid-body
// integrator.  Moved out of hp1.cu verbatim so the SAME TEXT can also be compiled for the host: AGX_DEV is device-only inside
// libaerial_gym_b200.so (the SASS of hp1.cu is unchanged by the move -- checked with cuobjdump) and host+device in the CPU shadow
// build of the test-suite (tests/csrc/host_shadow.cu), which steps these functions against the reference-generated fixtures
// without a GPU.
#pragma once
#include "../../include/aerial_gym_b200.h"
#include "agx_math.cuh"

// Loops over the motors inside one physics sub-step (hp1.cu, between the AGX_SHADOW markers): fully unrolled by default.
// -DAGX_HP1_ROLL_MOTORS keeps them rolled -- a round-2 experiment for the instruction-fetch side of the latency-bound step
// (tools/build_variant.py); the arithmetic is the same either way.
#ifdef AGX_HP1_ROLL_MOTORS
#define AGX_HP1_MOTOR_UNROLL _Pragma("unroll 1")
#else
#define AGX_HP1_MOTOR_UNROLL _Pragma("unroll")
#endif

namespace agx {

struct EnvState {
    V3 x;
    Q4 q;
    V3 v;
    V3 w;
};
struct Derived {
    V3 euler;
    Q4 qveh;
    V3 vveh;
    V3 vb;
    V3 wb;
};
struct Gains {
    V3 kp, kv, kr, kw;
};

AGX_DEV V3 ld3c(const float* c) { return V3{c[0], c[1], c[2]}; }

// Compile-time specialisation of the step's RUN-TIME switches.  The generic kernel reads the controller id and the motor-model /
// integrator flags from the launch constants and branches (uniformly) over all eight controller families: ~6,600 SASS instructions of
// which a warp executes ~2,000, spread over 105 KB of text -- more than the 32 KB L1.5 instruction cache, and every warp streams
// through it exactly once (ncu: stall_no_instruction leads).  HpSpec<S> with S >= 0 fixes controller (bits 0..3) and the flags
// USE_RPS | MOTOR_RK4 | DISCRETE_MIX | GYROSCOPIC (bits 4..7) at compile time, so an instantiation contains only the code its robot
// runs, laid out contiguously; S = -1 keeps every switch dynamic.  Same arithmetic either way (the shadow build and the generic
// kernel are the S = -1 text); hp1.cu picks the instantiation that matches cfg, falling back to the generic one.
template <int S>
struct HpSpec {
    static constexpr bool kFixed = S >= 0;
    AGX_DEV static int controller(const AgxHp1Config& cfg) {
        if constexpr (kFixed) return S & 0xf;
        else return cfg.controller;
    }
    template <int F>
    AGX_DEV static bool flag(const AgxHp1Config& cfg) {
        static_assert((F & ~0xf) == 0, "only the four low flags are specialised");
        if constexpr (kFixed) return ((S >> 4) & F) != 0;
        else return (cfg.flags & F) != 0;
    }
};
using HpSpecDyn = HpSpec<-1>;
constexpr int hp_spec_id(int controller, int flags) { return (controller & 0xf) | ((flags & 0xf) << 4); }

AGX_DEV EnvState unpack(const float r[13]) {
    EnvState s;
    s.x = V3{r[0], r[1], r[2]};
    s.q = Q4{r[3], r[4], r[5], r[6]};
    s.v = V3{r[7], r[8], r[9]};
    s.w = V3{r[10], r[11], r[12]};
    return s;
}
AGX_DEV void pack(const EnvState& s, float r[13]) {
    r[0] = s.x.x; r[1] = s.x.y; r[2] = s.x.z;
    r[3] = s.q.x; r[4] = s.q.y; r[5] = s.q.z; r[6] = s.q.w;
    r[7] = s.v.x; r[8] = s.v.y; r[9] = s.v.z;
    r[10] = s.w.x; r[11] = s.w.y; r[12] = s.w.z;
}

// ---- a1: BaseMultirotor.update_states   robots/base_multirotor.py:287-294 ----------------
// -DAGX_HP1_NOINLINE_DERIVED (experiment): one out-of-line copy shared by the call at the start of the step and the refresh
#if defined(AGX_HP1_NOINLINE_DERIVED) && !defined(AGX_HOST_SHADOW)
__device__ __noinline__ Derived update_states(const EnvState& s) {
#else
AGX_DEV Derived update_states(const EnvState& s) {
#endif
    Derived d;
    V3 e = euler_xyz_0_2pi(s.q);
    d.euler = V3{ssa_0_2pi(e.x), ssa_0_2pi(e.y), ssa_0_2pi(e.z)};
    d.qveh = quat_from_yaw(e.z);  // vehicle_frame_quat_from_quat uses the [0,2pi) yaw
    d.vveh = quat_rotate_inverse(d.qveh, s.v);
    d.vb = quat_rotate_inverse(s.q, s.v);
    d.wb = quat_rotate_inverse(s.q, s.w);
    return d;
}

// ---- a2: compute_acceleration           control/controllers/base_lee_controller.py:120-134
AGX_DEV V3 compute_acceleration(const EnvState& s, const Derived& d, const Gains& g, V3 sp_pos,
                                                   V3 sp_vel) {
    V3 pos_err = sp_pos - s.x;
    V3 vel_err = quat_rotate(d.qveh, sp_vel) - s.v;
    return g.kp * pos_err + g.kv * vel_err;
}
// ---- a5: euler_rates_to_body_rates      base_lee_controller.py:200-215 (yaw-rate only column
// plus the generic roll/pitch entries the reference writes) --------------------------------
AGX_DEV V3 euler_rates_to_body_rates(V3 euler, V3 rates) {
    float sp, cp, sr, cr;
    sincos_(euler.y, &sp, &cp);
    sincos_(euler.x, &sr, &cr);
    // T = [[1,0,-sp],[0,cr,sr*cp],[0,-sr,cr*cp]]
    return V3{rates.x + (-sp) * rates.z, cr * rates.y + (sr * cp) * rates.z, (-sr) * rates.y + (cr * cp) * rates.z};
}
// ---- a6: compute_body_torque            base_lee_controller.py:136-154 -------------------
AGX_DEV V3 compute_body_torque(const AgxHp1Config& cfg, const EnvState& s, const Derived& d,
                                                  const Gains& g, Q4 q_des, V3 w_des) {
    w_des.z = fminf(fmaxf(w_des.z, -cfg.max_yaw_rate), cfg.max_yaw_rate);
    Q4 qe = quat_mul(quat_conj(s.q), q_des);
    M33 R = quat_to_matrix(qe);
    // 0.5 * vee(R^T - R): [-S12, S02, -S01], S = R^T - R
    V3 rot_err{0.5f * -(R.m[7] - R.m[5]), 0.5f * (R.m[6] - R.m[2]), 0.5f * -(R.m[3] - R.m[1])};
    V3 angvel_err = d.wb - quat_rotate(qe, w_des);
    const float* J = cfg.inertia;
    V3 W = d.wb;
    V3 JW{J[0] * W.x + J[1] * W.y + J[2] * W.z, J[3] * W.x + J[4] * W.y + J[5] * W.z,
          J[6] * W.x + J[7] * W.y + J[8] * W.z};
    V3 ff = cross(W, JW);
    return neg(g.kr * rot_err) - g.kw * angvel_err + ff;
}
// ---- a3: calculate_desired_orientation_for_position_velocity_control  :173-194 ----------
AGX_DEV Q4 desired_orientation_pos_vel(V3 f, float yaw) {
    float fn = norm3(f);
    V3 b3{f.x / fn, f.y / fn, f.z / fn};
    float sy, cy;
    sincos_(yaw, &sy, &cy);
    V3 b2 = cross(b3, V3{cy, sy, 0.0f});
    float n2 = norm3(b2);
    b2 = V3{b2.x / n2, b2.y / n2, b2.z / n2};
    V3 b1 = cross(b2, b3);
    return matrix_cols_to_quat(b1, b2, b3);
}
// ---- a4: calculate_desired_orientation_from_forces_and_yaw  :157-169 --------------------
AGX_DEV Q4 desired_orientation_forces_yaw(V3 f, float yaw) {
    float pitch = atan2f(f.x, f.z);
    float roll = atan2f(-f.y, sqrtf(f.z * f.z + f.x * f.x));
    return quat_from_euler(roll, pitch, yaw);
}

// ---- a7: controller dispatch -> wrench[6] (or motor refs for AGX_CTRL_NONE) --------------
template <class SP = HpSpecDyn>
AGX_DEV void controller_wrench(const AgxHp1Config& cfg, const EnvState& s, const Derived& d,
                                                  const Gains& g, const float* act, float wr[6]) {
    const V3 grav = ld3c(cfg.gravity);
    const float m = cfg.mass;
#pragma unroll
    for (int i = 0; i < 6; ++i) wr[i] = 0.0f;
    const int c = SP::controller(cfg);
    const V3 zero3{0.0f, 0.0f, 0.0f};
    if (c == AGX_CTRL_ATTITUDE) {  // controllers/attitude_control.py:16-43
        wr[2] = (act[0] + 1.0f) * m * norm3(grav);
        V3 w_des = euler_rates_to_body_rates(d.euler, V3{0.0f, 0.0f, act[3]});
        Q4 q_des = quat_from_euler(act[1], act[2], d.euler.z);
        V3 t = compute_body_torque(cfg, s, d, g, q_des, w_des);
        wr[3] = t.x; wr[4] = t.y; wr[5] = t.z;
        return;
    }
    if (c == AGX_CTRL_RATES) {  // controllers/rates_control.py:23-26 (intent; reference line raises)
        wr[2] = (act[0] - grav.z) * m;
        V3 t = compute_body_torque(cfg, s, d, g, s.q, V3{act[1], act[2], act[3]});
        wr[3] = t.x; wr[4] = t.y; wr[5] = t.z;
        return;
    }
    if (c == AGX_CTRL_FULLY_ACTUATED) {  // controllers/fully_actuated_control.py:14-32
        float qn = fmaxf(sqrtf(act[3] * act[3] + act[4] * act[4] + act[5] * act[5] + act[6] * act[6]), 1e-9f);
        Q4 q_des{act[3] / qn, act[4] / qn, act[5] / qn, act[6] / qn};
        V3 accel = compute_acceleration(s, d, g, V3{act[0], act[1], act[2]}, zero3);
        V3 forces = (accel - grav) * m;
        V3 fb = quat_rotate_inverse(s.q, forces);
        wr[0] = fb.x; wr[1] = fb.y; wr[2] = fb.z;
        V3 t = compute_body_torque(cfg, s, d, g, q_des, zero3);
        wr[3] = t.x; wr[4] = t.y; wr[5] = t.z;
        return;
    }
    // thrust-vectoring family: position / velocity / velocity-steering / acceleration
    V3 accel;
    if (c == AGX_CTRL_POSITION) accel = compute_acceleration(s, d, g, V3{act[0], act[1], act[2]}, zero3);
    else if (c == AGX_CTRL_ACCELERATION) accel = V3{act[0], act[1], act[2]};
    else accel = compute_acceleration(s, d, g, s.x, V3{act[0], act[1], act[2]});
    V3 forces = (accel - grav) * m;
    M33 R = quat_to_matrix(s.q);
    wr[2] = forces.x * R.m[2] + forces.y * R.m[5] + forces.z * R.m[8];
    Q4 q_des;
    V3 w_des = zero3;
    if (c == AGX_CTRL_POSITION || c == AGX_CTRL_VELOCITY_STEERING) {
        q_des = desired_orientation_pos_vel(forces, act[3]);
    } else if (c == AGX_CTRL_VELOCITY) {
        q_des = desired_orientation_pos_vel(forces, d.euler.z);
        w_des = euler_rates_to_body_rates(d.euler, V3{0.0f, 0.0f, act[3]});
    } else {  // acceleration
        q_des = desired_orientation_forces_yaw(forces, d.euler.z);
        w_des = euler_rates_to_body_rates(d.euler, V3{0.0f, 0.0f, act[3]});
    }
    V3 t = compute_body_torque(cfg, s, d, g, q_des, w_des);
    wr[3] = t.x; wr[4] = t.y; wr[5] = t.z;
}

// ---- a9: MotorModel.update_motor_thrusts   control/motor_model.py:88-251 -----------------
AGX_DEV float motor_rate(float err, float mix, float max_rate) {
    return fmaxf(fminf(mix * err, max_rate), -max_rate);  // tensor_clamp, motor_model.py:160-162
}
AGX_DEV float motor_rk4(float ref, float cur, float mix, float max_rate, float dt) {
    float k1 = motor_rate(ref - cur, mix, max_rate);
    float k2 = motor_rate(ref - (cur + 0.5f * dt * k1), mix, max_rate);
    float k3 = motor_rate(ref - (cur + 0.5f * dt * k2), mix, max_rate);
    float k4 = motor_rate(ref - (cur + dt * k3), mix, max_rate);
    return (dt / 6.0f) * (k1 + 2.0f * k2 + 2.0f * k3 + k4);
}
template <class SP = HpSpecDyn>
AGX_DEV float motor_update(const AgxHp1Config& cfg, float cur, float ref_in, float tau_inc,
                                              float tau_dec, float k) {
    const float dt = cfg.dt;
    float ref = fminf(fmaxf(ref_in, cfg.min_thrust), cfg.max_thrust);
    float err = ref - cur;
    bool decreasing = (cur > 0.0f && err < 0.0f) || (cur < 0.0f && err > 0.0f);  // sign(f)*sign(err) < 0
    float tau = decreasing ? tau_dec : tau_inc;
    float mix = SP::template flag<AGX_F_DISCRETE_MIX>(cfg) ? 1.0f / (dt + tau) : 1.0f / tau;
    const bool rk4 = SP::template flag<AGX_F_MOTOR_RK4>(cfg);
    if (SP::template flag<AGX_F_USE_RPS>(cfg)) {
        float rpm = sqrtf(cur / k);
        float rpm_ref = sqrtf(ref / k);
        if (rk4) rpm += motor_rk4(rpm_ref, rpm, mix, cfg.max_thrust_rate, dt);
        else rpm += motor_rate(rpm_ref - rpm, mix, cfg.max_thrust_rate) * dt;
        return k * (rpm * rpm);
    }
    if (rk4) return cur + motor_rk4(ref, cur, mix, cfg.max_thrust_rate, dt);
    return cur + motor_rate(err, mix, cfg.max_thrust_rate) * dt;
}

// ---- a13: rigid-body integrator -- OUR SPEC (DESIGN.md "Integrator spec"; oracle
// rigid_body_integrate).  Replaces gym.simulate (env_manager/IGE_env_manager.py:477). --------
template <class SP = HpSpecDyn>
AGX_DEV void integrate(const AgxHp1Config& cfg, EnvState& s, V3 F, V3 T) {
    const float dt = cfg.dt;
    V3 a = quat_rotate(s.q, F) * (1.0f / cfg.mass) + ld3c(cfg.gravity);
    V3 v = (s.v + a * dt) * fmaxf(0.0f, 1.0f - dt * cfg.linear_damping);
    float vn = norm3(v);
    if (vn > cfg.max_linear_velocity) v = v * (cfg.max_linear_velocity / vn);
    V3 W = quat_rotate_inverse(s.q, s.w);
    const float* J = cfg.inertia;
    const float* Ji = cfg.inertia_inv;
    V3 rhs = T;
    if (SP::template flag<AGX_F_GYROSCOPIC>(cfg)) {
        V3 JW{J[0] * W.x + J[1] * W.y + J[2] * W.z, J[3] * W.x + J[4] * W.y + J[5] * W.z,
              J[6] * W.x + J[7] * W.y + J[8] * W.z};
        rhs = T - cross(W, JW);
    }
    V3 al{Ji[0] * rhs.x + Ji[1] * rhs.y + Ji[2] * rhs.z, Ji[3] * rhs.x + Ji[4] * rhs.y + Ji[5] * rhs.z,
          Ji[6] * rhs.x + Ji[7] * rhs.y + Ji[8] * rhs.z};
    V3 Wn = W + al * dt;
    V3 w = quat_rotate(s.q, Wn) * fmaxf(0.0f, 1.0f - dt * cfg.angular_damping);
    float wn = norm3(w);
    if (wn > cfg.max_angular_velocity) {
        w = w * (cfg.max_angular_velocity / wn);
        wn = norm3(w);
    }
    s.x = s.x + v * dt;
    float sh, ch;
    sincos_(0.5f * dt * wn, &sh, &ch);
    float so = (wn > 0.0f) ? sh / wn : 0.0f;
    Q4 dq{w.x * so, w.y * so, w.z * so, ch};
    Q4 qn = quat_mul(dq, s.q);
    float inv = 1.0f / sqrtf(qn.x * qn.x + qn.y * qn.y + qn.z * qn.z + qn.w * qn.w);
    s.q = Q4{qn.x * inv, qn.y * inv, qn.z * inv, qn.w * inv};
    s.v = v;
    s.w = w;
}

// L2-only load (the chained steps of hp1.cu never read state through an L1 line); a plain load on the host
AGX_DEV V3 ld3cg(const float* p) {
#ifdef __CUDA_ARCH__
    return V3{__ldcg(p), __ldcg(p + 1), __ldcg(p + 2)};
#else
    return V3{p[0], p[1], p[2]};
#endif
}

// per-env parameters held in registers for the whole step (a15 writes them on reset)
template <int M>
struct EnvParams {
    float thrust[M], tau_inc[M], tau_dec[M], k[M];
    Gains g;
    V3 bmin, bmax;
};
AGX_DEV float lerp_u(float lo, float hi, float u) { return (hi - lo) * u + lo; }  // math.py:51-54

// ---- a15: reset arithmetic (uniforms in registers) ----------------------------------------
template <int M>
AGX_DEV void apply_reset(const AgxHp1Config& cfg, const float us[13], const float ubl[3],
                                            const float ubh[3], const float ug[12], const float um[4 * M],
                                            EnvState& s, EnvParams<M>& p) {
    // IGE_env_manager.py:513-519
    p.bmin = V3{lerp_u(cfg.bounds_lo_min[0], cfg.bounds_lo_max[0], ubl[0]), lerp_u(cfg.bounds_lo_min[1], cfg.bounds_lo_max[1], ubl[1]),
                lerp_u(cfg.bounds_lo_min[2], cfg.bounds_lo_max[2], ubl[2])};
    p.bmax = V3{lerp_u(cfg.bounds_hi_min[0], cfg.bounds_hi_max[0], ubh[0]), lerp_u(cfg.bounds_hi_min[1], cfg.bounds_hi_max[1], ubh[1]),
                lerp_u(cfg.bounds_hi_min[2], cfg.bounds_hi_max[2], ubh[2])};
    // base_multirotor.py:177-199
    float rs[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) rs[j] = lerp_u(cfg.min_init_state[j], cfg.max_init_state[j], us[j]);
    s.x = V3{p.bmin.x + (p.bmax.x - p.bmin.x) * rs[0], p.bmin.y + (p.bmax.y - p.bmin.y) * rs[1],
             p.bmin.z + (p.bmax.z - p.bmin.z) * rs[2]};
    s.q = quat_from_euler(rs[3], rs[4], rs[5]);
    s.v = V3{rs[7], rs[8], rs[9]};
    s.w = V3{rs[10], rs[11], rs[12]};
    if (cfg.flags & AGX_F_RANDOMIZE_GAINS) {  // base_lee_controller.py:101-118
        p.g.kp = V3{lerp_u(cfg.K_pos_min[0], cfg.K_pos_max[0], ug[0]), lerp_u(cfg.K_pos_min[1], cfg.K_pos_max[1], ug[1]), lerp_u(cfg.K_pos_min[2], cfg.K_pos_max[2], ug[2])};
        p.g.kv = V3{lerp_u(cfg.K_vel_min[0], cfg.K_vel_max[0], ug[3]), lerp_u(cfg.K_vel_min[1], cfg.K_vel_max[1], ug[4]), lerp_u(cfg.K_vel_min[2], cfg.K_vel_max[2], ug[5])};
        p.g.kr = V3{lerp_u(cfg.K_rot_min[0], cfg.K_rot_max[0], ug[6]), lerp_u(cfg.K_rot_min[1], cfg.K_rot_max[1], ug[7]), lerp_u(cfg.K_rot_min[2], cfg.K_rot_max[2], ug[8])};
        p.g.kw = V3{lerp_u(cfg.K_angvel_min[0], cfg.K_angvel_max[0], ug[9]), lerp_u(cfg.K_angvel_min[1], cfg.K_angvel_max[1], ug[10]), lerp_u(cfg.K_angvel_min[2], cfg.K_angvel_max[2], ug[11])};
    }
    // motor_model.py:140-154
#pragma unroll
    for (int i = 0; i < M; ++i) {
        p.tau_inc[i] = lerp_u(cfg.tau_inc_range[0], cfg.tau_inc_range[1], um[4 * i + 0]);
        p.tau_dec[i] = lerp_u(cfg.tau_dec_range[0], cfg.tau_dec_range[1], um[4 * i + 1]);
        p.thrust[i] = lerp_u(cfg.min_thrust, cfg.max_thrust, um[4 * i + 2]);
        if (cfg.flags & AGX_F_USE_RPS) p.k[i] = lerp_u(cfg.k_thrust_range[0], cfg.k_thrust_range[1], um[4 * i + 3]);
    }
}

// device-RNG draw layout (oracle/philox.py restates it):
//   block 0..2 -> state[0..11]; block 3 -> state[12], bounds_lo[0..2]; block 4 -> bounds_hi[0..2], -
//   block 5..8 -> K_pos, K_vel, K_rot, K_angvel (xyz, -); block 9+i -> motor i: tau_inc, tau_dec, thrust, k
template <int M>
AGX_DEV void device_rng_reset(const AgxHp1Config& cfg, uint32_t env_gid, uint32_t episode, EnvState& s,
                                                 EnvParams<M>& p) {
    const uint32_t k0 = (uint32_t)(cfg.seed & 0xffffffffu), k1 = (uint32_t)(cfg.seed >> 32);
    float us[13], ubl[3] = {0.f, 0.f, 0.f}, ubh[3] = {0.f, 0.f, 0.f}, ug[12], um[4 * M];
    U4 b;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        b = philox4x32_10(U4{env_gid, episode, (uint32_t)i, 0u}, k0, k1);
        us[4 * i] = u01(b.x); us[4 * i + 1] = u01(b.y); us[4 * i + 2] = u01(b.z); us[4 * i + 3] = u01(b.w);
    }
    b = philox4x32_10(U4{env_gid, episode, 3u, 0u}, k0, k1);
    us[12] = u01(b.x); ubl[0] = u01(b.y); ubl[1] = u01(b.z); ubl[2] = u01(b.w);
    // blocks whose draws cannot matter are skipped (a degenerate range maps every u to the same value)
    const bool bounds_hi_random = cfg.bounds_hi_min[0] != cfg.bounds_hi_max[0] || cfg.bounds_hi_min[1] != cfg.bounds_hi_max[1] ||
                                  cfg.bounds_hi_min[2] != cfg.bounds_hi_max[2];
    if (bounds_hi_random) {
        b = philox4x32_10(U4{env_gid, episode, 4u, 0u}, k0, k1);
        ubh[0] = u01(b.x); ubh[1] = u01(b.y); ubh[2] = u01(b.z);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) ug[i] = 0.0f;
    if (cfg.flags & AGX_F_RANDOMIZE_GAINS) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            b = philox4x32_10(U4{env_gid, episode, (uint32_t)(5 + i), 0u}, k0, k1);
            ug[3 * i] = u01(b.x); ug[3 * i + 1] = u01(b.y); ug[3 * i + 2] = u01(b.z);
        }
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
        b = philox4x32_10(U4{env_gid, episode, (uint32_t)(9 + i), 0u}, k0, k1);
        um[4 * i] = u01(b.x); um[4 * i + 1] = u01(b.y); um[4 * i + 2] = u01(b.z); um[4 * i + 3] = u01(b.w);
    }
    apply_reset<M>(cfg, us, ubl, ubh, ug, um, s, p);
}

// Warp-cooperative form of device_rng_reset for the fused step: lane j computes Philox block j of the
// resetting env (9 + M <= 17 blocks), the uniforms go through the warp's shared-memory tile and only
// the owning lane applies them.  Same blocks, same uniforms, same results as the scalar form, but a
// reset costs the warp one Philox evaluation instead of 8-17 (resets are rare and divergent: a warp
// with ONE resetting env used to run ~850 extra instructions for it, the critical path of the launch).
template <int M>
AGX_DEV void coop_rng_draw(const AgxHp1Config& cfg, uint32_t env_gid, uint32_t episode, int lane, float* tile) {
    const uint32_t k0 = (uint32_t)(cfg.seed & 0xffffffffu), k1 = (uint32_t)(cfg.seed >> 32);
    if (lane < 9 + M) {
        const bool bounds_hi_random = cfg.bounds_hi_min[0] != cfg.bounds_hi_max[0] || cfg.bounds_hi_min[1] != cfg.bounds_hi_max[1] ||
                                      cfg.bounds_hi_min[2] != cfg.bounds_hi_max[2];
        const bool needed = lane < 4 || lane >= 9 || (lane == 4 ? bounds_hi_random : (cfg.flags & AGX_F_RANDOMIZE_GAINS) != 0);
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
        if (needed) {
            U4 b = philox4x32_10(U4{env_gid, episode, (uint32_t)lane, 0u}, k0, k1);
            u = make_float4(u01(b.x), u01(b.y), u01(b.z), u01(b.w));
        }
        reinterpret_cast<float4*>(tile)[lane] = u;
    }
}
template <int M>
AGX_DEV void apply_reset_from_tile(const AgxHp1Config& cfg, const float* tile, EnvState& s, EnvParams<M>& p) {
    float us[13], ubl[3], ubh[3], ug[12], um[4 * M];
#pragma unroll
    for (int i = 0; i < 13; ++i) us[i] = tile[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { ubl[i] = tile[13 + i]; ubh[i] = tile[16 + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ug[3 * i] = tile[20 + 4 * i]; ug[3 * i + 1] = tile[21 + 4 * i]; ug[3 * i + 2] = tile[22 + 4 * i];
    }
#pragma unroll
    for (int i = 0; i < 4 * M; ++i) um[i] = tile[36 + i];
    apply_reset<M>(cfg, us, ubl, ubh, ug, um, s, p);
}

// ---- a16: observation of the position task
AGX_DEV void make_obs(const EnvState& s, const Derived& d, V3 tgt, float o[13]) {
    // process_obs_for_task, task/position_setpoint_task/position_setpoint_task.py:194-203
    o[0] = tgt.x - s.x.x; o[1] = tgt.y - s.x.y; o[2] = tgt.z - s.x.z;
    o[3] = s.q.x; o[4] = s.q.y; o[5] = s.q.z; o[6] = s.q.w;
    o[7] = d.vb.x; o[8] = d.vb.y; o[9] = d.vb.z;
    o[10] = d.wb.x; o[11] = d.wb.y; o[12] = d.wb.z;
}

}  // namespace agx
