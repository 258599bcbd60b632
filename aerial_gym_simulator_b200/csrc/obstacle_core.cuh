// Kinematic obstacles of the "dynamic_env" environment: ObstacleManager.pre_physics_step
// (env_manager/obstacle_manager.py:40-44) overwrites every obstacle's twist with env_actions before each physics step and
// PhysX (gravity disabled, free base: config/asset_config/dynamic_env_object_config.py:28,41) moves it.  PhysX is not
// observable here, so the advance below is OUR SPEC -- the same semi-implicit step as the robot integrator:
//     v' = v max(0, 1 - dt c_lin),  w' = w max(0, 1 - dt c_ang),  x' = x + dt v',  q' = normalize(dq(w' dt) (x) q)
// AGX_DEV: device-only in the product, host+device in the CPU shadow build (tests/csrc/host_shadow.cu).
#pragma once
#include "agx_math.cuh"

namespace agx {

// one obstacle row [x y z qx qy qz qw vx vy vz wx wy wz]; tw = 6 floats (linear, angular; world frame) or nullptr
AGX_DEV void obstacle_advance(float* row, const float* tw, float dt, int substeps, float lin_damp, float ang_damp) {
    V3 x = ld3(row);
    Q4 q{row[3], row[4], row[5], row[6]};
    V3 v = ld3(row + 7), w = ld3(row + 10);
    const float kl = fmaxf(0.0f, 1.0f - dt * lin_damp), ka = fmaxf(0.0f, 1.0f - dt * ang_damp);
    for (int s = 0; s < substeps; ++s) {
        if (tw) {  // obstacle_manager.py:43-44, once per physics step
            v = ld3(tw);
            w = ld3(tw + 3);
        }
        v = v * kl;
        w = w * ka;
        x = x + v * dt;
        const float wn = norm3(w);
        float sh, ch;
        sincos_(0.5f * dt * wn, &sh, &ch);
        const float so = (wn > 0.0f) ? sh / wn : 0.0f;
        const Q4 qn = quat_mul(Q4{w.x * so, w.y * so, w.z * so, ch}, q);
        const float inv = 1.0f / sqrtf(qn.x * qn.x + qn.y * qn.y + qn.z * qn.z + qn.w * qn.w);
        q = Q4{qn.x * inv, qn.y * inv, qn.z * inv, qn.w * inv};
    }
    row[0] = x.x; row[1] = x.y; row[2] = x.z;
    row[3] = q.x; row[4] = q.y; row[5] = q.z; row[6] = q.w;
    row[7] = v.x; row[8] = v.y; row[9] = v.z;
    row[10] = w.x; row[11] = w.y; row[12] = w.z;
}

// obstacle i of the flattened [N*A] list
AGX_DEV void obstacle_step_item(long long i, float* state, int stride, const float* twist, float dt, int substeps, float lin_damp,
                                float ang_damp) {
    obstacle_advance(state + (size_t)i * stride, twist ? twist + (size_t)i * 6 : nullptr, dt, substeps, lin_damp, ang_damp);
}

}  // namespace agx
