/*
 * aerial_gym_b200 -- C ABI of the B200-native hot paths (libaerial_gym_b200.so).
 *
 * The reference (ntnu-arl/aerial_gym_simulator, commit f0d0f05) has NO plugin / FFI boundary
 * for these paths: they sit behind plain Python objects.  The entry
 * points below are therefore the narrowest C surface that replaces, one for one, the Python
 * call sites listed beside each function.  Conventions:
 *
 *   - plain C, POD structs, raw device pointers + sizes; no torch / C++ types;
 *   - the caller (PyTorch) owns every buffer; the library never allocates user-visible
 *     memory, never frees, never synchronises the stream unless the function says so;
 *   - every function returns 0 on success or a negative AGX_E_* code; agx_last_error()
 *     returns a thread-local message; no exceptions cross the boundary;
 *   - kernels are launched on the cudaStream_t passed in (as void*; 0 = legacy default).
 *
 * HP1 = rigid-body integrator + rotor/motor model + Lee controller + allocation.
 * HP2 = depth / segmentation / LiDAR ray-caster.
 */
#ifndef AERIAL_GYM_B200_H_
#define AERIAL_GYM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGX_ABI_VERSION 1
#define AGX_MAX_MOTORS 8

/* error codes */
#define AGX_OK 0
#define AGX_E_INVALID (-1)   /* bad argument / unsupported configuration */
#define AGX_E_CUDA (-2)      /* CUDA runtime error (message in agx_last_error) */
#define AGX_E_NULL (-3)      /* required pointer is NULL */

/* controller ids -- one per reference controller class (aerial_gym/control/__init__.py:42-99) */
#define AGX_CTRL_NONE 0              /* NoControl                       controllers/no_control.py:29-30 */
#define AGX_CTRL_ATTITUDE 1          /* LeeAttitudeController           controllers/attitude_control.py:16-43 */
#define AGX_CTRL_POSITION 2          /* LeePositionController           controllers/position_control.py:21-51 */
#define AGX_CTRL_VELOCITY 3          /* LeeVelocityController           controllers/velocity_control.py:17-53 */
#define AGX_CTRL_ACCELERATION 4      /* LeeAccelerationController       controllers/acceleration_control.py:16-46 */
#define AGX_CTRL_RATES 5             /* LeeRatesController              controllers/rates_control.py:16-28 */
#define AGX_CTRL_FULLY_ACTUATED 6    /* FullyActuatedController         controllers/fully_actuated_control.py:14-32 */
#define AGX_CTRL_VELOCITY_STEERING 7 /* LeeVelocitySteeringAngleController controllers/velocity_steeing_angle_controller.py:16-50 */

/* AgxHp1Config.flags */
#define AGX_F_USE_RPS 0x1            /* motor_model.py:98 (cfg.use_rps) */
#define AGX_F_MOTOR_RK4 0x2          /* motor_model.py:13-19 (default rk4) */
#define AGX_F_DISCRETE_MIX 0x4       /* motor_model.py:201-210 */
#define AGX_F_GYROSCOPIC 0x8         /* integrator spec: include -W x JW */
#define AGX_F_RANDOMIZE_GAINS 0x10   /* base_lee_controller.py:101-118 (cfg.randomize_params) */
#define AGX_F_DEVICE_RNG_RESET 0x20  /* fused step resets finished envs in-kernel (Philox4x32-10) */
#define AGX_F_STRICT_STALE_OBS 0x40  /* reproduce the stale-derived-state quirk */

typedef struct AgxHp1Config {
    int32_t num_envs;
    int32_t num_motors;        /* M in [1, AGX_MAX_MOTORS] */
    int32_t controller;        /* AGX_CTRL_* */
    int32_t num_actions;       /* columns of the action tensor */
    int32_t physics_steps;     /* physics steps per call (env_manager.py:417-428) */
    int32_t flags;             /* AGX_F_* */
    int32_t episode_len_steps; /* position task: truncation when sim_steps > this */
    int32_t env_id_offset;     /* global id of local env 0 (multi-GPU sharding; keys the device RNG) */
    uint64_t seed;             /* device RNG key */
    float dt;
    float gravity[3];
    float mass;
    float inertia[9];          /* row-major, about the COM, base frame */
    float inertia_inv[9];
    float alloc_pinv[AGX_MAX_MOTORS * 6]; /* [M][6] pinv(A), control_allocation.py:46-48 */
    float wrench_map[6 * AGX_MAX_MOTORS]; /* [6][M] motor thrust -> base-frame wrench about the COM */
    float com[3];
    float min_thrust, max_thrust, max_thrust_rate;
    float max_yaw_rate;
    float drag_lin1[3], drag_lin2[3], drag_ang1[3], drag_ang2[3]; /* base_multirotor.py:260-285 */
    float linear_damping, angular_damping;   /* base_quad_config.py:93-94 */
    float max_linear_velocity, max_angular_velocity;
    /* per-env parameter arrays may be NULL in AgxHp1Buffers; these constants are used instead */
    float K_pos[3], K_vel[3], K_rot[3], K_angvel[3];
    float tau_inc, tau_dec, k_thrust;
    float crash_distance;      /* position task: crash when |target - x| > this (8.0) */
    /* reset ranges (uniform: lo + (hi - lo) * u) */
    float min_init_state[13], max_init_state[13]; /* base_quad_config.py:30-59 */
    float bounds_lo_min[3], bounds_lo_max[3];     /* env_config/empty_env.py:27-31 */
    float bounds_hi_min[3], bounds_hi_max[3];
    float tau_inc_range[2], tau_dec_range[2], k_thrust_range[2];
    float K_pos_min[3], K_pos_max[3], K_vel_min[3], K_vel_max[3];
    float K_rot_min[3], K_rot_max[3], K_angvel_min[3], K_angvel_max[3];
} AgxHp1Config;

/* All pointers are DEVICE pointers.  [N,...] arrays are dense row-major fp32 unless noted. */
typedef struct AgxHp1Buffers {
    /* state (read + written) */
    float* root_state;        /* [N,13] x y z qx qy qz qw vx vy vz wx wy wz  (IGE_env_manager.py:347-358) */
    float* motor_thrust;      /* [N,M]  MotorModel.current_motor_thrust */
    int32_t* sim_steps;       /* [N]    env_manager.py:78-80 (may be NULL for agx_hp1_physics_step) */
    /* inputs */
    const float* actions;     /* [N,num_actions] */
    const float* disturbance; /* [N,6] gated body-0 wrench or NULL (base_multirotor.py:213-234) */
    const float* target_position; /* [N,3] or NULL (= zeros) */
    /* per-env parameters; NULL = use the AgxHp1Config constant.  Written on reset. */
    float* tau_inc;           /* [N,M] */
    float* tau_dec;           /* [N,M] */
    float* k_thrust;          /* [N,M] */
    float* K_pos;             /* [N,3] */
    float* K_vel;
    float* K_rot;
    float* K_angvel;
    float* bounds_min;        /* [N,3] env bounds (reset); may be NULL = bounds_lo_min / bounds_hi_min */
    float* bounds_max;
    /* derived states (BaseMultirotor.update_states); each may be NULL = not materialised */
    float* euler;             /* [N,3] */
    float* vehicle_orientation; /* [N,4] */
    float* vehicle_linvel;    /* [N,3] */
    float* body_linvel;       /* [N,3] */
    float* body_angvel;       /* [N,3] */
    float* body_wrench;       /* [N,6] net base-frame wrench about the COM of the last physics step, or NULL */
    /* position-task outputs */
    float* obs;               /* [N,13] */
    float* reward;            /* [N] */
    uint8_t* terminations;    /* [N] bool ("crashes") */
    uint8_t* truncations;     /* [N] bool */
    uint8_t* reset_mask;      /* [N] bool, envs reset (or to be reset) this step; may be NULL */
    int32_t* any_reset;       /* [2] device scratch: [0] flag, [1] block-arrival counter; zero-initialised by caller once */
    uint32_t* episode_count;  /* [N] device-RNG counter word, incremented per reset */
} AgxHp1Buffers;

/* explicit uniform draws u in [0,1) for agx_hp1_reset, in the reference's call order */
typedef struct AgxHp1ResetDraws {
    const float* bounds_lo;   /* [N,3]  IGE_env_manager.py:513-516 */
    const float* bounds_hi;   /* [N,3]  IGE_env_manager.py:517-519 */
    const float* state;       /* [N,13] base_multirotor.py:182 */
    const float* K_pos;       /* [N,3] each, or NULL when gains are not randomised */
    const float* K_vel;
    const float* K_rot;
    const float* K_angvel;
    const float* tau_inc;     /* [N,M] motor_model.py:141-143 */
    const float* tau_dec;     /* [N,M] motor_model.py:145-147 */
    const float* thrust;      /* [N,M] motor_model.py:148-150 */
    const float* k_thrust;    /* [N,M] motor_model.py:151-154 (NULL unless use_rps) */
} AgxHp1ResetDraws;

int agx_abi_version(void);
const char* agx_last_error(void);
/* sizeof() of the ABI structs as compiled into the library (binding self-check):
 * which = 0 AgxHp1Config, 1 AgxHp1Buffers, 2 AgxHp1ResetDraws, 3 AgxHp2Scene, 4 AgxHp2Sensor */
uint64_t agx_sizeof(int which);

/* Physics only: `physics_steps` x (update_states -> controller -> allocation -> motor -> drag ->
 * disturbance -> integrate).  Replaces the body of EnvManager.step's loop
 * (env_manager/env_manager.py:426-428 = robots/base_multirotor.py:296-307 + gym.simulate,
 * IGE_env_manager.py:444-449,477,486-495).  Derived-state arrays, when given, receive the
 * PRE-physics values of the last step (the reference's staleness). */
int agx_hp1_physics_step(const AgxHp1Config* cfg, const AgxHp1Buffers* buf, void* stream);

/* Fused PositionSetpointTask.step (task/position_setpoint_task/position_setpoint_task.py:152-182):
 * physics, sim_steps += 1, reward + crash (:245-282), truncation (:172-174), optional in-kernel
 * reset (AGX_F_DEVICE_RNG_RESET), observation (:194-203).  Without AGX_F_DEVICE_RNG_RESET the
 * envs to reset are only flagged (reset_mask, any_reset[0]) and the caller follows with
 * agx_hp1_reset + agx_hp1_refresh. */
int agx_hp1_position_task_step(const AgxHp1Config* cfg, const AgxHp1Buffers* buf, void* stream);

/* Masked re-initialisation, EnvManager.reset_idx for a robot-only scene
 * (env_manager/env_manager.py:273-301 -> IGE_env_manager.py:513-519, base_multirotor.py:177-205,
 * base_lee_controller.py:101-118, motor_model.py:140-154).  `mask` [N] bool selects envs.
 * draws != NULL: uniforms supplied by the caller (torch RNG, reference order);
 * draws == NULL: device RNG (same stream layout as the fused step). */
int agx_hp1_reset(const AgxHp1Config* cfg, const AgxHp1Buffers* buf, const uint8_t* mask,
                  const AgxHp1ResetDraws* draws, void* stream);

/* BaseMultirotor.update_states for ALL envs (base_multirotor.py:287-294, called at :204-205)
 * plus the position-task observation when buf->obs != NULL.  only_if_flag != 0: the pass is a
 * no-op unless buf->any_reset[0] != 0 (and it clears the flag). */
int agx_hp1_refresh(const AgxHp1Config* cfg, const AgxHp1Buffers* buf, int only_if_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AERIAL_GYM_B200_H_ */
