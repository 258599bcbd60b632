"""Oracle model tables for the golden-fixture robots (test helper).

Values restate the reference configs: config/robot_config/{base_quad,base_octarotor,lmf2}_config.py,
config/controller_config/{lee_controller_config,lee_controller_config_octarotor,
fully_actuated_controller_rov,lmf2_controller_config}.py."""
import math

import numpy as np

from oracle import hp1_oracle as O

QUAD_ALLOC = [
    [0.0, 0.0, 0.0, 0.0],
    [0.0, 0.0, 0.0, 0.0],
    [1.0, 1.0, 1.0, 1.0],
    [-0.13, -0.13, 0.13, 0.13],
    [-0.13, 0.13, 0.13, -0.13],
    [-0.01, 0.01, -0.01, 0.01],
]
LMF2_ALLOC = [row[:] for row in QUAD_ALLOC]
LMF2_ALLOC[5] = [-0.07, 0.07, -0.07, 0.07]
OCTA_ALLOC = [
    [-0.78867513, 0.21132487, -0.21132487, 0.78867513, 0.78867513, -0.21132487, 0.21132487, -0.78867513],
    [0.21132487, 0.78867513, -0.78867513, -0.21132487, -0.21132487, -0.78867513, 0.78867513, 0.21132487],
    [0.57735027, -0.57735027, -0.57735027, 0.57735027, 0.57735027, -0.57735027, -0.57735027, 0.57735027],
    [0.14226497, -0.21547005, 0.25773503, 0.01547005, -0.01547005, -0.25773503, 0.21547005, -0.14226497],
    [-0.25773503, 0.01547005, 0.14226497, 0.21547005, -0.21547005, -0.14226497, -0.01547005, 0.25773503],
    [0.11547005, -0.23094011, -0.11547005, 0.23094011, -0.23094011, 0.11547005, 0.23094011, -0.11547005],
]

LEE = dict(
    K_pos_range=((2.0, 2.0, 1.0), (3.0, 3.0, 2.0)),
    K_vel_range=((2.0, 2.0, 2.0), (3.0, 3.0, 3.0)),
    K_rot_range=((0.8, 0.8, 0.4), (1.2, 1.2, 0.6)),
    K_angvel_range=((0.1, 0.1, 0.1), (0.2, 0.2, 0.2)),
    randomize_params=False,
)
LEE_OCTA = dict(
    K_pos_range=((2.0, 2.0, 1.0), (3.0, 3.0, 2.0)),
    K_vel_range=((2.0, 2.0, 2.0), (3.0, 3.0, 3.0)),
    K_rot_range=((10.8, 10.8, 5.4), (10.2, 10.2, 5.6)),
    K_angvel_range=((2.1, 2.1, 2.1), (2.2, 2.2, 2.2)),
    randomize_params=True,
)
ROV_FA = dict(
    K_pos_range=((1.0, 1.0, 1.0), (1.0, 1.0, 1.0)),
    K_vel_range=((8.0, 8.0, 8.0), (8.0, 8.0, 8.0)),
    K_rot_range=((2.2, 2.2, 2.6), (2.2, 2.2, 2.6)),
    K_angvel_range=((2.1, 2.1, 2.1), (2.2, 2.2, 2.2)),
    randomize_params=True,
)
LMF2 = dict(
    K_pos_range=((2.0, 2.0, 1.0), (2.0, 2.0, 1.0)),
    K_vel_range=((2.7, 2.7, 1.7), (3.3, 3.3, 1.3)),
    K_rot_range=((1.6, 1.6, 0.25), (1.85, 1.85, 0.4)),
    K_angvel_range=((0.4, 0.4, 0.075), (0.5, 0.5, 0.09)),
    randomize_params=True,
)

CTRL_IDS = {
    "lee_attitude_control": O.CTRL_ATTITUDE,
    "lee_position_control": O.CTRL_POSITION,
    "lee_velocity_control": O.CTRL_VELOCITY,
    "lee_acceleration_control": O.CTRL_ACCELERATION,
    "no_control": O.CTRL_NONE,
    "octarotor_velocity_control": O.CTRL_VELOCITY,
    "rov_fully_actuated_control": O.CTRL_FULLY_ACTUATED,
    "lmf2_velocity_control": O.CTRL_VELOCITY,
}
CTRL_CFG = {
    "lee_attitude_control": LEE, "lee_position_control": LEE, "lee_velocity_control": LEE,
    "lee_acceleration_control": LEE, "no_control": LEE,
    "octarotor_velocity_control": LEE_OCTA, "rov_fully_actuated_control": ROV_FA,
    "lmf2_velocity_control": LMF2,
}


def oracle_model(robot: str, controller: str, mass, inertia) -> O.Hp1Model:
    kw = dict(controller=CTRL_IDS[controller], mass=float(mass), inertia=np.asarray(inertia))
    kw.update(CTRL_CFG[controller])
    if robot in ("base_quadrotor", "base_quad_root_link_control"):
        kw.update(
            num_motors=4, allocation_matrix=QUAD_ALLOC, motor_directions=[1, -1, 1, -1],
            force_application_level="motor_link" if robot == "base_quadrotor" else "root_link",
            use_rps=True, min_thrust=0.0, thrust_to_torque_ratio=0.01,
        )
        if robot == "base_quadrotor":
            kw.update(max_thrust=2.0, tau_inc_range=(0.04, 0.04), tau_dec_range=(0.04, 0.04))
        else:
            kw.update(max_thrust=10.0, tau_inc_range=(0.01, 0.03), tau_dec_range=(0.005, 0.005),
                      k_thrust_range=(0.00001826312, 0.00001826312))
    elif robot == "base_octarotor":
        kw.update(
            num_motors=8, allocation_matrix=OCTA_ALLOC, motor_directions=[1, -1, 1, -1, 1, -1, 1, -1],
            force_application_level="motor_link", use_rps=False, min_thrust=-6.25, max_thrust=6.25,
            tau_inc_range=(0.01, 0.03), tau_dec_range=(0.005, 0.005), thrust_to_torque_ratio=0.01,
            link_r=np.zeros((8, 3)), link_R=np.tile(np.eye(3), (8, 1, 1)),
            enable_disturbance=False,
            min_init_state=(0, 0, 0, 0, 0, -math.pi, 1.0, -0.2, -0.2, -0.2, -0.2, -0.2, -0.2),
            max_init_state=(1.0, 1.0, 1.0, 0, 0, math.pi, 1.0, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2),
        )
    elif robot == "lmf2":
        kw.update(
            num_motors=4, allocation_matrix=LMF2_ALLOC, motor_directions=[1, -1, 1, -1],
            force_application_level="base_link", use_rps=True, min_thrust=0.1, max_thrust=10.0,
            tau_inc_range=(0.05, 0.08), tau_dec_range=(0.005, 0.005), thrust_to_torque_ratio=0.07,
        )
    else:
        raise KeyError(robot)
    return O.Hp1Model(**kw)
