"""More golden vectors from the REFERENCE'S OWN BaseMultirotor.step (this container only), for robots / controllers the first set
(make_golden.py) does not reach:

    python tests/golden/make_golden_registry.py        -> tests/golden/hp1_regstep_<tag>.npz

Difference to make_golden.py: the consumer builds its model from THIS package's registries (robot_registry.make_robot(...).make_spec:
config mirror + URDF pipeline) instead of from a hand-written table in the tests, so the fixtures also pin the config -> spec path for
every robot family.  mass / inertia handed to the reference (it would read them from Isaac Gym) are this package's URDF composites,
stored in the fixture.  The steering-angle controller is imported by the reference's control/__init__.py but never registered; it is
registered here under the name this package uses.  (LeeRatesController cannot be run: rates_control.py:25 mixes an [N] with an
[N,3] tensor and raises -- DESIGN 5.)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as MG  # noqa: E402  (installs the reference loader, imports the reference control stack)

from aerial_gym.config.controller_config.lee_controller_config import control as ref_lee_cfg  # noqa: E402
from aerial_gym.control.controllers.velocity_steeing_angle_controller import LeeVelocitySteeringAngleController  # noqa: E402
from aerial_gym.registry.controller_registry import controller_registry as ref_controllers  # noqa: E402

import aerial_gym_simulator_b200.robots  # noqa: E402,F401
from aerial_gym_simulator_b200.config.env_config import EmptyEnvCfg  # noqa: E402
from aerial_gym_simulator_b200.registry._core import robot_registry as our_robots  # noqa: E402

#        tag                      robot                controller                              actions seed
CASES = [
    ("quad_velocity_steering", "base_quadrotor", "lee_velocity_steering_angle_control", 4, 52),
    ("magpie_acceleration", "magpie", "magpie_acceleration_control", 4, 53),
    ("magpie_velocity", "magpie", "magpie_velocity_control", 4, 54),
    ("x500_position", "x500", "lee_position_control", 4, 55),
    ("lmf1_attitude", "lmf1", "lee_attitude_control", 4, 56),
    ("lmf2_position", "lmf2", "lmf2_position_control", 4, 57),
    ("lmf2_acceleration", "lmf2", "lmf2_acceleration_control", 4, 58),
    ("tinyprop_no_control", "tinyprop", "no_control", 4, 59),
    ("tinyprop_position", "tinyprop", "lee_position_control", 4, 60),
    ("random_velocity", "base_random", "lee_velocity_control", 4, 61),
    ("morphy_stiff_attitude", "morphy_stiff", "lee_attitude_control", 4, 62),
    ("octa_position", "base_octarotor", "octarotor_position_control", 4, 63),
    ("octa_acceleration", "base_octarotor", "octarotor_acceleration_control", 4, 64),
    ("rov_fully_actuated", "base_rov", "rov_fully_actuated_control", 7, 65),  # the reference's BaseROV class (robots/base_rov.py)
]

if __name__ == "__main__":
    if "lee_velocity_steering_angle_control" not in ref_controllers.get_controller_names():
        ref_controllers.register_controller("lee_velocity_steering_angle_control", LeeVelocitySteeringAngleController, ref_lee_cfg)
    for tag, robot, controller, num_actions, seed in CASES:
        ours, _ = our_robots.make_robot(robot, controller, EmptyEnvCfg, "cpu")
        MG.ROBOT_MASS_INERTIA[robot] = (float(ours.robot_mass), np.asarray(ours.robot_inertia, dtype=np.float64))
        if controller == "no_control":
            num_actions = ours.num_actions
        MG.gen_step_fixture(tag, robot, controller, max(ours.num_bodies, 1), num_actions, seed)
        os.replace(os.path.join(HERE, f"hp1_step_{tag}.npz"), os.path.join(HERE, f"hp1_regstep_{tag}.npz"))
