"""Controller registry surface (reference: aerial_gym/control/__init__.py:42-99).

The controllers themselves are not Python here: every class below is a descriptor carrying the
AGX_CTRL_* id that selects the fused code path of the HP1 kernel (csrc/hp1.cu: controller_wrench).
Calling one directly raises -- the arithmetic only exists on the GPU."""
from .. import _lib
from ..config import controller_config as cc
from ..registry._core import controller_registry


class FusedController:
    CONTROLLER_ID = None

    def __init__(self, config, num_envs, device, mode="robot"):
        self.cfg, self.num_envs, self.device, self.mode = config, num_envs, device, mode

    def init_tensors(self, global_tensor_dict=None):
        pass

    def reset(self):
        pass

    def reset_idx(self, env_ids):
        pass

    def randomize_params(self, env_ids):
        pass  # done by agx_hp1_reset / the fused step (base_lee_controller.py:101-118)

    def update(self, command_actions):
        raise RuntimeError(
            f"{type(self).__name__} is fused into the HP1 CUDA kernel (controller id {self.CONTROLLER_ID}); "
            "step the environment (EnvManager.step / task.step) instead of calling the controller")

    __call__ = update


def _ctrl(name, cid):
    return type(name, (FusedController,), {"CONTROLLER_ID": cid})


NoControl = _ctrl("NoControl", _lib.CTRL_NONE)
LeeAttitudeController = _ctrl("LeeAttitudeController", _lib.CTRL_ATTITUDE)
LeePositionController = _ctrl("LeePositionController", _lib.CTRL_POSITION)
LeeVelocityController = _ctrl("LeeVelocityController", _lib.CTRL_VELOCITY)
LeeAccelerationController = _ctrl("LeeAccelerationController", _lib.CTRL_ACCELERATION)
LeeRatesController = _ctrl("LeeRatesController", _lib.CTRL_RATES)
FullyActuatedController = _ctrl("FullyActuatedController", _lib.CTRL_FULLY_ACTUATED)
LeeVelocitySteeringAngleController = _ctrl("LeeVelocitySteeringAngleController", _lib.CTRL_VELOCITY_STEERING)

controller_registry.register_controller("no_control", NoControl, cc.no_control_config)
_FAMILY = (("position", LeePositionController), ("velocity", LeeVelocityController), ("attitude", LeeAttitudeController),
           ("rates", LeeRatesController), ("acceleration", LeeAccelerationController))


def register_robot_controllers(robot_name=None, controller_config=None):
    for kind, cls in _FAMILY:
        controller_registry.register_controller(f"{robot_name}_{kind}_control", cls, controller_config)


register_robot_controllers("lee", cc.lee_controller_config)
register_robot_controllers("magpie", cc.magpie_controller_config)
register_robot_controllers("lmf2", cc.lmf2_controller_config)
register_robot_controllers("octarotor", cc.lee_controller_config_octarotor)
controller_registry.register_controller("rov_fully_actuated_control", FullyActuatedController, cc.fully_actuated_controller_config)
controller_registry.register_controller("lee_velocity_steering_angle_control", LeeVelocitySteeringAngleController,
                                        cc.lee_controller_config)
