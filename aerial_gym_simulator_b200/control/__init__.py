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

    # ---- gains (base_lee_controller.py:33-62, 78-85): the reference keeps [N,3] tensors on the controller object and lets callers
    # read and overwrite them (examples/tune_controllers.py).  Here they live in the engine: as per-env arrays when they are
    # randomised (or the env was built with args={"per_env_params": "all"}), else as one constant per axis in AgxHp1Config.
    _GAINS = {"K_pos_tensor_current": "K_pos", "K_linvel_tensor_current": "K_vel", "K_rot_tensor_current": "K_rot",
              "K_angvel_tensor_current": "K_angvel"}

    def bind_engine(self, engine):
        self._engine = engine

    def __getattr__(self, name):
        gains = type(self)._GAINS
        if name in gains:
            eng = self.__dict__.get("_engine")
            if eng is None:
                raise AttributeError(f"{name}: the controller is not attached to an engine yet (build the env first)")
            arr = getattr(eng, gains[name])
            if arr is not None:
                return arr  # the live [N,3] array the kernel reads
            import torch
            const = torch.tensor(list(getattr(eng.cfg, gains[name])), dtype=torch.float32, device=eng.root_state.device)
            return const.expand(self.num_envs, 3)  # a read-only view of the constant; write through set_controller_gains
        if name.startswith("K_") and name.endswith(("_tensor_min", "_tensor_max")):
            import torch
            key = name.replace("K_linvel", "K_vel")
            return torch.tensor(getattr(self.cfg, key), dtype=torch.float32, device=self.device).expand(self.num_envs, -1)
        raise AttributeError(name)

    def set_controller_gains(self, K_pos, K_vel, K_rot, K_angvel):
        import torch
        eng = self.__dict__.get("_engine")
        if eng is None:
            raise RuntimeError("set_controller_gains: the controller is not attached to an engine yet (build the env first)")
        for field, value in (("K_pos", K_pos), ("K_vel", K_vel), ("K_rot", K_rot), ("K_angvel", K_angvel)):
            value = torch.as_tensor(value, dtype=torch.float32)
            arr = getattr(eng, field)
            if arr is not None:
                arr[:] = value.to(arr.device)
                continue
            rows = value.reshape(-1, 3) if value.numel() > 1 else value.reshape(1, 1).expand(1, 3)
            if not bool((rows == rows[0]).all()):
                raise RuntimeError(f"set_controller_gains: {field} differs between envs but this env keeps one constant per axis; "
                                   "build it with args={'per_env_params': 'all'} for per-env gains")
            for i in range(3):
                getattr(eng.cfg, field)[i] = float(rows[0, i])

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
