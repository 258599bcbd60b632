"""Drop-in import names.  ``install()`` makes the reference's import paths resolve to this
package so its trainers / examples run unchanged:

    import aerial_gym_simulator_b200.compat as compat; compat.install()
    import isaacgym                                   # shim: gymapi / gymutil / gymtorch names only
    from aerial_gym.registry.task_registry import task_registry
    from aerial_gym.sim.sim_builder import SimBuilder
    from aerial_gym.utils.helpers import parse_arguments

(reference: aerial_gym/__init__.py imports isaacgym first; rl_training/* import the paths above.)"""
import sys
import types


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if getattr(sys.modules.get("aerial_gym"), "_b200_compat", False):
        return
    from . import config, control, env_manager, registry, robots, sensors, sim, task, utils
    from .sensors import imu_sensor
    from .task import (lidar_navigation_task, navigation_task, position_setpoint_task,
                       position_setpoint_task_sim2real as s2r_task, position_setpoint_task_sim2real_end_to_end as e2e_task,
                       radar_navigation_task)
    from .utils import vae_encoder
    from .config import (PACKAGE_DIRECTORY, asset_config, controller_config, env_config, robot_config, sensor_config,
                         sim_config, task_config)
    import importlib

    R = {n: importlib.import_module(f"{__package__}.registry.{n}") for n in
         ("task_registry", "robot_registry", "controller_registry", "env_registry", "sim_registry")}
    from .sim import sim_builder
    from .utils import helpers, logging, math

    # ---- isaacgym shim -----------------------------------------------------------------------
    gymapi = _module("isaacgym.gymapi", SIM_PHYSX=1, SIM_FLEX=0, LOCAL_SPACE=1, ENV_SPACE=0, GLOBAL_SPACE=2,
                     UP_AXIS_Y=0, UP_AXIS_Z=1, STATE_ALL=3)
    gymutil = _module("isaacgym.gymutil", parse_device_str=helpers.parse_device_str)
    gymtorch = _module("isaacgym.gymtorch", wrap_tensor=lambda t: t, unwrap_tensor=lambda t: t)
    _module("isaacgym", gymapi=gymapi, gymutil=gymutil, gymtorch=gymtorch, __path__=[])
    # ---- aerial_gym alias ----------------------------------------------------------------------
    pkg = _module("aerial_gym", AERIAL_GYM_DIRECTORY=PACKAGE_DIRECTORY, _b200_compat=True, __path__=[])
    table = {
        "registry": registry, "registry.task_registry": R["task_registry"], "registry.robot_registry": R["robot_registry"],
        "registry.controller_registry": R["controller_registry"], "registry.env_registry": R["env_registry"],
        "registry.sim_registry": R["sim_registry"], "sim": sim, "sim.sim_builder": sim_builder, "task": task,
        "env_manager": env_manager, "robots": robots, "control": control, "utils": utils, "utils.helpers": helpers,
        "utils.logging": logging, "utils.math": math, "config": config, "config.sim_config": sim_config,
        "config.env_config": env_config, "config.robot_config": robot_config, "config.controller_config": controller_config,
        "config.sensor_config": sensor_config, "config.asset_config": asset_config, "config.task_config": task_config,
        "sensors": sensors, "sensors.imu_sensor": imu_sensor,
        "task.navigation_task": _module("aerial_gym.task.navigation_task", navigation_task=navigation_task, __path__=[]),
        "task.navigation_task.navigation_task": navigation_task,
        "task.lidar_navigation_task": _module("aerial_gym.task.lidar_navigation_task", lidar_navigation_task=lidar_navigation_task,
                                              __path__=[]),
        "task.lidar_navigation_task.lidar_navigation_task": lidar_navigation_task,
        "task.position_setpoint_task": _module("aerial_gym.task.position_setpoint_task", position_setpoint_task=position_setpoint_task,
                                               __path__=[]),
        "task.position_setpoint_task.position_setpoint_task": position_setpoint_task,
        "task.position_setpoint_task_sim2real_end_to_end": _module(
            "aerial_gym.task.position_setpoint_task_sim2real_end_to_end", position_setpoint_task_sim2real_end_to_end=e2e_task, __path__=[]),
        "task.position_setpoint_task_sim2real_end_to_end.position_setpoint_task_sim2real_end_to_end": e2e_task,
        "task.position_setpoint_task_sim2real_px4": _module(
            "aerial_gym.task.position_setpoint_task_sim2real_px4", position_setpoint_task_sim2real_px4=e2e_task, __path__=[]),
        "task.position_setpoint_task_sim2real_px4.position_setpoint_task_sim2real_px4": e2e_task,
        "task.radar_navigation_task": _module("aerial_gym.task.radar_navigation_task", radar_navigation_task=radar_navigation_task, __path__=[]),
        "task.radar_navigation_task.radar_navigation_task": radar_navigation_task,
        "task.position_setpoint_task_sim2real": _module(
            "aerial_gym.task.position_setpoint_task_sim2real", position_setpoint_task_sim2real=s2r_task, __path__=[]),
        "task.position_setpoint_task_sim2real.position_setpoint_task_sim2real": s2r_task,
        "task.position_setpoint_task_acceleration_sim2real": _module(
            "aerial_gym.task.position_setpoint_task_acceleration_sim2real", position_setpoint_task_acceleration_sim2real=s2r_task, __path__=[]),
        "task.position_setpoint_task_acceleration_sim2real.position_setpoint_task_acceleration_sim2real": s2r_task,
        "utils.vae": _module("aerial_gym.utils.vae", vae_image_encoder=vae_encoder, __path__=[]),
        "utils.vae.vae_image_encoder": vae_encoder,
    }
    for name, mod in table.items():
        sys.modules["aerial_gym." + name] = mod
        if "." not in name:
            setattr(pkg, name, mod)
    # the reference's one-class-per-file config paths: aerial_gym.config.<kind>_config[.<sub>].<file> (compat_paths.py)
    from .compat_paths import CONFIG_MODULES
    flat = {"sim_config": sim_config, "env_config": env_config, "robot_config": robot_config, "controller_config": controller_config,
            "sensor_config": sensor_config, "asset_config": asset_config, "task_config": task_config}
    for rel, names in CONFIG_MODULES.items():
        parts = rel.split(".")
        for i in range(2, len(parts)):  # intermediate packages (sensor_config.lidar_config, ...)
            mid = "aerial_gym.config." + ".".join(parts[:i])
            if mid not in sys.modules:
                _module(mid, __path__=[])
        attrs = {}
        for name, where in names.items():
            mod_name, attr = where.split(".")
            attrs[name] = getattr(flat[mod_name], attr)
        _module("aerial_gym.config." + rel, **attrs)
