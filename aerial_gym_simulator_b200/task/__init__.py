"""Task registry surface (reference: aerial_gym/task/__init__.py)."""
from ..config.task_config import (radar_navigation_task_config, lidar_navigation_task_config, navigation_task_config, position_setpoint_task_config,
                                  position_setpoint_task_sim2real_end_to_end_config, position_setpoint_task_sim2real_px4_config,
                                  position_setpoint_task_sim2real_config, position_setpoint_task_acceleration_sim2real_config)
from ..registry._core import task_registry
from .lidar_navigation_task import LiDARNavigationTask
from .navigation_task import NavigationTask
from .position_setpoint_task import PositionSetpointTask
from .radar_navigation_task import RadarNavigationTask
from .position_setpoint_task_sim2real import PositionSetpointTaskAccelerationSim2Real, PositionSetpointTaskSim2Real
from .position_setpoint_task_sim2real_end_to_end import PositionSetpointTaskSim2RealEndToEnd, PositionSetpointTaskSim2RealPX4

task_registry.register_task("position_setpoint_task", PositionSetpointTask, position_setpoint_task_config)
task_registry.register_task("navigation_task", NavigationTask, navigation_task_config)
task_registry.register_task("lidar_navigation_task", LiDARNavigationTask, lidar_navigation_task_config)
task_registry.register_task("position_setpoint_task_sim2real_end_to_end", PositionSetpointTaskSim2RealEndToEnd,
                            position_setpoint_task_sim2real_end_to_end_config)
task_registry.register_task("position_setpoint_task_sim2real_px4", PositionSetpointTaskSim2RealPX4, position_setpoint_task_sim2real_px4_config)
task_registry.register_task("radar_navigation_task", RadarNavigationTask, radar_navigation_task_config)
task_registry.register_task("position_setpoint_task_sim2real", PositionSetpointTaskSim2Real, position_setpoint_task_sim2real_config)
task_registry.register_task("position_setpoint_task_acceleration_sim2real", PositionSetpointTaskAccelerationSim2Real,
                            position_setpoint_task_acceleration_sim2real_config)
