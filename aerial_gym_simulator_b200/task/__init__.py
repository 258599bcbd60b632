"""Task registry surface (reference: aerial_gym/task/__init__.py)."""
from ..config.task_config import position_setpoint_task_config
from ..registry._core import task_registry
from .position_setpoint_task import PositionSetpointTask

task_registry.register_task("position_setpoint_task", PositionSetpointTask, position_setpoint_task_config)
