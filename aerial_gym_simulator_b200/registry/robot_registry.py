from ._core import robot_registry  # noqa: F401
