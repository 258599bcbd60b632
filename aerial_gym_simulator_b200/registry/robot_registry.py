from . import robot_registry  # noqa: F401  (reference import path: aerial_gym.registry.robot_registry)
