from . import controller_registry  # noqa: F401  (reference import path: aerial_gym.registry.controller_registry)
