from . import task_registry  # noqa: F401  (reference import path: aerial_gym.registry.task_registry)
