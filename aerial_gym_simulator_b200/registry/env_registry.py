from . import env_config_registry  # noqa: F401  (reference import path: aerial_gym.registry.env_registry)
