from . import sim_config_registry  # noqa: F401  (reference import path: aerial_gym.registry.sim_registry)
