from ._core import sim_config_registry  # noqa: F401
