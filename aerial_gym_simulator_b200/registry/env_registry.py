from ._core import env_config_registry  # noqa: F401
