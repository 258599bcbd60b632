"""aerial_gym/sim/__init__.py surface: SimBuilder plus the names the reference's package exports (the registrations themselves
happen in env_manager/__init__.py)."""
from .. import env_manager as _env_manager  # noqa: F401  (registers the sim configs)
from ..config.sim_config import BaseSimConfig, BaseSimHeadlessConfig, SimCfg2Ms, SimCfg4Ms  # noqa: F401
from ..registry._core import sim_config_registry  # noqa: F401
from .sim_builder import SimBuilder  # noqa: F401
