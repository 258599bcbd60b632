from .sim_builder import SimBuilder  # noqa: F401
