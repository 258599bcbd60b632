from ._core import controller_registry  # noqa: F401
