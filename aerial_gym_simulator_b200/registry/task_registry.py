from ._core import task_registry  # noqa: F401
