"""Registries (reference: aerial_gym/registry/*.py).  Import the instances from the submodules,
exactly like the reference does: ``from ...registry.task_registry import task_registry``."""
