"""Parity: python/paddle/distributed/communication/group.py."""
from ..collective import Group, barrier, get_backend, get_group, new_group, wait  # noqa: F401
from ..env import destroy_process_group, is_initialized  # noqa: F401


def _get_global_group():
    return get_group(0)
