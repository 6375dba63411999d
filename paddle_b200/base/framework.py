"""paddle.base.framework names."""
from ..framework.place import _default_device as _current_expected_place  # noqa: F401
from ..static import Program, Variable, default_main_program, default_startup_program, in_dynamic_mode, program_guard  # noqa: F401
from ..tensor import EagerParamBase, Parameter  # noqa: F401


def in_dygraph_mode():
    return in_dynamic_mode()


def in_pir_mode():
    return False


def in_dynamic_or_pir_mode():
    return True
