"""paddle.base (legacy fluid namespace): the commonly imported names. Parity: python/paddle/base/__init__.py."""
from ..framework import unique_name  # noqa: F401
from ..framework.place import CPUPlace, CUDAPinnedPlace, CUDAPlace  # noqa: F401
from ..nn.layer import ParamAttr  # noqa: F401
from ..static import (Executor, Program, Variable, default_main_program, default_startup_program, global_scope, program_guard,  # noqa: F401
                      scope_guard)
from ..tensor import Tensor  # noqa: F401
from . import core, dygraph, framework  # noqa: F401


def in_dygraph_mode():
    from ..static import in_dynamic_mode

    return in_dynamic_mode()


def create_lod_tensor(data, recursive_seq_lens, place=None):
    from ..static.nn.sequence import create_lod_tensor as _c

    return _c(data, recursive_seq_lens, place)
