"""Tensor op namespace + Tensor method patching.

Parity: python/paddle/tensor/__init__.py (tensor_method_func list) and
python/paddle/base/dygraph/math_op_patch.py.
"""
from __future__ import annotations

import torch

from ..tensor import Tensor
from . import creation, einsum as _einsum_mod, linalg, logic, manipulation, math, random, search, stat
from .creation import *  # noqa: F401,F403
from .einsum import einsum  # noqa: F401
from .linalg import *  # noqa: F401,F403
from .logic import *  # noqa: F401,F403
from .manipulation import *  # noqa: F401,F403
from .math import *  # noqa: F401,F403
from .random import *  # noqa: F401,F403
from .search import *  # noqa: F401,F403
from .stat import *  # noqa: F401,F403
from .extras import (LazyGuard, check_shape, disable_signal_handler, masked_scatter_, pdist, set_printoptions, sgn, sinc_, t_, transpose_,  # noqa: F401
                     tril_, triu_)

# names that must NOT become Tensor methods (list-first-arg functions, property clashes, torch-internal contracts)
_NO_METHOD = {
    "shape", "numel", "concat", "stack", "hstack", "vstack", "dstack", "column_stack", "row_stack", "add_n",
    "meshgrid", "broadcast_tensors", "multiplex", "broadcast_shape", "block_diag", "cartesian_prod", "is_tensor",
    "where", "where_", "einsum", "multi_dot", "scatter_nd", "to_tensor", "zeros", "ones", "full", "empty", "arange",
    "linspace", "logspace", "eye", "tril_indices", "triu_indices", "create_parameter", "create_tensor", "range",
    "from_numpy", "fill_constant", "complex", "polar", "assign", "rand", "randn", "standard_normal", "normal",
    "uniform", "randint", "randperm", "log_normal", "binomial", "fp8_fp8_half_gemm_fused", "histogramdd", "tolist",
    "lu_solve", "atleast_1d", "atleast_2d", "atleast_3d", "rank",
}


def _patch():
    for mod in (math, manipulation, logic, search, stat, linalg, creation, random):
        for name in getattr(mod, "__all__", []):
            if name in _NO_METHOD:
                continue
            fn = getattr(mod, name)
            if callable(fn):
                setattr(Tensor, name, fn)
    Tensor.where = lambda self, x=None, y=None, name=None: search.where(self, x, y)
    Tensor.einsum = None
    del Tensor.einsum
    # operators with paddle scalar/ndarray tolerance are inherited from torch.Tensor; only the few that differ:
    Tensor.__matmul__ = lambda a, b: linalg.matmul(a, b)
    Tensor.__rmatmul__ = lambda a, b: linalg.matmul(b, a)
    Tensor.__floordiv__ = lambda a, b: math.floor_divide(a, b)
    Tensor.__mod__ = lambda a, b: math.remainder(a, b)
    Tensor.__invert__ = lambda a: torch.logical_not(a) if a.dtype == torch.bool else torch.bitwise_not(a)


_patch()
from . import method_extras  # noqa: E402,F401
