"""Common utilities shared by every layer: error types + enforce helpers, DDim, flags.
Parity: paddle/common/{enforce.h,errors.h,ddim.h,flags.cc} (L1 of SURVEY.md)."""
from ..framework.flags import flag, get_flags, set_flags  # noqa: F401
from .ddim import DDim, flatten_to_1d, flatten_to_2d, make_ddim, product, slice_ddim, stride, stride_numel, vectorize  # noqa: F401
from .errors import *  # noqa: F401,F403
from .errors import __all__ as _err_all

__all__ = ["DDim", "make_ddim", "product", "slice_ddim", "flatten_to_1d", "flatten_to_2d", "stride", "stride_numel", "vectorize",
           "flag", "get_flags", "set_flags"] + list(_err_all)
