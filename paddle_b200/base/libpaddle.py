"""`paddle.base.libpaddle`: in the reference this is the compiled core module; the names user code reaches through it resolve here -
everything of `base.core` plus the `pir` namespace (Program, PassManager, parse, translate_to_pir) backed by the native IR (csrc/runtime/ir.cpp)."""
from .core import *  # noqa: F401,F403
from .. import pir  # noqa: F401
