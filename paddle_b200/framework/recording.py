"""Static-graph recording hook shared by `static` (owner) and the kernel wrappers.

`TorchFunctionMode` cannot see `Tensor.as_subclass` or `autograd.Function.apply`, which is exactly what the hand-written
kernel wrappers use; `@recordable` makes such a wrapper appear in the tape as ONE node (run un-traced, recorded whole)."""
from __future__ import annotations

import functools

import torch

current = [None]   # the static.Program being recorded, if any


def recordable(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        prog = current[0]
        if prog is None or not (prog._touches_program(args) or prog._touches_program(kwargs)):
            return fn(*args, **kwargs)
        with torch._C.DisableTorchFunction():
            out = fn(*args, **kwargs)
        prog._record(wrapper, args, kwargs, out)
        return out

    return wrapper
