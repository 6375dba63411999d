"""Static-graph recording hook shared by `static` (owner) and the kernel wrappers.

`TorchFunctionMode` cannot see `Tensor.as_subclass` or `autograd.Function.apply`, which is exactly what the hand-written
kernel wrappers use; `@recordable` makes such a wrapper appear in the tape as ONE node (run un-traced, recorded whole)."""
from __future__ import annotations

import functools

import torch

current = [None]   # the static.Program being recorded, if any
_inside = [0]      # > 0 while the body of a recordable runs: nested recordables belong to the enclosing node


def _rewrap(o):
    """Bodies run with torch-function dispatch off, so ops that rely on it to keep the paddle Tensor type hand back base tensors."""
    if isinstance(o, torch.Tensor):
        if type(o) is torch.Tensor:
            from ..tensor import Tensor

            return o.as_subclass(Tensor)
        return o
    if isinstance(o, (list, tuple)):
        return type(o)(_rewrap(i) for i in o)
    return o


def recordable(fn=None, *, always=False):
    """Decorator.  `always`: record the call even when no argument is a value of the program - for ops whose result must be produced anew at every
    run although nothing feeds them (random sampling: a program would otherwise bake ONE sample in)."""
    if fn is None:
        return lambda f: recordable(f, always=always)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        prog = current[0]
        if prog is None or _inside[0] or not (always or prog._touches_program(args) or prog._touches_program(kwargs)):
            return fn(*args, **kwargs)
        _inside[0] += 1
        try:
            with torch._C.DisableTorchFunction():
                out = fn(*args, **kwargs)
        finally:
            _inside[0] -= 1
        out = _rewrap(out)
        prog._record(wrapper, args, kwargs, out)
        return out

    return wrapper


def make_recordable(namespace, names, always=False):
    """Wrap the named module-level functions of `namespace` (a module's globals()) as recordable ops.  For modules whose public functions compute on
    raw tensors (fast paths into kernels / index arithmetic) - without this a static program would silently bake its placeholder values in."""
    for name in names:
        fn = namespace.get(name)
        if callable(fn) and not isinstance(fn, type) and not getattr(fn, "_b200_recordable", False):
            w = recordable(fn, always=always)
            w._b200_recordable = True
            namespace[name] = w
