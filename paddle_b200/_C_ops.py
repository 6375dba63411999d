"""`paddle._C_ops`: final-state op entry points called positionally (`_C_ops.matmul(x, y, False, False)`).

Parity: python/paddle/_C_ops.py (the generated eager op bindings of core.eager.ops).  An attribute lookup resolves, in order: a kernel of the
KernelFactory (`kernels/registry.py`: selection by the first tensor's backend / dtype), the op schema registry (`ops/schema.py`, every public op
of `paddle_b200.ops.*`), then the functional namespaces.  Arguments are positional in schema order, as in the reference."""
from __future__ import annotations


def _resolve(name):
    from .kernels.registry import KernelFactory

    f = KernelFactory.instance()
    if f.has_kernel(name):
        def op(*args, **kwargs):
            return f.dispatch(name, *args, **kwargs)

        op.__name__ = name
        return op
    from .ops import schema

    try:
        fn = schema.get(name).func
        if fn is not None:
            return fn
    except KeyError:
        pass
    import importlib

    for mod in ("nn.functional", "incubate.nn.functional", "linalg", "fft", "signal", "sparse", "geometric"):
        try:
            m = importlib.import_module("paddle_b200." + mod)
        except ImportError:
            continue
        if hasattr(m, name):
            return getattr(m, name)
    base = name[:-1] if name.endswith("_") else None
    if base:
        from .tensor import Tensor

        if hasattr(Tensor, name):
            return getattr(Tensor, name)
    raise AttributeError(f"paddle_b200._C_ops has no op '{name}'")


_cache = {}


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    if name not in _cache:
        _cache[name] = _resolve(name)
    return _cache[name]
