"""paddle.utils. Parity: python/paddle/utils/__init__.py."""
from __future__ import annotations

import functools
import importlib
import warnings

from ..framework import unique_name  # noqa: F401
from . import cpp_extension, dlpack, download  # noqa: F401


def deprecated(update_to="", since="", reason="", level=0):
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*a, **k):
            msg = f"API '{fn.__module__}.{fn.__name__}' is deprecated since {since}" + (f", use '{update_to}' instead" if update_to else "") + (f". {reason}" if reason else "")
            if level >= 2:
                raise RuntimeError(msg)
            if level == 1 or level == 0:
                warnings.warn(msg, category=DeprecationWarning, stacklevel=2)
            return fn(*a, **k)

        return wrapper

    return deco


def try_import(module_name, err_msg=None):
    try:
        return importlib.import_module(module_name)
    except ImportError as e:
        raise ImportError(err_msg or f"Failed importing {module_name}. This likely means that some paddle modules require additional dependencies.") from e


def require_version(min_version, max_version=None):
    from .. import __version__

    def key(v):
        return tuple(int(x) for x in str(v).split(".")[:3] if x.isdigit())

    cur = key(__version__)
    if cur == (0, 0, 0):   # development build: every requirement is accepted, as in the reference
        return
    if cur < key(min_version) or (max_version is not None and cur > key(max_version)):
        raise Exception(f"paddle version {__version__} does not satisfy the requirement [{min_version}, {max_version or 'inf'}]")


def run_check():
    """paddle.utils.run_check(): trains a tiny model on every visible device through the native kernels."""
    import torch

    from .. import nn, optimizer, randn, set_device

    devs = ["cpu"] + ([f"gpu:{i}" for i in range(torch.cuda.device_count())] if torch.cuda.is_available() else [])
    for d in devs[-1:]:
        set_device(d)
        m = nn.Linear(8, 8)
        o = optimizer.SGD(0.1, parameters=m.parameters())
        x = randn([4, 8])
        l = (m(x) ** 2).mean()
        l.backward()
        o.step()
        print(f"paddle_b200 works on {d}.")
    if torch.cuda.is_available():
        from .._build import load

        print("native sm_100a extension:", "loaded" if load() is not None else "NOT BUILT")
    print("PaddlePaddle-compatible paddle_b200 is installed successfully!")


def flops(net, input_size, custom_ops=None, print_detail=False):
    from ..hapi.summary import flops as f

    return f(net, input_size, custom_ops, print_detail)
