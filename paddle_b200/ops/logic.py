"""Logic / compare / bitwise ops. Parity: python/paddle/tensor/logic.py."""
from __future__ import annotations

import torch

from ._helpers import T, raw, scalar_or_tensor, wrap


def _cmp(fn):
    def op(x, y, name=None):
        return fn(T(x), scalar_or_tensor(y))

    return op


equal = _cmp(torch.eq)
not_equal = _cmp(torch.ne)
greater_than = _cmp(torch.gt)
greater_equal = _cmp(torch.ge)
less_than = _cmp(torch.lt)
less_equal = _cmp(torch.le)
less = less_than
greater = greater_than


def equal_all(x, y, name=None):
    x, y = raw(x), raw(y)
    return wrap(torch.tensor(x.shape == y.shape and bool(torch.equal(x, y))))


def logical_and(x, y, out=None, name=None):
    return torch.logical_and(T(x), T(y))


def logical_or(x, y, out=None, name=None):
    return torch.logical_or(T(x), T(y))


def logical_xor(x, y, out=None, name=None):
    return torch.logical_xor(T(x), T(y))


def logical_not(x, out=None, name=None):
    return torch.logical_not(T(x))


def bitwise_and(x, y, out=None, name=None):
    return torch.bitwise_and(T(x), scalar_or_tensor(y))


def bitwise_or(x, y, out=None, name=None):
    return torch.bitwise_or(T(x), scalar_or_tensor(y))


def bitwise_xor(x, y, out=None, name=None):
    return torch.bitwise_xor(T(x), scalar_or_tensor(y))


def bitwise_not(x, out=None, name=None):
    return torch.bitwise_not(T(x))


bitwise_invert = bitwise_not


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def is_complex(x):
    return torch.is_complex(T(x))


def is_floating_point(x):
    return torch.is_floating_point(T(x))


def is_integer(x):
    x = T(x)
    return not torch.is_floating_point(x) and not torch.is_complex(x) and x.dtype != torch.bool


def isin(x, test_x, assume_unique=False, invert=False, name=None):
    return torch.isin(T(x), T(test_x), assume_unique=assume_unique, invert=invert)


def _mk_inplace(fn):
    def op(x, *a, **k):
        out = fn(x, *a, **k)
        with torch.no_grad():
            torch.Tensor.copy_(x, out)
        return x

    return op


for _n in ["equal", "not_equal", "greater_than", "greater_equal", "less_than", "less_equal", "logical_and",
           "logical_or", "logical_xor", "logical_not", "bitwise_and", "bitwise_or", "bitwise_xor", "bitwise_not"]:
    globals()[_n + "_"] = _mk_inplace(globals()[_n])

__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "T", "raw", "scalar_or_tensor", "wrap", "annotations")]


# static programs record these as single ops (their bodies compute on raw tensors / read values; framework/recording.py)
from ..framework.recording import make_recordable as _make_recordable  # noqa: E402

_make_recordable(globals(), ['equal_all'])
