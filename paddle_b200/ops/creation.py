"""Creation ops. Parity: python/paddle/tensor/creation.py."""
from __future__ import annotations

import numpy as np
import torch

from ..framework import dtype as _dt
from ._helpers import T, ax, dev, dt, raw, shp, to_int, to_tensor, wrap

__all__ = [
    "to_tensor", "zeros", "ones", "full", "empty", "zeros_like", "ones_like", "full_like", "empty_like",
    "arange", "linspace", "logspace", "eye", "tril", "triu", "tril_indices", "triu_indices", "meshgrid",
    "diag", "diagflat", "diag_embed", "assign", "clone", "complex", "polar", "cauchy_", "geometric_",
    "create_parameter", "create_tensor", "range", "vander", "from_numpy", "fill_constant",
]


from ..framework.recording import recordable as _recordable  # noqa: E402

def _fd(dtype):
    return dt(dtype, _dt.default_dtype())


@_recordable
def zeros(shape, dtype=None, name=None):
    return wrap(torch.zeros(shp(shape), dtype=_fd(dtype), device=dev()))


@_recordable
def ones(shape, dtype=None, name=None):
    return wrap(torch.ones(shp(shape), dtype=_fd(dtype), device=dev()))


@_recordable
def full(shape, fill_value, dtype=None, name=None):
    fill_value = to_int(fill_value) if not isinstance(fill_value, torch.Tensor) else fill_value.item()
    if dtype is None:
        if isinstance(fill_value, bool):
            d = torch.bool
        elif isinstance(fill_value, int):
            d = torch.int64
        elif isinstance(fill_value, type(1j)):   # `complex` is shadowed by paddle.complex below
            d = torch.complex64
        else:
            d = _dt.default_dtype()
    else:
        d = dt(dtype)
    return wrap(torch.full(shp(shape), fill_value, dtype=d, device=dev()))


def fill_constant(shape, dtype, value, force_cpu=False, out=None, name=None):
    return full(shape, value, dtype)


@_recordable
def empty(shape, dtype=None, name=None):
    return wrap(torch.empty(shp(shape), dtype=_fd(dtype), device=dev()))


@_recordable
def zeros_like(x, dtype=None, name=None):
    return wrap(torch.zeros_like(raw(x), dtype=dt(dtype)))


@_recordable
def ones_like(x, dtype=None, name=None):
    return wrap(torch.ones_like(raw(x), dtype=dt(dtype)))


@_recordable
def full_like(x, fill_value, dtype=None, name=None):
    return wrap(torch.full_like(raw(x), to_int(fill_value), dtype=dt(dtype)))


@_recordable
def empty_like(x, dtype=None, name=None):
    return wrap(torch.empty_like(raw(x), dtype=dt(dtype)))


@_recordable
def arange(start=0, end=None, step=1, dtype=None, name=None):
    start, end, step = to_int(start), to_int(end), to_int(step)
    if end is None:
        start, end = 0, start
    if dtype is None:
        d = torch.int64 if all(isinstance(v, (int, np.integer)) for v in (start, end, step)) else _dt.default_dtype()
    else:
        d = dt(dtype)
    return wrap(torch.arange(start, end, step, dtype=d, device=dev()))


range = arange  # noqa: A001


@_recordable
def linspace(start, stop, num, dtype=None, name=None):
    return wrap(torch.linspace(to_int(start), to_int(stop), int(to_int(num)), dtype=_fd(dtype), device=dev()))


def logspace(start, stop, num, base=10.0, dtype=None, name=None):
    return wrap(torch.logspace(to_int(start), to_int(stop), int(to_int(num)), base=to_int(base), dtype=_fd(dtype), device=dev()))


@_recordable
def eye(num_rows, num_columns=None, dtype=None, name=None):
    n = int(to_int(num_rows))
    m = n if num_columns is None else int(to_int(num_columns))
    return wrap(torch.eye(n, m, dtype=_fd(dtype), device=dev()))


def tril(x, diagonal=0, name=None):
    return torch.tril(T(x), diagonal)


def triu(x, diagonal=0, name=None):
    return torch.triu(T(x), diagonal)


@_recordable
def tril_indices(row, col=None, offset=0, dtype="int64"):
    col = row if col is None else col
    return wrap(torch.tril_indices(row, col, offset, dtype=dt(dtype), device=dev()))


@_recordable
def triu_indices(row, col=None, offset=0, dtype="int64"):
    col = row if col is None else col
    return wrap(torch.triu_indices(row, col, offset, dtype=dt(dtype), device=dev()))


def meshgrid(*args, **kwargs):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    return list(torch.meshgrid(*[T(a) for a in args], indexing="ij"))


def diag(x, offset=0, padding_value=0, name=None):
    x = T(x)
    if x.dim() == 1 and padding_value != 0:
        n = x.size(0) + abs(offset)
        out = torch.full((n, n), padding_value, dtype=x.dtype, device=x.device)
        idx = torch.arange(x.size(0), device=x.device)
        r, c = (idx, idx + offset) if offset >= 0 else (idx - offset, idx)
        out[r, c] = x
        return wrap(out)
    return torch.diag(x, offset)


def diagflat(x, offset=0, name=None):
    return torch.diagflat(T(x), offset)


def diag_embed(input, offset=0, dim1=-2, dim2=-1):
    return torch.diag_embed(T(input), offset, dim1, dim2)


def assign(x, output=None):
    x = T(x) if not isinstance(x, (list, tuple, np.ndarray, int, float, bool)) else to_tensor(np.asarray(x))
    if output is None:
        return torch.clone(x)
    with torch.no_grad():
        torch.Tensor.copy_(output, x)
    return output


def clone(x, name=None):
    return torch.clone(T(x))


def complex(real, imag, name=None):  # noqa: A001
    return torch.complex(T(real), T(imag))


def polar(abs, angle, name=None):  # noqa: A002
    return torch.polar(T(abs), T(angle))


def cauchy_(x, loc=0, scale=1, name=None):
    return torch.Tensor.cauchy_(x, loc, scale)


def geometric_(x, probs, name=None):
    return torch.Tensor.geometric_(x, probs)


def vander(x, n=None, increasing=False, name=None):
    return torch.vander(T(x), N=n, increasing=increasing)


def from_numpy(a):
    return to_tensor(a)


def create_parameter(shape, dtype, name=None, attr=None, is_bias=False, default_initializer=None):
    """paddle.create_parameter. Parity: python/paddle/tensor/creation.py:create_parameter."""
    from ..nn.layer import _make_parameter

    return _make_parameter(shp(shape), dt(dtype), attr=attr, is_bias=is_bias, default_initializer=default_initializer, name=name)


def create_tensor(dtype, name=None, persistable=False):
    t = wrap(torch.empty(0, dtype=dt(dtype), device=dev()))
    t.persistable = persistable
    if name:
        t.name = name
    return t
