"""Shared helpers for the op modules."""
from __future__ import annotations

import numpy as np
import torch

from ..framework import dtype as _dt
from ..framework import place as _place
from ..tensor import Tensor, as_tensor, to_tensor  # noqa: F401


def T(x):
    """Coerce to Tensor (python scalars / numpy / lists allowed)."""
    if isinstance(x, Tensor):
        return x
    if isinstance(x, torch.Tensor):
        return x.as_subclass(Tensor)
    return to_tensor(x)


def raw(x) -> torch.Tensor:
    """Strip to the base torch.Tensor (for python-level torch functions)."""
    if isinstance(x, Tensor):
        return x.as_subclass(torch.Tensor)
    if isinstance(x, torch.Tensor):
        return x
    return to_tensor(x).as_subclass(torch.Tensor)


def wrap(x):
    if isinstance(x, torch.Tensor):
        return x if isinstance(x, Tensor) else x.as_subclass(Tensor)
    if isinstance(x, (tuple, list)):
        return type(x)(wrap(i) for i in x) if type(x) in (tuple, list) else tuple(wrap(i) for i in x)
    return x


def to_int(v):
    if isinstance(v, torch.Tensor):
        return int(v.item())
    if isinstance(v, np.generic):
        return int(v)
    return v


def ax(axis):
    """Normalise an axis argument to int | tuple[int] | None."""
    if axis is None:
        return None
    if isinstance(axis, torch.Tensor):
        axis = axis.tolist()
    if isinstance(axis, (list, tuple)):
        return tuple(to_int(a) for a in axis)
    return int(axis)


def shp(shape):
    """Normalise a shape argument to a list of ints."""
    if shape is None:
        return None
    if isinstance(shape, torch.Tensor):
        return [int(s) for s in shape.tolist()] if shape.dim() > 0 else [int(shape.item())]
    if isinstance(shape, (int, np.integer)):
        return [int(shape)]
    return [to_int(s) for s in shape]


def dt(dtype, default=None):
    d = _dt.convert_dtype(dtype)
    return d if d is not None else default


def dev(place=None):
    return _place.to_torch_device(place)


def scalar_or_tensor(v):
    """Keep python scalars as-is (torch handles them with proper type promotion)."""
    if isinstance(v, (int, float, bool, complex)):
        return v
    if isinstance(v, np.generic):
        return v.item()
    return T(v)


def binary_args(x, y):
    """paddle binary ops accept tensor/scalar mixes; scalar-scalar becomes tensors."""
    xs, ys = scalar_or_tensor(x), scalar_or_tensor(y)
    if not isinstance(xs, torch.Tensor) and not isinstance(ys, torch.Tensor):
        xs = to_tensor(xs)
    return xs, ys
