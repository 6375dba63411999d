"""Remaining top-level names of python/paddle/__init__.py (__all__): in-place variants, pdist, sgn, printing / misc switches."""
from __future__ import annotations

import contextlib

import torch

from ._helpers import T, wrap


def _inplace(x, out):
    with torch.no_grad():
        if tuple(out.shape) != tuple(x.shape):
            with torch._C.DisableTorchFunctionSubclass():
                x.set_(out.detach().contiguous())   # shape-changing in-place ops rebind the storage (paddle semantics)
        else:
            torch.Tensor.copy_(x, out)
    return x


def t_(input, name=None):
    x = T(input)
    return _inplace(x, x.t().contiguous()) if x.dim() == 2 else x


def transpose_(x, perm, name=None):
    x = T(x)
    return _inplace(x, torch.permute(x, list(perm)).contiguous())


def triu_(x, diagonal=0, name=None):
    x = T(x)
    return _inplace(x, torch.triu(x, diagonal))


def tril_(x, diagonal=0, name=None):
    x = T(x)
    return _inplace(x, torch.tril(x, diagonal))


def sinc_(x, name=None):
    x = T(x)
    return _inplace(x, torch.sinc(x))


def masked_scatter_(x, mask, value, name=None):
    x = T(x)
    return _inplace(x, torch.masked_scatter(x, T(mask).to(torch.bool).expand_as(x), T(value).to(x.dtype)))


def sgn(x, name=None):
    """sign for real tensors, x / |x| for complex ones. Parity: tensor/math.py:sgn."""
    return torch.sgn(T(x))


def pdist(x, p=2.0, compute_mode="use_mm_for_euclid_dist_if_necessary", name=None):
    """Condensed pairwise distances between the rows of x [N, D] -> [N*(N-1)/2]. Parity: nn/functional/distance.py:pdist."""
    return torch.nn.functional.pdist(T(x), p=float(p))


def set_printoptions(precision=None, threshold=None, edgeitems=None, sci_mode=None, linewidth=None):
    kw = {k: v for k, v in dict(precision=precision, threshold=threshold, edgeitems=edgeitems, sci_mode=sci_mode, linewidth=linewidth).items() if v is not None}
    torch.set_printoptions(**kw)


def disable_signal_handler():
    """The reference installs C++ signal handlers for stack dumps; there is nothing to uninstall here."""


def check_shape(shape, op_name="", expected_shape_type=(list, tuple), expected_element_type=(int,), expected_tensor_dtype=("int32", "int64")):
    if isinstance(shape, torch.Tensor):
        if str(shape.dtype).split(".")[-1] not in expected_tensor_dtype:
            raise TypeError(f"{op_name}: shape tensor must be int32/int64")
        return
    if not isinstance(shape, expected_shape_type):
        raise TypeError(f"{op_name}: shape must be a list / tuple / Tensor, got {type(shape)}")
    for s in shape:
        if not isinstance(s, (*expected_element_type, torch.Tensor)):
            raise TypeError(f"{op_name}: shape elements must be int or Tensor, got {type(s)}")


class LazyGuard(contextlib.ContextDecorator):
    """Parity: python/paddle/lazy_init.py:LazyGuard. Parameters created inside the guard are allocated on the meta device and
    materialised by `layer.to(device)` / set_state_dict, so a model larger than host memory can be declared first."""

    def __enter__(self):
        from ..nn import layer as _layer

        self._prev = getattr(_layer, "_lazy_init", [False])
        _layer._lazy_init = [True]
        return self

    def __exit__(self, *a):
        from ..nn import layer as _layer

        _layer._lazy_init = self._prev
        return False
