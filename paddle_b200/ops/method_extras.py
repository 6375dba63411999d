"""Remaining Tensor attributes of the reference that are not plain op functions.
Parity: python/paddle/base/dygraph/tensor_patch_methods.py (apply / _grad_ivar / to_dlpack / ...), paddle/fluid/pybind/eager_method.cc
(_share_buffer_to, _is_shared_buffer_with, _clear_data, _slice, inplace_version, strides / offset, rows / is_selected_rows) and the
function-style entries of tensor_method_func (concat, stack, add_n, ...) that the reference also exposes as methods."""
from __future__ import annotations

import hashlib

import torch

from ..tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor)


def _apply(self, func):
    """Out-of-place user function on the tensor (the reference refuses tensors that require grad; so do we)."""
    if not self.stop_gradient:
        raise RuntimeError("Cannot apply function on a tensor that stop_gradient=False.")
    return func(self)


def _apply_(self, func):
    if not self.stop_gradient:
        raise RuntimeError("Cannot apply function on a tensor that stop_gradient=False.")
    out = func(self)
    with torch.no_grad():
        torch.Tensor.copy_(self, _raw(out))
    return self


def _fill_diagonal_tensor(self, y, offset=0, dim1=0, dim2=1, name=None):
    out = _raw(self).clone()
    torch.diagonal(out, offset, dim1, dim2).copy_(_raw(y))
    return out.as_subclass(Tensor)


def _fill_diagonal_tensor_(self, y, offset=0, dim1=0, dim2=1, name=None):
    with torch.no_grad():
        torch.diagonal(_raw(self), offset, dim1, dim2).copy_(_raw(y))
    return self


def _share_buffer_to(self, dst):
    """dst becomes an alias of this tensor's storage."""
    torch.Tensor.set_(dst, _raw(self).detach())
    return dst


def _is_shared_buffer_with(self, other):
    a, b = _raw(self), _raw(other)
    return a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr() and a.untyped_storage().data_ptr() != 0


def _clear_data(self):
    torch.Tensor.set_(self, torch.empty(0, dtype=self.dtype, device=self.device))


def _slice(self, begin, end):
    return _raw(self)[begin:end].as_subclass(Tensor)


def _md5sum(self):
    return hashlib.md5(_raw(self).detach().cpu().contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def _to_dlpack(self):
    return torch.utils.dlpack.to_dlpack(_raw(self).detach())


def _retain_grads(self):
    if self.grad_fn is not None:
        torch.Tensor.retain_grad(self)
    return self


def _matrix_transpose(self, name=None):
    return torch.transpose(self, -2, -1)               # on the tensor itself: recorded as an op in static programs


def _is_same_shape(self, other):
    return list(self.shape) == list(other.shape)


def patch():
    from . import creation, linalg, manipulation, math, search   # noqa: F401
    from .. import ops

    for name in ("add_n", "atleast_1d", "atleast_2d", "atleast_3d", "block_diag", "broadcast_shape", "broadcast_tensors", "concat", "stack", "create_parameter",
                 "create_tensor", "histogramdd", "is_tensor", "multi_dot", "multiplex", "polar", "scatter_nd", "where_"):
        fn = getattr(ops, name, None)
        if fn is not None and not hasattr(Tensor, name):
            setattr(Tensor, name, fn)
    Tensor.apply, Tensor.apply_ = _apply, _apply_
    Tensor.fill_diagonal_tensor, Tensor.fill_diagonal_tensor_ = _fill_diagonal_tensor, _fill_diagonal_tensor_
    Tensor._share_buffer_to, Tensor._is_shared_buffer_with, Tensor._clear_data, Tensor._slice = _share_buffer_to, _is_shared_buffer_with, _clear_data, _slice
    Tensor._md5sum, Tensor.to_dlpack, Tensor.retain_grads = _md5sum, _to_dlpack, _retain_grads
    Tensor.matrix_transpose, Tensor.is_same_shape = _matrix_transpose, _is_same_shape
    Tensor._grad_ivar = lambda self: self.grad
    Tensor._numel = lambda self: int(torch.Tensor.numel(self))
    Tensor._use_gpudnn = lambda self, use=True: self
    Tensor.inplace_version = property(lambda self: int(torch.Tensor._version.__get__(self)))
    Tensor._inplace_version = lambda self: int(torch.Tensor._version.__get__(self))
    Tensor.strides = property(lambda self: list(torch.Tensor.stride(self)))
    Tensor.offset = property(lambda self: int(torch.Tensor.storage_offset(self)) * self.element_size())
    Tensor.is_selected_rows = lambda self: False
    Tensor.rows = lambda self: (_ for _ in ()).throw(RuntimeError("rows() is only defined for SelectedRows tensors"))
    if not hasattr(Tensor, "nnz"):
        Tensor.nnz = lambda self: int(_raw(self)._nnz()) if _raw(self).layout != torch.strided else int(torch.count_nonzero(_raw(self)))


patch()
