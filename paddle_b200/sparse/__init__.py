"""paddle.sparse. Parity: python/paddle/sparse/{creation,unary,binary,multiary}.py (COO / CSR tensors)."""
from __future__ import annotations

import numpy as np
import torch

from ..framework.dtype import convert_dtype
from ..tensor import Tensor
from . import nn  # noqa: F401


def _raw(t):
    if isinstance(t, torch.Tensor):
        return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t
    return torch.as_tensor(np.asarray(t))


def _w(t):
    # sparse layouts cannot be re-typed with as_subclass (no storage): sparse tensors stay torch.Tensor with the paddle
    # sparse methods patched on (crows/cols/nnz/to_dense -> paddle Tensor)
    if isinstance(t, torch.Tensor) and t.layout != torch.strided:
        return t
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def sparse_coo_tensor(indices, values, shape=None, dtype=None, place=None, stop_gradient=True):
    i, v = _raw(indices).long(), _raw(values)
    if dtype is not None:
        v = v.to(convert_dtype(dtype))
    if shape is None:
        shape = (i.max(1).values + 1).tolist() + list(v.shape[1:])
    t = torch.sparse_coo_tensor(i, v, tuple(shape)).coalesce()
    t.requires_grad_(not stop_gradient and v.is_floating_point())
    return _w(t)


def sparse_csr_tensor(crows, cols, values, shape, dtype=None, place=None, stop_gradient=True):
    v = _raw(values)
    if dtype is not None:
        v = v.to(convert_dtype(dtype))
    t = torch.sparse_csr_tensor(_raw(crows).long(), _raw(cols).long(), v, tuple(shape))
    return _w(t)


def _unary(fn):
    def op(x, name=None):
        x = _raw(x)
        if x.layout == torch.sparse_coo:
            x = x.coalesce()
            return _w(torch.sparse_coo_tensor(x.indices(), fn(x.values()), x.shape))
        if x.layout == torch.sparse_csr:
            return _w(torch.sparse_csr_tensor(x.crow_indices(), x.col_indices(), fn(x.values()), x.shape))
        return _w(fn(x))

    return op


sin, tan, asin, atan, sinh, tanh, asinh, atanh = map(_unary, (torch.sin, torch.tan, torch.asin, torch.atan, torch.sinh, torch.tanh, torch.asinh, torch.atanh))
sqrt, square, log1p, abs, neg, expm1, rad2deg, deg2rad = map(_unary, (torch.sqrt, torch.square, torch.log1p, torch.abs, torch.neg, torch.expm1, torch.rad2deg, torch.deg2rad))
isnan = _unary(torch.isnan)


def pow(x, factor, name=None):  # noqa: A001
    return _unary(lambda v: torch.pow(v, factor))(x)


def cast(x, index_dtype=None, value_dtype=None, name=None):
    x = _raw(x)
    if x.layout == torch.sparse_coo:
        x = x.coalesce()
        i = x.indices().to(convert_dtype(index_dtype)) if index_dtype else x.indices()
        v = x.values().to(convert_dtype(value_dtype)) if value_dtype else x.values()
        return _w(torch.sparse_coo_tensor(i.long(), v, x.shape))
    v = x.values().to(convert_dtype(value_dtype)) if value_dtype else x.values()
    return _w(torch.sparse_csr_tensor(x.crow_indices(), x.col_indices(), v, x.shape))


def coalesce(x, name=None):
    return _w(_raw(x).coalesce())


def _dense(x):
    x = _raw(x)
    return x.to_dense() if x.layout != torch.strided else x


def _same_layout(dense, like):
    like = _raw(like)
    if like.layout == torch.sparse_coo:
        return _w(dense.to_sparse_coo())
    if like.layout == torch.sparse_csr:
        return _w(dense.to_sparse_csr())
    return _w(dense)


def add(x, y, name=None):
    xr, yr = _raw(x), _raw(y)
    if xr.layout == torch.sparse_coo and yr.layout == torch.sparse_coo:
        return _w((xr + yr).coalesce())
    return _same_layout(_dense(x) + _dense(y), x)


def subtract(x, y, name=None):
    xr, yr = _raw(x), _raw(y)
    if xr.layout == torch.sparse_coo and yr.layout == torch.sparse_coo:
        return _w((xr - yr).coalesce())
    return _same_layout(_dense(x) - _dense(y), x)


def multiply(x, y, name=None):
    xr = _raw(x)
    if isinstance(y, (int, float)):
        return _unary(lambda v: v * y)(x)
    yr = _raw(y)
    if xr.layout == torch.sparse_coo and yr.layout == torch.sparse_coo:
        return _w((xr * yr).coalesce())
    return _same_layout(_dense(x) * _dense(y), x)


def divide(x, y, name=None):
    if isinstance(y, (int, float)):
        return _unary(lambda v: v / y)(x)
    d = _dense(x) / _dense(y)
    return _same_layout(torch.nan_to_num(d, nan=0.0, posinf=0.0, neginf=0.0) if False else d, x)


def matmul(x, y, name=None):
    xr, yr = _raw(x), _raw(y)
    if xr.layout == torch.sparse_coo and yr.layout == torch.strided:
        return _w(torch.sparse.mm(xr, yr) if xr.dim() == 2 else torch.bmm(xr, yr))
    if xr.layout == torch.sparse_csr and yr.layout == torch.strided:
        return _w(xr @ yr)
    return _w(_dense(x) @ _dense(y))


def masked_matmul(x, y, mask, name=None):
    m = _raw(mask)
    out = _dense(x) @ _dense(y)
    if m.layout == torch.sparse_coo:
        m = m.coalesce()
        idx = m.indices()
        return _w(torch.sparse_coo_tensor(idx, out[tuple(idx)], m.shape))
    crow, col = m.crow_indices(), m.col_indices()
    rows = torch.repeat_interleave(torch.arange(m.shape[-2], device=col.device), crow[1:] - crow[:-1])
    return _w(torch.sparse_csr_tensor(crow, col, out[rows, col], m.shape))


def mv(x, vec, name=None):
    return _w(_raw(x) @ _raw(vec))


def addmm(input, x, y, beta=1.0, alpha=1.0, name=None):
    return _w(beta * _dense(input) + alpha * (_dense(x) @ _dense(y)))


def transpose(x, perm, name=None):
    xr = _raw(x)
    if xr.layout == torch.sparse_coo:
        xr = xr.coalesce()
        return _w(torch.sparse_coo_tensor(xr.indices()[list(perm)], xr.values(), [xr.shape[p] for p in perm]).coalesce())
    return _same_layout(_dense(x).permute(*perm), x)


def reshape(x, shape, name=None):
    return _same_layout(_dense(x).reshape(shape), x)


def sum(x, axis=None, dtype=None, keepdim=False, name=None):  # noqa: A001
    d = _dense(x)
    out = d.sum() if axis is None else d.sum(axis, keepdim=keepdim)
    if dtype is not None:
        out = out.to(convert_dtype(dtype))
    return _same_layout(out, x) if out.dim() > 0 else _w(out)


def slice(x, axes, starts, ends, name=None):  # noqa: A001
    d = _dense(x)
    import builtins

    idx = [builtins.slice(None)] * d.dim()
    for a, s, e in zip(axes, starts, ends):
        idx[a] = builtins.slice(s, e)
    return _same_layout(d[tuple(idx)], x)


def is_same_shape(x, y):
    return list(_raw(x).shape) == list(_raw(y).shape)


def mask_as(x, mask, name=None):
    m = _raw(mask)
    d = _dense(x)
    if m.layout == torch.sparse_coo:
        m = m.coalesce()
        return _w(torch.sparse_coo_tensor(m.indices(), d[tuple(m.indices())], m.shape))
    crow, col = m.crow_indices(), m.col_indices()
    rows = torch.repeat_interleave(torch.arange(m.shape[-2], device=col.device), crow[1:] - crow[:-1])
    return _w(torch.sparse_csr_tensor(crow, col, d[rows, col], m.shape))


def pca_lowrank(x, q=None, center=True, niter=2, name=None):
    u, s, v = torch.pca_lowrank(_dense(x).as_subclass(torch.Tensor), q=q, center=center, niter=niter)
    return _w(u), _w(s), _w(v)


# Tensor conversions (Tensor.to_sparse_coo / to_sparse_csr / to_dense / indices / values / crows / cols)
def _patch_tensor():
    Tensor.to_sparse_coo = lambda self, sparse_dim=None: _raw(self).to_sparse_coo() if sparse_dim is None else _raw(self).to_sparse(sparse_dim)
    Tensor.to_sparse_csr = lambda self: _raw(self).to_sparse_csr()
    _orig_to_dense = torch.Tensor.to_dense

    def to_dense(self, *a, **k):
        out = _orig_to_dense(self, *a, **k)
        return out.as_subclass(Tensor) if type(out) is torch.Tensor else out

    torch.Tensor.to_dense = to_dense
    torch.Tensor.crows = lambda self: self.crow_indices().as_subclass(Tensor)
    torch.Tensor.cols = lambda self: self.col_indices().as_subclass(Tensor)
    torch.Tensor.nnz = lambda self: int(self._nnz())
    torch.Tensor.is_sparse_coo = lambda self: self.layout == torch.sparse_coo
    torch.Tensor.is_sparse_csr = lambda self: self.layout == torch.sparse_csr


_patch_tensor()
