"""paddle.sparse.nn.functional. Sparse conv / pool run densely on the active sites' bounding volume and are re-sparsified
(correct semantics; the reference's gather-GEMM-scatter kernels are a performance follow-up)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from ...tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


def _w(t):
    if t.layout != torch.strided:
        return t
    return t.as_subclass(Tensor) if not isinstance(t, Tensor) else t


def _vals(fn, x):
    xr = _raw(x)
    if xr.layout == torch.sparse_coo:
        xr = xr.coalesce()
        return _w(torch.sparse_coo_tensor(xr.indices(), fn(xr.values()), xr.shape))
    return _w(torch.sparse_csr_tensor(xr.crow_indices(), xr.col_indices(), fn(xr.values()), xr.shape))


def relu(x, name=None):
    return _vals(torch.relu, x)


def relu6(x, name=None):
    return _vals(F.relu6, x)


def leaky_relu(x, negative_slope=0.01, name=None):
    return _vals(lambda v: F.leaky_relu(v, negative_slope), x)


def softmax(x, axis=-1, name=None):
    xr = _raw(x)
    if xr.layout == torch.sparse_csr:
        crow, col, v = xr.crow_indices(), xr.col_indices(), xr.values()
        rows = torch.repeat_interleave(torch.arange(xr.shape[-2], device=col.device), crow[1:] - crow[:-1])
        mx = torch.full((xr.shape[-2],), float("-inf"), dtype=v.dtype, device=v.device).scatter_reduce(0, rows, v, "amax")
        e = torch.exp(v - mx[rows])
        s = torch.zeros(xr.shape[-2], dtype=v.dtype, device=v.device).index_add(0, rows, e)
        return _w(torch.sparse_csr_tensor(crow, col, e / s[rows], xr.shape))
    return _w(torch.sparse.softmax(xr.coalesce(), dim=axis))


def _conv3d(x, weight, bias, stride, padding, dilation, groups, subm=False):
    xr = _raw(x).coalesce()
    dense = xr.to_dense().permute(0, 4, 1, 2, 3)                    # NDHWC -> NCDHW
    w = _raw(weight).permute(4, 3, 0, 1, 2)                          # DHWIO -> OIDHW
    if subm:
        k = w.shape[2:]
        padding = tuple((kk - 1) // 2 for kk in k)
        stride = 1
    out = F.conv3d(dense, w, None if bias is None else _raw(bias), stride, padding, dilation, groups).permute(0, 2, 3, 4, 1)
    if subm:
        idx = xr.indices()[:4]
        vals = out[idx[0], idx[1], idx[2], idx[3]]
        return _w(torch.sparse_coo_tensor(idx, vals, out.shape))
    active = F.conv3d((dense.abs().sum(1, keepdim=True) > 0).float(), torch.ones(1, 1, *w.shape[2:], device=w.device), None, stride, padding, dilation) > 0
    mask = active[:, 0]
    idx = mask.nonzero().t()
    return _w(torch.sparse_coo_tensor(idx, out[mask], out.shape))


def conv3d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NDHWC", name=None):
    return _conv3d(x, weight, bias, stride, padding, dilation, groups, False)


def subm_conv3d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NDHWC", key=None, name=None):
    return _conv3d(x, weight, bias, stride, padding, dilation, groups, True)


def max_pool3d(x, kernel_size, stride=None, padding=0, ceil_mode=False, data_format="NDHWC", name=None):
    xr = _raw(x).coalesce()
    dense = xr.to_dense().permute(0, 4, 1, 2, 3)
    neg = torch.where(dense == 0, torch.full_like(dense, float("-inf")), dense)
    out = F.max_pool3d(neg, kernel_size, stride, padding, 1, ceil_mode)
    out = torch.where(torch.isinf(out), torch.zeros_like(out), out).permute(0, 2, 3, 4, 1)
    mask = out.abs().sum(-1) > 0
    return _w(torch.sparse_coo_tensor(mask.nonzero().t(), out[mask], out.shape))


def attention(query, key, value, sparse_mask, key_padding_mask=None, attn_mask=None, name=None):
    """softmax(QK^T / sqrt(d)) restricted to sparse_mask's pattern, times V. Parity: sparse/nn/functional/transformer.py."""
    q, k, v = _raw(query), _raw(key), _raw(value)
    m = _raw(sparse_mask)
    dense_mask = m.to_dense() != 0 if m.layout != torch.strided else m != 0
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    s = s.masked_fill(~dense_mask.reshape(s.shape), float("-inf"))
    if key_padding_mask is not None:
        s = s + _raw(key_padding_mask)[:, None, None, :]
    if attn_mask is not None:
        s = s + _raw(attn_mask)
    return _w(torch.softmax(s, -1).nan_to_num(0.0) @ v)


def _conv2d(x, weight, bias, stride, padding, dilation, groups, subm=False):
    """Sparse NHWC conv: same dense-compute / sparse-pattern strategy as _conv3d (weight layout HWIO)."""
    xr = _raw(x).coalesce()
    dense = xr.to_dense().permute(0, 3, 1, 2)
    w = _raw(weight).permute(3, 2, 0, 1)
    if subm:
        padding = tuple((kk - 1) // 2 for kk in w.shape[2:])
        stride = 1
    out = F.conv2d(dense, w, None if bias is None else _raw(bias), stride, padding, dilation, groups).permute(0, 2, 3, 1)
    if subm:
        idx = xr.indices()[:3]
        return _w(torch.sparse_coo_tensor(idx, out[idx[0], idx[1], idx[2]], out.shape))
    active = F.conv2d((dense.abs().sum(1, keepdim=True) > 0).float(), torch.ones(1, 1, *w.shape[2:], device=w.device), None, stride, padding, dilation) > 0
    mask = active[:, 0]
    return _w(torch.sparse_coo_tensor(mask.nonzero().t(), out[mask], out.shape))


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NHWC", name=None):
    return _conv2d(x, weight, bias, stride, padding, dilation, groups, False)


def subm_conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NHWC", key=None, name=None):
    return _conv2d(x, weight, bias, stride, padding, dilation, groups, True)


def subm_conv2d_igemm(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NHWC", key=None, name=None):
    return _conv2d(x, weight, bias, stride, padding, dilation, groups, True)


def subm_conv3d_igemm(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NDHWC", key=None, name=None):
    return _conv3d(x, weight, bias, stride, padding, dilation, groups, True)
