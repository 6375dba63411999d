"""paddle.sparse.nn.functional. Sparse conv / pool use a gather-GEMM-scatter rulebook over the ACTIVE sites (never a dense volume):
memory and work scale with nnz * kernel volume. Parity: paddle/phi/kernels/sparse/gpu/conv_kernel.cu, pool_kernel.cu."""
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


def _tup(v, n):
    return tuple(int(i) for i in v) if isinstance(v, (list, tuple)) else (int(v),) * n


def _rulebook(idx, spatial, ksize, stride, padding, dilation, subm):
    """Gather-GEMM-scatter rulebook of an N-d sparse convolution (reference: paddle/phi/kernels/sparse/gpu/conv_kernel.cu - product
    rulebook + unique output coordinates).  idx: [1 + nd, nnz] coordinates (batch first) of the active input sites.
    Returns (out_idx [1 + nd, n_out], out_spatial, pairs) with pairs[k] = (in_rows, out_rows) for kernel offset k: the work is
    O(nnz * prod(ksize)), memory never depends on the volume."""
    nd = len(spatial)
    dev = idx.device
    out_spatial = tuple((spatial[d] + 2 * padding[d] - dilation[d] * (ksize[d] - 1) - 1) // stride[d] + 1 for d in range(nd)) if not subm else tuple(spatial)
    offs = torch.stack(torch.meshgrid(*[torch.arange(k, device=dev) for k in ksize], indexing="ij"), -1).reshape(-1, nd)     # [K, nd]
    b, pos = idx[0], idx[1:].t()                                                                                                # [nnz], [nnz, nd]
    pad_t, str_t, dil_t = (torch.tensor(v, device=dev) for v in (padding, stride, dilation))
    osz = torch.tensor(out_spatial, device=dev)
    # output position o satisfies o * stride - pad + k * dilation = i  ->  o = (i + pad - k * dilation) / stride (exact, in range)
    num = pos[None, :, :] + pad_t - offs[:, None, :] * dil_t                                                                   # [K, nnz, nd]
    o = torch.div(num, str_t, rounding_mode="floor")
    ok = ((num % str_t) == 0).all(-1) & (o >= 0).all(-1) & (o < osz).all(-1)                                                    # [K, nnz]
    mult = torch.ones(nd, dtype=torch.long, device=dev)
    for d in range(nd - 2, -1, -1):
        mult[d] = mult[d + 1] * out_spatial[d + 1]
    vol = int(mult[0]) * out_spatial[0]
    key = b[None, :] * vol + (o * mult).sum(-1)                                                                                 # [K, nnz] linear output site
    if subm:       # outputs live exactly on the input sites
        in_key = b * vol + (pos * mult).sum(-1)
        sorted_key, order = torch.sort(in_key)
        slot = torch.searchsorted(sorted_key, key.clamp(min=0))
        slot_c = slot.clamp(max=sorted_key.numel() - 1)
        hit = ok & (sorted_key[slot_c] == key)
        out_rows_all = order[slot_c]
        out_idx = idx
    else:
        uniq = torch.unique(key[ok])
        slot = torch.searchsorted(uniq, key.clamp(min=0)).clamp(max=max(uniq.numel() - 1, 0))
        hit = ok
        out_rows_all = slot
        ob = torch.div(uniq, vol, rounding_mode="floor")
        rem = uniq - ob * vol
        coords = []
        for d in range(nd):
            coords.append(torch.div(rem, mult[d], rounding_mode="floor"))
            rem = rem - coords[-1] * mult[d]
        out_idx = torch.stack([ob] + coords, 0)
    rows = torch.arange(pos.shape[0], device=dev)
    pairs = []
    for k in range(offs.shape[0]):
        m = hit[k]
        pairs.append((rows[m], out_rows_all[k][m]))
    return out_idx, out_spatial, pairs


def _sparse_conv(x, weight, bias, stride, padding, dilation, groups, subm, nd):
    """Sparse N-d convolution on COO input [N, *spatial, C] with a [*ksize, C_in / groups, C_out] weight: per kernel offset one gather,
    one GEMM on the gathered features (tcgen05 GEMM for bf16 / fp16 CUDA features), one scatter-add.  Autograd flows through the
    index_select / matmul / index_add chain."""
    xr = _raw(x).coalesce()
    w = _raw(weight)
    ksize = tuple(w.shape[:nd])
    stride, padding, dilation = _tup(stride, nd), _tup(padding, nd), _tup(dilation, nd)
    if subm:
        stride, padding = (1,) * nd, tuple(dilation[d] * (ksize[d] - 1) // 2 for d in range(nd))
    idx, feats = xr.indices(), xr.values()                                              # [1 + nd, nnz], [nnz, C]
    spatial = tuple(xr.shape[1:1 + nd])
    out_idx, out_spatial, pairs = _rulebook(idx, spatial, ksize, stride, padding, dilation, subm)
    cin, cout = w.shape[nd] * groups, w.shape[nd + 1]
    wk = w.reshape(-1, w.shape[nd], cout)                                               # [K, C_in / groups, C_out]
    n_out = out_idx.shape[1]
    out = feats.new_zeros((n_out, cout))
    cg_in, cg_out = cin // groups, cout // groups
    for k, (rin, rout) in enumerate(pairs):
        if rin.numel() == 0:
            continue
        g = feats.index_select(0, rin)
        if groups == 1:
            y = g @ wk[k]
        else:
            y = torch.cat([g[:, gi * cg_in:(gi + 1) * cg_in] @ wk[k][:, gi * cg_out:(gi + 1) * cg_out] for gi in range(groups)], 1)
        out = out.index_add(0, rout, y)
    if bias is not None:
        out = out + _raw(bias)
    shape = (xr.shape[0],) + tuple(out_spatial) + (cout,)
    return _w(torch.sparse_coo_tensor(out_idx, out, shape))


def _conv3d(x, weight, bias, stride, padding, dilation, groups, subm=False):
    return _sparse_conv(x, weight, bias, stride, padding, dilation, groups, subm, 3)


def conv3d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NDHWC", name=None):
    return _conv3d(x, weight, bias, stride, padding, dilation, groups, False)


def subm_conv3d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NDHWC", key=None, name=None):
    return _conv3d(x, weight, bias, stride, padding, dilation, groups, True)


def max_pool3d(x, kernel_size, stride=None, padding=0, ceil_mode=False, data_format="NDHWC", name=None):
    """Sparse max pooling through the same rulebook: every active output site takes the maximum over the active inputs it covers."""
    xr = _raw(x).coalesce()
    ksize = _tup(kernel_size, 3)
    stride = _tup(stride if stride is not None else kernel_size, 3)
    padding = _tup(padding, 3)
    idx, feats = xr.indices(), xr.values()
    spatial = tuple(xr.shape[1:4])
    if ceil_mode:
        out_sp = tuple(-(-(spatial[d] + 2 * padding[d] - ksize[d]) // stride[d]) + 1 for d in range(3))
    out_idx, out_spatial, pairs = _rulebook(idx, spatial, ksize, stride, padding, (1, 1, 1), False)
    n_out = out_idx.shape[1]
    out = feats.new_full((n_out, feats.shape[1]), float("-inf"))
    for rin, rout in pairs:
        if rin.numel():
            out = out.scatter_reduce(0, rout[:, None].expand(-1, feats.shape[1]), feats.index_select(0, rin), "amax", include_self=True)
    out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
    return _w(torch.sparse_coo_tensor(out_idx, out, (xr.shape[0],) + tuple(out_spatial) + (feats.shape[1],)))


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
    """Sparse NHWC conv (weight layout HWIO) through the rulebook path."""
    return _sparse_conv(x, weight, bias, stride, padding, dilation, groups, subm, 2)


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NHWC", name=None):
    return _conv2d(x, weight, bias, stride, padding, dilation, groups, False)


def subm_conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NHWC", key=None, name=None):
    return _conv2d(x, weight, bias, stride, padding, dilation, groups, True)


def subm_conv2d_igemm(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NHWC", key=None, name=None):
    return _conv2d(x, weight, bias, stride, padding, dilation, groups, True)


def subm_conv3d_igemm(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NDHWC", key=None, name=None):
    return _conv3d(x, weight, bias, stride, padding, dilation, groups, True)
