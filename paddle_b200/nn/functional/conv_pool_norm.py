"""conv / pooling / normalisation functionals.

Parity: python/paddle/nn/functional/conv.py, pooling.py, norm.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ...ops._helpers import T, raw, to_int, wrap


def _tup(v, n):
    if isinstance(v, (list, tuple)):
        v = [int(to_int(i)) for i in v]
        return tuple(v) if len(v) == n else tuple(v)
    return (int(v),) * n


def _cl(data_format):
    return data_format in ("NLC", "NHWC", "NDHWC")


def _to_cf(x, channel_last):
    if not channel_last:
        return x
    nd = x.dim()
    return x.permute(0, nd - 1, *range(1, nd - 1))


def _from_cf(x, channel_last):
    if not channel_last:
        return x
    nd = x.dim()
    return x.permute(0, *range(2, nd), 1)


def _conv_padding(x, padding, n, kernel, stride, dilation):
    """Returns (x_maybe_padded, torch_padding)."""
    if isinstance(padding, str):
        p = padding.lower()
        if p == "valid":
            return x, 0
        if p == "same":
            if all(s == 1 for s in stride):
                return x, "same"
            pads = []
            for i in range(n):
                size = x.size(2 + i)
                out = -(-size // stride[i])
                total = max(0, (out - 1) * stride[i] + (kernel[i] - 1) * dilation[i] + 1 - size)
                pads.append((total // 2, total - total // 2))
            flat = []
            for lo, hi in reversed(pads):
                flat += [lo, hi]
            return F.pad(x, flat), 0
        raise ValueError(padding)
    if isinstance(padding, (list, tuple)):
        padding = [to_int(p) if not isinstance(p, (list, tuple)) else p for p in padding]
        if len(padding) == n and all(isinstance(p, int) for p in padding):
            return x, tuple(padding)
        if len(padding) == 2 * n and all(isinstance(p, int) for p in padding):
            pairs = [(padding[2 * i], padding[2 * i + 1]) for i in range(n)]
            if all(lo == hi for lo, hi in pairs):
                return x, tuple(lo for lo, _ in pairs)
            flat = []
            for lo, hi in reversed(pairs):
                flat += [lo, hi]
            return F.pad(x, flat), 0
        if len(padding) == n + 2:  # per-dim pairs incl. batch/channel
            pairs = [tuple(p) for p in padding if isinstance(p, (list, tuple))]
            pairs = pairs[-n:]
            flat = []
            for lo, hi in reversed(pairs):
                flat += [lo, hi]
            return F.pad(x, flat), 0
    return x, _tup(padding, n)


def _conv(n, x, weight, bias, stride, padding, dilation, groups, data_format):
    x, w = T(x), T(weight)
    cl = _cl(data_format)
    x = _to_cf(x, cl)
    stride, dilation = _tup(stride, n), _tup(dilation, n)
    x, pad = _conv_padding(x, padding, n, w.shape[2:], stride, dilation)
    fn = (F.conv1d, F.conv2d, F.conv3d)[n - 1]
    from ...amp.auto_cast import fp32_guard

    ctx, (x, w, bias) = fp32_guard(f"conv{n}d", x, w, None if bias is None else T(bias))
    with ctx:
        out = fn(x, w, bias, stride, pad, dilation, groups)
    return _from_cf(out, cl)


def conv1d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NCL", name=None):
    return _conv(1, x, weight, bias, stride, padding, dilation, groups, data_format)


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NCHW", name=None):
    return _conv(2, x, weight, bias, stride, padding, dilation, groups, data_format)


def conv3d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, data_format="NCDHW", name=None):
    return _conv(3, x, weight, bias, stride, padding, dilation, groups, data_format)


def _conv_t(n, x, weight, bias, stride, padding, output_padding, dilation, groups, output_size, data_format):
    x, w = T(x), T(weight)
    cl = _cl(data_format)
    x = _to_cf(x, cl)
    stride, dilation = _tup(stride, n), _tup(dilation, n)
    if isinstance(padding, str):
        padding = 0 if padding.lower() == "valid" else tuple(((w.shape[2 + i] - 1) * dilation[i]) // 2 for i in range(n))
    pad = _tup(padding, n) if not (isinstance(padding, (list, tuple)) and len(padding) == 2 * n) else tuple(int(padding[2 * i]) for i in range(n))
    opad = _tup(output_padding, n)
    if output_size is not None:
        osz = [int(to_int(s)) for s in (output_size if isinstance(output_size, (list, tuple)) else [output_size] * n)]
        opad = tuple(osz[i] - ((x.size(2 + i) - 1) * stride[i] - 2 * pad[i] + dilation[i] * (w.shape[2 + i] - 1) + 1) for i in range(n))
    fn = (F.conv_transpose1d, F.conv_transpose2d, F.conv_transpose3d)[n - 1]
    out = fn(x, w, None if bias is None else T(bias), stride, pad, opad, groups, dilation)
    return _from_cf(out, cl)


def conv1d_transpose(x, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1, output_size=None, data_format="NCL", name=None):
    return _conv_t(1, x, weight, bias, stride, padding, output_padding, dilation, groups, output_size, data_format)


def conv2d_transpose(x, weight, bias=None, stride=1, padding=0, output_padding=0, dilation=1, groups=1, output_size=None, data_format="NCHW", name=None):
    return _conv_t(2, x, weight, bias, stride, padding, output_padding, dilation, groups, output_size, data_format)


def conv3d_transpose(x, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1, output_size=None, data_format="NCDHW", name=None):
    return _conv_t(3, x, weight, bias, stride, padding, output_padding, dilation, groups, output_size, data_format)


# ----------------------------------------------------------------------------- pooling
def _pool_pad(padding, n):
    if isinstance(padding, str):
        return 0 if padding.lower() == "valid" else "same"
    if isinstance(padding, (list, tuple)) and len(padding) == 2 * n:
        return tuple(int(padding[2 * i]) for i in range(n))
    return _tup(padding, n)


def _same_pad(x, k, s, n):
    flat = []
    for i in reversed(range(n)):
        size = x.size(2 + i)
        out = -(-size // s[i])
        total = max(0, (out - 1) * s[i] + k[i] - size)
        flat += [total // 2, total - total // 2]
    return flat


def _avg_pool(n, x, kernel_size, stride, padding, exclusive, ceil_mode, divisor_override, data_format):
    x = T(x)
    cl = _cl(data_format)
    x = _to_cf(x, cl)
    k = _tup(kernel_size, n)
    s = k if stride is None else _tup(stride, n)
    p = _pool_pad(padding, n)
    if p == "same":
        x, p = F.pad(x, _same_pad(x, k, s, n)), 0
    fn = (F.avg_pool1d, F.avg_pool2d, F.avg_pool3d)[n - 1]
    if n == 1:
        out = fn(x, k, s, p, ceil_mode, not exclusive)
    else:
        out = fn(x, k, s, p, ceil_mode, not exclusive, divisor_override)
    return _from_cf(out, cl)


def avg_pool1d(x, kernel_size, stride=None, padding=0, exclusive=True, ceil_mode=False, name=None):
    return _avg_pool(1, x, kernel_size, stride, padding, exclusive, ceil_mode, None, "NCL")


def avg_pool2d(x, kernel_size, stride=None, padding=0, ceil_mode=False, exclusive=True, divisor_override=None, data_format="NCHW", name=None):
    return _avg_pool(2, x, kernel_size, stride, padding, exclusive, ceil_mode, divisor_override, data_format)


def avg_pool3d(x, kernel_size, stride=None, padding=0, ceil_mode=False, exclusive=True, divisor_override=None, data_format="NCDHW", name=None):
    return _avg_pool(3, x, kernel_size, stride, padding, exclusive, ceil_mode, divisor_override, data_format)


def _max_pool(n, x, kernel_size, stride, padding, return_mask, ceil_mode, data_format):
    x = T(x)
    cl = _cl(data_format)
    x = _to_cf(x, cl)
    k = _tup(kernel_size, n)
    s = k if stride is None else _tup(stride, n)
    p = _pool_pad(padding, n)
    if p == "same":
        x, p = F.pad(x, _same_pad(x, k, s, n), value=float("-inf")), 0
    fn = (F.max_pool1d, F.max_pool2d, F.max_pool3d)[n - 1]
    if return_mask:
        out, idx = fn(x, k, s, p, 1, ceil_mode, True)
        return _from_cf(out, cl), _from_cf(idx, cl)
    return _from_cf(fn(x, k, s, p, 1, ceil_mode, False), cl)


def max_pool1d(x, kernel_size, stride=None, padding=0, return_mask=False, ceil_mode=False, name=None):
    return _max_pool(1, x, kernel_size, stride, padding, return_mask, ceil_mode, "NCL")


def max_pool2d(x, kernel_size, stride=None, padding=0, return_mask=False, ceil_mode=False, data_format="NCHW", name=None):
    return _max_pool(2, x, kernel_size, stride, padding, return_mask, ceil_mode, data_format)


def max_pool3d(x, kernel_size, stride=None, padding=0, return_mask=False, ceil_mode=False, data_format="NCDHW", name=None):
    return _max_pool(3, x, kernel_size, stride, padding, return_mask, ceil_mode, data_format)


def _osz(v, n):
    if isinstance(v, (list, tuple)):
        return tuple(None if i is None else int(to_int(i)) for i in v)
    return (int(v),) * n


def adaptive_avg_pool1d(x, output_size, name=None):
    return F.adaptive_avg_pool1d(T(x), int(output_size))


def adaptive_avg_pool2d(x, output_size, data_format="NCHW", name=None):
    cl = _cl(data_format)
    return _from_cf(F.adaptive_avg_pool2d(_to_cf(T(x), cl), _osz(output_size, 2)), cl)


def adaptive_avg_pool3d(x, output_size, data_format="NCDHW", name=None):
    cl = _cl(data_format)
    return _from_cf(F.adaptive_avg_pool3d(_to_cf(T(x), cl), _osz(output_size, 3)), cl)


def adaptive_max_pool1d(x, output_size, return_mask=False, name=None):
    r = F.adaptive_max_pool1d(T(x), int(output_size), return_mask)
    return r


def adaptive_max_pool2d(x, output_size, return_mask=False, name=None):
    return F.adaptive_max_pool2d(T(x), _osz(output_size, 2), return_mask)


def adaptive_max_pool3d(x, output_size, return_mask=False, name=None):
    return F.adaptive_max_pool3d(T(x), _osz(output_size, 3), return_mask)


def lp_pool1d(x, norm_type, kernel_size, stride=None, padding=0, ceil_mode=False, data_format="NCL", name=None):
    cl = _cl(data_format)
    return _from_cf(F.lp_pool1d(_to_cf(T(x), cl), float(norm_type), int(kernel_size) if not isinstance(kernel_size, (list, tuple)) else kernel_size[0], stride if stride is None or isinstance(stride, int) else stride[0], ceil_mode), cl)


def lp_pool2d(x, norm_type, kernel_size, stride=None, padding=0, ceil_mode=False, data_format="NCHW", name=None):
    cl = _cl(data_format)
    return _from_cf(F.lp_pool2d(_to_cf(T(x), cl), float(norm_type), _tup(kernel_size, 2), None if stride is None else _tup(stride, 2), ceil_mode), cl)


def max_unpool1d(x, indices, kernel_size, stride=None, padding=0, data_format="NCL", output_size=None, name=None):
    return F.max_unpool1d(T(x), T(indices).long(), kernel_size, stride, padding, output_size)


def max_unpool2d(x, indices, kernel_size, stride=None, padding=0, data_format="NCHW", output_size=None, name=None):
    return F.max_unpool2d(T(x), T(indices).long(), kernel_size, stride, padding, output_size)


def max_unpool3d(x, indices, kernel_size, stride=None, padding=0, data_format="NCDHW", output_size=None, name=None):
    return F.max_unpool3d(T(x), T(indices).long(), kernel_size, stride, padding, output_size)


def fractional_max_pool2d(x, output_size, kernel_size=None, random_u=None, return_mask=False, name=None):
    x = T(x)
    ks = _tup(kernel_size, 2) if kernel_size is not None else (2, 2)
    samples = None if random_u is None else torch.full((x.size(0), x.size(1), 2), float(random_u), dtype=x.dtype, device=x.device)
    return F.fractional_max_pool2d(x, ks, output_size=_osz(output_size, 2), return_indices=return_mask, _random_samples=samples)


def fractional_max_pool3d(x, output_size, kernel_size=None, random_u=None, return_mask=False, name=None):
    x = T(x)
    ks = _tup(kernel_size, 3) if kernel_size is not None else (2, 2, 2)
    samples = None if random_u is None else torch.full((x.size(0), x.size(1), 3), float(random_u), dtype=x.dtype, device=x.device)
    return F.fractional_max_pool3d(x, ks, output_size=_osz(output_size, 3), return_indices=return_mask, _random_samples=samples)


# ----------------------------------------------------------------------------- normalisation
def batch_norm(x, running_mean, running_var, weight=None, bias=None, training=False, momentum=0.9, epsilon=1e-05,
               data_format="NCHW", use_global_stats=None, name=None):
    x = T(x)
    cl = _cl(data_format) and x.dim() > 2
    xx = _to_cf(x, cl)
    use_batch = training and not use_global_stats
    out = F.batch_norm(xx, running_mean, running_var, weight, bias, use_batch, 1.0 - momentum, epsilon)
    return _from_cf(out, cl)


def layer_norm(x, normalized_shape, weight=None, bias=None, epsilon=1e-05, name=None):
    from ...kernels import norm as K

    if isinstance(normalized_shape, int):
        normalized_shape = [normalized_shape]
    return K.layer_norm(T(x), list(normalized_shape), weight, bias, epsilon)


def rms_norm(x, weight=None, epsilon=1e-6, bias=None, name=None):
    from ...kernels import norm as K

    return K.rms_norm(T(x), weight, epsilon, bias)


def instance_norm(x, running_mean=None, running_var=None, weight=None, bias=None, use_input_stats=True,
                  momentum=0.9, eps=1e-05, data_format="NCHW", name=None):
    x = T(x)
    cl = _cl(data_format)
    out = F.instance_norm(_to_cf(x, cl), running_mean, running_var, weight, bias, use_input_stats, 1.0 - momentum, eps)
    return _from_cf(out, cl)


def group_norm(x, num_groups, epsilon=1e-05, weight=None, bias=None, data_format="NCHW", name=None):
    x = T(x)
    cl = _cl(data_format)
    return _from_cf(F.group_norm(_to_cf(x, cl), num_groups, weight, bias, epsilon), cl)


def local_response_norm(x, size, alpha=1e-4, beta=0.75, k=1.0, data_format="NCHW", name=None):
    x = T(x)
    cl = _cl(data_format)
    return _from_cf(F.local_response_norm(_to_cf(x, cl), size, alpha * size if False else alpha, beta, k), cl)


__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "F", "T", "raw", "to_int", "wrap", "annotations")]
