"""Math ops. Parity: python/paddle/tensor/math.py, ops.py."""
from __future__ import annotations

import math as _math

import numpy as np
import torch

from ..framework import dtype as _dt
from ._helpers import T, ax, binary_args, dt, raw, scalar_or_tensor, shp, to_int, to_tensor, wrap

# --------------------------------------------------------------------------- unary table
_UNARY = {
    "abs": torch.abs, "acos": torch.acos, "acosh": torch.acosh, "asin": torch.asin, "asinh": torch.asinh,
    "atan": torch.atan, "atanh": torch.atanh, "ceil": torch.ceil, "cos": torch.cos, "cosh": torch.cosh,
    "exp": torch.exp, "expm1": torch.expm1, "floor": torch.floor, "log": torch.log, "log2": torch.log2,
    "log10": torch.log10, "log1p": torch.log1p, "neg": torch.neg, "reciprocal": torch.reciprocal,
    "round": torch.round, "rsqrt": torch.rsqrt, "sign": torch.sign, "sin": torch.sin, "sinh": torch.sinh,
    "sqrt": torch.sqrt, "square": torch.square, "tan": torch.tan, "tanh": torch.tanh, "trunc": torch.trunc,
    "erf": torch.erf, "erfinv": torch.erfinv, "lgamma": torch.lgamma, "digamma": torch.digamma,
    "sigmoid": torch.sigmoid, "frac": torch.frac, "angle": torch.angle, "conj": torch.conj,
    "real": torch.real, "imag": torch.imag, "i0": torch.i0, "i0e": torch.special.i0e, "i1": torch.special.i1,
    "i1e": torch.special.i1e, "isfinite": torch.isfinite, "isinf": torch.isinf, "isnan": torch.isnan,
    "exp2": torch.exp2, "sinc": torch.sinc, "signbit": torch.signbit, "deg2rad": torch.deg2rad,
    "rad2deg": torch.rad2deg, "positive": torch.positive, "negative": torch.neg, "gammaln": torch.lgamma,
    "isneginf": torch.isneginf, "isposinf": torch.isposinf, "isreal": torch.isreal, "logit_": None,
}
_UNARY.pop("logit_")


from ..framework.recording import recordable as _recordable  # noqa: E402

def _mk_unary(name, fn):
    def op(x, name=None):
        return fn(T(x))

    op.__name__ = name
    op.__doc__ = f"paddle.{name}(x). Parity: python/paddle/tensor/ops.py / math.py."
    return op


for _n, _f in _UNARY.items():
    globals()[_n] = _mk_unary(_n, _f)


def _mk_inplace(name):
    def op(x, *a, **k):
        out = globals()[name](x, *a, **k)
        with torch.no_grad():
            torch.Tensor.copy_(x, out)
        return x

    op.__name__ = name + "_"
    return op


# --------------------------------------------------------------------------- binary
def add(x, y, name=None):
    x, y = binary_args(x, y)
    return torch.add(x, y)


def subtract(x, y, name=None):
    x, y = binary_args(x, y)
    return torch.sub(x, y)


def multiply(x, y, name=None):
    x, y = binary_args(x, y)
    return torch.mul(x, y)


def divide(x, y, name=None):
    x, y = binary_args(x, y)
    return torch.true_divide(x, y)


def floor_divide(x, y, name=None):
    x, y = binary_args(x, y)
    return torch.div(x, y, rounding_mode="floor")


def remainder(x, y, name=None):
    x, y = binary_args(x, y)
    return torch.remainder(x, y)


mod = remainder
floor_mod = remainder


def pow(x, y, name=None):  # noqa: A001
    x, y = binary_args(x, y)
    return torch.pow(x, y)


def maximum(x, y, name=None):
    return torch.maximum(T(x), T(y))


def minimum(x, y, name=None):
    return torch.minimum(T(x), T(y))


def fmax(x, y, name=None):
    return torch.fmax(T(x), T(y))


def fmin(x, y, name=None):
    return torch.fmin(T(x), T(y))


def atan2(x, y, name=None):
    return torch.atan2(T(x), T(y))


def hypot(x, y, name=None):
    return torch.hypot(T(x), T(y))


def copysign(x, y, name=None):
    return torch.copysign(T(x), scalar_or_tensor(y))


def nextafter(x, y, name=None):
    return torch.nextafter(T(x), T(y))


def ldexp(x, y, name=None):
    return torch.ldexp(T(x), T(y))


def heaviside(x, y, name=None):
    return torch.heaviside(T(x), T(y))


def gcd(x, y, name=None):
    return torch.gcd(T(x), T(y))


def lcm(x, y, name=None):
    return torch.lcm(T(x), T(y))


def logaddexp(x, y, name=None):
    return torch.logaddexp(T(x), T(y))


def inner(x, y, name=None):
    return torch.inner(T(x), T(y))


def outer(x, y, name=None):
    return torch.outer(T(x).reshape(-1), T(y).reshape(-1))


def kron(x, y, name=None):
    return torch.kron(T(x), T(y))


@_recordable
def scale(x, scale=1.0, bias=0.0, bias_after_scale=True, act=None, name=None):
    x = T(x)
    s = scale if not isinstance(scale, torch.Tensor) else scale
    out = x * s + bias if bias_after_scale else (x + bias) * s
    if act is not None:
        from ..nn import functional as F

        out = getattr(F, act)(out)
    return out


def stanh(x, scale_a=0.67, scale_b=1.7159, name=None):
    return scale_b * torch.tanh(scale_a * T(x))


def multiplex(inputs, index, name=None):
    stacked = torch.stack([T(i) for i in inputs], 0)
    idx = T(index).reshape(-1).long()
    return stacked[idx, torch.arange(stacked.size(1), device=stacked.device)]


@_recordable
def clip(x, min=None, max=None, name=None):  # noqa: A002
    x = T(x)
    mn = min.item() if isinstance(min, torch.Tensor) and min.numel() == 1 else min
    mx = max.item() if isinstance(max, torch.Tensor) and max.numel() == 1 else max
    if mn is None and mx is None:
        return x.clone()
    return torch.clamp(x, mn, mx)


def lerp(x, y, weight, name=None):
    x = T(x)
    w = weight if isinstance(weight, torch.Tensor) else float(weight)
    return torch.lerp(x, T(y), w)


def logit(x, eps=None, name=None):
    return torch.logit(T(x), eps)


def nan_to_num(x, nan=0.0, posinf=None, neginf=None, name=None):
    return torch.nan_to_num(T(x), nan, posinf, neginf)


def addmm(input, x, y, beta=1.0, alpha=1.0, name=None):
    return torch.addmm(T(input), T(x), T(y), beta=beta, alpha=alpha)


def baddbmm(input, x, y, beta=1.0, alpha=1.0, name=None):
    return torch.baddbmm(T(input), T(x), T(y), beta=beta, alpha=alpha)


def trace(x, offset=0, axis1=0, axis2=1, name=None):
    return torch.diagonal(T(x), offset, axis1, axis2).sum(-1)


def diagonal(x, offset=0, axis1=0, axis2=1, name=None):
    return torch.diagonal(T(x), offset, axis1, axis2)


def polygamma(x, n, name=None):
    return torch.polygamma(n, T(x))


def multigammaln(x, p, name=None):
    return torch.special.multigammaln(T(x), p)


def gammainc(x, y, name=None):
    return torch.special.gammainc(T(x), T(y))


def gammaincc(x, y, name=None):
    return torch.special.gammaincc(T(x), T(y))


def increment(x, value=1.0, name=None):
    with torch.no_grad():
        x.add_(value)
    return x


def add_n(inputs, name=None):
    if isinstance(inputs, torch.Tensor):
        return inputs.clone()
    out = T(inputs[0])
    for t in inputs[1:]:
        out = out + T(t)
    return out


def frexp(x, name=None):
    m, e = torch.frexp(T(x))
    return m, e.to(x.dtype)


def cummax(x, axis=None, dtype="int64", name=None):
    x = T(x)
    if axis is None:
        x, axis = x.reshape(-1), 0
    v, i = torch.cummax(x, axis)
    return v, i.to(dt(dtype))


def cummin(x, axis=None, dtype="int64", name=None):
    x = T(x)
    if axis is None:
        x, axis = x.reshape(-1), 0
    v, i = torch.cummin(x, axis)
    return v, i.to(dt(dtype))


def cumsum(x, axis=None, dtype=None, name=None):
    x = T(x)
    if axis is None:
        x, axis = x.reshape(-1), 0
    return torch.cumsum(x, int(to_int(axis)), dtype=dt(dtype))


def cumprod(x, dim=None, dtype=None, name=None):
    x = T(x)
    if dim is None:
        x, dim = x.reshape(-1), 0
    return torch.cumprod(x, int(to_int(dim)), dtype=dt(dtype))


def logcumsumexp(x, axis=None, dtype=None, name=None):
    x = T(x)
    if dtype is not None:
        x = x.to(dt(dtype))
    if axis is None:
        x, axis = x.reshape(-1), 0
    return torch.logcumsumexp(x, axis)


def diff(x, n=1, axis=-1, prepend=None, append=None, name=None):
    return torch.diff(T(x), n, axis, prepend=None if prepend is None else T(prepend), append=None if append is None else T(append))


def trapezoid(y, x=None, dx=None, axis=-1, name=None):
    if x is not None:
        return torch.trapezoid(T(y), x=T(x), dim=axis)
    return torch.trapezoid(T(y), dx=1.0 if dx is None else dx, dim=axis)


def cumulative_trapezoid(y, x=None, dx=None, axis=-1, name=None):
    if x is not None:
        return torch.cumulative_trapezoid(T(y), x=T(x), dim=axis)
    return torch.cumulative_trapezoid(T(y), dx=1.0 if dx is None else dx, dim=axis)


# --------------------------------------------------------------------------- reductions
def _reduce(fn, x, axis, keepdim, dtype=None):
    x = T(x)
    a = ax(axis)
    kw = {}
    if dtype is not None:
        kw["dtype"] = dt(dtype)
    if a is None or (isinstance(a, tuple) and len(a) == 0):
        out = fn(x, **kw)
        if keepdim:
            out = out.reshape([1] * x.dim())
        return out
    return fn(x, a, keepdim=keepdim, **kw)


def sum(x, axis=None, dtype=None, keepdim=False, name=None):  # noqa: A001
    x = T(x)
    if dtype is None and x.dtype == torch.bool:
        dtype = torch.int64
    return _reduce(torch.sum, x, axis, keepdim, dtype)


def nansum(x, axis=None, dtype=None, keepdim=False, name=None):
    return _reduce(torch.nansum, x, axis, keepdim, dtype)


def mean(x, axis=None, keepdim=False, name=None):
    return _reduce(torch.mean, x, axis, keepdim)


def nanmean(x, axis=None, keepdim=False, name=None):
    return _reduce(torch.nanmean, x, axis, keepdim)


def prod(x, axis=None, keepdim=False, dtype=None, name=None):
    x = T(x)
    a = ax(axis)
    if dtype is not None:
        x = x.to(dt(dtype))
    if a is None:
        out = torch.prod(x)
        return out.reshape([1] * x.dim()) if keepdim else out
    if isinstance(a, tuple):
        for d in sorted([d % x.dim() for d in a], reverse=True):
            x = torch.prod(x, d, keepdim=keepdim)
        return x
    return torch.prod(x, a, keepdim=keepdim)


def max(x, axis=None, keepdim=False, name=None):  # noqa: A001
    return _reduce(torch.amax, x, axis, keepdim)


def min(x, axis=None, keepdim=False, name=None):  # noqa: A001
    return _reduce(torch.amin, x, axis, keepdim)


amax = max
amin = min


def logsumexp(x, axis=None, keepdim=False, name=None):
    x = T(x)
    a = ax(axis)
    if a is None:
        a = tuple(range(x.dim()))
    return torch.logsumexp(x, a, keepdim=keepdim)


def all(x, axis=None, keepdim=False, name=None):  # noqa: A001
    return _reduce(torch.all, T(x).bool(), axis, keepdim)


def any(x, axis=None, keepdim=False, name=None):  # noqa: A001
    return _reduce(torch.any, T(x).bool(), axis, keepdim)


def count_nonzero(x, axis=None, keepdim=False, name=None):
    x = T(x)
    a = ax(axis)
    out = torch.count_nonzero(x, a)
    if keepdim:
        out = out.reshape([1] * x.dim()) if a is None else torch.sum((x != 0), a, keepdim=True)
    return out


def broadcast_shape(x_shape, y_shape):
    return list(torch.broadcast_shapes(tuple(x_shape), tuple(y_shape)))


def isclose(x, y, rtol=1e-05, atol=1e-08, equal_nan=False, name=None):
    return torch.isclose(T(x), T(y), rtol=rtol, atol=atol, equal_nan=equal_nan)


def allclose(x, y, rtol=1e-05, atol=1e-08, equal_nan=False, name=None):
    return wrap(torch.tensor(torch.allclose(raw(x), raw(y), rtol=rtol, atol=atol, equal_nan=equal_nan)))


def renorm(x, p, axis, max_norm):
    return torch.renorm(T(x), p, axis, max_norm)


def take(x, index, mode="raise", name=None):
    x, index = T(x), T(index)
    flat = x.reshape(-1)
    n = flat.numel()
    if mode == "wrap":
        index = index % n
    elif mode == "clip":
        index = index.clamp(0, n - 1)
    else:
        index = torch.where(index < 0, index + n, index)
    return flat[index]


def combinations(x, r=2, with_replacement=False, name=None):
    return torch.combinations(T(x), r, with_replacement)


def round(x, decimals=0, name=None):
    return torch.round(T(x), decimals=int(decimals))


def round_(x, decimals=0, name=None):
    with torch.no_grad():
        torch.Tensor.round_(x, decimals=int(decimals))
    return x


def trunc(input, name=None):
    return torch.trunc(T(input))


def cartesian_prod(x, name=None):
    return torch.cartesian_prod(*[T(i) for i in x])


def _store(out, res):
    if out is None:
        return res
    T(out).copy_(res)
    return out


def bitwise_left_shift(x, y, is_arithmetic=True, out=None, name=None):
    return _store(out, torch.bitwise_left_shift(T(x), T(y)))


def bitwise_right_shift(x, y, is_arithmetic=True, out=None, name=None):
    x, y = T(x), T(y)
    if is_arithmetic or x.dtype in (torch.uint8, torch.bool):
        return _store(out, torch.bitwise_right_shift(x, y))
    # logical shift of a signed integer: shift the sign-extended value, then clear the bits that were shifted in
    bits = x.element_size() * 8
    shifted = torch.bitwise_right_shift(x, y)
    ones = torch.full_like(x, -1)
    mask = torch.where(y > 0, ~torch.bitwise_left_shift(ones, (bits - y).clamp(min=0, max=bits - 1)), ones)
    return _store(out, torch.where(y >= bits, torch.zeros_like(x), shifted & mask))


def reduce_as(x, target, name=None):
    x, target = T(x), T(target)
    return x.sum_to_size(*target.size())


def histogram(input, bins=100, min=0, max=0, weight=None, density=False, name=None):  # noqa: A002
    x = T(input)
    xf = x.float() if not x.is_floating_point() else x
    if min == 0 and max == 0:
        min, max = xf.min().item(), xf.max().item()
    h = torch.histc(raw(xf), bins=bins, min=min, max=max) if weight is None else torch.histogram(
        raw(xf).cpu(), bins=bins, range=(float(min), float(max)), weight=raw(weight).cpu().to(xf.dtype), density=density)[0].to(x.device)
    if weight is None and density:
        h = h / (h.sum() * (max - min) / bins)
    return wrap(h if density or weight is not None else h.to(torch.int64))


def histogramdd(x, bins=10, ranges=None, density=False, weights=None, name=None):
    h, edges = torch.histogramdd(raw(x).cpu(), bins=bins, range=ranges, weight=None if weights is None else raw(weights).cpu(), density=density)
    return wrap(h), [wrap(e) for e in edges]


def histogram_bin_edges(input, bins=100, min=0, max=0, name=None):  # noqa: A002
    x = T(input).float()
    if min == 0 and max == 0:
        min, max = x.min().item(), x.max().item()
    return wrap(torch.linspace(min, max, bins + 1, device=x.device))


def bincount(x, weights=None, minlength=0, name=None):
    return torch.bincount(T(x), None if weights is None else T(weights), minlength)


_INPLACE = ["abs", "acos", "asin", "atan", "ceil", "cos", "cosh", "exp", "expm1", "floor", "log", "log2", "log10",
            "log1p", "neg", "reciprocal", "round", "rsqrt", "sin", "sinh", "sqrt", "square", "tan", "tanh", "trunc",
            "erf", "erfinv", "lgamma", "digamma", "sigmoid", "frac", "add", "subtract", "multiply", "divide",
            "remainder", "mod", "floor_divide", "pow", "clip", "scale", "lerp", "cumsum", "cumprod", "nan_to_num",
            "logit", "i0", "acosh", "asinh", "atanh", "floor_mod", "gcd", "lcm", "hypot", "ldexp", "copysign",
            "polygamma", "multigammaln", "bitwise_left_shift", "bitwise_right_shift", "renorm", "addmm", "baddbmm",
            "gammaln", "gammainc", "gammaincc"]
for _n in _INPLACE:
    globals()[_n + "_"] = _mk_inplace(_n)

__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in (
    "np", "torch", "T", "ax", "binary_args", "dt", "raw", "scalar_or_tensor", "shp", "to_int", "to_tensor", "wrap", "annotations")]


# static programs record these as single ops (their bodies compute on raw tensors / read values; framework/recording.py)
from ..framework.recording import make_recordable as _make_recordable  # noqa: E402

_make_recordable(globals(), ['allclose', 'histogramdd'])
