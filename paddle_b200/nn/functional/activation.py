"""Activation functionals. Parity: python/paddle/nn/functional/activation.py."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ...framework import dtype as _dt
from ...ops._helpers import T, raw, wrap


from ...framework.recording import recordable as _recordable  # noqa: E402

def relu(x, name=None):
    return F.relu(T(x))


def relu_(x, name=None):
    return F.relu_(x)


def relu6(x, name=None):
    return F.relu6(T(x))


def leaky_relu(x, negative_slope=0.01, name=None):
    return F.leaky_relu(T(x), negative_slope)


def leaky_relu_(x, negative_slope=0.01, name=None):
    return F.leaky_relu_(x, negative_slope)


def prelu(x, weight, data_format="NCHW", name=None):
    x, w = T(x), T(weight)
    if data_format in ("NHWC", "NLC", "NDHWC") and w.numel() > 1:
        return torch.where(x >= 0, x, x * w.reshape([1] * (x.dim() - 1) + [-1]))
    return F.prelu(x, w)


@_recordable
def rrelu(x, lower=1.0 / 8, upper=1.0 / 3, training=True, name=None):
    return F.rrelu(T(x), lower, upper, training)


def elu(x, alpha=1.0, name=None):
    return F.elu(T(x), alpha)


def elu_(x, alpha=1.0, name=None):
    return F.elu_(x, alpha)


def selu(x, scale=1.0507009873554804934193349852946, alpha=1.6732632423543772848170429916717, name=None):
    x = T(x)
    return scale * torch.where(x > 0, x, alpha * (torch.exp(x) - 1))


def celu(x, alpha=1.0, name=None):
    return F.celu(T(x), alpha)


def gelu(x, approximate=False, name=None):
    return F.gelu(T(x), approximate="tanh" if approximate else "none")


def silu(x, name=None):
    return F.silu(T(x))


swish = silu


def mish(x, name=None):
    return F.mish(T(x))


def sigmoid(x, name=None):
    return torch.sigmoid(T(x))


def hardsigmoid(x, slope=0.1666667, offset=0.5, name=None):
    return torch.clamp(T(x) * slope + offset, 0.0, 1.0)


def hardswish(x, name=None):
    return F.hardswish(T(x))


def hardtanh(x, min=-1.0, max=1.0, name=None):  # noqa: A002
    return F.hardtanh(T(x), min, max)


def hardtanh_(x, min=-1.0, max=1.0, name=None):  # noqa: A002
    return F.hardtanh_(x, min, max)


def hardshrink(x, threshold=0.5, name=None):
    return F.hardshrink(T(x), threshold)


def softshrink(x, threshold=0.5, name=None):
    return F.softshrink(T(x), threshold)


def tanhshrink(x, name=None):
    return F.tanhshrink(T(x))


def tanh(x, name=None):
    return torch.tanh(T(x))


def tanh_(x, name=None):
    return torch.tanh_(x)


def softplus(x, beta=1, threshold=20, name=None):
    return F.softplus(T(x), beta, threshold)


def softsign(x, name=None):
    return F.softsign(T(x))


def log_sigmoid(x, name=None):
    return F.logsigmoid(T(x))


def thresholded_relu(x, threshold=1.0, value=0.0, name=None):
    x = T(x)
    return torch.where(x > threshold, x, torch.full_like(x, value))


def thresholded_relu_(x, threshold=1.0, value=0.0, name=None):
    out = thresholded_relu(x, threshold, value)
    with torch.no_grad():
        torch.Tensor.copy_(x, out)
    return x


def maxout(x, groups, axis=1, name=None):
    x = T(x)
    axis = axis % x.dim()
    s = list(x.size())
    s[axis:axis + 1] = [s[axis] // groups, groups]
    return torch.amax(x.reshape(s), dim=axis + 1)


def softmax(x, axis=-1, dtype=None, name=None):
    from ...amp.auto_cast import black_dtype

    x = T(x)
    return F.softmax(x, dim=axis, dtype=black_dtype("softmax", x, _dt.convert_dtype(dtype)))


def softmax_(x, axis=-1, dtype=None, name=None):
    out = softmax(x, axis, dtype)
    with torch.no_grad():
        torch.Tensor.copy_(x, out)
    return x


def log_softmax(x, axis=-1, dtype=None, name=None):
    from ...amp.auto_cast import black_dtype

    x = T(x)
    return F.log_softmax(x, dim=axis, dtype=black_dtype("softmax", x, _dt.convert_dtype(dtype)))


@_recordable
def gumbel_softmax(x, temperature=1.0, hard=False, axis=-1, name=None):
    return F.gumbel_softmax(T(x), tau=temperature, hard=hard, dim=axis)


def glu(x, axis=-1, name=None):
    return F.glu(T(x), dim=axis)


def swiglu(x, y=None, name=None):
    """silu(x) * y. Parity: python/paddle/incubate/nn/functional/swiglu.py."""
    from ...kernels import activation as K

    return K.swiglu(x, y)


__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "F", "T", "raw", "wrap", "annotations")]
