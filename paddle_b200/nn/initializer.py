"""Initializers. Parity: python/paddle/nn/initializer/*.py.

Paddle weight layout conventions used for fan computation: Linear weight is
[in, out]; Conv weight is [out_c, in_c/groups, *k].
"""
from __future__ import annotations

import math

import numpy as np
import torch

__all__ = ["Initializer", "Constant", "Normal", "TruncatedNormal", "Uniform", "XavierNormal", "XavierUniform",
           "KaimingNormal", "KaimingUniform", "Assign", "Orthogonal", "Dirac", "Bilinear", "calculate_gain",
           "set_global_initializer", "MSRA", "NumpyArrayInitializer"]

_global_weight_init = None
_global_bias_init = None


def set_global_initializer(weight_init, bias_init=None):
    global _global_weight_init, _global_bias_init
    _global_weight_init, _global_bias_init = weight_init, bias_init


def _global_initializer(is_bias):
    return _global_bias_init if is_bias else _global_weight_init


def calculate_gain(nonlinearity, param=None):
    if nonlinearity in ("linear", "conv1d", "conv2d", "conv3d", "conv1d_transpose", "conv2d_transpose",
                        "conv3d_transpose", "sigmoid"):
        return 1.0
    if nonlinearity == "tanh":
        return 5.0 / 3
    if nonlinearity == "relu":
        return math.sqrt(2.0)
    if nonlinearity == "leaky_relu":
        slope = 0.01 if param is None else param
        return math.sqrt(2.0 / (1 + slope ** 2))
    if nonlinearity == "selu":
        return 3.0 / 4
    raise ValueError(f"unsupported nonlinearity {nonlinearity}")


def _fans(t):
    shape = list(t.size())
    if len(shape) == 0:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]  # paddle Linear: [in, out]
    rf = int(np.prod(shape[2:]))
    return shape[1] * rf, shape[0] * rf


class Initializer:
    def __call__(self, param, block=None):
        with torch.no_grad():
            self._init(param.as_subclass(torch.Tensor) if isinstance(param, torch.Tensor) else param)
        return param

    def _init(self, t):
        raise NotImplementedError


class Constant(Initializer):
    def __init__(self, value=0.0):
        self.value = value

    def _init(self, t):
        t.fill_(self.value)


class Normal(Initializer):
    def __init__(self, mean=0.0, std=1.0, name=None):
        self.mean, self.std = mean, std

    def _init(self, t):
        if t.dtype in (torch.bfloat16, torch.float16) and t.device.type == "cpu":
            t.copy_(torch.empty(t.shape, dtype=torch.float32).normal_(self.mean, self.std))
        else:
            t.normal_(self.mean, self.std)


class TruncatedNormal(Initializer):
    def __init__(self, mean=0.0, std=1.0, a=-2.0, b=2.0, name=None):
        self.mean, self.std, self.a, self.b = mean, std, a, b

    def _init(self, t):
        tmp = torch.empty(t.shape, dtype=torch.float32, device=t.device)
        torch.nn.init.trunc_normal_(tmp, self.mean, self.std, self.mean + self.a * self.std, self.mean + self.b * self.std)
        t.copy_(tmp)


class Uniform(Initializer):
    def __init__(self, low=-1.0, high=1.0, name=None):
        self.low, self.high = low, high

    def _init(self, t):
        tmp = torch.empty(t.shape, dtype=torch.float32, device=t.device).uniform_(self.low, self.high)
        t.copy_(tmp)


class XavierNormal(Initializer):
    def __init__(self, fan_in=None, fan_out=None, gain=1.0, name=None):
        self.fan_in, self.fan_out, self.gain = fan_in, fan_out, gain

    def _init(self, t):
        fi, fo = _fans(t)
        fi, fo = self.fan_in or fi, self.fan_out or fo
        Normal(0.0, self.gain * math.sqrt(2.0 / (fi + fo)))._init(t)


class XavierUniform(Initializer):
    def __init__(self, fan_in=None, fan_out=None, gain=1.0, name=None):
        self.fan_in, self.fan_out, self.gain = fan_in, fan_out, gain

    def _init(self, t):
        fi, fo = _fans(t)
        fi, fo = self.fan_in or fi, self.fan_out or fo
        lim = self.gain * math.sqrt(6.0 / (fi + fo))
        Uniform(-lim, lim)._init(t)


class KaimingNormal(Initializer):
    def __init__(self, fan_in=None, negative_slope=0.0, nonlinearity="relu", mode="fan_in"):
        self.fan_in, self.negative_slope, self.nonlinearity, self.mode = fan_in, negative_slope, nonlinearity, mode

    def _init(self, t):
        fi, fo = _fans(t)
        fan = self.fan_in or (fi if self.mode == "fan_in" else fo)
        gain = calculate_gain(self.nonlinearity, self.negative_slope)
        Normal(0.0, gain / math.sqrt(fan))._init(t)


class KaimingUniform(Initializer):
    def __init__(self, fan_in=None, negative_slope=0.0, nonlinearity="relu", mode="fan_in"):
        self.fan_in, self.negative_slope, self.nonlinearity, self.mode = fan_in, negative_slope, nonlinearity, mode

    def _init(self, t):
        fi, fo = _fans(t)
        fan = self.fan_in or (fi if self.mode == "fan_in" else fo)
        gain = calculate_gain(self.nonlinearity, self.negative_slope)
        lim = gain * math.sqrt(3.0 / fan)
        Uniform(-lim, lim)._init(t)


MSRA = KaimingNormal


class Assign(Initializer):
    def __init__(self, value, name=None):
        self.value = value

    def _init(self, t):
        v = self.value
        if isinstance(v, torch.Tensor):
            v = v.detach().as_subclass(torch.Tensor)
        else:
            v = torch.as_tensor(np.asarray(v))
        t.copy_(v.to(device=t.device, dtype=t.dtype).reshape(t.shape))


NumpyArrayInitializer = Assign


class Orthogonal(Initializer):
    def __init__(self, gain=1.0, name=None):
        self.gain = gain

    def _init(self, t):
        tmp = torch.empty(t.shape, dtype=torch.float32, device=t.device)
        torch.nn.init.orthogonal_(tmp, self.gain)
        t.copy_(tmp)


class Dirac(Initializer):
    def __init__(self, groups=1, name=None):
        self.groups = groups

    def _init(self, t):
        tmp = torch.empty(t.shape, dtype=torch.float32, device=t.device)
        torch.nn.init.dirac_(tmp, self.groups)
        t.copy_(tmp)


class Bilinear(Initializer):
    """Bilinear upsampling kernel for transposed conv weights [C, 1, k, k]."""

    def _init(self, t):
        shape = t.shape
        if len(shape) != 4 or shape[2] != shape[3]:
            raise ValueError("Bilinear initializer expects a 4-D square kernel")
        k = shape[3]
        f = math.ceil(k / 2.0)
        c = (2 * f - 1 - f % 2) / (2.0 * f)
        idx = torch.arange(k, dtype=torch.float32)
        w1 = 1 - (idx / f - c).abs()
        w = (w1[:, None] * w1[None, :]).to(t.dtype).to(t.device)
        t.copy_(w.expand(shape))
