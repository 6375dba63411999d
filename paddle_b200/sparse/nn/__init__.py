"""paddle.sparse.nn. Parity: python/paddle/sparse/nn/ (ReLU, Softmax, Conv3D, SubmConv3D, BatchNorm, MaxPool3D, attention)."""
from __future__ import annotations

import torch

from ...nn import initializer as I
from ...nn.layer import Layer
from ...tensor import Tensor
from . import functional  # noqa: F401


def _raw(t):
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


def _w(t):
    if t.layout != torch.strided:
        return t
    return t.as_subclass(Tensor) if not isinstance(t, Tensor) else t


class ReLU(Layer):
    def forward(self, x):
        return functional.relu(x)


class ReLU6(Layer):
    def forward(self, x):
        return functional.relu6(x)


class LeakyReLU(Layer):
    def __init__(self, negative_slope=0.01, name=None):
        super().__init__()
        self.s = negative_slope

    def forward(self, x):
        return functional.leaky_relu(x, self.s)


class Softmax(Layer):
    def __init__(self, axis=-1, name=None):
        super().__init__()
        self.axis = axis

    def forward(self, x):
        return functional.softmax(x, self.axis)


class BatchNorm(Layer):
    """BatchNorm over the channel (last) dim of a sparse NDHWC tensor's non-zero values."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-05, weight_attr=None, bias_attr=None, data_format="NDHWC", use_global_stats=None, name=None):
        super().__init__()
        self._m, self._eps = momentum, epsilon
        self.weight = self.create_parameter([num_features], attr=weight_attr, default_initializer=I.Constant(1.0))
        self.bias = self.create_parameter([num_features], attr=bias_attr, is_bias=True)
        self.register_buffer("_mean", torch.zeros(num_features))
        self.register_buffer("_variance", torch.ones(num_features))

    def forward(self, x):
        xr = _raw(x).coalesce()
        v = torch.nn.functional.batch_norm(xr.values(), self._mean, self._variance, self.weight, self.bias, self.training, 1 - self._m, self._eps)
        return _w(torch.sparse_coo_tensor(xr.indices(), v, xr.shape))


SyncBatchNorm = BatchNorm


class _SpConv3D(Layer):
    _subm = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, padding_mode="zeros", key=None,
                 weight_attr=None, bias_attr=None, data_format="NDHWC"):
        super().__init__()
        k = (kernel_size,) * 3 if isinstance(kernel_size, int) else tuple(kernel_size)
        self._cfg = (stride, padding, dilation, groups)
        self.weight = self.create_parameter([*k, in_channels // groups, out_channels], attr=weight_attr)   # DHWIO like the reference
        self.bias = self.create_parameter([out_channels], attr=bias_attr, is_bias=True)

    def forward(self, x):
        return functional._conv3d(x, self.weight, self.bias, *self._cfg, subm=self._subm)


class Conv3D(_SpConv3D):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, padding_mode="zeros",
                 weight_attr=None, bias_attr=None, data_format="NDHWC"):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, None, weight_attr, bias_attr, data_format)


class SubmConv3D(_SpConv3D):
    _subm = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, padding_mode="zeros", key=None,
                 weight_attr=None, bias_attr=None, data_format="NDHWC", backend=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, key, weight_attr, bias_attr, data_format)
        self._backend = backend


class MaxPool3D(Layer):
    def __init__(self, kernel_size, stride=None, padding=0, return_mask=False, ceil_mode=False, data_format="NDHWC", name=None):
        super().__init__()
        self._cfg = (kernel_size, stride, padding, ceil_mode)

    def forward(self, x):
        return functional.max_pool3d(x, *self._cfg)


class _SpConv2D(Layer):
    _subm = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, padding_mode="zeros", key=None,
                 weight_attr=None, bias_attr=None, data_format="NHWC"):
        super().__init__()
        k = (kernel_size,) * 2 if isinstance(kernel_size, int) else tuple(kernel_size)
        self._cfg = (stride, padding, dilation, groups)
        self.weight = self.create_parameter([*k, in_channels // groups, out_channels], attr=weight_attr)   # HWIO like the reference
        self.bias = self.create_parameter([out_channels], attr=bias_attr, is_bias=True)

    def forward(self, x):
        return functional._conv2d(x, self.weight, self.bias, *self._cfg, subm=self._subm)


class Conv2D(_SpConv2D):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, padding_mode="zeros",
                 weight_attr=None, bias_attr=None, data_format="NHWC"):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, None, weight_attr, bias_attr, data_format)


class SubmConv2D(_SpConv2D):
    _subm = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, padding_mode="zeros", key=None,
                 weight_attr=None, bias_attr=None, data_format="NHWC", backend=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, key, weight_attr, bias_attr, data_format)
        self._backend = backend
