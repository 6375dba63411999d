"""Conv / Norm / Pooling layers. Parity: python/paddle/nn/layer/conv.py, norm.py, pooling.py."""
from __future__ import annotations

import numpy as np
import torch

from . import functional as F
from . import initializer as I
from .layer import Layer


def _ntuple(v, n):
    return tuple(v) if isinstance(v, (list, tuple)) else (v,) * n


class _ConvNd(Layer):
    _n = 2
    _transposed = False
    _default_fmt = "NCHW"

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 padding_mode="zeros", weight_attr=None, bias_attr=None, data_format=None, output_padding=0):
        super().__init__()
        n = self._n
        self._in_channels, self._out_channels = in_channels, out_channels
        self._kernel_size = _ntuple(kernel_size, n)
        self._stride, self._padding, self._dilation = stride, padding, dilation
        self._groups, self._padding_mode = groups, padding_mode
        self._data_format = data_format or self._default_fmt
        self._output_padding = output_padding
        if self._transposed:
            shape = [in_channels, out_channels // groups, *self._kernel_size]
        else:
            shape = [out_channels, in_channels // groups, *self._kernel_size]
        fan_in = (in_channels // groups) * int(np.prod(self._kernel_size))
        std = (2.0 / fan_in) ** 0.5
        self.weight = self.create_parameter(shape, attr=weight_attr, default_initializer=I.Normal(0.0, std))
        self.bias = self.create_parameter([out_channels], attr=bias_attr, is_bias=True)

    def extra_repr(self):
        return (f"{self._in_channels}, {self._out_channels}, kernel_size={list(self._kernel_size)}, stride={self._stride}, "
                f"padding={self._padding}, dilation={self._dilation}, groups={self._groups}, data_format={self._data_format}")

    def _pad_input(self, x):
        if self._padding_mode == "zeros" or isinstance(self._padding, str):
            return x, self._padding
        p = _ntuple(self._padding, self._n)
        flat = []
        for v in reversed(p):
            flat += [v, v]
        mode = {"reflect": "reflect", "replicate": "replicate", "circular": "circular"}[self._padding_mode]
        return F.pad(x, flat, mode=mode, data_format=self._data_format), 0


class Conv1D(_ConvNd):
    _n, _default_fmt = 1, "NCL"

    def forward(self, x):
        x, p = self._pad_input(x)
        return F.conv1d(x, self.weight, self.bias, self._stride, p, self._dilation, self._groups, self._data_format)


class Conv2D(_ConvNd):
    _n, _default_fmt = 2, "NCHW"

    def forward(self, x):
        x, p = self._pad_input(x)
        return F.conv2d(x, self.weight, self.bias, self._stride, p, self._dilation, self._groups, self._data_format)


class Conv3D(_ConvNd):
    _n, _default_fmt = 3, "NCDHW"

    def forward(self, x):
        x, p = self._pad_input(x)
        return F.conv3d(x, self.weight, self.bias, self._stride, p, self._dilation, self._groups, self._data_format)


class Conv1DTranspose(_ConvNd):
    _n, _default_fmt, _transposed = 1, "NCL", True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, groups=1, dilation=1,
                 weight_attr=None, bias_attr=None, data_format="NCL"):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, "zeros", weight_attr, bias_attr, data_format, output_padding)

    def forward(self, x, output_size=None):
        return F.conv1d_transpose(x, self.weight, self.bias, self._stride, self._padding, self._output_padding, self._groups, self._dilation, output_size, self._data_format)


class Conv2DTranspose(_ConvNd):
    _n, _default_fmt, _transposed = 2, "NCHW", True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, dilation=1, groups=1,
                 weight_attr=None, bias_attr=None, data_format="NCHW"):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, "zeros", weight_attr, bias_attr, data_format, output_padding)

    def forward(self, x, output_size=None):
        return F.conv2d_transpose(x, self.weight, self.bias, self._stride, self._padding, self._output_padding, self._dilation, self._groups, output_size, self._data_format)


class Conv3DTranspose(_ConvNd):
    _n, _default_fmt, _transposed = 3, "NCDHW", True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, dilation=1, groups=1,
                 weight_attr=None, bias_attr=None, data_format="NCDHW"):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, "zeros", weight_attr, bias_attr, data_format, output_padding)

    def forward(self, x, output_size=None):
        return F.conv3d_transpose(x, self.weight, self.bias, self._stride, self._padding, self._output_padding, self._groups, self._dilation, output_size, self._data_format)


# ----------------------------------------------------------------------------------------------- norm
class _BatchNormBase(Layer):
    def __init__(self, num_features, momentum=0.9, epsilon=1e-05, weight_attr=None, bias_attr=None, data_format="NCHW",
                 use_global_stats=None, name=None):
        super().__init__()
        self._num_features, self._momentum, self._epsilon = num_features, momentum, epsilon
        self._data_format, self._use_global_stats = data_format, use_global_stats
        self.weight = self.create_parameter([num_features], attr=weight_attr, default_initializer=I.Constant(1.0))
        self.bias = self.create_parameter([num_features], attr=bias_attr, is_bias=True)
        self.register_buffer("_mean", torch.zeros(num_features, dtype=torch.float32))
        self.register_buffer("_variance", torch.ones(num_features, dtype=torch.float32))

    def forward(self, x):
        return F.batch_norm(x, self._mean, self._variance, self.weight, self.bias, self.training, self._momentum, self._epsilon,
                            self._data_format, self._use_global_stats)

    def extra_repr(self):
        return f"num_features={self._num_features}, momentum={self._momentum}, epsilon={self._epsilon}"


class BatchNorm1D(_BatchNormBase):
    def __init__(self, num_features, momentum=0.9, epsilon=1e-05, weight_attr=None, bias_attr=None, data_format="NCL", use_global_stats=None, name=None):
        super().__init__(num_features, momentum, epsilon, weight_attr, bias_attr, data_format, use_global_stats)


class BatchNorm2D(_BatchNormBase):
    pass


class BatchNorm3D(_BatchNormBase):
    def __init__(self, num_features, momentum=0.9, epsilon=1e-05, weight_attr=None, bias_attr=None, data_format="NCDHW", use_global_stats=None, name=None):
        super().__init__(num_features, momentum, epsilon, weight_attr, bias_attr, data_format, use_global_stats)


class BatchNorm(_BatchNormBase):
    """Legacy paddle.nn.BatchNorm(num_channels, act=None, ...)."""

    def __init__(self, num_channels, act=None, is_test=False, momentum=0.9, epsilon=1e-05, param_attr=None, bias_attr=None,
                 dtype="float32", data_layout="NCHW", in_place=False, moving_mean_name=None, moving_variance_name=None,
                 do_model_average_for_mean_and_var=True, use_global_stats=False, trainable_statistics=False):
        super().__init__(num_channels, momentum, epsilon, param_attr, bias_attr, data_layout, use_global_stats or None)
        self._act = act

    def forward(self, x):
        y = super().forward(x)
        return getattr(F, self._act)(y) if self._act else y


class SyncBatchNorm(_BatchNormBase):
    """Cross-replica batch norm: statistics all-reduced over the data-parallel group.
    Parity: python/paddle/nn/layer/norm.py:SyncBatchNorm (sync_batch_norm kernel)."""

    def forward(self, x):
        import torch.distributed as dist

        if not (self.training and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return super().forward(x)
        cl = self._data_format in ("NLC", "NHWC", "NDHWC")
        xr = x.as_subclass(torch.Tensor)
        dims = [d for d in range(xr.dim()) if d != (xr.dim() - 1 if cl else 1)]
        xf = xr.float()
        n = torch.tensor([xf.numel() / xf.shape[-1 if cl else 1]], device=xf.device)
        stats = torch.cat([xf.sum(dims), (xf * xf).sum(dims), n])
        stats = _SyncStats.apply(stats)
        c = self._num_features
        total = stats[-1]
        mean = stats[:c] / total
        var = stats[c:2 * c] / total - mean * mean
        with torch.no_grad():
            m = 1.0 - self._momentum
            self._mean.mul_(self._momentum).add_(mean.detach() * m)
            self._variance.mul_(self._momentum).add_(var.detach() * total / (total - 1).clamp(min=1) * m)
        shape = [1] * xr.dim()
        shape[-1 if cl else 1] = c
        y = (xf - mean.reshape(shape)) * torch.rsqrt(var.reshape(shape) + self._epsilon)
        y = y * self.weight.float().reshape(shape) + self.bias.float().reshape(shape)
        from ..tensor import Tensor

        return y.to(x.dtype).as_subclass(Tensor)

    @classmethod
    def convert_sync_batchnorm(cls, layer):
        out = layer
        if isinstance(layer, _BatchNormBase) and not isinstance(layer, SyncBatchNorm):
            out = SyncBatchNorm(layer._num_features, layer._momentum, layer._epsilon, data_format=layer._data_format)
            out.weight, out.bias = layer.weight, layer.bias
            out._buffers["_mean"], out._buffers["_variance"] = layer._mean, layer._variance
        for name, sub in list(layer._sub_layers.items()):
            if sub is not None:
                out._sub_layers[name] = cls.convert_sync_batchnorm(sub)
        return out


class _SyncStats(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stats):
        import torch.distributed as dist

        out = stats.clone()
        dist.all_reduce(out)
        return out

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist

        g = g.clone()
        dist.all_reduce(g)
        return g


class LayerNorm(Layer):
    def __init__(self, normalized_shape, epsilon=1e-05, weight_attr=None, bias_attr=None, name=None):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = [normalized_shape]
        self._normalized_shape, self._epsilon = list(normalized_shape), epsilon
        self.weight = self.create_parameter(self._normalized_shape, attr=weight_attr, default_initializer=I.Constant(1.0))
        self.bias = self.create_parameter(self._normalized_shape, attr=bias_attr, is_bias=True)

    def forward(self, x):
        return F.layer_norm(x, self._normalized_shape, self.weight, self.bias, self._epsilon)

    def extra_repr(self):
        return f"normalized_shape={self._normalized_shape}, epsilon={self._epsilon}"


class RMSNorm(Layer):
    """Parity: paddle.incubate.nn.functional.fused_rms_norm / python/paddle/nn/layer/norm.py:RMSNorm."""

    def __init__(self, hidden_size, epsilon=1e-6, weight_attr=None, name=None):
        super().__init__()
        self._epsilon = epsilon
        self.weight = self.create_parameter([hidden_size], attr=weight_attr, default_initializer=I.Constant(1.0))

    def forward(self, x):
        return F.rms_norm(x, self.weight, self._epsilon)


class GroupNorm(Layer):
    def __init__(self, num_groups, num_channels, epsilon=1e-05, weight_attr=None, bias_attr=None, data_format="NCHW", name=None):
        super().__init__()
        self._num_groups, self._epsilon, self._data_format = num_groups, epsilon, data_format
        self.weight = self.create_parameter([num_channels], attr=weight_attr, default_initializer=I.Constant(1.0))
        self.bias = self.create_parameter([num_channels], attr=bias_attr, is_bias=True)

    def forward(self, x):
        return F.group_norm(x, self._num_groups, self._epsilon, self.weight, self.bias, self._data_format)


class _InstanceNormBase(Layer):
    def __init__(self, num_features, epsilon=1e-05, momentum=0.9, weight_attr=None, bias_attr=None, data_format="NCHW", name=None):
        super().__init__()
        self._epsilon, self._data_format = epsilon, data_format
        if weight_attr is False or bias_attr is False:
            self.scale = self.bias = None
        else:
            self.scale = self.create_parameter([num_features], attr=weight_attr, default_initializer=I.Constant(1.0))
            self.bias = self.create_parameter([num_features], attr=bias_attr, is_bias=True)

    def forward(self, x):
        return F.instance_norm(x, None, None, self.scale, self.bias, True, 0.9, self._epsilon, self._data_format)


class InstanceNorm1D(_InstanceNormBase):
    pass


class InstanceNorm2D(_InstanceNormBase):
    pass


class InstanceNorm3D(_InstanceNormBase):
    pass


class LocalResponseNorm(Layer):
    def __init__(self, size, alpha=1e-4, beta=0.75, k=1.0, data_format="NCHW", name=None):
        super().__init__()
        self.size, self.alpha, self.beta, self.k, self.data_format = size, alpha, beta, k, data_format

    def forward(self, x):
        return F.local_response_norm(x, self.size, self.alpha, self.beta, self.k, self.data_format)


class SpectralNorm(Layer):
    def __init__(self, weight_shape, dim=0, power_iters=1, eps=1e-12, dtype="float32", epsilon=None):
        super().__init__()
        self._dim, self._power_iters, self._epsilon = dim, power_iters, (eps if epsilon is None else epsilon)
        h = weight_shape[dim]
        w = int(np.prod(weight_shape)) // h
        self.weight_u = self.create_parameter([h], default_initializer=I.Normal(0, 1))
        self.weight_v = self.create_parameter([w], default_initializer=I.Normal(0, 1))
        self.weight_u.stop_gradient = True
        self.weight_v.stop_gradient = True

    def forward(self, weight):
        wr = weight.as_subclass(torch.Tensor)
        perm = [self._dim] + [i for i in range(wr.dim()) if i != self._dim]
        mat = wr.permute(*perm).reshape(wr.shape[self._dim], -1)
        u, v = self.weight_u.as_subclass(torch.Tensor), self.weight_v.as_subclass(torch.Tensor)
        with torch.no_grad():
            for _ in range(self._power_iters):
                v = torch.nn.functional.normalize(mat.t() @ u, dim=0, eps=self._epsilon)
                u = torch.nn.functional.normalize(mat @ v, dim=0, eps=self._epsilon)
            self.weight_u.copy_(u)
            self.weight_v.copy_(v)
        sigma = u @ (mat @ v)
        return weight / sigma


# ----------------------------------------------------------------------------------------------- pooling
def _pool_layer(name, fn, arg_names, defaults):
    def __init__(self, *args, **kwargs):
        Layer.__init__(self)
        vals = dict(defaults)
        for k, v in zip(arg_names, args):
            vals[k] = v
        for k, v in kwargs.items():
            if k != "name":
                vals[k] = v
        self._cfg = vals

    def forward(self, x):
        return fn(x, **self._cfg)

    return type(name, (Layer,), {"__init__": __init__, "forward": forward})


AvgPool1D = _pool_layer("AvgPool1D", F.avg_pool1d, ["kernel_size", "stride", "padding", "exclusive", "ceil_mode"],
                        dict(stride=None, padding=0, exclusive=True, ceil_mode=False))
AvgPool2D = _pool_layer("AvgPool2D", F.avg_pool2d, ["kernel_size", "stride", "padding", "ceil_mode", "exclusive", "divisor_override", "data_format"],
                        dict(stride=None, padding=0, ceil_mode=False, exclusive=True, divisor_override=None, data_format="NCHW"))
AvgPool3D = _pool_layer("AvgPool3D", F.avg_pool3d, ["kernel_size", "stride", "padding", "ceil_mode", "exclusive", "divisor_override", "data_format"],
                        dict(stride=None, padding=0, ceil_mode=False, exclusive=True, divisor_override=None, data_format="NCDHW"))
MaxPool1D = _pool_layer("MaxPool1D", F.max_pool1d, ["kernel_size", "stride", "padding", "return_mask", "ceil_mode"],
                        dict(stride=None, padding=0, return_mask=False, ceil_mode=False))
MaxPool2D = _pool_layer("MaxPool2D", F.max_pool2d, ["kernel_size", "stride", "padding", "return_mask", "ceil_mode", "data_format"],
                        dict(stride=None, padding=0, return_mask=False, ceil_mode=False, data_format="NCHW"))
MaxPool3D = _pool_layer("MaxPool3D", F.max_pool3d, ["kernel_size", "stride", "padding", "return_mask", "ceil_mode", "data_format"],
                        dict(stride=None, padding=0, return_mask=False, ceil_mode=False, data_format="NCDHW"))
AdaptiveAvgPool1D = _pool_layer("AdaptiveAvgPool1D", F.adaptive_avg_pool1d, ["output_size"], {})
AdaptiveAvgPool2D = _pool_layer("AdaptiveAvgPool2D", F.adaptive_avg_pool2d, ["output_size", "data_format"], dict(data_format="NCHW"))
AdaptiveAvgPool3D = _pool_layer("AdaptiveAvgPool3D", F.adaptive_avg_pool3d, ["output_size", "data_format"], dict(data_format="NCDHW"))
AdaptiveMaxPool1D = _pool_layer("AdaptiveMaxPool1D", F.adaptive_max_pool1d, ["output_size", "return_mask"], dict(return_mask=False))
AdaptiveMaxPool2D = _pool_layer("AdaptiveMaxPool2D", F.adaptive_max_pool2d, ["output_size", "return_mask"], dict(return_mask=False))
AdaptiveMaxPool3D = _pool_layer("AdaptiveMaxPool3D", F.adaptive_max_pool3d, ["output_size", "return_mask"], dict(return_mask=False))
LPPool1D = _pool_layer("LPPool1D", F.lp_pool1d, ["norm_type", "kernel_size", "stride", "padding", "ceil_mode", "data_format"],
                       dict(stride=None, padding=0, ceil_mode=False, data_format="NCL"))
LPPool2D = _pool_layer("LPPool2D", F.lp_pool2d, ["norm_type", "kernel_size", "stride", "padding", "ceil_mode", "data_format"],
                       dict(stride=None, padding=0, ceil_mode=False, data_format="NCHW"))
MaxUnPool1D = _pool_layer("MaxUnPool1D", lambda x, indices=None, **k: F.max_unpool1d(x, indices, **k), ["kernel_size", "stride", "padding", "data_format", "output_size"],
                          dict(stride=None, padding=0, data_format="NCL", output_size=None))
MaxUnPool2D = _pool_layer("MaxUnPool2D", lambda x, indices=None, **k: F.max_unpool2d(x, indices, **k), ["kernel_size", "stride", "padding", "data_format", "output_size"],
                          dict(stride=None, padding=0, data_format="NCHW", output_size=None))
MaxUnPool3D = _pool_layer("MaxUnPool3D", lambda x, indices=None, **k: F.max_unpool3d(x, indices, **k), ["kernel_size", "stride", "padding", "data_format", "output_size"],
                          dict(stride=None, padding=0, data_format="NCDHW", output_size=None))
FractionalMaxPool2D = _pool_layer("FractionalMaxPool2D", F.fractional_max_pool2d, ["output_size", "kernel_size", "random_u", "return_mask"],
                                  dict(kernel_size=None, random_u=None, return_mask=False))
FractionalMaxPool3D = _pool_layer("FractionalMaxPool3D", F.fractional_max_pool3d, ["output_size", "kernel_size", "random_u", "return_mask"],
                                  dict(kernel_size=None, random_u=None, return_mask=False))


def _unpool_forward(cls, fn):
    def forward(self, x, indices):
        return fn(x, indices, **self._cfg)

    cls.forward = forward


_unpool_forward(MaxUnPool1D, F.max_unpool1d)
_unpool_forward(MaxUnPool2D, F.max_unpool2d)
_unpool_forward(MaxUnPool3D, F.max_unpool3d)
