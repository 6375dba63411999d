"""paddle.quantization. Parity: python/paddle/quantization/{config,qat,ptq,quantize,base_quanter,base_observer,factory}.py,
quanters/abs_max.py, observers/abs_max.py."""
from __future__ import annotations

import copy

import torch

from ..nn.layer import Layer
from ..tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


class _FakeQuant(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, qmax):
        s = scale.clamp(min=1e-9)
        return torch.round(x / s * qmax).clamp(-qmax - 1, qmax) * s / qmax

    @staticmethod
    def backward(ctx, g):
        return g, None, None  # straight-through estimator


class BaseQuanter(Layer):
    def scales(self):
        raise NotImplementedError

    def zero_points(self):
        return None

    def quant_axis(self):
        return -1

    def bit_length(self):
        return 8


class BaseObserver(BaseQuanter):
    def cal_thresholds(self):
        pass


class QuanterFactory:
    def __init__(self, cls, *args, **kwargs):
        self._cls, self._args, self._kwargs = cls, args, kwargs

    def _instance(self, layer):
        return self._cls(layer, *self._args, **self._kwargs)


ObserverFactory = QuanterFactory


class FakeQuanterWithAbsMaxObserverLayer(BaseQuanter):
    def __init__(self, layer=None, moving_rate=0.9, bit_length=8, dtype="float32", name=None):
        super().__init__()
        self._rate, self._bits = moving_rate, bit_length
        self.register_buffer("_scale", torch.ones(1))
        self.register_buffer("_state", torch.zeros(1))
        self.register_buffer("_accum", torch.zeros(1))

    def forward(self, x):
        xr = _raw(x)
        if self.training:
            with torch.no_grad():
                cur = xr.detach().abs().max().float().reshape(1)
                self._state.mul_(self._rate).add_(1.0)
                self._accum.mul_(self._rate).add_(cur)
                self._scale.copy_(self._accum / self._state)
        qmax = float(2 ** (self._bits - 1) - 1)
        return _FakeQuant.apply(xr, _raw(self._scale).to(xr.dtype), qmax).as_subclass(Tensor)

    def scales(self):
        return self._scale

    def bit_length(self):
        return self._bits


def FakeQuanterWithAbsMaxObserver(moving_rate=0.9, bit_length=8, dtype="float32", name=None):
    return QuanterFactory(FakeQuanterWithAbsMaxObserverLayer, moving_rate=moving_rate, bit_length=bit_length, dtype=dtype)


class AbsmaxObserverLayer(BaseObserver):
    def __init__(self, layer=None, quant_bits=8):
        super().__init__()
        self._bits = quant_bits
        self._max = 1e-7

    def forward(self, x):
        self._max = max(self._max, float(_raw(x).detach().abs().max()))
        return x

    def scales(self):
        return torch.tensor([self._max]).as_subclass(Tensor)

    def bit_length(self):
        return self._bits


def AbsmaxObserver(quant_bits=8):
    return QuanterFactory(AbsmaxObserverLayer, quant_bits=quant_bits)


class SingleLayerConfig:
    def __init__(self, activation, weight):
        self.activation, self.weight = activation, weight


class QuantConfig:
    def __init__(self, activation=None, weight=None):
        self._global = SingleLayerConfig(activation, weight)
        self._by_layer, self._by_name, self._by_type = {}, {}, {}
        self._qat_mapping = {}

    def add_layer_config(self, layer, activation=None, weight=None):
        for l in (layer if isinstance(layer, (list, tuple)) else [layer]):
            self._by_layer[id(l)] = SingleLayerConfig(activation, weight)

    def add_name_config(self, layer_name, activation=None, weight=None):
        for n in (layer_name if isinstance(layer_name, (list, tuple)) else [layer_name]):
            self._by_name[n] = SingleLayerConfig(activation, weight)

    def add_type_config(self, layer_type, activation=None, weight=None):
        for t in (layer_type if isinstance(layer_type, (list, tuple)) else [layer_type]):
            self._by_type[t] = SingleLayerConfig(activation, weight)

    def add_qat_layer_mapping(self, source, target):
        self._qat_mapping[source] = target

    def _config_for(self, layer, name):
        if id(layer) in self._by_layer:
            return self._by_layer[id(layer)]
        if name in self._by_name:
            return self._by_name[name]
        for t, c in self._by_type.items():
            if isinstance(layer, t):
                return c
        if self._global.activation is not None or self._global.weight is not None:
            return self._global
        return None


class _QuantedWrapper(Layer):
    """Wraps Linear / Conv layers: fake-quantises the input activation and the weight."""

    def __init__(self, layer, cfg):
        super().__init__()
        self._layer = layer
        self.activation_quanter = cfg.activation._instance(layer) if cfg.activation is not None else None
        self.weight_quanter = cfg.weight._instance(layer) if cfg.weight is not None else None

    def forward(self, x):
        if self.activation_quanter is not None:
            x = self.activation_quanter(x)
        if self.weight_quanter is not None and hasattr(self._layer, "weight"):
            w = self._layer.weight
            qw = self.weight_quanter(w)
            saved = self._layer._parameters["weight"]
            self._layer._parameters.pop("weight")
            object.__setattr__(self._layer, "weight", qw)
            try:
                return self._layer(x)
            finally:
                self._layer.__dict__.pop("weight", None)
                self._layer._parameters["weight"] = saved
        return self._layer(x)


class Quantization:
    def __init__(self, config):
        self._config = config

    def _convert(self, model, inplace):
        from .. import nn

        m = model if inplace else copy.deepcopy(model)

        def walk(parent, prefix):
            for name, sub in list(parent._sub_layers.items()):
                full = prefix + ("." if prefix else "") + name
                cfg = self._config._config_for(sub, full)
                if cfg is not None and isinstance(sub, (nn.Linear, nn.Conv2D, nn.Conv1D, nn.Conv3D)):
                    parent._sub_layers[name] = _QuantedWrapper(sub, cfg)
                else:
                    walk(sub, full)

        walk(m, "")
        return m

    def quantize(self, model, inplace=False):
        return self._convert(model, inplace)

    def convert(self, model, inplace=False, remain_weight=False):
        """Freeze: fold the fake-quantised weights into the layers, drop the quanters."""
        m = model if inplace else copy.deepcopy(model)

        def walk(parent):
            for name, sub in list(parent._sub_layers.items()):
                if isinstance(sub, _QuantedWrapper):
                    if sub.weight_quanter is not None and hasattr(sub._layer, "weight") and not remain_weight:
                        with torch.no_grad():
                            sub.weight_quanter.eval()
                            sub._layer.weight.set_value(sub.weight_quanter(sub._layer.weight))
                    parent._sub_layers[name] = sub._layer
                else:
                    walk(sub)

        walk(m)
        return m


class QAT(Quantization):
    pass


class PTQ(Quantization):
    def quantize(self, model, inplace=False):
        model.eval()
        return self._convert(model, inplace)


def quanter(class_name):
    name = class_name
    def deco(cls):
        globals()[name] = lambda *a, **k: QuanterFactory(cls, *a, **k)
        return cls

    return deco


__all__ = ["QuantConfig", "BaseQuanter", "BaseObserver", "quanter", "QAT", "PTQ"]
