from .. import AbsmaxObserver, AbsmaxObserverLayer  # noqa: F401

import torch  # noqa: E402

from .. import BaseObserver, QuanterFactory  # noqa: E402
from ...tensor import Tensor  # noqa: E402


class GroupWiseWeightObserverLayer(BaseObserver):
    """Abs-max scales per group of `group_size` input channels (weight-only int4/int8 quantisation of Linear weights [in, out]).
    Parity: python/paddle/quantization/observers/groupwise.py."""

    def __init__(self, layer=None, quant_bits=8, group_size=128):
        super().__init__()
        self._bits, self.group_size = quant_bits, group_size
        self._scale = None

    def forward(self, x):
        w = x.as_subclass(torch.Tensor).detach().float()
        k = w.shape[0]
        g = self.group_size if (self.group_size > 0 and k % self.group_size == 0) else k
        self._scale = w.reshape(k // g, g, *w.shape[1:]).abs().amax(1).clamp_min(1e-8)     # [in / group, out]
        return x

    def cal_thresholds(self):
        pass

    def scales(self):
        return None if self._scale is None else self._scale.as_subclass(Tensor)

    def zero_points(self):
        return None if self._scale is None else torch.zeros_like(self._scale).as_subclass(Tensor)

    def bit_length(self):
        return self._bits

    def quant_axis(self):
        return 0

    def min_value(self):
        return 0.0

    def max_value(self):
        return float(self._scale.max()) if self._scale is not None else 0.0


def GroupWiseWeightObserver(quant_bits=8, group_size=128):
    return QuanterFactory(GroupWiseWeightObserverLayer, quant_bits=quant_bits, group_size=group_size)
