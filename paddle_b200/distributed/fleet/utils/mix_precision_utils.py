"""Parity: fleet/utils/mix_precision_utils.py (MixPrecisionLayer / MixPrecisionOptimizer / MixPrecisionScaler): keep fp32 main
gradients next to low-precision parameters.  Here the flat arenas already hold fp32 master weights (multi_precision) and the
fused AdamW reads bf16 gradients directly, so these wrappers only have to switch those options on."""


class MixPrecisionLayer:
    def __new__(cls, layers, dtype="float16"):
        layers._cast_floating(dtype) if hasattr(layers, "_cast_floating") else None
        return layers


class MixPrecisionOptimizer:
    def __new__(cls, optimizer):
        optimizer._multi_precision = True
        return optimizer


class MixPrecisionScaler:
    def __new__(cls, scaler):
        return scaler


def unscale_method(self, optimizer):
    return self._unscale(optimizer) if hasattr(self, "_unscale") else None
