"""paddle.amp. Parity: python/paddle/amp/__init__.py."""
from . import debugging  # noqa: F401
from .auto_cast import amp_guard, auto_cast, decorate, is_bfloat16_supported, is_float16_supported, white_list, black_list  # noqa: F401
from .grad_scaler import AmpScaler, GradScaler, OptimizerState  # noqa: F401

__all__ = ["auto_cast", "GradScaler", "decorate", "is_float16_supported", "is_bfloat16_supported"]
