"""paddle.hapi: high-level Model API. Parity: python/paddle/hapi/{model,model_summary,dynamic_flops}.py."""
from .model import Model  # noqa: F401
from .summary import flops, summary  # noqa: F401
from .. import callbacks  # noqa: F401
