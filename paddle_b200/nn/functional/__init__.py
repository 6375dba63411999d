"""paddle.nn.functional. Parity: python/paddle/nn/functional/__init__.py."""
from .activation import *  # noqa: F401,F403
from .attention import *  # noqa: F401,F403
from .common import *  # noqa: F401,F403
from .conv_pool_norm import *  # noqa: F401,F403
from .loss import *  # noqa: F401,F403
from . import activation, attention, common, conv_pool_norm, loss  # noqa: F401
from . import attention as flash_attention_module  # noqa: F401
