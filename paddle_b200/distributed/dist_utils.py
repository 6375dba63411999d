"""paddle.distributed.utils."""
from .fleet.hybrid import fused_allreduce_gradients  # noqa: F401
from .fleet.mp_layers import AllGatherOp, GatherOp, ReduceScatterOp, ScatterOp  # noqa: F401


def global_scatter(*a, **k):
    from ..incubate.moe import global_scatter as f

    return f(*a, **k)


def global_gather(*a, **k):
    from ..incubate.moe import global_gather as f

    return f(*a, **k)
