"""paddle.distributed.communication namespace. Parity: python/paddle/distributed/communication/__init__.py."""
from ..collective import *  # noqa: F401,F403
from ..collective import (ReduceOp, all_gather, all_gather_object, all_reduce, alltoall, alltoall_single, barrier, batch_isend_irecv, broadcast,  # noqa: F401
                          broadcast_object_list, gather, irecv, isend, recv, reduce, reduce_scatter, scatter, scatter_object_list, send, wait)
from .. import stream  # noqa: F401
from . import group  # noqa: F401
