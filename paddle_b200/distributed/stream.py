"""paddle.distributed.stream.* (collectives with explicit sync_op / use_calc_stream). Parity: communication/stream/*.py."""
from .collective import (all_gather, all_reduce, alltoall, alltoall_single, broadcast, gather, recv, reduce, reduce_scatter, scatter, send)  # noqa: F401
