"""Parity: fleet/utils/sequence_parallel_utils.py."""
from ..mp_layers import (AllGatherOp, ColumnSequenceParallelLinear, GatherOp, ReduceScatterOp, RowSequenceParallelLinear, ScatterOp,  # noqa: F401
                         is_sequence_parallel_parameter, mark_as_sequence_parallel_parameter, register_sequence_parallel_allreduce_hooks)


def scatter(x, group=None):
    return ScatterOp.apply(x, group)


def all_gather(x, group=None):
    return AllGatherOp.apply(x, group)


def reduce_scatter(x, group=None):
    return ReduceScatterOp.apply(x, group)
