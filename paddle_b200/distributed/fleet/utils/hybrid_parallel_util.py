"""Parity: fleet/utils/hybrid_parallel_util.py."""
from ..hybrid import (broadcast_dp_parameters, broadcast_input_data, broadcast_mp_parameters, broadcast_sep_parameters,  # noqa: F401
                      broadcast_sharding_parameters, fused_allreduce_gradients)


def sharding_reduce_gradients(parameter_list, hcg):
    """Average gradients over the sharding group (stage-1 style, every rank keeps the full gradient)."""
    from ..hybrid import _allreduce_tensors, _n
    import torch

    group = hcg.get_sharding_parallel_group()
    grads = [torch.Tensor.grad.__get__(p) for p in parameter_list if torch.Tensor.grad.__get__(p) is not None]
    _allreduce_tensors(grads, group, 1.0 / _n(group))
