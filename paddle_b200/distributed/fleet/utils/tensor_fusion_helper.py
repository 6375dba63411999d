"""Parity: fleet/utils/tensor_fusion_helper.py (fused_parameters, FusedCommBuffer): here parameters and gradients are fused into
flat arenas once (parallel/arena.py) and buckets are ranges of the gradient slab."""
from ....parallel.arena import ParamArena


class HOOK_ACTION:
    ALL_REDUCE = 0
    REDUCE = 1
    REDUCE_SCATTER = 2


def fused_parameters(parameters, use_main_grad=False, fuse_param=True, comm_overlap=False, comm_group=None, act=None, dst=-1, acc_step=1,
                     scale_after_comm=False, group_size=256 << 20, apply_decay_param_fun=None, **kw):
    """Returns (decay_fused, all_fused, all_buffers): one slab per (dtype, decay class)."""
    fn = (lambda p: 1 if apply_decay_param_fun(p.name) else 0) if apply_decay_param_fun is not None else None
    arena = ParamArena(list(parameters), group_fn=fn)
    slabs = arena.all_slabs()
    decay = [s.data for k, s in arena.slabs.items() if k[2] == 1] if fn is not None else [s.data for s in slabs]
    return decay, [s.data for s in slabs], arena.buckets(group_size)


def obtain_storage(parameters, **kw):
    return ParamArena(list(parameters))
