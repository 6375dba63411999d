"""paddle.distributed.stream.* — collectives with explicit `sync_op` / `use_calc_stream`. Parity: communication/stream/*.py.

`use_calc_stream=True` asks for the collective to be issued on the compute stream (no event hand-over afterwards); with one
process per GPU and NCCL / peer-memory kernels launched on the current stream that is what `sync_op=True` already does, so the flag
only validates its contract (it requires sync_op=True, like the reference)."""
from . import collective as _c


def _check(sync_op, use_calc_stream):
    if use_calc_stream and not sync_op:
        raise RuntimeError("use_calc_stream can only be true in sync op behavior.")


def all_reduce(tensor, op=_c.ReduceOp.SUM, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    return _c.all_reduce(tensor, op, group, sync_op)


def all_gather(tensor_or_tensor_list, tensor, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    if isinstance(tensor_or_tensor_list, (list, tuple)):
        return _c.all_gather(tensor_or_tensor_list, tensor, group, sync_op)
    return _c.all_gather_into_tensor(tensor_or_tensor_list, tensor, group, sync_op)


def alltoall(out_tensor_or_tensor_list, in_tensor_or_tensor_list, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    if isinstance(out_tensor_or_tensor_list, (list, tuple)):
        return _c._alltoall_out_first(out_tensor_or_tensor_list, in_tensor_or_tensor_list, group, sync_op)
    return _c._alltoall_single_out_first(out_tensor_or_tensor_list, in_tensor_or_tensor_list, None, None, group, sync_op)


def alltoall_single(out_tensor, in_tensor, out_split_sizes=None, in_split_sizes=None, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    return _c._alltoall_single_out_first(out_tensor, in_tensor, in_split_sizes, out_split_sizes, group, sync_op)


def broadcast(tensor, src=0, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    return _c.broadcast(tensor, src, group, sync_op)


def reduce(tensor, dst=0, op=_c.ReduceOp.SUM, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    return _c.reduce(tensor, dst, op, group, sync_op)


def reduce_scatter(tensor, tensor_or_tensor_list, op=_c.ReduceOp.SUM, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    if isinstance(tensor_or_tensor_list, (list, tuple)):
        return _c.reduce_scatter(tensor, tensor_or_tensor_list, op, group, sync_op)
    return _c.reduce_scatter_tensor(tensor, tensor_or_tensor_list, op, group, sync_op)


def scatter(tensor, tensor_or_tensor_list=None, src=0, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    if tensor_or_tensor_list is not None and not isinstance(tensor_or_tensor_list, (list, tuple)):
        from . import env

        n = len(group.ranks) if group is not None and hasattr(group, "ranks") else env.get_world_size()
        tensor_or_tensor_list = list(tensor_or_tensor_list.chunk(n, 0))
    return _c.scatter(tensor, tensor_or_tensor_list, src, group, sync_op)


def gather(tensor, gather_list=None, dst=0, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    return _c.gather(tensor, gather_list, dst, group, sync_op)


def send(tensor, dst=0, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    return _c.send(tensor, dst, group, sync_op)


def recv(tensor, src=0, group=None, sync_op=True, use_calc_stream=False):
    _check(sync_op, use_calc_stream)
    return _c.recv(tensor, src, group, sync_op)
