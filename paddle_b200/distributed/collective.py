"""Collective communication API. Parity: python/paddle/distributed/communication/*.py, collective.py (new_group),
paddle/fluid/distributed/collective/process_group_nccl.cc.  Plumbing = torch.distributed process groups."""
from __future__ import annotations

import pickle

import numpy as np
import torch
import torch.distributed as dist

from ..tensor import Tensor
from . import env


class ReduceOp:
    SUM = 0
    MAX = 1
    MIN = 2
    PROD = 3
    AVG = 4


_OP = {ReduceOp.SUM: dist.ReduceOp.SUM, ReduceOp.MAX: dist.ReduceOp.MAX, ReduceOp.MIN: dist.ReduceOp.MIN,
       ReduceOp.PROD: dist.ReduceOp.PRODUCT, ReduceOp.AVG: dist.ReduceOp.AVG}


class Group:
    """Process group handle. Parity: communication/group.py:Group."""

    def __init__(self, rank_in_group, id, ranks, pg=None, name=None):  # noqa: A002  (reference parameter name)
        self._rank_in_group, self._id, self._ranks, self.pg, self._name = rank_in_group, id, list(ranks), pg, name

    @property
    def rank(self):
        return self._rank_in_group

    @property
    def ranks(self):
        return self._ranks

    @property
    def nranks(self):
        return len(self._ranks)

    world_size = nranks

    @property
    def id(self):
        return self._id

    @property
    def name(self):
        return self._name or f"group_{self._id}"

    @property
    def process_group(self):
        return self.pg

    def is_member(self):
        return self._rank_in_group >= 0

    def get_group_rank(self, rank):
        return self._ranks.index(rank) if rank in self._ranks else -1

    def __repr__(self):
        return f"Group(rank={self.rank}, nranks={self.nranks}, id={self.id}, ranks={self._ranks})"


_groups = {}
_next_gid = [1]


def _global_group():
    g = _groups.get(0)
    if g is None or g.nranks != env.get_world_size():
        g = Group(env.get_rank(), 0, list(range(env.get_world_size())), None, "global")
        _groups[0] = g
    return g


def new_group(ranks=None, backend=None, timeout=None):
    if not env.is_initialized():
        env.init_parallel_env()
    world = env.get_world_size()
    ranks = sorted(ranks) if ranks is not None else list(range(world))
    pg = dist.new_group(ranks=ranks, backend=backend)
    me = env.get_rank()
    gid = _next_gid[0]
    _next_gid[0] += 1
    g = Group(ranks.index(me) if me in ranks else -1, gid, ranks, pg)
    _groups[gid] = g
    return g


def get_group(id=0):  # noqa: A002
    return _global_group() if id == 0 else _groups.get(id)


def _pg(group):
    if group is None:
        return None
    return group.pg if isinstance(group, Group) else group


def _grank(group, rank):
    """global rank -> kept as global (torch takes global ranks for src/dst)."""
    return rank


def is_available():
    return dist.is_available()


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


class _Task:
    def __init__(self, work=None):
        self._work = work

    def wait(self):
        if self._work is not None:
            self._work.wait()
        return True

    def is_completed(self):
        return True if self._work is None else self._work.is_completed()


def _single():
    return not env.is_initialized() or env.get_world_size() == 1


def all_reduce(tensor, op=ReduceOp.SUM, group=None, sync_op=True):
    if _single() or (isinstance(group, Group) and group.nranks == 1):
        return _Task()
    w = dist.all_reduce(_raw(tensor), op=_OP[op], group=_pg(group), async_op=not sync_op)
    return _Task(w)


def all_gather(tensor_list, tensor, group=None, sync_op=True):
    n = group.nranks if isinstance(group, Group) else env.get_world_size()
    t = _raw(tensor)
    if _single() or n == 1:
        tensor_list.clear()
        tensor_list.append(tensor.clone() if isinstance(tensor, torch.Tensor) else tensor)
        return _Task()
    outs = [torch.empty_like(t) for _ in range(n)]
    w = dist.all_gather(outs, t.contiguous(), group=_pg(group), async_op=not sync_op)
    tensor_list.clear()
    tensor_list.extend(o.as_subclass(Tensor) for o in outs)
    return _Task(w)


def all_gather_into_tensor(out, tensor, group=None, sync_op=True):
    if _single():
        _raw(out).copy_(_raw(tensor).reshape(_raw(out).shape))
        return _Task()
    return _Task(dist.all_gather_into_tensor(_raw(out), _raw(tensor).contiguous(), group=_pg(group), async_op=not sync_op))


def all_gather_object(object_list, obj, group=None):
    n = group.nranks if isinstance(group, Group) else env.get_world_size()
    if _single() or n == 1:
        object_list.clear()
        object_list.append(obj)
        return
    outs = [None] * n
    dist.all_gather_object(outs, obj, group=_pg(group))
    object_list.clear()
    object_list.extend(outs)


def broadcast(tensor, src, group=None, sync_op=True):
    if _single() or (isinstance(group, Group) and group.nranks == 1):
        return _Task()
    return _Task(dist.broadcast(_raw(tensor), src=src, group=_pg(group), async_op=not sync_op))


def broadcast_object_list(object_list, src, group=None):
    if _single():
        return
    dist.broadcast_object_list(object_list, src=src, group=_pg(group))


def reduce(tensor, dst, op=ReduceOp.SUM, group=None, sync_op=True):
    if _single():
        return _Task()
    return _Task(dist.reduce(_raw(tensor), dst=dst, op=_OP[op], group=_pg(group), async_op=not sync_op))


def reduce_scatter(tensor, tensor_list, op=ReduceOp.SUM, group=None, sync_op=True):
    if _single():
        _raw(tensor).copy_(_raw(tensor_list[0]))
        return _Task()
    ins = [_raw(t).contiguous() for t in tensor_list]
    return _Task(dist.reduce_scatter(_raw(tensor), ins, op=_OP[op], group=_pg(group), async_op=not sync_op))


def reduce_scatter_tensor(out, tensor, op=ReduceOp.SUM, group=None, sync_op=True):
    if _single():
        _raw(out).copy_(_raw(tensor).reshape(_raw(out).shape))
        return _Task()
    return _Task(dist.reduce_scatter_tensor(_raw(out), _raw(tensor).contiguous(), op=_OP[op], group=_pg(group), async_op=not sync_op))


def scatter(tensor, tensor_list=None, src=0, group=None, sync_op=True):
    if _single():
        if tensor_list:
            _raw(tensor).copy_(_raw(tensor_list[0]))
        return _Task()
    me = env.get_rank()
    ins = [_raw(t).contiguous() for t in tensor_list] if (me == src and tensor_list) else None
    return _Task(dist.scatter(_raw(tensor), ins, src=src, group=_pg(group), async_op=not sync_op))


def scatter_object_list(out_object_list, in_object_list=None, src=0, group=None):
    if _single():
        out_object_list.clear()
        out_object_list.append(in_object_list[0])
        return
    out = [None]
    dist.scatter_object_list(out, in_object_list if env.get_rank() == src else None, src=src, group=_pg(group))
    out_object_list.clear()
    out_object_list.extend(out)


def gather(tensor, gather_list=None, dst=0, group=None, sync_op=True):
    if _single():
        if gather_list is not None:
            gather_list.clear()
            gather_list.append(tensor)
        return _Task()
    me = env.get_rank()
    n = group.nranks if isinstance(group, Group) else env.get_world_size()
    outs = [torch.empty_like(_raw(tensor)) for _ in range(n)] if me == dst else None
    w = dist.gather(_raw(tensor).contiguous(), outs, dst=dst, group=_pg(group), async_op=not sync_op)
    if me == dst and gather_list is not None:
        gather_list.clear()
        gather_list.extend(o.as_subclass(Tensor) for o in outs)
    return _Task(w)


def _alltoall_out_first(out_tensor_list, in_tensor_list, group=None, sync_op=True):
    if _single():
        out_tensor_list.clear()
        out_tensor_list.extend(in_tensor_list)
        return _Task()
    ins = [_raw(t).contiguous() for t in in_tensor_list]
    outs = [torch.empty_like(t) for t in ins] if not out_tensor_list else [_raw(t) for t in out_tensor_list]
    pg = _pg(group)
    backend = dist.get_backend(pg)
    if backend == "gloo":  # gloo has no all_to_all: compose from gathers
        n = len(ins)
        me = dist.get_rank(pg)
        for r in range(n):
            lst = [torch.empty_like(ins[r]) for _ in range(n)] if me == r else None
            gr = dist.get_global_rank(pg, r) if pg is not None else r
            dist.gather(ins[r], lst, dst=gr, group=pg)
            if me == r:
                for j in range(n):
                    outs[j].copy_(lst[j])
        w = None
    else:
        w = dist.all_to_all(outs, ins, group=pg, async_op=not sync_op)
    if not out_tensor_list:
        out_tensor_list.extend(o.as_subclass(Tensor) for o in outs)
    return _Task(w)


def _alltoall_single_out_first(out_tensor, in_tensor, in_split_sizes=None, out_split_sizes=None, group=None, sync_op=True):
    if _single():
        _raw(out_tensor).copy_(_raw(in_tensor))
        return _Task()
    pg = _pg(group)
    if dist.get_backend(pg) == "gloo":
        n = dist.get_world_size(pg)
        i = _raw(in_tensor)
        ins = list(i.split(in_split_sizes, 0)) if in_split_sizes else list(i.chunk(n, 0))
        o = _raw(out_tensor)
        outs = list(o.split(out_split_sizes, 0)) if out_split_sizes else list(o.chunk(n, 0))
        tmp = [torch.empty_like(x) for x in outs]
        _alltoall_out_first(tmp, ins, group)
        for a, b in zip(outs, tmp):
            a.copy_(b)
        return _Task()
    return _Task(dist.all_to_all_single(_raw(out_tensor), _raw(in_tensor).contiguous(), out_split_sizes, in_split_sizes, group=pg, async_op=not sync_op))


def alltoall(in_tensor_list, out_tensor_list, group=None, sync_op=True):
    """Parity: distributed/communication/all_to_all.py:alltoall — the *input* list comes first (the stream.* variant takes the output first)."""
    return _alltoall_out_first(out_tensor_list, in_tensor_list, group, sync_op)


def alltoall_single(in_tensor, out_tensor, in_split_sizes=None, out_split_sizes=None, group=None, sync_op=True):
    """Parity: distributed/communication/all_to_all.py:alltoall_single (input first)."""
    return _alltoall_single_out_first(out_tensor, in_tensor, in_split_sizes, out_split_sizes, group, sync_op)


def send(tensor, dst=0, group=None, sync_op=True):
    if _single():
        return _Task()
    if sync_op:
        dist.send(_raw(tensor).contiguous(), dst=dst, group=_pg(group))
        return _Task()
    return _Task(dist.isend(_raw(tensor).contiguous(), dst=dst, group=_pg(group)))


def recv(tensor, src=0, group=None, sync_op=True):
    if _single():
        return _Task()
    if sync_op:
        dist.recv(_raw(tensor), src=src, group=_pg(group))
        return _Task()
    return _Task(dist.irecv(_raw(tensor), src=src, group=_pg(group)))


def isend(tensor, dst, group=None):
    return send(tensor, dst, group, sync_op=False)


def irecv(tensor, src=None, group=None):
    return recv(tensor, src, group, sync_op=False)


class P2POp:
    def __init__(self, op, tensor, peer, group=None):
        self.op, self.tensor, self.peer, self.group = op, tensor, peer, group


def batch_isend_irecv(p2p_op_list):
    if _single():
        return []
    ops = []
    for o in p2p_op_list:
        fn = dist.isend if o.op in (isend, send, dist.isend) else dist.irecv
        ops.append(dist.P2POp(fn, _raw(o.tensor), o.peer, _pg(o.group)))
    works = dist.batch_isend_irecv(ops)
    return [_Task(w) for w in works]


def barrier(group=None):
    if _single():
        return
    if torch.cuda.is_available() and dist.get_backend(_pg(group)) == "nccl":
        dist.barrier(group=_pg(group), device_ids=[torch.cuda.current_device()])
    else:
        dist.barrier(group=_pg(group))


def wait(tensor, group=None, use_calc_stream=True):
    if torch.cuda.is_available():
        torch.cuda.current_stream().synchronize() if not use_calc_stream else None


def get_backend(group=None):
    return dist.get_backend(_pg(group)) if env.is_initialized() else "none"


def split(x, size, operation, axis=0, num_partitions=1, gather_out=True, weight_attr=None, bias_attr=None, name=None):
    """paddle.distributed.split: build a model-parallel linear/embedding on the fly. Parity: collective.py:split."""
    from .fleet import mp_layers as L

    if operation == "embedding":
        layer = L.VocabParallelEmbedding(size[0], size[1], weight_attr=weight_attr)
        return layer(x)
    if operation == "linear":
        if axis == 0:
            layer = L.RowParallelLinear(size[0], size[1], weight_attr=weight_attr, has_bias=bias_attr is not False, input_is_parallel=False)
        else:
            layer = L.ColumnParallelLinear(size[0], size[1], weight_attr=weight_attr, has_bias=bias_attr is not False, gather_output=gather_out)
        return layer(x)
    raise ValueError(f"unsupported operation {operation}")
