"""Tensor(model)-parallel layers and ops.

Parity: python/paddle/distributed/fleet/layers/mpu/mp_layers.py (VocabParallelEmbedding, ColumnParallelLinear,
RowParallelLinear, ParallelCrossEntropy), mp_ops.py (_c_identity, _c_concat, _c_split, _mp_allreduce),
fleet/utils/sequence_parallel_utils.py (ScatterOp, GatherOp, AllGatherOp, ReduceScatterOp,
ColumnSequenceParallelLinear, RowSequenceParallelLinear).

B200 design: when the symmetric peer heap is up (``parallel.symm``) and the operands qualify, Row/Column parallel
linears run the fused kernels of csrc/comm/ (GEMM -> reduce-scatter / all-reduce, all-gather -> GEMM over NVSwitch peer
memory).  Otherwise the collective is a torch.distributed call (NCCL on GPU, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...nn import functional as F
from ...nn import initializer as I
from ...nn.layer import Layer
from ...tensor import Tensor
from .. import collective as C
from . import topology as topo
from .random import get_rng_state_tracker


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def _mp_group(group=None):
    if group is not None:
        return group
    hcg = topo.get_hybrid_communicate_group()
    return hcg.get_model_parallel_group() if hcg is not None else None


def _pg(group):
    return group.pg if isinstance(group, C.Group) else group


def _nranks(group):
    if group is None:
        return 1
    return group.nranks if isinstance(group, C.Group) else dist.get_world_size(group)


def _rank(group):
    if group is None:
        return 0
    return group.rank if isinstance(group, C.Group) else dist.get_rank(group)


# ------------------------------------------------------------------------------------------------ autograd collectives
class _Identity(torch.autograd.Function):
    """fwd: identity; bwd: all-reduce (input of a column-parallel region)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if _nranks(ctx.group) > 1:
            from .. import comm_timer as CT

            g = g.contiguous()
            with CT.region("mp_all_reduce"):
                dist.all_reduce(g, group=_pg(ctx.group))
        return g, None


class _AllReduce(torch.autograd.Function):
    """fwd: all-reduce; bwd: identity (output of a row-parallel region)."""

    @staticmethod
    def forward(ctx, x, group):
        if _nranks(group) > 1:
            from .. import comm_timer as CT

            x = x.contiguous().clone()
            with CT.region("mp_all_reduce"):
                dist.all_reduce(x, group=_pg(group))
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None


def _gather_last(x, group):
    n = _nranks(group)
    outs = [torch.empty_like(x) for _ in range(n)]
    dist.all_gather(outs, x.contiguous(), group=_pg(group))
    return torch.cat(outs, -1)


class _Concat(torch.autograd.Function):
    """fwd: all-gather along last dim; bwd: take own slice."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return _gather_last(x, group) if _nranks(group) > 1 else x

    @staticmethod
    def backward(ctx, g):
        n = _nranks(ctx.group)
        if n == 1:
            return g, None
        return g.chunk(n, -1)[_rank(ctx.group)].contiguous(), None


class _Split(torch.autograd.Function):
    """fwd: take own slice of last dim; bwd: all-gather."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        n = _nranks(group)
        return x.chunk(n, -1)[_rank(group)].contiguous() if n > 1 else x

    @staticmethod
    def backward(ctx, g):
        return (_gather_last(g, ctx.group) if _nranks(ctx.group) > 1 else g), None


def _all_gather_dim0(x, group):
    n = _nranks(group)
    from .. import comm_timer as CT

    out = torch.empty((x.shape[0] * n, *x.shape[1:]), dtype=x.dtype, device=x.device)
    with CT.region("mp_all_gather"):
        dist.all_gather_into_tensor(out, x.contiguous(), group=_pg(group))
    return out


def _reduce_scatter_dim0(x, group):
    n = _nranks(group)
    out = torch.empty((x.shape[0] // n, *x.shape[1:]), dtype=x.dtype, device=x.device)
    x = x.contiguous()
    if dist.get_backend(_pg(group)) == "gloo":  # gloo lacks reduce_scatter: all-reduce then slice
        y = x.clone()
        dist.all_reduce(y, group=_pg(group))
        out.copy_(y.chunk(n, 0)[_rank(group)])
    else:
        from .. import comm_timer as CT

        with CT.region("mp_reduce_scatter"):
            dist.reduce_scatter_tensor(out, x, group=_pg(group))
    return out


class ScatterOp(torch.autograd.Function):
    """Sequence-parallel entry: split dim0 (sequence) across mp ranks; bwd all-gathers."""

    @staticmethod
    def forward(ctx, x, group=None):
        group = _mp_group(group)
        ctx.group = group
        n = _nranks(group)
        return x.chunk(n, 0)[_rank(group)].contiguous() if n > 1 else x

    @staticmethod
    def backward(ctx, g):
        return (_all_gather_dim0(g, ctx.group) if _nranks(ctx.group) > 1 else g), None


class GatherOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group=None):
        group = _mp_group(group)
        ctx.group = group
        return _all_gather_dim0(x, group) if _nranks(group) > 1 else x

    @staticmethod
    def backward(ctx, g):
        n = _nranks(ctx.group)
        return (g.chunk(n, 0)[_rank(ctx.group)].contiguous() if n > 1 else g), None


class AllGatherOp(torch.autograd.Function):
    """fwd all-gather(dim0), bwd reduce-scatter(dim0)."""

    @staticmethod
    def forward(ctx, x, group=None):
        group = _mp_group(group)
        ctx.group = group
        return _all_gather_dim0(x, group) if _nranks(group) > 1 else x

    @staticmethod
    def backward(ctx, g):
        return (_reduce_scatter_dim0(g, ctx.group) if _nranks(ctx.group) > 1 else g), None


class ReduceScatterOp(torch.autograd.Function):
    """fwd reduce-scatter(dim0), bwd all-gather(dim0)."""

    @staticmethod
    def forward(ctx, x, group=None):
        group = _mp_group(group)
        ctx.group = group
        return _reduce_scatter_dim0(x, group) if _nranks(group) > 1 else x

    @staticmethod
    def backward(ctx, g):
        return (_all_gather_dim0(g, ctx.group) if _nranks(ctx.group) > 1 else g), None


def _c_identity(tensor, group=None, skip_c_identity_dynamic=False):
    return _w(_Identity.apply(_raw(tensor), _mp_group(group)))


def _mp_allreduce(tensor, op=None, group=None, use_calc_stream=True, use_model_parallel=True, skip_c_identity_dynamic=False):
    return _w(_AllReduce.apply(_raw(tensor), _mp_group(group)))


def _c_concat(tensor, group=None):
    return _w(_Concat.apply(_raw(tensor), _mp_group(group)))


def _c_split(tensor, group=None):
    return _w(_Split.apply(_raw(tensor), _mp_group(group)))


def _mark_dist_shard(p, axis, rank, world):
    """Checkpoint metadata of a tensor-parallel parameter: (global shape, offsets of the local block) - read by
    distributed/checkpoint.py so that save / load re-shard across mp degrees instead of treating every rank's block as the whole."""
    if world <= 1:
        return
    local = [int(d) for d in torch.Tensor.size(p)]
    gshape = list(local)
    gshape[axis] = local[axis] * world
    offs = [0] * len(local)
    offs[axis] = local[axis] * rank
    p.__dict__["_dist_shard"] = (tuple(gshape), tuple(offs))
    p.split_axis = axis


def mark_as_sequence_parallel_parameter(p):
    p.sequence_parallel = True


def is_sequence_parallel_parameter(p):
    return getattr(p, "sequence_parallel", False)


def register_sequence_parallel_allreduce_hooks(model, accumulation_steps=1, fuse_sequence_parallel_allreduce=False):
    """Norm weights/biases replicated across mp ranks see only 1/mp of the sequence under sequence parallelism:
    their grads must be all-reduced over the mp group."""
    group = _mp_group()
    if _nranks(group) <= 1:
        return
    for p in model.parameters():
        if is_sequence_parallel_parameter(p):
            def hook(g, _grp=group, _p=p):
                if _p.__dict__.get("_sp_reduce_in_optimizer", False):
                    return g  # the hybrid optimizer reduces the whole sequence-parallel gradient slab in one collective
                g = g.contiguous()
                dist.all_reduce(g, group=_pg(_grp))
                return g

            p.register_hook(hook)


# ------------------------------------------------------------------------------------------------ layers
class VocabParallelEmbedding(Layer):
    def __init__(self, num_embeddings, embedding_dim, weight_attr=None, mp_group=None, name=None):
        super().__init__()
        self.group = _mp_group(mp_group)
        self.world_size, self.rank = _nranks(self.group), _rank(self.group)
        assert num_embeddings % self.world_size == 0, "vocab size must be divisible by the mp degree"
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.per_part = num_embeddings // self.world_size
        self.vocab_start = self.rank * self.per_part
        with get_rng_state_tracker().rng_state():
            self.weight = self.create_parameter([self.per_part, embedding_dim], attr=weight_attr, default_initializer=I.Normal(0.0, 0.02))
        self.weight.is_distributed = self.world_size > 1
        _mark_dist_shard(self.weight, 0, self.rank, self.world_size)

    def forward(self, x):
        if self.world_size == 1:
            return F.embedding(x, self.weight)
        ids = _raw(x).long()
        local = ids - self.vocab_start
        inside = (local >= 0) & (local < self.per_part)
        emb = torch.nn.functional.embedding(local.clamp(0, self.per_part - 1), _raw(self.weight))
        emb = emb * inside.unsqueeze(-1).to(emb.dtype)
        return _w(_AllReduce.apply(emb, self.group))


class ColumnParallelLinear(Layer):
    """Y = X W, W split along columns: W = [W_1 .. W_p]. Parity: mp_layers.py:ColumnParallelLinear."""

    def __init__(self, in_features, out_features, weight_attr=None, has_bias=None, gather_output=True, fuse_matmul_bias=False,
                 mp_group=None, name=None):
        super().__init__()
        self.group = _mp_group(mp_group)
        self.world_size, self.rank = _nranks(self.group), _rank(self.group)
        assert out_features % self.world_size == 0
        self.in_features, self.out_features = in_features, out_features
        self.out_per_part = out_features // self.world_size
        self.gather_output = gather_output
        with get_rng_state_tracker().rng_state():
            self.weight = self.create_parameter([in_features, self.out_per_part], attr=weight_attr)
        self.weight.is_distributed = self.world_size > 1
        _mark_dist_shard(self.weight, 1, self.rank, self.world_size)
        self.bias = self.create_parameter([self.out_per_part], is_bias=True) if has_bias else None
        if self.bias is not None:
            self.bias.is_distributed = self.world_size > 1
            _mark_dist_shard(self.bias, 0, self.rank, self.world_size)

    def forward(self, x):
        if self.world_size > 1:
            x = _w(_Identity.apply(_raw(x), self.group))
        y = F.linear(x, self.weight, self.bias)
        if self.gather_output and self.world_size > 1:
            y = _w(_Concat.apply(_raw(y), self.group))
        return y


class RowParallelLinear(Layer):
    """Y = X W, W split along rows, X split along its last dim. Parity: mp_layers.py:RowParallelLinear."""

    def __init__(self, in_features, out_features, weight_attr=None, has_bias=True, input_is_parallel=False, fuse_matmul_bias=False,
                 mp_group=None, name=None):
        super().__init__()
        self.group = _mp_group(mp_group)
        self.world_size, self.rank = _nranks(self.group), _rank(self.group)
        assert in_features % self.world_size == 0
        self.in_features, self.out_features = in_features, out_features
        self.in_per_part = in_features // self.world_size
        self.input_is_parallel = input_is_parallel
        with get_rng_state_tracker().rng_state():
            self.weight = self.create_parameter([self.in_per_part, out_features], attr=weight_attr)
        self.weight.is_distributed = self.world_size > 1
        _mark_dist_shard(self.weight, 0, self.rank, self.world_size)
        self.bias = self.create_parameter([out_features], is_bias=True) if has_bias else None

    def forward(self, x):
        if not self.input_is_parallel and self.world_size > 1:
            x = _w(_Split.apply(_raw(x), self.group))
        if self.world_size > 1:
            from ...parallel import fused_mp

            y = fused_mp.row_parallel_linear(x, self.weight, self.group)  # GEMM -> all-reduce (fused over P2P when possible)
            if self.bias is not None:
                y = y + self.bias
            return y
        return F.linear(x, self.weight, self.bias)


class ColumnSequenceParallelLinear(Layer):
    """Sequence-parallel column linear: all-gather(seq) -> GEMM. Input [s/p, b, h] -> output [s, b, out/p]."""

    def __init__(self, in_features, out_features, weight_attr=None, has_bias=None, gather_output=False, fuse_matmul_bias=False,
                 mp_group=None, name=None):
        super().__init__()
        self.group = _mp_group(mp_group)
        self.world_size, self.rank = _nranks(self.group), _rank(self.group)
        assert out_features % self.world_size == 0 and not gather_output
        self.out_per_part = out_features // self.world_size
        with get_rng_state_tracker().rng_state():
            self.weight = self.create_parameter([in_features, self.out_per_part], attr=weight_attr)
        self.weight.is_distributed = self.world_size > 1
        _mark_dist_shard(self.weight, 1, self.rank, self.world_size)
        self.bias = self.create_parameter([self.out_per_part], is_bias=True) if has_bias else None
        if self.bias is not None:
            self.bias.is_distributed = self.world_size > 1
            _mark_dist_shard(self.bias, 0, self.rank, self.world_size)

    def forward(self, x):
        if self.world_size > 1:
            from ...parallel import fused_mp

            y = fused_mp.allgather_linear(x, self.weight, self.group)  # all-gather -> GEMM (fused over P2P when possible)
            return y + self.bias if self.bias is not None else y
        return F.linear(x, self.weight, self.bias)


class RowSequenceParallelLinear(Layer):
    """Sequence-parallel row linear: GEMM -> reduce-scatter(seq). Input [s, b, in/p] -> output [s/p, b, out]."""

    def __init__(self, in_features, out_features, weight_attr=None, has_bias=True, input_is_parallel=True, fuse_matmul_bias=False,
                 mp_group=None, name=None):
        super().__init__()
        self.group = _mp_group(mp_group)
        self.world_size, self.rank = _nranks(self.group), _rank(self.group)
        assert in_features % self.world_size == 0 and input_is_parallel
        self.in_per_part = in_features // self.world_size
        with get_rng_state_tracker().rng_state():
            self.weight = self.create_parameter([self.in_per_part, out_features], attr=weight_attr)
        self.weight.is_distributed = self.world_size > 1
        _mark_dist_shard(self.weight, 0, self.rank, self.world_size)
        self.bias = self.create_parameter([out_features], is_bias=True) if has_bias else None
        if self.bias is not None:
            mark_as_sequence_parallel_parameter(self.bias)

    def forward(self, x):
        if self.world_size > 1:
            from ...parallel import fused_mp

            y = fused_mp.linear_reduce_scatter(x, self.weight, self.group)
            return y + self.bias if self.bias is not None else y
        return F.linear(x, self.weight, self.bias)


class ParallelCrossEntropy(Layer):
    """Vocab-parallel softmax cross-entropy. Parity: mp_layers.py:ParallelCrossEntropy (c_softmax_with_cross_entropy)."""

    def __init__(self, mp_group=None, name=None, ignore_index=-100):
        super().__init__()
        self.group = _mp_group(mp_group)
        self.world_size, self.rank = _nranks(self.group), _rank(self.group)
        self.ignore_index = ignore_index

    def forward(self, input, label):
        from ...kernels import loss as KL

        x = _raw(input)
        lab = _raw(label)
        if lab.dim() == x.dim():
            lab = lab.squeeze(-1)
        if self.world_size == 1:
            loss = KL.softmax_cross_entropy(x.reshape(-1, x.shape[-1]), lab.reshape(-1), self.ignore_index)
        else:
            v = x.shape[-1]
            loss = KL.vocab_parallel_cross_entropy(x.reshape(-1, v), lab.reshape(-1), self.rank * v, _pg(self.group), self.ignore_index)
        return _w(_raw(loss).reshape(*lab.shape, 1))
