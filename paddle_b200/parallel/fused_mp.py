"""Dispatch for the compute+collective hot ops of tensor parallelism.

  row_parallel_linear   : Y = all_reduce(X_local @ W_local)                  (RowParallelLinear)
  linear_reduce_scatter : Y = reduce_scatter_seq(X @ W_local)                (RowSequenceParallelLinear)
  allgather_linear      : Y = all_gather_seq(X_local) @ W_local              (ColumnSequenceParallelLinear)

Fast path (GPU, symmetric heap initialised, bf16/fp16), see parallel/symm.py for the exact launch sequence of each op:
  * all-gather -> GEMM: two staging copies of the local shard (into the symmetric shard buffer the peers read, and into the gathered
    operand), then ONE CTA-pair GEMM launch whose copy warps pull the peers' shards over NVLink while the tensor cores run;
  * GEMM -> reduce-scatter: ONE GEMM launch whose TMA-store epilogue writes every output row into its owner's staging slot (peer
    HBM) + a small `reduce_slots` kernel that sums `world` local slots;
  * row-parallel all-reduce: GEMM into the symmetric buffer, then the two-shot peer-memory all-reduce kernel, then one copy out.
Fallback: GEMM kernel + torch.distributed collective (bracketed by distributed.comm_timer for the exposed-communication report).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..kernels import wgrad as WG
from ..tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def _pg(group):
    return getattr(group, "pg", group)


def _n(group):
    return group.nranks if hasattr(group, "nranks") else dist.get_world_size(group)


def _r(group):
    return group.rank if hasattr(group, "rank") and not callable(group.rank) else dist.get_rank(_pg(group))


def _fused_ctx(x, group):
    """Returns the symmetric-memory context for `group` if the fused kernels can be used for x, else None."""
    if not x.is_cuda or x.dtype not in (torch.bfloat16, torch.float16):
        return None
    from ..framework.flags import flag

    if not flag("FLAGS_b200_p2p_collectives", True):
        return None
    from . import symm

    return symm.context_for(group)


def _mm(x, w):
    from ..kernels import gemm as KG

    return _raw(KG.linear(x, w, None))


def _all_gather0(x, group):
    from ..distributed import comm_timer as CT

    out = torch.empty((x.shape[0] * _n(group), *x.shape[1:]), dtype=x.dtype, device=x.device)
    with CT.region("mp_all_gather"):
        dist.all_gather_into_tensor(out, x.contiguous(), group=_pg(group))
    return out


def _reduce_scatter0(x, group):
    n = _n(group)
    out = torch.empty((x.shape[0] // n, *x.shape[1:]), dtype=x.dtype, device=x.device)
    if dist.get_backend(_pg(group)) == "gloo":
        y = x.contiguous().clone()
        dist.all_reduce(y, group=_pg(group))
        out.copy_(y.chunk(n, 0)[_r(group)])
    else:
        from ..distributed import comm_timer as CT

        with CT.region("mp_reduce_scatter"):
            dist.reduce_scatter_tensor(out, x.contiguous(), group=_pg(group))
    return out


class _RowParallelLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, group, sink=None):
        ctx.save_for_backward(x, w)
        ctx.group, ctx.sink = group, sink
        sc = _fused_ctx(x, group)
        if sc is not None:
            return sc.gemm_allreduce(x, w)
        from ..distributed import comm_timer as CT

        y = torch.matmul(x, w) if not x.is_cuda else _mm(x.detach(), w.detach())
        with CT.region("mp_all_reduce"):
            dist.all_reduce(y, group=_pg(group))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        from ..kernels import gemm as KG

        dy2 = dy.reshape(-1, dy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dx = KG.gemm(dy2, w, b_is_nk=True).reshape(x.shape) if ctx.needs_input_grad[0] else None
        dw = WG.emit(ctx.sink, x2, dy2) if ctx.needs_input_grad[1] else None
        return dx, dw, None, None


def row_parallel_linear(x, w, group):
    return _w(_RowParallelLinear.apply(_raw(x), _raw(w), group, WG.sink_for(w)))


class _LinearReduceScatter(torch.autograd.Function):
    """x: [S, B, K_local] -> y: [S/p, B, N]"""

    @staticmethod
    def forward(ctx, x, w, group, sink=None):
        ctx.save_for_backward(x, w)
        ctx.group, ctx.sink = group, sink
        sc = _fused_ctx(x, group)
        if sc is not None:
            return sc.gemm_reduce_scatter(x, w)
        y = torch.matmul(x, w) if not x.is_cuda else _mm(x.detach(), w.detach())
        return _reduce_scatter0(y, group)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        from ..kernels import gemm as KG

        sc = _fused_ctx(dy, ctx.group)
        dx = dw = None
        # dX = all_gather(dY) @ W^T  (fused all-gather -> GEMM) ; dW = X^T @ all_gather(dY)
        if sc is not None:
            dx, dy_full = sc.allgather_gemm(dy, w, b_is_nk=True, return_gathered=True)
        else:
            dy_full = _all_gather0(dy, ctx.group)
            dx = KG.gemm(dy_full.reshape(-1, dy_full.shape[-1]), w, b_is_nk=True).reshape(x.shape)
        if ctx.needs_input_grad[1]:
            dw = WG.emit(ctx.sink, x.reshape(-1, x.shape[-1]), dy_full.reshape(-1, dy_full.shape[-1]))
        return dx, dw, None, None


def linear_reduce_scatter(x, w, group):
    return _w(_LinearReduceScatter.apply(_raw(x), _raw(w), group, WG.sink_for(w)))


class _AllGatherLinear(torch.autograd.Function):
    """x: [S/p, B, K] -> y: [S, B, N_local]"""

    @staticmethod
    def forward(ctx, x, w, group, sink=None):
        ctx.group, ctx.sink = group, sink
        sc = _fused_ctx(x, group)
        if sc is not None:
            y, x_full = sc.allgather_gemm(x, w, b_is_nk=False, return_gathered=True)
        else:
            x_full = _all_gather0(x, group)
            y = torch.matmul(x_full, w) if not x.is_cuda else _mm(x_full, w.detach())
        ctx.save_for_backward(x_full, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_full, w = ctx.saved_tensors
        from ..kernels import gemm as KG

        sc = _fused_ctx(dy, ctx.group)
        dy2 = dy.reshape(-1, dy.shape[-1])
        if sc is not None:
            dx = sc.gemm_reduce_scatter(dy, w, b_is_nk=True)   # dX = reduce_scatter(dY @ W^T), fused
        else:
            dx_full = KG.gemm(dy2, w, b_is_nk=True).reshape(x_full.shape)
            dx = _reduce_scatter0(dx_full, ctx.group)
        dw = WG.emit(ctx.sink, x_full.reshape(-1, x_full.shape[-1]), dy2) if ctx.needs_input_grad[1] else None
        return dx, dw, None, None


def allgather_linear(x, w, group):
    return _w(_AllGatherLinear.apply(_raw(x), _raw(w), group, WG.sink_for(w)))
