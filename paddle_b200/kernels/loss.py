"""Fused softmax cross-entropy (+ vocab-parallel). Parity: phi cross_entropy_with_softmax, c_softmax_with_cross_entropy."""
from __future__ import annotations

import torch

from ..framework.recording import recordable
import torch.nn.functional as F

from . import ext, raw, use_fused, wrap
from ..framework.flags import flag


class _SoftmaxCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index, inplace_bwd):
        lg = logits.contiguous()
        loss, lse = ext().softmax_ce_fwd(lg, labels, ignore_index)
        ctx.save_for_backward(lg, labels, lse)
        ctx.ignore_index, ctx.inplace = ignore_index, inplace_bwd
        return loss

    @staticmethod
    def backward(ctx, dloss):
        lg, labels, lse = ctx.saved_tensors
        return ext().softmax_ce_bwd(lg, labels, lse, dloss.contiguous(), ctx.ignore_index, ctx.inplace), None, None, None


@recordable
def softmax_cross_entropy(logits, labels, ignore_index=-100, inplace_backward=False):
    """Per-row loss (fp32) for logits [N, V] and int labels [N]."""
    logits, labels = raw(logits), raw(labels)
    if use_fused(logits) and logits.dtype in (torch.float32, torch.float16, torch.bfloat16):
        return wrap(_SoftmaxCE.apply(logits, labels, int(ignore_index), bool(inplace_backward)))
    return wrap(F.cross_entropy(logits.float(), labels.long(), ignore_index=ignore_index, reduction="none"))


class _VocabParallelCE(torch.autograd.Function):
    """logits are sharded along the vocab axis across `group`; labels are global ids."""

    @staticmethod
    def forward(ctx, logits, labels, vocab_start, group, ignore_index):
        import torch.distributed as dist

        lg = logits.contiguous()
        fused = use_fused(lg)
        if fused:
            row_max = ext().vp_ce_max(lg)
        else:
            row_max = lg.float().max(-1).values
        dist.all_reduce(row_max, op=dist.ReduceOp.MAX, group=group)
        if fused:
            sumexp, tgt = ext().vp_ce_sumexp(lg, labels, row_max, vocab_start)
        else:
            sumexp = torch.exp(lg.float() - row_max[:, None]).sum(-1)
            local = labels - vocab_start
            inside = (local >= 0) & (local < lg.shape[-1])
            tgt = torch.where(inside, lg.float().gather(1, local.clamp(0, lg.shape[-1] - 1)[:, None]).squeeze(1), torch.zeros_like(sumexp))
        packed = torch.stack([sumexp, tgt])
        dist.all_reduce(packed, group=group)
        sumexp, tgt = packed[0], packed[1]
        loss = torch.log(sumexp) + row_max - tgt
        loss = torch.where(labels == ignore_index, torch.zeros_like(loss), loss)
        ctx.save_for_backward(lg, labels, row_max, sumexp)
        ctx.vocab_start, ctx.ignore_index, ctx.fused = vocab_start, ignore_index, fused
        return loss

    @staticmethod
    def backward(ctx, dloss):
        lg, labels, row_max, sumexp = ctx.saved_tensors
        if ctx.fused:
            g = ext().vp_ce_bwd(lg, labels, row_max, sumexp, dloss.contiguous(), ctx.vocab_start, ctx.ignore_index, False)
        else:
            p = torch.exp(lg.float() - row_max[:, None]) / sumexp[:, None]
            local = labels - ctx.vocab_start
            inside = (local >= 0) & (local < lg.shape[-1])
            oh = torch.zeros_like(p)
            oh[inside, local[inside]] = 1.0
            d = torch.where(labels == ctx.ignore_index, torch.zeros_like(dloss), dloss).float()
            g = ((p - oh) * d[:, None]).to(lg.dtype)
        return g, None, None, None, None


@recordable
def vocab_parallel_cross_entropy(logits, labels, vocab_start, group, ignore_index=-100):
    return wrap(_VocabParallelCE.apply(raw(logits), raw(labels).long(), int(vocab_start), group, int(ignore_index)))
