"""paddle.incubate.layers: legacy CTR / ranking helper ops. Parity: python/paddle/incubate/layers/nn.py (the subset that is pure tensor
algebra: shuffle_batch, partial_concat, partial_sum, batch_fc, fused_embedding_seq_pool, pow2_decay_with_linear_warmup, fused_bn_add_act)."""
from __future__ import annotations

import torch

from ..nn import functional as F
from ..nn.layer import _make_parameter
from ..tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor)


def shuffle_batch(x, seed=None):
    """Random permutation of the rows (all leading dims flattened) — negative sampling inside a batch."""
    r = _raw(x)
    lead = r.shape[:-1]
    g = torch.Generator(device="cpu")
    if seed is not None:
        g.manual_seed(int(seed))
    else:
        g.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
    perm = torch.randperm(int(torch.tensor(lead).prod()) if lead else 1, generator=g).to(r.device)
    return r.reshape(-1, r.shape[-1])[perm].reshape(r.shape).as_subclass(Tensor)


def partial_concat(input, start_index=0, length=-1):
    """Concat the column range [start, start+length) of every [N, D_i] input."""
    outs = []
    for t in input:
        r = _raw(t)
        s = start_index if start_index >= 0 else r.shape[1] + start_index
        e = r.shape[1] if length < 0 else s + length
        outs.append(r[:, s:e])
    return torch.cat(outs, 1).as_subclass(Tensor)


def partial_sum(input, start_index=0, length=-1):
    outs = None
    for t in input:
        r = _raw(t)
        s = start_index if start_index >= 0 else r.shape[1] + start_index
        e = r.shape[1] if length < 0 else s + length
        outs = r[:, s:e] if outs is None else outs + r[:, s:e]
    return outs.as_subclass(Tensor)


def batch_fc(input, param_size, param_attr, bias_size, bias_attr, act=None):
    """input [S, B, In] x weight [S, In, Out] + bias [S, Out]: one independent FC per slot."""
    w = _make_parameter(list(param_size), "float32", param_attr)
    b = _make_parameter(list(bias_size), "float32", bias_attr, is_bias=True)
    out = torch.bmm(_raw(input), _raw(w)) + _raw(b).unsqueeze(1)
    out = out.as_subclass(Tensor)
    return getattr(F, act)(out) if act else out


def fused_embedding_seq_pool(input, size, is_sparse=False, padding_idx=None, combiner="sum", param_attr=None, dtype="float32"):
    """Embedding lookup of a LoD id tensor followed by a per-sequence sum."""
    from ..static.nn.sequence import _lod_of, sequence_pool

    w = _make_parameter(list(size), dtype, param_attr)
    ids = _raw(input).reshape(-1).long()
    rows = F.embedding(ids.as_subclass(Tensor), w, padding_idx=padding_idx)
    rows.set_lod(_lod_of(input))
    return sequence_pool(rows, combiner)


def pow2_decay_with_linear_warmup(warmup_steps, total_steps, base_lr, end_lr, dtype="float32", name=None):
    """LR schedule object: linear warm-up to base_lr, then (1 - progress)^2 decay to end_lr."""
    from ..optimizer.lr import LambdaDecay

    def factor(step):
        if step < warmup_steps:
            return (step + 1) / float(warmup_steps)
        if step >= total_steps:
            return end_lr / base_lr
        p = 1.0 - (step - warmup_steps) / float(max(1, total_steps - warmup_steps))
        return (end_lr + (base_lr - end_lr) * p * p) / base_lr

    return LambdaDecay(base_lr, factor)


def fused_bn_add_act(x, y, momentum=0.9, epsilon=1e-05, param_attr=None, bias_attr=None, moving_mean_name=None, moving_variance_name=None, act=None, name=None):
    """act(batch_norm(x) + y) on NHWC tensors."""
    r = _raw(x)
    c = r.shape[-1]
    w = _make_parameter([c], "float32", param_attr, default_initializer=None)
    with torch.no_grad():
        _raw(w).fill_(1.0)
    b = _make_parameter([c], "float32", bias_attr, is_bias=True)
    out = torch.nn.functional.batch_norm(r.movedim(-1, 1), None, None, _raw(w).to(r.dtype), _raw(b).to(r.dtype), True, 1 - momentum, epsilon).movedim(1, -1) + _raw(y)
    out = out.as_subclass(Tensor)
    return getattr(F, act or "relu")(out)
