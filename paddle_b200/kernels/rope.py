"""Rotary position embedding. Parity: python/paddle/incubate/nn/functional/fused_rotary_position_embedding.py."""
from __future__ import annotations

import torch

from ..framework.recording import recordable

from . import ext, raw, use_fused, wrap

_table_cache = {}


def rope_tables(seq_len, dim, base=10000.0, device="cpu"):
    """fp32 cos/sin tables [seq_len, dim/2]."""
    key = (int(seq_len), int(dim), float(base), str(device))
    t = _table_cache.get(key)
    if t is None:
        inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device) / dim))
        ang = torch.outer(torch.arange(seq_len, dtype=torch.float32, device=device), inv)
        t = (ang.cos().contiguous(), ang.sin().contiguous())
        if len(_table_cache) > 64:
            _table_cache.clear()
        _table_cache[key] = t
    return t


class _Rope(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos_t, sin_t, pos_ids, seq, neox):
        ctx.save_for_backward(cos_t, sin_t, pos_ids)
        ctx.seq, ctx.neox = seq, neox
        return ext().rope(x.contiguous(), cos_t, sin_t, pos_ids, seq, neox, False)

    @staticmethod
    def backward(ctx, dy):
        cos_t, sin_t, pos_ids = ctx.saved_tensors
        return ext().rope(dy.contiguous(), cos_t, sin_t, pos_ids, ctx.seq, ctx.neox, True), None, None, None, None, None


def rope_ref(x, cos_t, sin_t, pos_ids, neox):
    # x: [B, S, H, D]; tables [P, D/2]
    b, s, h, d = x.shape
    if pos_ids is None:
        cos, sin = cos_t[:s], sin_t[:s]
        cos, sin = cos[None, :, None, :], sin[None, :, None, :]
    else:
        cos, sin = cos_t[pos_ids][:, :, None, :], sin_t[pos_ids][:, :, None, :]
    xf = x.float()
    if neox:
        a, bb = xf[..., : d // 2], xf[..., d // 2:]
        out = torch.cat([a * cos - bb * sin, bb * cos + a * sin], -1)
    else:
        a, bb = xf[..., 0::2], xf[..., 1::2]
        out = torch.stack([a * cos - bb * sin, bb * cos + a * sin], -1).flatten(-2)
    return out.to(x.dtype)


class _RopePacked(torch.autograd.Function):
    """In-place rotary on the q,k heads of a fused [B, S, (q|k|v heads), D] tensor."""

    @staticmethod
    def forward(ctx, qkv, cos_t, sin_t, pos_ids, seq, rope_heads, total_heads, dim, neox):
        ctx.save_for_backward(cos_t, sin_t, pos_ids)
        ctx.cfg = (seq, rope_heads, total_heads, dim, neox)
        ext().rope_packed_(qkv, cos_t, sin_t, pos_ids, seq, rope_heads, total_heads, dim, neox, False)
        ctx.mark_dirty(qkv)
        return qkv

    @staticmethod
    def backward(ctx, dy):
        cos_t, sin_t, pos_ids = ctx.saved_tensors
        seq, rope_heads, total_heads, dim, neox = ctx.cfg
        g = dy.contiguous()
        if g.data_ptr() == dy.data_ptr():
            g = g.clone()
        ext().rope_packed_(g, cos_t, sin_t, pos_ids, seq, rope_heads, total_heads, dim, neox, True)
        return g, None, None, None, None, None, None, None, None


@recordable
def apply_rope_packed(qkv, cos_t, sin_t, rope_heads, total_heads, dim, position_ids=None, neox=True):
    """qkv: [B, S, total_heads*dim] (or [B,S,total_heads,dim]); rotates heads [0, rope_heads) in place."""
    q = raw(qkv)
    cos_t, sin_t, position_ids = raw(cos_t), raw(sin_t), raw(position_ids)
    if use_fused(q) and q.is_contiguous() and q.dtype in (torch.float32, torch.float16, torch.bfloat16) and (dim // 2) % (16 // q.element_size()) == 0:
        pid = position_ids.reshape(-1).contiguous() if position_ids is not None else None
        if q._is_view() and q.requires_grad:
            q = q.clone()  # autograd forbids in-place updates of views produced inside custom Functions (e.g. fused all-gather GEMM)
        return wrap(_RopePacked.apply(q, cos_t.float().contiguous(), sin_t.float().contiguous(), pid, int(q.shape[1]),
                                      int(rope_heads), int(total_heads), int(dim), bool(neox)))
    b, s = q.shape[0], q.shape[1]
    x4 = q.reshape(b, s, total_heads, dim)
    rot = rope_ref(x4[:, :, :rope_heads], cos_t.float(), sin_t.float(), position_ids, neox)
    return wrap(torch.cat([rot, x4[:, :, rope_heads:]], 2).reshape(q.shape))


@recordable
def apply_rope(x, cos_t, sin_t, position_ids=None, neox=True):
    """x: [B, S, H, D]; cos/sin fp32 [P, D/2]."""
    x = raw(x)
    cos_t, sin_t, position_ids = raw(cos_t), raw(sin_t), raw(position_ids)
    d = x.shape[-1]
    if use_fused(x) and x.dtype in (torch.float32, torch.float16, torch.bfloat16) and (d // 2) % (16 // x.element_size()) == 0:
        pid = position_ids.reshape(-1).contiguous() if position_ids is not None else None
        return wrap(_Rope.apply(x, cos_t.float().contiguous(), sin_t.float().contiguous(), pid, int(x.shape[1]), bool(neox)))
    return wrap(rope_ref(x, cos_t.float(), sin_t.float(), position_ids, neox))
