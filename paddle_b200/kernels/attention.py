"""Attention entry point ([B,S,H,D] layout). Parity: paddle flash_attention / scaled_dot_product_attention.

Dispatch: the sm_100a flash kernel (csrc/attention.cu) when available for the shape, else PyTorch SDPA (library path).
"""
from __future__ import annotations

import math

import torch

from ..framework.recording import recordable
import torch.nn.functional as F

from . import raw, use_fused, wrap


def attention_ref(q, k, v, mask=None, dropout_p=0.0, causal=False, scale=None):
    b, sq, h, d = q.shape
    hk = k.shape[2]
    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    if hk != h:
        rep = h // hk
        kt = kt.repeat_interleave(rep, 1)
        vt = vt.repeat_interleave(rep, 1)
    if mask is not None and mask.dtype != torch.bool and mask.dtype != q.dtype:
        mask = mask.to(q.dtype)
    out = F.scaled_dot_product_attention(qt, kt, vt, attn_mask=mask, dropout_p=dropout_p, is_causal=causal and mask is None, scale=scale)
    return out.transpose(1, 2)


@recordable
def attention(q, k, v, mask=None, dropout_p=0.0, causal=False, scale=None):
    q, k, v, mask = raw(q), raw(k), raw(v), raw(mask)
    if causal and mask is not None:
        sq, sk = q.shape[1], k.shape[1]
        cm = torch.ones(sq, sk, dtype=torch.bool, device=q.device).tril(sk - sq)
        mask = (mask & cm) if mask.dtype == torch.bool else mask.masked_fill(~cm, float("-inf"))
        causal = False
    return wrap(attention_ref(q, k, v, mask, dropout_p, causal, scale))
