"""Paged (block) KV-cache attention for serving. Parity: paddle/phi/kernels/fusion/gpu/block_multi_head_attention_kernel.cu
(python/paddle/incubate/nn/functional/block_multihead_attention.py)."""
from __future__ import annotations

import math

import torch

from ...tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def block_attention(qkv, key_cache, value_cache, seq_lens_encoder, seq_lens_decoder, seq_lens_this_time, cu_seqlens_q, block_tables, block_size):
    """qkv: [total_tokens, (H + 2*H_kv) * D] packed over the batch; caches [num_blocks, H_kv, block_size, D]."""
    qkv, kc, vc = _raw(qkv), _raw(key_cache), _raw(value_cache)
    nkv, d = kc.shape[1], kc.shape[3]
    nh = qkv.shape[1] // d - 2 * nkv
    enc, dec, now = _raw(seq_lens_encoder).reshape(-1).tolist(), _raw(seq_lens_decoder).reshape(-1).tolist(), _raw(seq_lens_this_time).reshape(-1).tolist()
    cu = _raw(cu_seqlens_q).reshape(-1).tolist()
    bt = _raw(block_tables)
    out = qkv.new_zeros((qkv.shape[0], nh * d))
    for b in range(len(now)):
        n = now[b]
        if n == 0:
            continue
        rows = qkv[cu[b]:cu[b] + n].reshape(n, nh + 2 * nkv, d)
        q, k, v = rows[:, :nh], rows[:, nh:nh + nkv], rows[:, nh + nkv:]
        past = dec[b] if enc[b] == 0 else 0
        for t in range(n):  # append new K/V into the paged cache
            pos = past + t
            blk, off = int(bt[b, pos // block_size]), pos % block_size
            kc[blk, :, off] = k[t]
            vc[blk, :, off] = v[t]
        total = past + n
        nblk = (total + block_size - 1) // block_size
        blks = bt[b, :nblk].long()
        K = kc[blks].permute(1, 0, 2, 3).reshape(nkv, nblk * block_size, d)[:, :total]
        V = vc[blks].permute(1, 0, 2, 3).reshape(nkv, nblk * block_size, d)[:, :total]
        rep = nh // nkv
        K, V = K.repeat_interleave(rep, 0), V.repeat_interleave(rep, 0)
        s = torch.einsum("nhd,hsd->hns", q.float(), K.float()) / math.sqrt(d)
        qpos = torch.arange(past, total, device=qkv.device)[None, :, None]
        kpos = torch.arange(total, device=qkv.device)[None, None, :]
        s = s.masked_fill(kpos > qpos, float("-inf"))
        o = torch.einsum("hns,hsd->nhd", torch.softmax(s, -1), V.float())
        out[cu[b]:cu[b] + n] = o.reshape(n, nh * d).to(out.dtype)
    return out.as_subclass(Tensor), qkv.as_subclass(Tensor), kc.as_subclass(Tensor), vc.as_subclass(Tensor)
