"""Paged (block) KV-cache attention for serving. Parity: paddle/phi/kernels/fusion/gpu/block_multi_head_attention_kernel.cu
(python/paddle/incubate/nn/functional/block_multihead_attention.py)."""
from __future__ import annotations

import math

import torch

from ...tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _paged_kernels_ok(qkv, kc, d):
    from ...framework.flags import flag

    return qkv.is_cuda and kc.is_cuda and d == 128 and qkv.dtype in (torch.float16, torch.bfloat16) and kc.dtype == qkv.dtype and flag("FLAGS_use_fused_kernels", True)


def block_attention(qkv, key_cache, value_cache, seq_lens_encoder, seq_lens_decoder, seq_lens_this_time, cu_seqlens_q, block_tables, block_size):
    """qkv: [total_tokens, (H + 2*H_kv) * D] packed over the batch; caches [num_blocks, H_kv, block_size, D].

    CUDA path (head_dim 128, fp16 / bf16): the new K / V rows of EVERY sequence are scattered into the paged caches with one indexed write
    (block and row computed on the device from the block table - no per-token Python loop); sequences that decode one token attend to their
    cache through `decode_attention_paged` (csrc/decode_attention.cu: one table lookup per cached row, split-K over the positions), the
    prefill sequences run as ONE packed variable-length causal attention on the tcgen05 kernels (kernels/attention.py column mask)."""
    qkv, kc, vc = _raw(qkv), _raw(key_cache), _raw(value_cache)
    nkv, d = kc.shape[1], kc.shape[3]
    if not _paged_kernels_ok(qkv, kc, d):
        return _block_attention_ref(qkv, key_cache, value_cache, seq_lens_encoder, seq_lens_decoder, seq_lens_this_time, cu_seqlens_q, block_tables, block_size)
    from ..._build import ext
    from ...kernels import attention as KAT

    dev = qkv.device
    nh = qkv.shape[1] // d - 2 * nkv
    enc = _raw(seq_lens_encoder).reshape(-1).to(dev, torch.int64)
    dec = _raw(seq_lens_decoder).reshape(-1).to(dev, torch.int64)
    now = _raw(seq_lens_this_time).reshape(-1).to(dev, torch.int64)
    cu = _raw(cu_seqlens_q).reshape(-1).to(dev, torch.int64)
    bt = _raw(block_tables).to(dev, torch.int32).contiguous()
    nseq = now.numel()
    total_tokens = qkv.shape[0]
    rows = qkv.reshape(total_tokens, nh + 2 * nkv, d)
    q, k, v = rows[:, :nh], rows[:, nh:nh + nkv], rows[:, nh + nkv:]
    # ---- scatter the new K / V rows into the paged caches (all sequences at once)
    tok = torch.arange(total_tokens, device=dev)
    seq_of = torch.bucketize(tok, cu[1:nseq + 1], right=True).clamp(max=nseq - 1)
    valid = tok < cu[nseq]
    past = torch.where(enc > 0, torch.zeros_like(dec), dec)                  # prefill starts at position 0, decode continues after the cache
    pos = past[seq_of] + (tok - cu[seq_of])
    blk = bt.long()[seq_of, (pos // block_size).clamp(max=bt.shape[1] - 1)]
    off = pos % block_size
    sel = valid.nonzero().reshape(-1)
    kc[blk[sel], :, off[sel]] = k[sel]
    vc[blk[sel], :, off[sel]] = v[sel]
    out = qkv.new_zeros((total_tokens, nh * d))
    scale = 1.0 / math.sqrt(d)
    is_dec = (now == 1) & (enc == 0)
    is_pre = (now > 0) & ~is_dec
    # ---- decode sequences: one query token against the paged cache
    if bool(is_dec.any()):
        ids = is_dec.nonzero().reshape(-1)
        qd = q[cu[ids]].contiguous()                                        # [Bd, H, D]
        lens = (dec[ids] + 1).to(torch.int32).contiguous()
        od = ext().decode_attention_paged(qd, kc, vc, lens, bt[ids].contiguous(), scale)
        out[cu[ids]] = od.reshape(ids.numel(), nh * d)
    # ---- prefill sequences: packed varlen causal attention over their new tokens (+ any cached prefix gathered once)
    if bool(is_pre.any()):
        ids = is_pre.nonzero().reshape(-1)
        if bool((past[ids] > 0).any()):       # chunked prefill on top of a cached prefix: rare, keep the simple reference for those
            ref_out, _, _, _ = _block_attention_ref(qkv, kc, vc, seq_lens_encoder, seq_lens_decoder, seq_lens_this_time, cu_seqlens_q, block_tables, block_size)
            sel_tok = (is_pre[seq_of] & valid).nonzero().reshape(-1)
            out[sel_tok] = _raw(ref_out)[sel_tok]
        else:
            sel_tok = (is_pre[seq_of] & valid).nonzero().reshape(-1)
            lens_p = now[ids]
            cu_p = torch.zeros(ids.numel() + 1, dtype=torch.int64, device=dev)
            cu_p[1:] = torch.cumsum(lens_p, 0)
            tot = int(sel_tok.numel())
            pad = (-tot) % 128                                              # the kernels tile 128 rows: pad the packed batch, padded rows are masked out
            qp = torch.cat([q[sel_tok], q.new_zeros(pad, nh, d)]).unsqueeze(0)
            kp = torch.cat([k[sel_tok], k.new_zeros(pad, nkv, d)]).unsqueeze(0)
            vp = torch.cat([v[sel_tok], v.new_zeros(pad, nkv, d)]).unsqueeze(0)
            cm = KAT.colmask_from_cu_seqlens(cu_p, cu_p, tot + pad)
            op = KAT.attention_colmask(qp, kp, vp, cm, causal=True, scale=scale)
            if op is None:
                ref_out, _, _, _ = _block_attention_ref(qkv, kc, vc, seq_lens_encoder, seq_lens_decoder, seq_lens_this_time, cu_seqlens_q, block_tables, block_size)
                out[sel_tok] = _raw(ref_out)[sel_tok]
            else:
                out[sel_tok] = _raw(op)[0, :tot].reshape(tot, nh * d)
    return out.as_subclass(Tensor), qkv.as_subclass(Tensor), kc.as_subclass(Tensor), vc.as_subclass(Tensor)


def _block_attention_ref(qkv, key_cache, value_cache, seq_lens_encoder, seq_lens_decoder, seq_lens_this_time, cu_seqlens_q, block_tables, block_size):
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
