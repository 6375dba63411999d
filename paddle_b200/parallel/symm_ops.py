"""Windowed peer-memory reduce-scatter / all-gather for arbitrary (non-symmetric) flat tensors.

Used by GroupSharded stage 2/3 (gradient reduce-scatter, parameter all-gather) and anything else that holds its data in
ordinary allocations: chunks are staged through a symmetric window ([world, w] layout), the NVLink traffic is done by the
kernels of csrc/comm/p2p_collectives.cu, and the results are copied out.  Tensors that already live in the symmetric
heap skip the staging copies (see SymmContext.allreduce_)."""
from __future__ import annotations

import os

import torch


def _window_elems(ctx, dtype, world):
    window = int(os.environ.get("B200_SYMM_WINDOW_MB", "256")) << 20
    esize = torch.empty(0, dtype=dtype).element_size()
    per = max(8, window // esize // world)
    return per // 8 * 8          # keep 16-byte vectors for every dtype


def reduce_scatter_into(ctx, out, flat):
    """out[shard] = sum over ranks of flat[rank*shard : (rank+1)*shard] (flat: [world*shard] on every rank)."""
    world, shard = ctx.world, out.numel()
    assert flat.numel() == world * shard
    vec = 16 // flat.element_size()
    if shard % vec:
        return False
    w = min(shard, _window_elems(ctx, flat.dtype, world))
    w = max(vec, w // vec * vec)
    buf, off = ctx.buffer(("rs_win", world), (world, w), flat.dtype)
    src = flat.view(world, shard)
    for lo in range(0, shard, w):
        hi = min(shard, lo + w)
        n = hi - lo
        if n == w:
            buf.copy_(src[:, lo:hi])
            ctx.heap.reduce_scatter(off, out[lo:hi], world * w, ctx.next_epoch())
        else:   # tail: use a dense [world, n] prefix of the window
            tail = buf.view(-1)[: world * n].view(world, n)
            tail.copy_(src[:, lo:hi])
            ctx.heap.reduce_scatter(off, out[lo:hi], world * n, ctx.next_epoch())
    return True


def all_gather_into(ctx, flat_out, shard_t):
    """flat_out[r*shard : (r+1)*shard] = rank r's shard_t."""
    world, shard = ctx.world, shard_t.numel()
    assert flat_out.numel() == world * shard
    vec = 16 // shard_t.element_size()
    if shard % vec:
        return False
    w = min(shard, _window_elems(ctx, shard_t.dtype, world))
    w = max(vec, w // vec * vec)
    buf, off = ctx.buffer(("ag_win", world), (world, w), shard_t.dtype)
    dst = flat_out.view(world, shard)
    flat_shard = shard_t.reshape(-1)
    for lo in range(0, shard, w):
        hi = min(shard, lo + w)
        n = hi - lo
        view = buf if n == w else buf.view(-1)[: world * n].view(world, n)
        view[ctx.rank].copy_(flat_shard[lo:hi])
        ctx.heap.allgather(off, n * shard_t.element_size(), ctx.next_epoch())
        dst[:, lo:hi].copy_(view)
    return True
