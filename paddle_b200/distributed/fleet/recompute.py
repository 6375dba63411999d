"""Activation recompute. Parity: python/paddle/distributed/fleet/recompute/recompute.py (recompute, recompute_sequential),
recompute_hybrid.py.  RNG state (incl. the model-parallel tracker) is stashed and replayed so dropout masks match."""
from __future__ import annotations

import torch
import torch.utils.checkpoint as cp

from ...tensor import Tensor
from .random import get_rng_state_tracker


def _wrap_out(o):
    if isinstance(o, torch.Tensor) and not isinstance(o, Tensor):
        return o.as_subclass(Tensor)
    if isinstance(o, (tuple, list)):
        return type(o)(_wrap_out(i) for i in o)
    return o


def recompute(function, *args, **kwargs):
    """Run `function(*args)` without storing activations; re-run it during backward."""
    preserve = kwargs.pop("preserve_rng_state", True)
    kwargs.pop("use_reentrant", None)
    kwargs.pop("offload_indices", None)
    tracker = get_rng_state_tracker()
    saved_tracker = tracker.get_states_tracker() if preserve else None
    first = [True]

    def run(*a):
        if not first[0] and saved_tracker is not None:
            cur = tracker.get_states_tracker()
            tracker.set_states_tracker(saved_tracker)
            try:
                return function(*a, **kwargs)
            finally:
                tracker.set_states_tracker(cur)
        first[0] = False
        return function(*a, **kwargs)

    if not torch.is_grad_enabled():
        return function(*args, **kwargs)
    # determinism_check="none": the metadata check compares `.shape` objects, and paddle Tensors report shapes as lists while the
    # recomputed plain tensors report torch.Size (equal values, different types)
    out = cp.checkpoint(run, *args, use_reentrant=False, preserve_rng_state=preserve, determinism_check="none")
    return _wrap_out(out)


def recompute_sequential(ctx, functions, *args, **kwargs):
    segments = ctx.get("segments", 1) if isinstance(ctx, dict) else 1
    preserve = ctx.get("preserve_rng_state", True) if isinstance(ctx, dict) else True
    layers = list(functions) if not callable(functions) else list(functions.children()) if hasattr(functions, "children") else [functions]
    n = len(layers)
    seg = max(1, n // max(1, segments))

    def run_range(lo, hi):
        def f(x):
            for l in layers[lo:hi]:
                x = l(x)
            return x

        return f

    x = args[0] if len(args) == 1 else args
    end = 0
    for lo in range(0, seg * (segments - 1), seg):
        end = lo + seg
        x = recompute(run_range(lo, end), x, preserve_rng_state=preserve)
    return run_range(end, n)(x)


def recompute_hybrid(ctx, function, *args, **kwargs):
    """Recompute whose *saved inputs* are additionally partitioned over the model-parallel group and / or parked in pinned host
    memory. ctx: {"mp_group": Group, "offload": bool, "partition": bool}.  Parity: recompute_hybrid.py (_HPRecomputeFunction).

    Only the tensors the checkpoint keeps for the backward re-run (the segment inputs) are affected: with `partition` every mp rank
    keeps 1/mp of each (they are replicated over the mp group in tensor-parallel blocks) and the full tensor is all-gathered right
    before the re-run; with `offload` the kept piece lives on the host and is copied back asynchronously."""
    import torch.distributed as dist

    ctx = ctx or {}
    group = ctx.get("mp_group")
    pg = getattr(group, "pg", group)
    world = dist.get_world_size(pg) if (ctx.get("partition") and pg is not None and dist.is_initialized()) else 1
    rank = dist.get_rank(pg) if world > 1 else 0
    offload = bool(ctx.get("offload"))
    if world <= 1 and not offload:
        return recompute(function, *args, **kwargs)

    def pack(t):
        if not isinstance(t, torch.Tensor) or not t.is_floating_point() or t.numel() < world or t.numel() % world:
            return ("raw", t)
        meta = (tuple(t.shape), t.dtype, t.device)
        piece = t.detach().reshape(-1).chunk(world)[rank].clone() if world > 1 else t.detach()
        if offload and piece.is_cuda:
            host = torch.empty(piece.shape, dtype=piece.dtype, pin_memory=True)
            host.copy_(piece, non_blocking=True)
            piece = host
        return ("packed", piece, meta)

    def unpack(obj):
        if obj[0] == "raw":
            return obj[1]
        _, piece, (shape, dtype, device) = obj
        piece = piece.to(device, non_blocking=True) if piece.device != device else piece
        if world > 1:
            full = torch.empty(piece.numel() * world, dtype=dtype, device=device)
            dist.all_gather_into_tensor(full, piece.contiguous(), group=pg) if dist.get_backend(pg) != "gloo" else \
                dist.all_gather(list(full.chunk(world)), piece.contiguous(), group=pg)
            piece = full
        return piece.reshape(shape)

    kwargs.pop("use_reentrant", None)
    preserve = kwargs.pop("preserve_rng_state", True)
    if not torch.is_grad_enabled():
        return function(*args, **kwargs)
    out = _HybridRecompute.apply(lambda *a: function(*a, **kwargs), preserve, pack, unpack, *args)
    return _wrap_out(out)


class _HybridRecompute(torch.autograd.Function):
    """Re-entrant checkpoint that owns its saved inputs, so they can be packed (sharded over mp / moved to pinned host memory)."""

    @staticmethod
    def forward(ctx, fn, preserve, pack, unpack, *args):
        ctx.fn, ctx.unpack, ctx.preserve = fn, unpack, preserve
        ctx.packed = [pack(a) if isinstance(a, torch.Tensor) else ("raw", a) for a in args]
        ctx.req = [isinstance(a, torch.Tensor) and a.requires_grad for a in args]
        if preserve:
            ctx.cpu_rng = torch.get_rng_state()
            ctx.cuda_rng = torch.cuda.get_rng_state() if torch.cuda.is_available() else None
            ctx.tracker = get_rng_state_tracker().get_states_tracker()
        with torch.no_grad():
            out = fn(*args)
        return out

    @staticmethod
    def backward(ctx, *grads):
        ins = []
        for obj, req in zip(ctx.packed, ctx.req):
            v = ctx.unpack(obj)
            if isinstance(v, torch.Tensor):
                v = v.detach().requires_grad_(req)
            ins.append(v)
        tracker = get_rng_state_tracker()
        devices = [torch.cuda.current_device()] if (ctx.preserve and torch.cuda.is_available()) else []
        with torch.random.fork_rng(devices=devices, enabled=ctx.preserve):
            if ctx.preserve:
                torch.set_rng_state(ctx.cpu_rng)
                if ctx.cuda_rng is not None:
                    torch.cuda.set_rng_state(ctx.cuda_rng)
                cur = tracker.get_states_tracker()
                tracker.set_states_tracker(ctx.tracker)
            try:
                with torch.enable_grad():
                    out = ctx.fn(*ins)
            finally:
                if ctx.preserve:
                    tracker.set_states_tracker(cur)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        pairs = [(o, g) for o, g in zip(outs, grads) if isinstance(o, torch.Tensor) and o.requires_grad and g is not None]
        if pairs:
            torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        return (None, None, None, None) + tuple((torch.Tensor.grad.__get__(v) if isinstance(v, torch.Tensor) and r else None) for v, r in zip(ins, ctx.req))
