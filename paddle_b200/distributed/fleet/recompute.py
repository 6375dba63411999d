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
    """Recompute with mp-partitioned saved activations / optional offload (ctx: mp_group, offload, partition)."""
    return recompute(function, *args, **kwargs)
