"""Exposed-communication accounting on the device.

`region(name)` brackets a collective that runs ON THE COMPUTE STREAM (or a point where the compute stream waits for a side
stream) with a CUDA event pair; `summary()` synchronises once and returns the summed milliseconds and call counts per name.
Everything between the two events is time in which the compute stream could not run model kernels, i.e. communication that
was not hidden.  Fused compute+collective kernels do not appear here (their transfer overlaps their own tiles); bench.py
calibrates those separately (fused kernel vs the same GEMM without the collective).  Disabled (the default) it costs one
attribute test per call.

Role in the reference: the per-op comm timers of python/paddle/distributed/fleet/utils/timer_helper.py.
"""
from __future__ import annotations

import contextlib

import torch

_state = {"on": False, "events": []}


def enable(flag=True):
    _state["on"] = bool(flag)
    _state["events"].clear()


def enabled():
    return _state["on"]


@contextlib.contextmanager
def region(name):
    if not _state["on"] or not torch.cuda.is_available():
        yield
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    try:
        yield
    finally:
        b.record()
        _state["events"].append((name, a, b))


def summary(reset=True):
    """{name: {"ms": total milliseconds, "calls": n}} since the last reset (synchronises the device)."""
    out = {}
    if _state["events"]:
        torch.cuda.synchronize()
    for name, a, b in _state["events"]:
        d = out.setdefault(name, {"ms": 0.0, "calls": 0})
        d["ms"] += a.elapsed_time(b)
        d["calls"] += 1
    if reset:
        _state["events"].clear()
    return out
