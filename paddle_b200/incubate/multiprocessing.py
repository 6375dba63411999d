"""paddle.incubate.multiprocessing: `multiprocessing` with tensors passed through shared memory instead of being pickled by value.
Parity: python/paddle/incubate/multiprocessing/{__init__,reductions}.py.

Everything of the standard module is re-exported; in addition the ForkingPickler learns to send `paddle_b200.Tensor` / `Parameter`
as a handle to their (shared-memory / CUDA-IPC) storage — the receiving process maps the same memory."""
from __future__ import annotations

import multiprocessing as _mp
from multiprocessing import *  # noqa: F401,F403
from multiprocessing.reduction import ForkingPickler

import torch
import torch.multiprocessing.reductions as _tr

__all__ = list(getattr(_mp, "__all__", [])) + ["init_reductions"]


def _rebuild(cls, rebuild_fn, rebuild_args, stop_gradient, name):
    t = rebuild_fn(*rebuild_args)
    out = t.as_subclass(cls)
    try:
        out.stop_gradient = stop_gradient
        if name is not None:
            out.name = name
    except Exception:  # noqa: BLE001  (attributes are best effort)
        pass
    return out


def _reduce(t):
    from ..tensor import Tensor

    raw = t.detach().as_subclass(torch.Tensor)
    fn, args = _tr.reduce_tensor(raw)          # moves CPU storage into shared memory / opens a CUDA IPC handle
    return _rebuild, (Tensor, fn, args, bool(getattr(t, "stop_gradient", True)), getattr(t, "name", None))


def init_reductions():
    from ..tensor import Parameter, Tensor

    _tr.init_reductions()
    ForkingPickler.register(Tensor, _reduce)
    ForkingPickler.register(Parameter, _reduce)


init_reductions()
