"""Search / sort ops. Parity: python/paddle/tensor/search.py."""
from __future__ import annotations

import torch

from ._helpers import T, ax, dt, raw, scalar_or_tensor, to_int, wrap


from ..framework.recording import recordable as _recordable  # noqa: E402

def argmax(x, axis=None, keepdim=False, dtype="int64", name=None):
    x = T(x)
    out = torch.argmax(x, dim=ax(axis), keepdim=keepdim if axis is not None else False)
    if axis is None and keepdim:
        out = out.reshape([1] * x.dim())
    return out.to(dt(dtype))


def argmin(x, axis=None, keepdim=False, dtype="int64", name=None):
    x = T(x)
    out = torch.argmin(x, dim=ax(axis), keepdim=keepdim if axis is not None else False)
    if axis is None and keepdim:
        out = out.reshape([1] * x.dim())
    return out.to(dt(dtype))


def argsort(x, axis=-1, descending=False, stable=False, name=None):
    return torch.argsort(T(x), dim=axis, descending=descending, stable=stable)


def sort(x, axis=-1, descending=False, stable=False, name=None):
    return torch.sort(T(x), dim=axis, descending=descending, stable=stable)[0]


@_recordable
def topk(x, k, axis=None, largest=True, sorted=True, name=None):
    x = T(x)
    v, i = torch.topk(x, int(to_int(k)), dim=-1 if axis is None else axis, largest=largest, sorted=sorted)
    return v, i


def kthvalue(x, k, axis=None, keepdim=False, name=None):
    v, i = torch.kthvalue(T(x), k, dim=-1 if axis is None else axis, keepdim=keepdim)
    return v, i


def mode(x, axis=-1, keepdim=False, name=None):
    v, i = torch.mode(T(x), dim=axis, keepdim=keepdim)
    return v, i


def where(condition, x=None, y=None, name=None):
    c = T(condition)
    if x is None and y is None:
        return tuple(torch.nonzero(c, as_tuple=True))
    return torch.where(c, scalar_or_tensor(x), scalar_or_tensor(y))


def where_(condition, x, y, name=None):
    out = where(condition, x, y)
    with torch.no_grad():
        torch.Tensor.copy_(x, out)
    return x


def nonzero(x, as_tuple=False):
    x = T(x)
    if as_tuple:
        return tuple(i.reshape(-1, 1) for i in torch.nonzero(x, as_tuple=True))
    return torch.nonzero(x)


def searchsorted(sorted_sequence, values, out_int32=False, right=False, name=None):
    return torch.searchsorted(T(sorted_sequence), T(values), out_int32=out_int32, right=right)


def bucketize(x, sorted_sequence, out_int32=False, right=False, name=None):
    return torch.bucketize(T(x), T(sorted_sequence), out_int32=out_int32, right=right)


def index_sample(x, index):
    return torch.gather(T(x), 1, T(index).long())


def top_p_sampling(x, ps, threshold=None, topp_seed=None, seed=-1, k=0, mode="truncated", return_top=False, name=None):
    """Nucleus sampling over the last axis. Parity: python/paddle/tensor/search.py:top_p_sampling."""
    x, ps = T(x), T(ps)
    probs, idx = torch.sort(x, dim=-1, descending=True)
    cum = torch.cumsum(probs, -1)
    keep = (cum - probs) < ps.reshape(-1, 1)
    probs = torch.where(keep, probs, torch.zeros_like(probs))
    probs = probs / probs.sum(-1, keepdim=True)
    g = None
    if seed is not None and seed >= 0:
        g = torch.Generator(device=x.device).manual_seed(int(seed))
    choice = torch.multinomial(raw(probs).float(), 1, generator=g)
    ids = torch.gather(raw(idx), -1, choice)
    scores = torch.gather(raw(x), -1, ids)
    return wrap(scores), wrap(ids)


__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "T", "ax", "dt", "raw", "scalar_or_tensor", "to_int", "wrap", "annotations")]
