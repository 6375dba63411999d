"""Statistics. Parity: python/paddle/tensor/stat.py."""
from __future__ import annotations

import torch

from ._helpers import T, ax, wrap


def var(x, axis=None, unbiased=True, keepdim=False, name=None):
    return torch.var(T(x), dim=ax(axis), correction=1 if unbiased else 0, keepdim=keepdim)


def std(x, axis=None, unbiased=True, keepdim=False, name=None):
    return torch.std(T(x), dim=ax(axis), correction=1 if unbiased else 0, keepdim=keepdim)


def median(x, axis=None, keepdim=False, mode="avg", name=None):
    x = T(x)
    if mode == "avg":
        q = torch.quantile(x.float() if x.dtype not in (torch.float32, torch.float64) else x, 0.5, dim=axis, keepdim=keepdim, interpolation="midpoint")
        return q
    if axis is None:
        v = torch.median(x.reshape(-1))
        return v.reshape([1] * x.dim()) if keepdim else v
    v, i = torch.median(x, dim=axis, keepdim=keepdim)
    return v, i


def nanmedian(x, axis=None, keepdim=False, mode="avg", name=None):
    x = T(x)
    if mode == "avg":
        return torch.nanquantile(x, 0.5, dim=axis, keepdim=keepdim, interpolation="midpoint")
    if axis is None:
        return torch.nanmedian(x)
    return torch.nanmedian(x, dim=axis, keepdim=keepdim)


def quantile(x, q, axis=None, keepdim=False, interpolation="linear", name=None):
    x = T(x)
    qq = torch.as_tensor(q, dtype=x.dtype, device=x.device) if not isinstance(q, torch.Tensor) else q.to(x.dtype)
    a = ax(axis)
    if isinstance(a, tuple):
        nd = x.dim()
        a = tuple(d % nd for d in a)
        keep = [d for d in range(nd) if d not in a]
        xp = x.permute(*keep, *a).flatten(len(keep))
        out = torch.quantile(xp, qq, dim=-1, keepdim=False, interpolation=interpolation)
        if keepdim:
            for d in sorted(a):
                out = out.unsqueeze(d + (1 if qq.dim() > 0 else 0))
        return out
    return torch.quantile(x, qq, dim=a, keepdim=keepdim, interpolation=interpolation)


def nanquantile(x, q, axis=None, keepdim=False, interpolation="linear", name=None):
    x = T(x)
    qq = torch.as_tensor(q, dtype=x.dtype, device=x.device) if not isinstance(q, torch.Tensor) else q.to(x.dtype)
    return torch.nanquantile(x, qq, dim=ax(axis), keepdim=keepdim, interpolation=interpolation)


__all__ = ["var", "std", "median", "nanmedian", "quantile", "nanquantile"]
