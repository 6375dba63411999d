"""paddle.geometric. Parity: python/paddle/geometric/{message_passing,math,reindex,sampling}.py."""
from __future__ import annotations

import numpy as np
import torch

from ..tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else torch.as_tensor(np.asarray(t))


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def _scatter(msg, dst, n, reduce_op):
    shape = (n, *msg.shape[1:])
    idx = dst.reshape(-1, *[1] * (msg.dim() - 1)).expand_as(msg)
    if reduce_op == "sum":
        return torch.zeros(shape, dtype=msg.dtype, device=msg.device).scatter_add(0, idx, msg)
    if reduce_op == "mean":
        s = torch.zeros(shape, dtype=msg.dtype, device=msg.device).scatter_add(0, idx, msg)
        c = torch.zeros(n, dtype=msg.dtype, device=msg.device).scatter_add(0, dst, torch.ones_like(dst, dtype=msg.dtype)).clamp(min=1)
        return s / c.reshape(-1, *[1] * (msg.dim() - 1))
    red = {"max": "amax", "min": "amin"}[reduce_op]
    out = torch.zeros(shape, dtype=msg.dtype, device=msg.device).scatter_reduce(0, idx, msg, red, include_self=False)
    return out


def send_u_recv(x, src_index, dst_index, reduce_op="sum", out_size=None, name=None):
    x, s, d = _raw(x), _raw(src_index).long(), _raw(dst_index).long()
    n = int(out_size) if out_size is not None and int(out_size) > 0 else x.shape[0]
    return _w(_scatter(x[s], d, n, reduce_op))


def _msg(a, b, op):
    return {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b}[op]


def send_ue_recv(x, y, src_index, dst_index, message_op="add", reduce_op="sum", out_size=None, name=None):
    x, y, s, d = _raw(x), _raw(y), _raw(src_index).long(), _raw(dst_index).long()
    n = int(out_size) if out_size is not None and int(out_size) > 0 else x.shape[0]
    return _w(_scatter(_msg(x[s], y, message_op), d, n, reduce_op))


def send_uv(x, y, src_index, dst_index, message_op="add", name=None):
    x, y, s, d = _raw(x), _raw(y), _raw(src_index).long(), _raw(dst_index).long()
    return _w(_msg(x[s], y[d], message_op))


def _segment(data, ids, op):
    data, ids = _raw(data), _raw(ids).long()
    n = int(ids.max().item()) + 1 if ids.numel() else 0
    return _w(_scatter(data, ids, n, op))


def segment_sum(data, segment_ids, name=None):
    return _segment(data, segment_ids, "sum")


def segment_mean(data, segment_ids, name=None):
    return _segment(data, segment_ids, "mean")


def segment_max(data, segment_ids, name=None):
    return _segment(data, segment_ids, "max")


def segment_min(data, segment_ids, name=None):
    return _segment(data, segment_ids, "min")


def reindex_graph(x, neighbors, count, value_buffer=None, index_buffer=None, name=None):
    x, nb, cnt = _raw(x), _raw(neighbors), _raw(count)
    nodes = torch.cat([x, nb])
    uniq, inv = torch.unique(nodes, return_inverse=True)
    # keep first-occurrence order (x first, then new neighbours)
    first = torch.full((uniq.numel(),), nodes.numel(), dtype=torch.int64, device=nodes.device)
    first.scatter_reduce_(0, inv, torch.arange(nodes.numel(), device=nodes.device), "amin")
    order = torch.argsort(first)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=order.device)
    new_ids = rank[inv]
    reindex_src = new_ids[x.numel():]
    reindex_dst = torch.repeat_interleave(new_ids[: x.numel()], cnt.long())
    return _w(reindex_src), _w(reindex_dst), _w(uniq[order])


def reindex_heter_graph(x, neighbors, count, value_buffer=None, index_buffer=None, name=None):
    """Several edge types over the same centre nodes: one shared id space, per-type (src, dst) lists concatenated."""
    nb = torch.cat([_raw(n) for n in neighbors])
    xr = _raw(x)
    nodes = torch.cat([xr, nb])
    uniq, inv = torch.unique(nodes, return_inverse=True)
    first = torch.full((uniq.numel(),), nodes.numel(), dtype=torch.int64, device=nodes.device)
    first.scatter_reduce_(0, inv, torch.arange(nodes.numel(), device=nodes.device), "amin")
    order = torch.argsort(first)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=order.device)
    new_ids = rank[inv]
    src, out_nodes = _w(new_ids[xr.numel():]), _w(uniq[order])
    dsts = [torch.repeat_interleave(new_ids[: xr.numel()], _raw(c).long()) for c in count]
    return src, _w(torch.cat(dsts)), out_nodes


def sample_neighbors(row, colptr, input_nodes, sample_size=-1, eids=None, return_eids=False, perm_buffer=None, name=None):
    row, colptr, nodes = _raw(row), _raw(colptr), _raw(input_nodes)
    outs, cnts, oe = [], [], []
    for n in nodes.tolist():
        lo, hi = int(colptr[n]), int(colptr[n + 1])
        idx = torch.arange(lo, hi)
        if 0 <= sample_size < idx.numel():
            idx = idx[torch.randperm(idx.numel())[:sample_size]]
        outs.append(row[idx])
        cnts.append(idx.numel())
        if return_eids and eids is not None:
            oe.append(_raw(eids)[idx])
    res = (_w(torch.cat(outs) if outs else row[:0]), _w(torch.tensor(cnts, dtype=torch.int32)))
    return res + (_w(torch.cat(oe)),) if return_eids and eids is not None else res


def weighted_sample_neighbors(row, colptr, edge_weight, input_nodes, sample_size=-1, eids=None, return_eids=False, name=None):
    row, colptr, w, nodes = _raw(row), _raw(colptr), _raw(edge_weight).float(), _raw(input_nodes)
    outs, cnts = [], []
    for n in nodes.tolist():
        lo, hi = int(colptr[n]), int(colptr[n + 1])
        idx = torch.arange(lo, hi)
        if 0 <= sample_size < idx.numel():
            idx = idx[torch.multinomial(w[lo:hi], sample_size, replacement=False)]
        outs.append(row[idx])
        cnts.append(idx.numel())
    return _w(torch.cat(outs) if outs else row[:0]), _w(torch.tensor(cnts, dtype=torch.int32))


# static programs record these as single ops (their bodies compute on raw tensors; framework/recording.py)
from ..framework.recording import make_recordable as _make_recordable  # noqa: E402

_make_recordable(globals(), ['send_u_recv', 'send_ue_recv', 'send_uv', 'segment_sum', 'segment_mean', 'segment_max', 'segment_min'])
