"""LoD sequence ops. Parity: python/paddle/static/nn/sequence_lod.py over paddle/phi/kernels/*sequence* (sequence_pool,
sequence_softmax, sequence_expand, sequence_pad / unpad, sequence_conv, ...).

A LoD tensor here is a Tensor of packed rows [total, ...] carrying level-0 offsets (`t.set_lod([[0, 2, 5]])` or
`base.create_lod_tensor(data, [[2, 3]])`).  Every op is a handful of vectorised gathers / segment reductions on the packed rows —
no per-sequence Python loops on the data path — and the result carries the LoD it should have."""
from __future__ import annotations

import torch

from ...nn import functional as F
from ...nn import initializer as I
from ...nn.layer import _make_parameter
from ...tensor import Tensor


# ---- LoD plumbing ---------------------------------------------------------------------------------------------------------
def _lod_of(t):
    lod = getattr(t, "__dict__", {}).get("_lod")
    if not lod:
        raise ValueError("this op needs a LoD tensor: call tensor.set_lod([[0, n1, n1+n2, ...]]) or base.create_lod_tensor first")
    return lod


def _offsets(t, level=-1):
    return torch.as_tensor(_lod_of(t)[level], dtype=torch.long, device=t.device)


def _with_lod(t, offsets_list):
    t = t if isinstance(t, Tensor) else t.as_subclass(Tensor)
    t.__dict__["_lod"] = [[int(v) for v in o] for o in offsets_list]
    return t


def create_lod_tensor(data, recursive_seq_lens, place=None):
    """Parity: python/paddle/base/lod_tensor.py:create_lod_tensor. `data`: ndarray / Tensor of packed rows, or a list of lists."""
    import numpy as np

    from ... import to_tensor

    if isinstance(data, list) and data and isinstance(data[0], (list, tuple)):
        recursive_seq_lens = [[len(s) for s in data]] if recursive_seq_lens is None else recursive_seq_lens
        data = np.concatenate([np.asarray(s).reshape(len(s), -1) for s in data], 0)
    t = data if isinstance(data, torch.Tensor) else to_tensor(np.asarray(data))
    t = t if isinstance(t, Tensor) else t.as_subclass(Tensor)
    return t.set_recursive_sequence_lengths(recursive_seq_lens)


def _seg_ids(off):
    lens = off[1:] - off[:-1]
    return torch.repeat_interleave(torch.arange(lens.numel(), device=off.device), lens), lens


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else torch.as_tensor(t)


# ---- ops -------------------------------------------------------------------------------------------------------------------
def sequence_pool(input, pool_type, is_test=False, pad_value=0.0):
    x, off = _raw(input), _offsets(input)
    seg, lens = _seg_ids(off)
    n = lens.numel()
    flat = x.reshape(x.shape[0], -1)
    kind = pool_type.lower()
    if kind in ("sum", "average", "sqrt"):
        out = torch.zeros(n, flat.shape[1], dtype=flat.dtype, device=x.device).index_add_(0, seg, flat)
        if kind == "average":
            out = out / lens.clamp(min=1).to(out.dtype)[:, None]
        elif kind == "sqrt":
            out = out / lens.clamp(min=1).to(out.dtype).sqrt()[:, None]
    elif kind == "max":
        out = torch.full((n, flat.shape[1]), float("-inf"), dtype=flat.dtype, device=x.device).scatter_reduce_(
            0, seg[:, None].expand_as(flat), flat, "amax", include_self=True)
    elif kind in ("first", "last"):
        idx = off[:-1] if kind == "first" else (off[1:] - 1)
        out = flat[idx.clamp(0, max(flat.shape[0] - 1, 0))]
    else:
        raise ValueError(f"sequence_pool: unknown pool_type {pool_type!r}")
    out = torch.where((lens == 0)[:, None], torch.full_like(out, pad_value), out)   # empty sequences
    return out.reshape(n, *x.shape[1:]).as_subclass(Tensor)


def sequence_first_step(input):
    return sequence_pool(input, "first")


def sequence_last_step(input):
    return sequence_pool(input, "last")


def sequence_softmax(input, use_cudnn=False, name=None):
    x, off = _raw(input), _offsets(input)
    seg, lens = _seg_ids(off)
    v = x.reshape(-1)
    mx = torch.full((lens.numel(),), float("-inf"), dtype=v.dtype, device=v.device).scatter_reduce_(0, seg, v, "amax", include_self=True)
    e = torch.exp(v - mx[seg])
    den = torch.zeros(lens.numel(), dtype=v.dtype, device=v.device).index_add_(0, seg, e)
    return _with_lod((e / den[seg]).reshape(x.shape), _lod_of(input))


def sequence_concat(input, name=None):
    """Concatenate the i-th sequences of every input: out_i = [x1_i; x2_i; ...]."""
    offs = [_offsets(t) for t in input]
    n = offs[0].numel() - 1
    lens = torch.stack([o[1:] - o[:-1] for o in offs], 1)                 # [n, k]
    out_off = torch.cat([lens.new_zeros(1), lens.sum(1).cumsum(0)])
    start_in_out = out_off[:-1, None] + torch.cat([lens.new_zeros(n, 1), lens.cumsum(1)[:, :-1]], 1)   # [n, k]
    rows = torch.cat([_raw(t) for t in input], 0)
    dest = []
    for k, o in enumerate(offs):
        seg, l = _seg_ids(o)
        dest.append(start_in_out[seg, k] + (torch.arange(seg.numel(), device=seg.device) - o[:-1][seg]))
    dest = torch.cat(dest)
    out = torch.empty_like(rows)
    out[dest] = rows
    return _with_lod(out, [out_off.tolist()])


def sequence_slice(input, offset, length, name=None):
    x, off = _raw(input), _offsets(input)
    o, l = _raw(offset).reshape(-1).long().to(x.device), _raw(length).reshape(-1).long().to(x.device)
    new_off = torch.cat([l.new_zeros(1), l.cumsum(0)])
    seg, _ = _seg_ids(new_off)
    src = off[:-1][seg] + o[seg] + (torch.arange(seg.numel(), device=x.device) - new_off[:-1][seg])
    return _with_lod(x[src], [new_off.tolist()])


def sequence_expand(x, y, ref_level=-1, name=None):
    """Repeat the i-th sequence (or row, when x has no LoD) of x as many times as y's i-th ref-level sequence is long."""
    xr = _raw(x)
    yoff = _offsets(y, ref_level)
    reps = yoff[1:] - yoff[:-1]
    if "_lod" in getattr(x, "__dict__", {}) and x.__dict__["_lod"]:
        xoff = _offsets(x)
        xl = xoff[1:] - xoff[:-1]
        seq_of_out = torch.repeat_interleave(torch.arange(reps.numel(), device=xr.device), reps)      # which x-sequence each copy is
        out_lens = xl[seq_of_out]
        out_off = torch.cat([out_lens.new_zeros(1), out_lens.cumsum(0)])
        seg, _ = _seg_ids(out_off)
        src = xoff[:-1][seq_of_out[seg]] + (torch.arange(seg.numel(), device=xr.device) - out_off[:-1][seg])
        return _with_lod(xr[src], [out_off.tolist()])
    return _with_lod(torch.repeat_interleave(xr, reps, 0), [yoff.tolist()])


def sequence_expand_as(x, y, name=None):
    yoff = _offsets(y)
    return _with_lod(torch.repeat_interleave(_raw(x), yoff[1:] - yoff[:-1], 0), [yoff.tolist()])


def sequence_pad(x, pad_value, maxlen=None, name=None):
    xr, off = _raw(x), _offsets(x)
    seg, lens = _seg_ids(off)
    n = lens.numel()
    L = int(maxlen) if maxlen is not None else int(lens.max()) if n else 0
    pv = _raw(pad_value).to(xr.dtype).to(xr.device)
    pv = pv.reshape(()) if pv.numel() == 1 else pv.reshape(xr.shape[1:])      # a scalar, or one value per feature
    out = pv.expand(n, L, *xr.shape[1:]).clone()
    pos = torch.arange(seg.numel(), device=xr.device) - off[:-1][seg]
    keep = pos < L
    out[seg[keep], pos[keep]] = xr[keep]
    return out.as_subclass(Tensor), lens.as_subclass(Tensor)


def sequence_unpad(x, length, name=None):
    xr, lens = _raw(x), _raw(length).reshape(-1).long()
    mask = torch.arange(xr.shape[1], device=xr.device)[None, :] < lens[:, None].to(xr.device)
    off = torch.cat([lens.new_zeros(1), lens.cumsum(0)])
    return _with_lod(xr[mask], [off.tolist()])


def sequence_reshape(input, new_dim):
    x, off = _raw(input), _offsets(input)
    d = x.shape[1]
    if ((off * d) % new_dim != 0).any():
        raise ValueError("sequence_reshape: every sequence's element count must be divisible by new_dim")
    return _with_lod(x.reshape(-1, new_dim), [(off * d // new_dim).tolist()])


def sequence_scatter(input, index, updates, name=None):
    """out[i, index_i[k]] += updates_i[k] for the k-th element of the i-th sequence."""
    x = _raw(input).clone()
    off = _offsets(index)
    seg, _ = _seg_ids(off)
    x.index_put_((seg, _raw(index).reshape(-1).long()), _raw(updates).reshape(-1).to(x.dtype), accumulate=True)
    return x.as_subclass(Tensor)


def sequence_enumerate(input, win_size, pad_value=0, name=None):
    x, off = _raw(input).reshape(-1), _offsets(input)
    seg, _ = _seg_ids(off)
    pos = torch.arange(x.numel(), device=x.device)
    idx = pos[:, None] + torch.arange(win_size, device=x.device)[None, :]
    valid = idx < off[1:][seg][:, None]
    out = torch.where(valid, x[idx.clamp(max=max(x.numel() - 1, 0))], torch.full_like(idx, pad_value).to(x.dtype))
    return _with_lod(out, _lod_of(input))


def sequence_reverse(x, name=None):
    xr, off = _raw(x), _offsets(x)
    seg, _ = _seg_ids(off)
    pos = torch.arange(xr.shape[0], device=xr.device)
    src = off[:-1][seg] + off[1:][seg] - 1 - pos
    return _with_lod(xr[src], _lod_of(x))


def sequence_conv(input, num_filters, filter_size=3, filter_stride=1, padding=True, padding_start=None, bias_attr=None, param_attr=None, act=None, name=None):
    """Context-window projection inside each sequence: row t sees rows [t+start, t+start+filter_size) of its own sequence (zeros
    outside), flattened and multiplied by a [filter_size * D, num_filters] weight."""
    x, off = _raw(input), _offsets(input)
    D = x.shape[1]
    start = -((filter_size - 1) // 2) if padding_start is None else int(padding_start)
    seg, _ = _seg_ids(off)
    pos = torch.arange(x.shape[0], device=x.device)
    idx = pos[:, None] + start + torch.arange(filter_size, device=x.device)[None, :]
    valid = (idx >= off[:-1][seg][:, None]) & (idx < off[1:][seg][:, None])
    ctx = x[idx.clamp(0, max(x.shape[0] - 1, 0))] * valid[..., None].to(x.dtype)        # [T, k, D]
    w = _make_parameter([filter_size * D, num_filters], str(x.dtype).replace("torch.", ""), param_attr, default_initializer=I.XavierUniform())
    out = ctx.reshape(x.shape[0], -1).as_subclass(Tensor) @ w
    if bias_attr is not False:
        out = out + _make_parameter([num_filters], str(x.dtype).replace("torch.", ""), bias_attr, is_bias=True)
    if act:
        out = getattr(F, act)(out)
    return _with_lod(out, _lod_of(input))
