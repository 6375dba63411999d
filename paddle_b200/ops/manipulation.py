"""Shape / indexing ops. Parity: python/paddle/tensor/manipulation.py."""
from __future__ import annotations

import builtins

import numpy as np
import torch

from ._helpers import T, ax, dt, raw, shp, to_int, to_tensor, wrap


from ..framework.recording import recordable as _recordable  # noqa: E402

@_recordable
def reshape(x, *shape, name=None):
    x = T(x)
    s = shp(shape[0] if len(shape) == 1 else list(shape))
    # paddle: 0 means "copy this dim from input"
    s = [x.size(i) if (v == 0 and i < x.dim()) else v for i, v in enumerate(s)]
    return torch.reshape(x, s)


def reshape_(x, shape, name=None):
    out = reshape(x, shape)
    x.data = raw(out)
    return x


@_recordable
def view(x, shape_or_dtype, name=None):
    x = T(x)
    if isinstance(shape_or_dtype, (list, tuple)):
        return torch.Tensor.view(x, *shp(shape_or_dtype))
    return torch.Tensor.view(x, dt(shape_or_dtype))


def view_as(x, other, name=None):
    return torch.Tensor.view(T(x), *other.size())


def transpose(x, *perm, name=None):
    x = T(x)
    if len(perm) == 2 and all(isinstance(p, int) for p in perm) and x.dim() != 2:
        return torch.transpose(x, perm[0], perm[1])  # torch-style (dim0, dim1)
    p = shp(perm[0] if len(perm) == 1 else list(perm))
    return torch.permute(x, tuple(p))


def t(input, name=None):
    x = T(input)
    if x.dim() > 2:
        raise ValueError(f"paddle.t only supports tensors of rank <= 2, got {x.dim()}; use paddle.transpose")
    return x if x.dim() < 2 else torch.transpose(x, 0, 1)


def moveaxis(x, source, destination, name=None):
    return torch.movedim(T(x), ax(source), ax(destination))


def swapaxes(x, axis0, axis1, name=None):
    return torch.transpose(T(x), axis0, axis1)


swapdims = swapaxes


def concat(x, axis=0, name=None):
    return torch.cat([T(i) for i in x], int(to_int(axis)))


def stack(x, axis=0, name=None):
    return torch.stack([T(i) for i in x], int(to_int(axis)))


def hstack(x, name=None):
    return torch.hstack([T(i) for i in x])


def vstack(x, name=None):
    return torch.vstack([T(i) for i in x])


def dstack(x, name=None):
    return torch.dstack([T(i) for i in x])


def column_stack(x, name=None):
    return torch.column_stack([T(i) for i in x])


def row_stack(x, name=None):
    return torch.vstack([T(i) for i in x])


@_recordable
def split(x, num_or_sections, axis=0, name=None):
    x = T(x)
    axis = int(to_int(axis))
    if isinstance(num_or_sections, int):
        n = x.size(axis)
        if n % num_or_sections != 0:
            raise ValueError(f"split: dim {axis} of size {n} is not divisible by {num_or_sections}")
        return list(torch.Tensor.split(x, n // num_or_sections, axis))
    secs = [to_int(s) for s in num_or_sections]
    if -1 in secs:
        known = builtins.sum(s for s in secs if s != -1)
        secs[secs.index(-1)] = x.size(axis) - known
    return list(torch.Tensor.split(x, secs, axis))


def tensor_split(x, num_or_indices, axis=0, name=None):
    return list(torch.tensor_split(T(x), num_or_indices if isinstance(num_or_indices, int) else list(num_or_indices), dim=axis))


def hsplit(x, num_or_indices, name=None):
    return list(torch.hsplit(T(x), num_or_indices))


def vsplit(x, num_or_indices, name=None):
    return list(torch.vsplit(T(x), num_or_indices))


def dsplit(x, num_or_indices, name=None):
    return list(torch.dsplit(T(x), num_or_indices))


def chunk(x, chunks, axis=0, name=None):
    return list(torch.chunk(T(x), chunks, int(to_int(axis))))


def unbind(input, axis=0):
    return list(torch.unbind(T(input), axis))


def unstack(x, axis=0, num=None):
    return list(torch.unbind(T(x), axis))


@_recordable
def squeeze(x, axis=None, name=None):
    x = T(x)
    a = ax(axis)
    if a is None:
        return torch.squeeze(x)
    if isinstance(a, int):
        a = (a,)
    a = tuple(d for d in a if x.size(d) == 1)
    return torch.squeeze(x, a) if a else x


@_recordable
def unsqueeze(x, axis, name=None):
    x = T(x)
    a = ax(axis)
    if isinstance(a, int):
        return torch.unsqueeze(x, a)
    for d in a:
        x = torch.unsqueeze(x, d)
    return x


def squeeze_(x, axis=None, name=None):
    out = squeeze(x, axis)
    x.data = raw(out)
    return x


def unsqueeze_(x, axis, name=None):
    out = unsqueeze(x, axis)
    x.data = raw(out)
    return x


@_recordable
def flatten(x, start_axis=0, stop_axis=-1, name=None):
    x = T(x)
    if x.dim() == 0:
        return x.reshape(1)
    return torch.flatten(x, start_axis, stop_axis)


def flatten_(x, start_axis=0, stop_axis=-1, name=None):
    out = flatten(x, start_axis, stop_axis)
    x.data = raw(out)
    return x


@_recordable
def unflatten(x, axis, shape, name=None):
    return torch.unflatten(T(x), axis, shp(shape))


@_recordable
def expand(x, *shape, name=None):
    return torch.Tensor.expand(T(x), *shp(shape[0] if len(shape) == 1 else list(shape)))


@_recordable
def expand_as(x, y, name=None):
    return torch.Tensor.expand(T(x), *y.size())


@_recordable
def broadcast_to(x, shape, name=None):
    return torch.broadcast_to(T(x), shp(shape))


def broadcast_tensors(input, name=None):
    return list(torch.broadcast_tensors(*[T(i) for i in input]))


@_recordable
def tile(x, *repeat_times, name=None):
    return torch.tile(T(x), tuple(shp(repeat_times[0] if len(repeat_times) == 1 else list(repeat_times))))


@_recordable
def repeat_interleave(x, repeats, axis=None, name=None):
    r = repeats if not isinstance(repeats, torch.Tensor) else T(repeats)
    return torch.repeat_interleave(T(x), r, dim=axis)


def flip(x, axis, name=None):
    a = ax(axis)
    return torch.flip(T(x), (a,) if isinstance(a, int) else a)


def reverse(x, axis, name=None):
    return flip(x, axis)


def rot90(x, k=1, axes=(0, 1), name=None):
    return torch.rot90(T(x), k, list(axes))


@_recordable
def roll(x, shifts, axis=None, name=None):
    return torch.roll(T(x), ax(shifts) if not isinstance(shifts, int) else shifts, ax(axis))


def cast(x, dtype):
    return T(x).to(dt(dtype))


def cast_(x, dtype):
    x.data = raw(x).to(dt(dtype))
    return x


@_recordable
def slice(input, axes, starts, ends):  # noqa: A001
    x = T(input)
    idx = [builtins.slice(None)] * x.dim()
    for a, s, e in zip(axes, starts, ends):
        idx[a] = builtins.slice(to_int(s), to_int(e))
    return x[tuple(idx)]


@_recordable
def strided_slice(x, axes, starts, ends, strides, name=None):
    x = T(x)
    idx = [builtins.slice(None)] * x.dim()
    for a, s, e, st in zip(axes, starts, ends, strides):
        s, e, st = to_int(s), to_int(e), to_int(st)
        if st < 0:
            n = x.size(a)
            s = s + n if s < 0 else builtins.min(s, n - 1)
            e = e + n if e < 0 else e
            ii = torch.arange(s, e, st, device=x.device)
            x = torch.index_select(x, a, ii)
        else:
            idx[a] = builtins.slice(s, e, st)
    return x[tuple(idx)]


@_recordable
def crop(x, shape=None, offsets=None, name=None):
    x = T(x)
    s = shp(shape) if shape is not None else list(x.size())
    o = shp(offsets) if offsets is not None else [0] * x.dim()
    idx = tuple(builtins.slice(oi, oi + (si if si != -1 else x.size(i) - oi)) for i, (oi, si) in enumerate(zip(o, s)))
    return x[idx]


def gather(x, index, axis=None, name=None):
    x, index = T(x), T(index)
    axis = 0 if axis is None else int(to_int(axis))
    if index.dim() == 0:
        index = index.reshape(1)
    return torch.index_select(x, axis, index.reshape(-1).long())


def gather_nd(x, index, name=None):
    x, index = T(x), T(index).long()
    k = index.size(-1)
    return x[tuple(index[..., i] for i in range(k))]


def scatter(x, index, updates, overwrite=True, name=None):
    x, index, updates = T(x), T(index).long().reshape(-1), T(updates)
    out = x.clone()
    if overwrite:
        out[index] = updates.to(out.dtype)
    else:
        out[index] = 0
        out = torch.index_add(out, 0, index, updates.to(out.dtype))
    return out


def scatter_(x, index, updates, overwrite=True, name=None):
    out = scatter(x, index, updates, overwrite)
    with torch.no_grad():
        torch.Tensor.copy_(x, out)
    return x


def scatter_nd_add(x, index, updates, name=None):
    x, index, updates = T(x), T(index).long(), T(updates)
    k = index.size(-1)
    return x.index_put(tuple(index[..., i] for i in range(k)), updates.to(x.dtype), accumulate=True)


def scatter_nd(index, updates, shape, name=None):
    updates = T(updates)
    return scatter_nd_add(torch.zeros(shp(shape), dtype=updates.dtype, device=updates.device), index, updates)


def index_select(x, index, axis=0, name=None):
    return torch.index_select(T(x), axis, T(index).long())


def index_sample(x, index):
    return torch.gather(T(x), 1, T(index).long())


def index_add(x, index, axis, value, name=None):
    return torch.index_add(T(x), axis, T(index).long(), T(value))


def index_add_(x, index, axis, value, name=None):
    return torch.Tensor.index_add_(x, axis, T(index).long(), T(value))


def index_put(x, indices, value, accumulate=False, name=None):
    return torch.index_put(T(x), tuple(T(i) for i in indices), T(value).to(x.dtype), accumulate)


def index_put_(x, indices, value, accumulate=False, name=None):
    return torch.Tensor.index_put_(x, tuple(T(i) for i in indices), T(value).to(x.dtype), accumulate)


def index_fill(x, index, axis, value, name=None):
    return torch.index_fill(T(x), axis, T(index).long(), to_int(value))


def index_fill_(x, index, axis, value, name=None):
    return torch.Tensor.index_fill_(x, axis, T(index).long(), to_int(value))


def masked_select(x, mask, name=None):
    return torch.masked_select(T(x), T(mask))


def masked_fill(x, mask, value, name=None):
    v = value if not isinstance(value, torch.Tensor) else (value.item() if value.numel() == 1 else value)
    return torch.masked_fill(T(x), T(mask), v) if not isinstance(v, torch.Tensor) else torch.where(T(mask), v, T(x))


def masked_fill_(x, mask, value, name=None):
    out = masked_fill(x, mask, value)
    with torch.no_grad():
        torch.Tensor.copy_(x, out)
    return x


def masked_scatter(x, mask, value, name=None):
    return torch.masked_scatter(T(x), T(mask), T(value))


def take_along_axis(arr, indices, axis, broadcast=True):
    arr, indices = T(arr), T(indices).long()
    if broadcast:
        shape = list(arr.size())
        shape[axis] = indices.size(axis)
        indices = torch.broadcast_to(indices, torch.broadcast_shapes(tuple(shape), tuple(indices.size()))) if indices.dim() == arr.dim() else indices
    return torch.take_along_dim(arr, indices, axis)


def put_along_axis(arr, indices, values, axis, reduce="assign", include_self=True, broadcast=True):
    arr, indices = T(arr), T(indices).long()
    values = T(values).to(arr.dtype) if isinstance(values, torch.Tensor) else torch.full_like(indices, values, dtype=arr.dtype)
    if values.size() != indices.size():
        values = torch.broadcast_to(values, indices.size())
    if reduce == "assign":
        return torch.scatter(arr, axis, indices, values)
    red = {"add": "sum", "mul": "prod", "multiply": "prod", "mean": "mean", "amin": "amin", "amax": "amax"}[reduce]
    return torch.scatter_reduce(arr, axis, indices, values, red, include_self=include_self)


def put_along_axis_(arr, indices, values, axis, reduce="assign", include_self=True, broadcast=True):
    out = put_along_axis(arr, indices, values, axis, reduce, include_self)
    with torch.no_grad():
        torch.Tensor.copy_(arr, out)
    return arr


def shard_index(input, index_num, nshards, shard_id, ignore_value=-1):
    x = T(input)
    size = (index_num + nshards - 1) // nshards
    lo = shard_id * size
    inside = (x >= lo) & (x < lo + size)
    return torch.where(inside, x - lo, torch.full_like(x, ignore_value))


def unique(x, return_index=False, return_inverse=False, return_counts=False, axis=None, dtype="int64", name=None):
    x = T(x)
    out, inv, cnt = torch.unique(raw(x), sorted=True, return_inverse=True, return_counts=True, dim=axis)
    res = [wrap(out)]
    if return_index:
        flat_inv = inv.reshape(-1)
        perm = torch.arange(flat_inv.numel() - 1, -1, -1, device=x.device)
        first = torch.empty(out.size(0) if axis is not None else out.numel(), dtype=torch.int64, device=x.device)
        first.scatter_(0, flat_inv.flip(0), perm)
        res.append(wrap(first.to(dt(dtype))))
    if return_inverse:
        res.append(wrap(inv.to(dt(dtype))))
    if return_counts:
        res.append(wrap(cnt.to(dt(dtype))))
    return res[0] if len(res) == 1 else tuple(res)


def unique_consecutive(x, return_inverse=False, return_counts=False, axis=None, dtype="int64", name=None):
    r = torch.unique_consecutive(raw(x), return_inverse=return_inverse, return_counts=return_counts, dim=axis)
    if not isinstance(r, tuple):
        return wrap(r)
    return tuple(wrap(v if i == 0 else v.to(dt(dtype))) for i, v in enumerate(r))


def as_complex(x, name=None):
    return torch.view_as_complex(T(x).contiguous())


def as_real(x, name=None):
    return torch.view_as_real(T(x))


def as_strided(x, shape, stride, offset=0, name=None):
    return torch.as_strided(T(x), shp(shape), shp(stride), offset)


def unfold(x, axis, size, step, name=None):
    return torch.Tensor.unfold(T(x), axis, size, step)


def atleast_1d(*inputs, name=None):
    r = [torch.atleast_1d(T(i)) for i in inputs]
    return r[0] if len(r) == 1 else r


def atleast_2d(*inputs, name=None):
    r = [torch.atleast_2d(T(i)) for i in inputs]
    return r[0] if len(r) == 1 else r


def atleast_3d(*inputs, name=None):
    r = [torch.atleast_3d(T(i)) for i in inputs]
    return r[0] if len(r) == 1 else r


def select_scatter(x, values, axis, index, name=None):
    return torch.select_scatter(T(x), T(values), axis, index)


def slice_scatter(x, value, axes, starts, ends, strides, name=None):
    out = T(x).clone()
    idx = [builtins.slice(None)] * out.dim()
    for a, s, e, st in zip(axes, starts, ends, strides):
        idx[a] = builtins.slice(to_int(s), to_int(e), to_int(st))
    out[tuple(idx)] = T(value)
    return out


def diagonal_scatter(x, y, offset=0, axis1=0, axis2=1, name=None):
    return torch.diagonal_scatter(T(x), T(y), offset, axis1, axis2)


def tolist(x):
    return torch.Tensor.tolist(T(x))


# shape / numel read metadata, which a recorded program must do at run time (dynamic batch): recorded as whole nodes
@_recordable
def numel(x, name=None):
    return wrap(torch.tensor(torch.Tensor.numel(T(x)), dtype=torch.int64))


@_recordable
def shape(input):  # noqa: A001
    return wrap(torch.tensor(list(torch.Tensor.size(T(input))), dtype=torch.int32))


def rank(input):
    return wrap(torch.tensor(T(input).dim(), dtype=torch.int32))


@_recordable
def is_empty(x, name=None):
    return wrap(torch.tensor(torch.Tensor.numel(T(x)) == 0))


def tensordot(x, y, axes=2, name=None):
    if isinstance(axes, torch.Tensor):
        axes = axes.tolist()
    if isinstance(axes, (list, tuple)) and len(axes) == 2 and isinstance(axes[0], (list, tuple)):
        axes = (list(axes[0]), list(axes[1]))
    elif isinstance(axes, (list, tuple)):
        axes = (list(axes), list(axes))
    return torch.tensordot(T(x), T(y), dims=axes)


def block_diag(inputs, name=None):
    return torch.block_diag(*[T(i) for i in inputs])


__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in (
    "np", "torch", "builtins", "T", "ax", "dt", "raw", "shp", "to_int", "to_tensor", "wrap", "annotations")]


# static programs record these as single ops (their bodies compute on raw tensors / read values; framework/recording.py)
from ..framework.recording import make_recordable as _make_recordable  # noqa: E402

_make_recordable(globals(), ['unique', 'unique_consecutive'])
