"""SPMD propagation rules beyond the core set of api.py (elementwise / matmul / reduce / embedding / loss).

Parity: paddle/phi/infermeta/spmd_rules/*.cc - reshape, concat, split, stack, slice, squeeze, unsqueeze, flatten, cumsum, argmax,
topk, tile, expand_as, triu, one_hot, gather, scatter, pad, flip, roll, flash_attention, conv2d, pool2d, dropout, where, numel, ...

Most of those rules share one shape: the op TOUCHES some dimensions of its input and leaves the others alone; when none of the touched
dimensions is sharded the op runs on the local shard unchanged and the output keeps the input's placements (with the dimension indices
remapped where the op inserts / removes dimensions).  `dimwise` captures that; every rule is a small description of "which dims are
touched" and "how output dims map to input dims".  When a touched dimension IS sharded the rule un-shards just that dimension
(reshard to Replicate along the offending mesh axes) instead of falling back to fully replicated execution.
"""
from __future__ import annotations

import torch

RULES = {}


def rule(*names):
    def deco(fn):
        for n in names:
            RULES[n] = fn
        return fn
    return deco


def _api():
    from . import api

    return api


def _norm(d, nd):
    return d % nd if nd else 0


def _unshard_dims(x, mesh, dims):
    """x with every mesh axis that shards one of `dims` turned into Replicate (Partial resolved first)."""
    A = _api()
    x = A._resolve_partial(x)
    nd = x.dim()
    dims = {_norm(d, nd) for d in dims}
    pl = list(x.placements)
    new = [A.Replicate() if isinstance(p, A.Shard) and _norm(p.dim, nd) in dims else p for p in pl]
    return A.reshard(x, mesh, new) if new != pl else x


def dimwise(func, args, kwargs, mesh, touched, remap=None, out_index=None):
    """Run `func` on the local shard of args[0] after un-sharding the `touched` dims. remap(in_dim) -> out_dim (or None if the dim
    disappears) gives the output placements; every other DistTensor argument is aligned to the first one (or replicated)."""
    A = _api()
    x = _unshard_dims(args[0], mesh, touched)
    nd = x.dim()
    rep = [A.Replicate()] * mesh.ndim

    def loc(a):
        if not A._is_dt(a):
            return a
        if a.dim() == nd and list(a.shape) == list(x.shape):
            return A._raw(A.reshard(A._resolve_partial(a), mesh, list(x.placements)))
        return A._raw(A.reshard(A._resolve_partial(a), mesh, rep))

    ret = func(A._raw(x), *A._map(args[1:], loc), **A._map(kwargs, loc))
    out_pl = []
    for p in x.placements:
        if isinstance(p, A.Shard):
            d = _norm(p.dim, nd)
            nd_out = remap(d) if remap is not None else d
            out_pl.append(A.Shard(nd_out) if nd_out is not None else A.Replicate())
        else:
            out_pl.append(p)
    return A._map(ret, lambda r: A._mk(r, mesh, out_pl) if isinstance(r, torch.Tensor) else r)


def _dim_arg(args, kwargs, pos, names=("dim", "axis"), default=-1):
    for n in names:
        if n in kwargs:
            return kwargs[n]
    return args[pos] if len(args) > pos else default


# ---- ops that work along ONE dim -------------------------------------------------------------------------------------------
@rule("cumsum", "cumprod", "cummax", "cummin", "logcumsumexp", "sort", "argsort", "topk", "kthvalue", "flip_one", "softmax_", "log_softmax_")
def _along_dim(func, name, args, kwargs, mesh):
    x = args[0]
    pos = 2 if name in ("topk", "kthvalue") else 1
    d = _dim_arg(args, kwargs, pos, default=-1)
    return dimwise(func, args, kwargs, mesh, [d])


@rule("argmax", "argmin", "amax", "amin", "max", "min", "prod", "any", "all", "logsumexp", "var", "std", "norm", "median", "nanmean", "nansum")
def _reduce_like(func, name, args, kwargs, mesh):
    A = _api()
    x = args[0]
    nd = x.dim()
    d = _dim_arg(args, kwargs, 1, default=None)
    if name in ("max", "min") and len(args) > 1 and isinstance(args[1], torch.Tensor):     # elementwise max(x, y)
        return A._rule_elementwise(func, {"max": "maximum", "min": "minimum"}[name], args, kwargs, [a for a in A._flatten(args, kwargs) if A._is_dt(a)], mesh)
    keep = bool(kwargs.get("keepdim", kwargs.get("keepdims", False)))
    if d is None:                                                    # full reduction: needs every dim local
        return dimwise(func, args, kwargs, mesh, list(range(nd)), remap=lambda i: None)
    dims = [d] if isinstance(d, int) else list(d)
    nds = sorted(_norm(v, nd) for v in dims)

    def remap(i):
        if i in nds:
            return None
        return i if keep else i - sum(1 for v in nds if v < i)

    return dimwise(func, args, kwargs, mesh, dims, remap=remap)


# ---- shape ops ---------------------------------------------------------------------------------------------------------------
@rule("unsqueeze")
def _unsqueeze(func, name, args, kwargs, mesh):
    nd = args[0].dim()
    d = _dim_arg(args, kwargs, 1)
    d = d % (nd + 1)
    return dimwise(func, args, kwargs, mesh, [], remap=lambda i: i + 1 if i >= d else i)


@rule("squeeze")
def _squeeze(func, name, args, kwargs, mesh):
    x = args[0]
    nd = x.dim()
    d = _dim_arg(args, kwargs, 1, default=None)
    dims = [i for i in range(nd) if x.shape[i] == 1] if d is None else [_norm(v, nd) for v in ([d] if isinstance(d, int) else d) if x.shape[_norm(v, nd)] == 1]
    return dimwise(func, args, kwargs, mesh, dims, remap=lambda i: i - sum(1 for v in dims if v < i))


@rule("flatten")
def _flatten_rule(func, name, args, kwargs, mesh):
    nd = args[0].dim()
    s = _norm(kwargs.get("start_dim", kwargs.get("start_axis", args[1] if len(args) > 1 else 0)), nd)
    e = _norm(kwargs.get("end_dim", kwargs.get("stop_axis", args[2] if len(args) > 2 else -1)), nd)
    # the first flattened dim may stay sharded (its shards are contiguous blocks of the merged dim); the others must be local
    return dimwise(func, args, kwargs, mesh, list(range(s + 1, e + 1)), remap=lambda i: i if i <= s else i - (e - s))


@rule("reshape", "view")
def _reshape(func, name, args, kwargs, mesh):
    A = _api()
    x = A._resolve_partial(args[0])
    shape = kwargs.get("shape", args[1] if len(args) == 2 and isinstance(args[1], (list, tuple, torch.Size)) else list(args[1:]))
    shape = [int(v) for v in shape]
    gshape = list(x.shape)
    nd = len(gshape)
    # leading dims that are copied unchanged keep their sharding; everything behind the first changed dim must be local
    keep = 0
    while keep < min(nd, len(shape)) and shape[keep] in (gshape[keep], 0) and gshape[keep] != 1:
        keep += 1
    sharded = {_norm(p.dim, nd) for p in x.placements if isinstance(p, A.Shard)}
    x = _unshard_dims(x, mesh, [d for d in sharded if d >= keep])
    local = list(A._RAW.size(x))
    new_local = [local[i] if i < keep else shape[i] for i in range(len(shape))]
    ret = A._raw(x).reshape(new_local)
    return A._mk(ret, mesh, [p if not (isinstance(p, A.Shard) and _norm(p.dim, nd) >= keep) else A.Replicate() for p in x.placements])


@rule("expand", "expand_as", "broadcast_to", "tile", "repeat")
def _expand(func, name, args, kwargs, mesh):
    A = _api()
    x = args[0]
    nd = x.dim()
    if name == "expand_as":
        tgt = list(args[1].shape)
        return dimwise(lambda a, *_r, **_k: a.expand(_local_shape_like(a, x, tgt)), (x,), {}, mesh, [i for i in range(nd) if x.shape[i] == 1],
                       remap=lambda i: i + (len(tgt) - nd))
    sizes = kwargs.get("shape", kwargs.get("size", args[1] if len(args) == 2 and isinstance(args[1], (list, tuple, torch.Size)) else list(args[1:])))
    sizes = [int(v) for v in sizes]
    extra = len(sizes) - nd
    if name in ("tile", "repeat"):      # a repeated dim must be local; dims repeated once keep their sharding
        touched = [i for i in range(nd) if sizes[i + extra] != 1] if extra >= 0 else list(range(nd))
        return dimwise(func, args, kwargs, mesh, touched, remap=lambda i: i + max(extra, 0))
    touched = [i for i in range(nd) if x.shape[i] == 1 and sizes[i + extra] not in (1, -1)]
    xs = _unshard_dims(x, mesh, touched)
    local = list(A._RAW.size(xs))
    tgt = [sizes[j] if j < extra else (local[j - extra] if sizes[j] in (-1, x.shape[j - extra]) else sizes[j]) for j in range(len(sizes))]
    ret = A._raw(xs).expand(tgt)
    return A._mk(ret, mesh, [A.Shard(_norm(p.dim, nd) + extra) if isinstance(p, A.Shard) else p for p in xs.placements])


def _local_shape_like(a, x, tgt):
    nd = x.dim()
    extra = len(tgt) - nd
    return [tgt[j] if j < extra or x.shape[j - extra] == 1 else a.shape[j - extra] for j in range(len(tgt))]


@rule("cat", "concat", "concatenate", "stack", "hstack", "vstack")
def _concat(func, name, args, kwargs, mesh):
    A = _api()
    ts = list(args[0])
    d = _dim_arg(args, kwargs, 1, default=0)
    first = next(t for t in ts if A._is_dt(t))
    nd = first.dim()
    lead = _unshard_dims(first, mesh, [d] if name in ("cat", "concat", "concatenate") else [])
    pl = list(lead.placements)
    locs = [A._raw(A.reshard(A._resolve_partial(t if A._is_dt(t) else A._to_dt(t, mesh)), mesh, pl)) for t in ts]
    ret = func(locs, *args[1:], **kwargs)
    if name == "stack":
        dd = d % (nd + 1)
        pl = [A.Shard(_norm(p.dim, nd) + 1) if isinstance(p, A.Shard) and _norm(p.dim, nd) >= dd else p for p in pl]
    return A._mk(ret, mesh, pl)


@rule("split", "chunk", "unbind", "tensor_split", "split_with_sizes")
def _split(func, name, args, kwargs, mesh):
    nd = args[0].dim()
    d = _dim_arg(args, kwargs, 2 if name in ("split", "chunk", "tensor_split", "split_with_sizes") else 1, default=0)
    if name == "unbind":
        dd = _norm(d, nd)
        return dimwise(func, args, kwargs, mesh, [d], remap=lambda i: i - 1 if i > dd else i)
    return dimwise(func, args, kwargs, mesh, [d])


@rule("narrow", "index_select", "gather", "take_along_dim", "scatter", "scatter_add", "index_add", "index_copy", "index_fill", "roll", "flip",
      "diff", "repeat_interleave", "unfold")
def _indexed(func, name, args, kwargs, mesh):
    x = args[0]
    nd = x.dim()
    if name == "flip" or name == "roll":
        dims = kwargs.get("dims", args[-1] if len(args) > 1 else list(range(nd)))
        dims = [dims] if isinstance(dims, int) else list(dims)
        return dimwise(func, args, kwargs, mesh, dims)
    d = _dim_arg(args, kwargs, 1, default=0)
    if name in ("gather", "take_along_dim", "scatter", "scatter_add"):
        # the index tensor has the input's rank: every dim but `d` may stay sharded if the index is sharded the same way
        return dimwise(func, args, kwargs, mesh, [d])
    return dimwise(func, args, kwargs, mesh, [d])


@rule("triu", "tril")
def _tri(func, name, args, kwargs, mesh):
    nd = args[0].dim()
    return dimwise(func, args, kwargs, mesh, [nd - 1, nd - 2])


@rule("one_hot")
def _one_hot(func, name, args, kwargs, mesh):
    return dimwise(func, args, kwargs, mesh, [])          # appends a trailing class dim; existing dims keep their placement


@rule("pad", "constant_pad_nd")
def _pad(func, name, args, kwargs, mesh):
    nd = args[0].dim()
    pads = kwargs.get("pad", args[1] if len(args) > 1 else [])
    touched = [nd - 1 - i // 2 for i in range(0, len(pads), 2) if pads[i] or pads[i + 1]] if isinstance(pads, (list, tuple)) else list(range(nd))
    return dimwise(func, args, kwargs, mesh, touched)


@rule("scaled_dot_product_attention")
def _sdpa(func, name, args, kwargs, mesh):
    """[B, H, S, D] attention: batch and head dims are embarrassingly parallel; sequence / feature dims must be local."""
    A = _api()
    q = A._resolve_partial(args[0])
    nd = q.dim()
    q = _unshard_dims(q, mesh, [nd - 1, nd - 2])
    pl = list(q.placements)
    kv = [A._raw(A.reshard(A._resolve_partial(t), mesh, pl)) if A._is_dt(t) else t for t in args[1:3]]
    rep = [A.Replicate()] * mesh.ndim
    rest = A._map(args[3:], lambda a: A._raw(A.reshard(a, mesh, rep)) if A._is_dt(a) else a)
    kw = A._map(kwargs, lambda a: A._raw(A.reshard(a, mesh, rep)) if A._is_dt(a) else a)
    return A._mk(func(A._raw(q), *kv, *rest, **kw), mesh, pl)


@rule("conv1d", "conv2d", "conv3d", "conv_transpose2d", "max_pool2d", "avg_pool2d", "adaptive_avg_pool2d", "adaptive_max_pool2d", "max_pool1d",
      "avg_pool1d", "max_pool3d", "avg_pool3d", "interpolate", "upsample", "batch_norm_eval", "pixel_shuffle", "unfold_im2col")
def _batch_parallel(func, name, args, kwargs, mesh):
    """Data-parallel only: the batch dim may be sharded, channel / spatial dims are gathered; weights are used replicated."""
    nd = args[0].dim()
    return dimwise(func, args, kwargs, mesh, list(range(1, nd)))


@rule("numel", "size", "shape")
def _meta(func, name, args, kwargs, mesh):
    A = _api()
    return func(A._raw(args[0]), *args[1:], **kwargs)


@rule("swiglu", "glu")
def _glu(func, name, args, kwargs, mesh):
    nd = args[0].dim()
    return dimwise(func, args, kwargs, mesh, [_dim_arg(args, kwargs, 1, default=nd - 1)])


def lookup(name):
    return RULES.get(name)
