"""DistTensor, reshard and the dygraph semi-auto API. Parity: python/paddle/distributed/auto_parallel/api.py
(shard_tensor:L200+, reshard, shard_layer, shard_optimizer, shard_dataloader, to_static/DistModel, Strategy) and the SPMD
rules of paddle/phi/infermeta/spmd_rules/{elementwise,matmul,reduction,embedding,transpose,softmax}.cc."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ...tensor import Parameter, Tensor
from .. import env
from .placement import Partial, Placement, Replicate, Shard
from .process_mesh import ProcessMesh, get_mesh

_RAW = torch.Tensor


def _raw(t):
    return t.as_subclass(_RAW) if isinstance(t, torch.Tensor) and type(t) is not _RAW else t


# --------------------------------------------------------------------------------------------------------------------
# autograd-aware collectives along one mesh dim
# --------------------------------------------------------------------------------------------------------------------
def _pg(mesh, mdim):
    g = mesh.group_along(mdim)
    return None if g is None else g.pg


def _coord(mesh, mdim):
    c = mesh.coord_of(env.get_rank())
    return 0 if c is None else c[mdim]


class _AllGather(torch.autograd.Function):   # Shard(d) -> Replicate ; bwd: take own slice
    @staticmethod
    def forward(ctx, x, pg, n, idx, d):
        ctx.a = (n, idx, d)
        parts = [torch.empty_like(x) for _ in range(n)]
        dist.all_gather(parts, x.contiguous(), group=pg)
        return torch.cat(parts, d)

    @staticmethod
    def backward(ctx, g):
        n, idx, d = ctx.a
        return g.chunk(n, d)[idx].contiguous(), None, None, None, None


class _Slice(torch.autograd.Function):       # Replicate -> Shard(d) ; bwd: all-gather
    @staticmethod
    def forward(ctx, x, pg, n, idx, d):
        ctx.a = (pg, n, d)
        return x.chunk(n, d)[idx].contiguous()

    @staticmethod
    def backward(ctx, g):
        pg, n, d = ctx.a
        parts = [torch.empty_like(g) for _ in range(n)]
        dist.all_gather(parts, g.contiguous(), group=pg)
        return torch.cat(parts, d), None, None, None, None


class _AllReduce(torch.autograd.Function):   # Partial -> Replicate ; bwd: identity
    @staticmethod
    def forward(ctx, x, pg, op):
        y = x.clone()
        dist.all_reduce(y, op=op, group=pg)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class _IdentityFwdAllReduceBwd(torch.autograd.Function):   # Replicate feeding a sharded compute
    @staticmethod
    def forward(ctx, x, pg):
        ctx.pg = pg
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        dist.all_reduce(g, group=ctx.pg)
        return g, None


class _AllToAll(torch.autograd.Function):    # Shard(a) -> Shard(b) ; bwd: Shard(b) -> Shard(a)
    @staticmethod
    def forward(ctx, x, pg, n, a, b):
        ctx.a = (pg, n, a, b)
        send = [c.contiguous() for c in x.chunk(n, b)]
        if dist.get_backend(pg) == "gloo":   # gloo has no all-to-all: gather everything, keep my column block
            me = dist.get_rank(pg)
            full = [torch.empty_like(x) for _ in range(n)]
            dist.all_gather(full, x.contiguous(), group=pg)
            return torch.cat([f.chunk(n, b)[me] for f in full], a)
        recv = [torch.empty_like(send[0]) for _ in range(n)]
        dist.all_to_all(recv, send, group=pg)
        return torch.cat(recv, a)

    @staticmethod
    def backward(ctx, g):
        pg, n, a, b = ctx.a
        return _AllToAll.apply(g, pg, n, b, a), None, None, None, None


_REDOP = {"sum": dist.ReduceOp.SUM, "avg": dist.ReduceOp.AVG, "mean": dist.ReduceOp.AVG, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN,
          "prod": dist.ReduceOp.PRODUCT}


def _all_reduce(x, pg, how):
    if how in ("avg", "mean") and dist.get_backend(pg) == "gloo":   # gloo has no AVG
        return _AllReduce.apply(x, pg, dist.ReduceOp.SUM) / dist.get_world_size(pg)
    return _AllReduce.apply(x, pg, _REDOP[how])


# --------------------------------------------------------------------------------------------------------------------
# DistTensor
# --------------------------------------------------------------------------------------------------------------------
class DistAttr:
    def __init__(self, mesh, sharding_specs):
        self.process_mesh, self.sharding_specs = mesh, sharding_specs

    def to_placements(self):
        pl = [Replicate() for _ in range(self.process_mesh.ndim)]
        for tdim, spec in enumerate(self.sharding_specs):
            if spec is not None:
                pl[self.process_mesh.dim_index(spec)] = Shard(tdim)
        return pl


def _norm_placements(mesh, placements):
    pl = list(placements) if placements is not None else []
    pl += [Replicate()] * (mesh.ndim - len(pl))
    return pl


def _local_shape(gshape, mesh, placements):
    s = list(gshape)
    for md, p in enumerate(placements):
        if isinstance(p, Shard):
            n = mesh.shape[md]
            d = p.dim % len(s)
            assert s[d] % n == 0, f"dim {d} of global shape {gshape} is not divisible by mesh dim {md} (size {n})"
            s[d] //= n
    return s


def _global_shape(lshape, mesh, placements):
    s = list(lshape)
    for md, p in enumerate(placements):
        if isinstance(p, Shard):
            s[p.dim % len(s)] *= mesh.shape[md]
    return s


def _offsets(gshape, mesh, placements):
    off = [0] * len(gshape)
    c = mesh.coord_of(env.get_rank()) or (0,) * mesh.ndim
    for md, p in enumerate(placements):
        if isinstance(p, Shard):
            d = p.dim % len(gshape)
            off[d] += c[md] * _stride_of(gshape, mesh, placements, md, d)
    return off


def _stride_of(gshape, mesh, placements, md, d):
    """Extent along tensor dim d owned by one coordinate step of mesh dim md (nested shardings split left to right)."""
    ext = gshape[d]
    for k in range(md + 1):
        p = placements[k]
        if isinstance(p, Shard) and p.dim % len(gshape) == d:
            ext //= mesh.shape[k]
    return ext


class DistTensor(Tensor):
    """Local shard + (process_mesh, placements). `shape` is the global shape (as in the reference); `_local_value()` gives
    the shard."""
    _is_dist = True

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return _dispatch(func, args, kwargs or {})

    @property
    def shape(self):
        return _global_shape(list(_RAW.size(self)), self.process_mesh, self.placements)

    @property
    def process_mesh(self):
        return self.__dict__["_dist_mesh"]

    @property
    def placements(self):
        return self.__dict__["_dist_placements"]

    @property
    def dist_attr(self):
        return self

    def is_dist(self):
        return True

    @property
    def _local_shape(self):
        return list(_RAW.size(self))

    def _local_value(self):
        with torch._C.DisableTorchFunctionSubclass():
            return self.as_subclass(Tensor)

    def __repr__(self):
        with torch._C.DisableTorchFunctionSubclass():
            return f"DistTensor(global_shape={self.shape}, placements={self.placements}, mesh={self.process_mesh.shape}, local=\n{_RAW.__repr__(_raw(self))})"


class DistParameter(DistTensor, Parameter):
    pass


def _mk(local, mesh, placements, like=None):
    if local is None or not isinstance(local, torch.Tensor):
        return local
    with torch._C.DisableTorchFunctionSubclass():
        t = _raw(local).as_subclass(DistTensor)
    t.__dict__["_dist_mesh"] = mesh
    t.__dict__["_dist_placements"] = list(placements)
    gs = _global_shape(list(_RAW.size(t)), mesh, placements)
    t.__dict__["_dist_shard"] = (gs, _offsets(gs, mesh, placements))
    return t


def _is_dt(x):
    return isinstance(x, DistTensor)


def shard_tensor(data, mesh=None, placements=None, dtype=None, place=None, stop_gradient=None, dist_attr=None):
    """Global value -> DistTensor (each rank keeps its shard)."""
    if dist_attr is not None:
        mesh, placements = dist_attr.process_mesh, dist_attr.to_placements()
    mesh = mesh or get_mesh()
    placements = _norm_placements(mesh, placements)
    src = data if isinstance(data, torch.Tensor) else torch.as_tensor(np.asarray(data))
    with torch._C.DisableTorchFunctionSubclass():
        local = _raw(src).detach()
        if dtype is not None:
            from ...framework.dtype import to_torch_dtype

            local = local.to(to_torch_dtype(dtype))
        c = mesh.coord_of(env.get_rank()) or (0,) * mesh.ndim
        for md, p in enumerate(placements):
            if isinstance(p, Shard):
                n = mesh.shape[md]
                d = p.dim % local.dim()
                assert local.shape[d] % n == 0, f"cannot shard dim {d} (size {local.shape[d]}) over {n} ranks evenly"
                local = local.chunk(n, d)[c[md]]
            elif isinstance(p, Partial) and c[md] != 0:
                local = torch.zeros_like(local)
        local = local.contiguous().clone() if local.data_ptr() == _raw(src).data_ptr() and any(isinstance(p, Shard) for p in placements) else local
    if isinstance(data, Parameter):
        with torch._C.DisableTorchFunctionSubclass():
            t = torch.Tensor._make_subclass(DistParameter, local, data.requires_grad)
        t.__dict__.update({k: v for k, v in data.__dict__.items()})
        t.__dict__["_dist_mesh"], t.__dict__["_dist_placements"] = mesh, placements
        gs = _global_shape(list(local.shape), mesh, placements)
        t.__dict__["_dist_shard"] = (gs, _offsets(gs, mesh, placements))
        if any(isinstance(p, Shard) for p in placements):
            t.__dict__["is_distributed"] = True
        return t
    t = _mk(local, mesh, placements)
    rg = (not stop_gradient) if stop_gradient is not None else (isinstance(data, torch.Tensor) and data.requires_grad)
    if rg and t.is_floating_point():
        with torch._C.DisableTorchFunctionSubclass():
            t.requires_grad_(True)
    return t


def dtensor_from_fn(fn, mesh, placements, *args, **kwargs):
    return shard_tensor(fn(*args, **kwargs), mesh, placements)


def dtensor_from_local(local, mesh, placements):
    return _mk(_raw(local), mesh, _norm_placements(mesh, placements))


def dtensor_to_local(t, mesh=None, placements=None):
    return t._local_value() if _is_dt(t) else t


def _reshard_local(x, mesh, src, dst):
    """x: raw local tensor with placements `src` -> raw local tensor with placements `dst` (same mesh)."""
    cur = list(src)
    # resolve right-to-left so nested shards of one tensor dim unwind in the right order
    for md in reversed(range(mesh.ndim)):
        s, d = cur[md], dst[md]
        if s == d:
            continue
        n = mesh.shape[md]
        if n == 1:
            cur[md] = d
            continue
        pg, idx = _pg(mesh, md), _coord(mesh, md)
        if pg is None:   # single-process run of an n>1 mesh is not meaningful
            raise RuntimeError("reshard needs an initialised parallel env covering the mesh")
        if isinstance(s, Partial):
            x = _all_reduce(x, pg, s.reduce_type)
            s = Replicate()
        if isinstance(s, Shard) and isinstance(d, Shard):
            x = _AllToAll.apply(x, pg, n, s.dim % x.dim(), d.dim % x.dim())
        elif isinstance(s, Shard):
            x = _AllGather.apply(x, pg, n, idx, s.dim % x.dim())
            s = Replicate()
            if isinstance(d, Partial):
                x = x if idx == 0 else x * 0
        elif isinstance(d, Shard):
            x = _Slice.apply(x, pg, n, idx, d.dim % x.dim())
        elif isinstance(d, Partial):
            x = x if idx == 0 else x * 0
        cur[md] = d
    return x


def reshard(dist_tensor, mesh=None, placements=None):
    if not _is_dt(dist_tensor):
        return shard_tensor(dist_tensor, mesh, placements)
    mesh = mesh or dist_tensor.process_mesh
    placements = _norm_placements(mesh, placements)
    x = _raw(dist_tensor)
    if mesh != dist_tensor.process_mesh:
        x = _cross_mesh(x, dist_tensor, mesh, placements)
        return _mk(x, mesh, placements)
    with torch._C.DisableTorchFunctionSubclass():
        y = _reshard_local(x, mesh, dist_tensor.placements, placements)
    return _mk(y, mesh, placements)


def _cross_mesh(x, t, mesh, placements):
    """Same-shaped meshes on different ranks (pipeline stage hand-off): coordinate-wise send/recv (forward only)."""
    src = t.process_mesh
    assert src.shape == mesh.shape, "cross-mesh reshard needs meshes of the same shape"
    me = env.get_rank()
    with torch._C.DisableTorchFunctionSubclass():
        full = _reshard_local(x, src, t.placements, placements) if src.coord_of(me) is not None else None
        for s_rank, d_rank in zip(src.process_ids, mesh.process_ids):
            if me == s_rank and me != d_rank:
                dist.send(full.contiguous(), d_rank)
            if me == d_rank and me != s_rank:
                buf = torch.empty(_local_shape(t.shape, mesh, placements), dtype=x.dtype, device=x.device)
                dist.recv(buf, s_rank)
                full = buf
    return full


def unshard_dtensor(dist_tensor):
    t = dist_tensor
    if not _is_dt(t):
        return t
    r = reshard(t, t.process_mesh, [Replicate()] * t.process_mesh.ndim)
    with torch._C.DisableTorchFunctionSubclass():
        return _raw(r).as_subclass(Tensor)


# --------------------------------------------------------------------------------------------------------------------
# sharding propagation
# --------------------------------------------------------------------------------------------------------------------
_EW = {"add", "sub", "mul", "div", "true_divide", "neg", "relu", "gelu", "silu", "sigmoid", "tanh", "exp", "log", "sqrt", "rsqrt", "pow",
       "abs", "clone", "detach", "to", "type", "float", "half", "bfloat16", "double", "contiguous", "dropout", "where", "maximum", "minimum",
       "clamp", "square", "sin", "cos", "erf", "floor", "ceil", "round", "sign", "reciprocal", "leaky_relu", "elu", "softplus", "mish", "hardswish",
       "eq", "ne", "lt", "le", "gt", "ge", "logical_and", "logical_or", "logical_not", "isnan", "isinf", "isfinite", "zeros_like", "ones_like",
       "masked_fill", "lerp", "addcmul", "addcdiv", "copy", "zero", "fill", "requires_grad", "bool", "int", "long", "tanhshrink", "relu6",
       "hardtanh", "celu", "selu", "log1p", "expm1", "exp2", "log2", "log10", "fmod", "remainder", "floor_divide", "bitwise_and", "bitwise_or",
       "bitwise_not", "rsub", "_to_copy", "nan_to_num", "clip", "positive", "conj", "real", "alias", "view_as", "detach_", "retain_grad", "cpu", "cuda"}
_LINEAR_UNARY = {"clone", "detach", "to", "type", "float", "half", "bfloat16", "double", "contiguous", "neg", "_to_copy", "alias", "cpu", "cuda",
                 "requires_grad", "retain_grad", "zero"}
_ROWWISE = {"softmax", "log_softmax", "layer_norm", "rms_norm", "normalize", "group_norm"}
_LOCAL_META = {"size", "dim", "ndimension", "stride", "is_contiguous", "data_ptr", "element_size", "is_floating_point", "is_complex", "numel",
               "storage_offset", "is_pinned", "get_device", "is_cuda", "nelement", "_is_view", "is_leaf", "untyped_storage", "__hash__", "__len__",
               "is_shared", "is_sparse", "is_quantized", "has_names", "is_inference", "_version"}


def _norm_name(func):
    n = getattr(func, "__name__", None) or str(func)
    n = n.strip("_")
    if n.startswith("r") and n[1:] in ("add", "sub", "mul", "div", "truediv", "pow", "matmul"):
        n = n[1:] if n != "rsub" else "rsub"
    if n.startswith("i") and n[1:] in ("add", "sub", "mul", "div", "truediv"):
        n = n[1:]
    return {"truediv": "div", "multiply": "mul", "subtract": "sub", "divide": "div"}.get(n, n)


def _flatten(args, kwargs):
    out = []

    def rec(a):
        if isinstance(a, (list, tuple)):
            for b in a:
                rec(b)
        elif isinstance(a, dict):
            for b in a.values():
                rec(b)
        else:
            out.append(a)
    rec(args)
    rec(kwargs)
    return out


def _map(a, f):
    if isinstance(a, (list, tuple)):
        return type(a)(_map(b, f) for b in a)
    if isinstance(a, dict):
        return {k: _map(v, f) for k, v in a.items()}
    return f(a)


def _resolve_partial(t):
    if _is_dt(t) and any(isinstance(p, Partial) for p in t.placements):
        return reshard(t, t.process_mesh, [Replicate() if isinstance(p, Partial) else p for p in t.placements])
    return t


def _align_to(arg, mesh, tgt_pl, tgt_ndim):
    """Local value of `arg` (DistTensor / plain tensor / scalar) that broadcasts against a target with placements tgt_pl."""
    if not isinstance(arg, torch.Tensor):
        return arg
    nd = arg.dim()
    want = []
    for p in tgt_pl:
        if isinstance(p, Shard):
            d = p.dim % tgt_ndim - (tgt_ndim - nd)
            gsz = (arg.shape if _is_dt(arg) else list(arg.shape))[d] if d >= 0 else 1
            want.append(Shard(d) if d >= 0 and gsz != 1 else Replicate())
        else:
            want.append(Replicate())
    if _is_dt(arg):
        a = _resolve_partial(arg)
        return _raw(reshard(a, mesh, want)) if list(a.placements) != want else _raw(a)
    x = _raw(arg)
    for md, p in enumerate(want):
        if isinstance(p, Shard) and mesh.shape[md] > 1:
            x = _Slice.apply(x, _pg(mesh, md), mesh.shape[md], _coord(mesh, md), p.dim)
    return x


def _fallback(func, args, kwargs, mesh):
    full = lambda a: _raw(reshard(_resolve_partial(a), mesh, [Replicate()] * mesh.ndim)) if _is_dt(a) else _raw(a)  # noqa: E731
    ret = func(*_map(args, full), **_map(kwargs, full))
    rep = [Replicate()] * mesh.ndim
    return _map(ret, lambda r: _mk(r, mesh, rep) if isinstance(r, torch.Tensor) else r)


def _dispatch(func, args, kwargs):
    dts = [a for a in _flatten(args, kwargs) if _is_dt(a)]
    with torch._C.DisableTorchFunctionSubclass():
        if not dts:
            return func(*args, **kwargs)
        mesh = dts[0].process_mesh
        name = _norm_name(func)
        # -- metadata / property access: local view ----------------------------------------------------------------
        if name in _LOCAL_META:
            return func(*_map(args, _raw), **_map(kwargs, _raw))
        if name in ("get", "set", "delete"):   # property descriptor (.grad, .dtype, .requires_grad, ...)
            ret = func(*args, **kwargs)   # on the tensor itself: `.grad` lives on the leaf, not on an alias
            if isinstance(ret, torch.Tensor) and not _is_dt(ret) and list(ret.shape) == list(_RAW.size(dts[0])):
                return _mk(ret, mesh, dts[0].placements)
            return ret
        if name == "backward":
            return func(*_map(args, _raw), **_map(kwargs, _raw))
        if name in ("item", "tolist", "numpy", "repr", "str", "format", "bool", "float", "int", "index") and len(dts) == 1 and name in ("item", "tolist", "numpy"):
            full = unshard_dtensor(_resolve_partial(dts[0]))
            return func(full, *args[1:], **kwargs)
        # -- rules -------------------------------------------------------------------------------------------------
        if name in _EW:
            return _rule_elementwise(func, name, args, kwargs, dts, mesh)
        if name in ("matmul", "mm", "bmm"):
            return _rule_matmul(args[0], args[1], mesh)
        if name == "linear":
            x, w = args[0], args[1]
            b = args[2] if len(args) > 2 else kwargs.get("bias")
            y = _rule_matmul(x, _rule_transpose(w, mesh, -1, -2) if _is_dt(w) else _raw(w).t(), mesh)
            return y if b is None else _dispatch(torch.add, (y, b), {})
        if name in ("sum", "mean"):
            return _rule_reduce(func, name, args, kwargs, mesh)
        if name in ("transpose", "t", "permute", "swapaxes"):
            return _rule_permute(func, name, args, kwargs, mesh)
        if name == "embedding":
            return _rule_embedding(func, args, kwargs, mesh)
        if name in _ROWWISE:
            return _rule_rowwise(func, args, kwargs, mesh)
        if name in ("cross_entropy", "nll_loss", "mse_loss", "l1_loss", "binary_cross_entropy_with_logits", "smooth_l1_loss"):
            return _rule_loss(func, args, kwargs, mesh)
        from . import spmd_rules

        r = spmd_rules.lookup(name)
        if r is not None and _is_dt(args[0] if not isinstance(args[0], (list, tuple)) else next((t for t in args[0] if _is_dt(t)), None)):
            try:
                return r(func, name, args, kwargs, mesh)
            except (NotImplementedError, StopIteration):
                pass
        return _fallback(func, args, kwargs, mesh)


def _rule_elementwise(func, name, args, kwargs, dts, mesh):
    if name not in _LINEAR_UNARY and not (name in ("mul", "div") and len(dts) == 1):
        args, kwargs = _map(args, _resolve_partial), _map(kwargs, _resolve_partial)
        dts = [a for a in _flatten(args, kwargs) if _is_dt(a)]
    tgt = max(dts, key=lambda t: (t.dim(), sum(isinstance(p, Shard) for p in t.placements)))
    pl, nd = list(tgt.placements), tgt.dim()
    loc = lambda a: _align_to(a, mesh, pl, nd) if isinstance(a, torch.Tensor) and a is not tgt else (_raw(a) if a is tgt else a)  # noqa: E731
    ret = func(*_map(args, loc), **_map(kwargs, loc))
    if name in ("copy", "zero", "fill", "detach_", "requires_grad") and ret is not None and isinstance(ret, torch.Tensor) and ret.data_ptr() == _raw(args[0]).data_ptr():
        return args[0]
    return _map(ret, lambda r: _mk(r, mesh, pl) if isinstance(r, torch.Tensor) else r)


def _pl_of(x, mesh):
    return list(x.placements) if _is_dt(x) else [Replicate()] * mesh.ndim


def _to_dt(x, mesh):
    return x if _is_dt(x) else _mk(_raw(x), mesh, [Replicate()] * mesh.ndim)


def _rule_matmul(x, w, mesh):
    x, w = _to_dt(_resolve_partial(x), mesh), _to_dt(_resolve_partial(w), mesh)
    xn, wn = x.dim(), w.dim()
    if wn == 1 or xn == 1:
        return _fallback(torch.matmul, (x, w), {}, mesh)
    px, pw = _pl_of(x, mesh), _pl_of(w, mesh)
    lx, lw = _raw(x), _raw(w)
    out_pl = []
    on = max(xn, wn)
    for md in range(mesh.ndim):
        a, b = px[md], pw[md]
        n = mesh.shape[md]
        pg, idx = (_pg(mesh, md), _coord(mesh, md)) if n > 1 else (None, 0)
        a_k = isinstance(a, Shard) and a.dim % xn == xn - 1
        b_k = isinstance(b, Shard) and b.dim % wn == wn - 2
        b_n = isinstance(b, Shard) and b.dim % wn == wn - 1
        a_m = isinstance(a, Shard) and not a_k      # batch or M dim of x
        if n == 1:
            out_pl.append(Replicate())
        elif a_k or b_k:                            # contraction dim sharded -> partial sums
            if not a_k:
                if isinstance(a, Shard):
                    lx = _AllGather.apply(lx, pg, n, idx, a.dim % xn)
                lx = _Slice.apply(lx, pg, n, idx, xn - 1)
            if not b_k:
                if isinstance(b, Shard):
                    lw = _AllGather.apply(lw, pg, n, idx, b.dim % wn)
                lw = _Slice.apply(lw, pg, n, idx, wn - 2)
            out_pl.append(Partial())
        elif b_n:                                   # column parallel
            if isinstance(a, Shard):
                lx = _AllGather.apply(lx, pg, n, idx, a.dim % xn)
            else:
                lx = _IdentityFwdAllReduceBwd.apply(lx, pg)
            out_pl.append(Shard(on - 1))
        elif a_m:                                   # data parallel on rows / batch
            if isinstance(b, Shard):
                lw = _AllGather.apply(lw, pg, n, idx, b.dim % wn)
            out_pl.append(Shard(a.dim % xn + (on - xn)))
        else:
            if isinstance(b, Shard):
                lw = _AllGather.apply(lw, pg, n, idx, b.dim % wn)
            out_pl.append(Replicate())
    return _mk(torch.matmul(lx, lw), mesh, out_pl)


def _rule_transpose(w, mesh, d0, d1):
    nd = w.dim()
    d0, d1 = d0 % nd, d1 % nd
    m = {d0: d1, d1: d0}
    pl = [Shard(m.get(p.dim % nd, p.dim % nd)) if isinstance(p, Shard) else p for p in w.placements]
    return _mk(_raw(w).transpose(d0, d1), mesh, pl)


def _rule_permute(func, name, args, kwargs, mesh):
    x = args[0]
    nd = x.dim()
    if name == "t":
        return _rule_transpose(x, mesh, 0, 1) if nd == 2 else x
    if name in ("transpose", "swapaxes"):
        d0 = args[1] if len(args) > 1 else kwargs.get("dim0", kwargs.get("axis0"))
        d1 = args[2] if len(args) > 2 else kwargs.get("dim1", kwargs.get("axis1"))
        return _rule_transpose(x, mesh, d0, d1)
    dims = args[1:] if len(args) > 2 or (len(args) == 2 and isinstance(args[1], int)) else (args[1] if len(args) > 1 else kwargs["dims"])
    dims = [d % nd for d in dims]
    inv = {old: new for new, old in enumerate(dims)}
    pl = [Shard(inv[p.dim % nd]) if isinstance(p, Shard) else p for p in x.placements]
    return _mk(_raw(x).permute(dims), mesh, pl)


def _rule_reduce(func, name, args, kwargs, mesh):
    x = _resolve_partial(args[0])
    nd = x.dim()
    dim = args[1] if len(args) > 1 else kwargs.get("dim", kwargs.get("axis"))
    keep = bool(args[2] if len(args) > 2 else kwargs.get("keepdim", False))
    dims = list(range(nd)) if dim is None else sorted({d % nd for d in ([dim] if isinstance(dim, int) else list(dim))})
    kw = {k: v for k, v in kwargs.items() if k in ("dtype",)}
    loc = func(_raw(x), dims, keep, **kw) if dim is not None else func(_raw(x), **kw)
    pl = []
    for p in x.placements:
        if isinstance(p, Shard):
            d = p.dim % nd
            if d in dims:
                pl.append(Partial("sum" if name == "sum" else "avg"))
            else:
                pl.append(Shard(d if keep else d - sum(1 for r in dims if r < d)))
        else:
            pl.append(p)
    return _mk(loc, mesh, pl)


def _rule_embedding(func, args, kwargs, mesh):
    ids, w = args[0], args[1]
    if not _is_dt(w) or not any(isinstance(p, Shard) for p in w.placements):
        ids_d = _to_dt(ids, mesh)
        out = func(_raw(ids_d), _raw(w), *args[2:], **kwargs)   # replicated table: its grad is averaged over the dp dims by shard_optimizer
        return _mk(out, mesh, list(ids_d.placements))
    ids_d = _to_dt(ids, mesh)
    il, wl = _raw(ids_d), _raw(w)
    out_pl = list(ids_d.placements)
    for md, p in enumerate(w.placements):
        n = mesh.shape[md]
        if not isinstance(p, Shard) or n == 1:
            continue
        pg, idx = _pg(mesh, md), _coord(mesh, md)
        if isinstance(out_pl[md], Shard):
            il = _AllGather.apply(il, pg, n, idx, out_pl[md].dim)
        if p.dim % 2 == 0:      # vocab parallel: masked lookup -> partial
            per = wl.shape[0]
            lo = idx * per
            mask = (il >= lo) & (il < lo + per)
            o = torch.nn.functional.embedding((il - lo).clamp(0, per - 1), wl)
            o = o * mask.unsqueeze(-1).to(o.dtype)
            return _mk(o, mesh, [Partial() if k == md else q for k, q in enumerate(out_pl)])
        out_pl[md] = Shard(il.dim())
    return _mk(func(il, wl, *args[2:], **kwargs), mesh, out_pl)


def _rule_rowwise(func, args, kwargs, mesh):
    x = _resolve_partial(args[0])
    nd = x.dim()
    if any(isinstance(p, Shard) and p.dim % nd == nd - 1 for p in x.placements):
        x = reshard(x, mesh, [Replicate() if isinstance(p, Shard) and p.dim % nd == nd - 1 else p for p in x.placements])
    rep = [Replicate()] * mesh.ndim
    loc = lambda a: _raw(reshard(_resolve_partial(a), mesh, rep)) if _is_dt(a) else a  # noqa: E731
    ret = func(_raw(x), *_map(args[1:], loc), **_map(kwargs, loc))
    return _mk(ret, mesh, list(x.placements))


def _rule_loss(func, args, kwargs, mesh):
    x = _resolve_partial(args[0])
    red = kwargs.get("reduction", "mean")
    batch_only = all((not isinstance(p, Shard)) or p.dim % x.dim() == 0 for p in x.placements)
    if not batch_only:
        return _fallback(func, args, kwargs, mesh)
    y = args[1]
    yl = _align_to(y, mesh, [p if isinstance(p, Shard) else Replicate() for p in x.placements], x.dim()) if isinstance(y, torch.Tensor) and y.dim() == x.dim() \
        else _raw(reshard(_to_dt(y, mesh), mesh, list(x.placements))) if isinstance(y, torch.Tensor) else y
    rep = [Replicate()] * mesh.ndim
    rest = _map(args[2:], lambda a: _raw(reshard(a, mesh, rep)) if _is_dt(a) else a)
    kw = _map(kwargs, lambda a: _raw(reshard(a, mesh, rep)) if _is_dt(a) else a)
    ret = func(_raw(x), yl, *rest, **kw)
    if red == "none":
        return _mk(ret, mesh, list(x.placements))
    how = "avg" if red == "mean" else "sum"
    return _mk(ret, mesh, [Partial(how) if isinstance(p, Shard) else p for p in x.placements])


# --------------------------------------------------------------------------------------------------------------------
# layer / optimizer / dataloader / model wrappers
# --------------------------------------------------------------------------------------------------------------------
def shard_layer(layer, process_mesh, shard_fn=None, input_fn=None, output_fn=None):
    """Convert the parameters of `layer` to DistTensors: `shard_fn(name, sublayer, mesh)` places them explicitly, everything
    left untouched becomes Replicate."""
    for name, sub in layer.named_sublayers(include_self=True):
        if shard_fn is not None:
            shard_fn(name, sub, process_mesh)
        for pname, p in list(sub._parameters.items()):
            if p is not None and not _is_dt(p):
                sub._parameters[pname] = shard_tensor(p, process_mesh, [Replicate()] * process_mesh.ndim)
    if input_fn is not None:
        layer.register_forward_pre_hook(lambda l, inp: input_fn(inp, process_mesh))
    if output_fn is not None:
        layer.register_forward_post_hook(lambda l, inp, out: output_fn(out, process_mesh))
    return layer


def _grad(p):
    with torch._C.DisableTorchFunctionSubclass():
        g = p.grad
    return None if g is None else _raw(g)


def _set_grad(p, g):
    with torch._C.DisableTorchFunctionSubclass():
        p.grad = g


class _ShardingStage:
    stage = 0

    def __init__(self, mesh=None, sharding_mesh_dim=None):
        if isinstance(mesh, (str, int)) and sharding_mesh_dim is None:      # older call style: ShardingStage1("dp", mesh)
            mesh, sharding_mesh_dim = None, mesh
        elif isinstance(mesh, (str, int)):
            mesh, sharding_mesh_dim = sharding_mesh_dim, mesh
        self.mesh_dim, self.mesh = sharding_mesh_dim, mesh

    def __call__(self, key, param, accumulator):
        return accumulator


class ShardingStage1(_ShardingStage):
    stage = 1


class ShardingStage2(_ShardingStage):
    stage = 2


class ShardingStage3(_ShardingStage):
    stage = 3


class _ShardOptimizer:
    """Optimizer over DistTensor parameters: synchronises the gradients of parameters that are replicated along mesh dims
    (data parallel average), optionally partitions optimizer state ownership (ZeRO-1/2 via `ShardingStageN`), then steps."""

    def __init__(self, optimizer, shard_fn=None, gradient_accumulation_steps=1):
        self._inner, self._shard_fn, self._acc, self._k = optimizer, shard_fn, max(1, int(gradient_accumulation_steps)), 0
        self._owner = None

    def __getattr__(self, n):
        return getattr(self.__dict__["_inner"], n)

    def _params(self):
        return [p for p in self._inner._parameter_list if not isinstance(p, dict)] if not (self._inner._parameter_list and isinstance(self._inner._parameter_list[0], dict)) \
            else [p for g in self._inner._parameter_list for p in g["params"]]

    def _sync_grads(self):
        for p in self._params():
            g = _grad(p)
            if g is None or not _is_dt(p):
                continue
            mesh = p.process_mesh
            for md, pl in enumerate(p.placements):
                if isinstance(pl, Replicate) and mesh.shape[md] > 1:
                    pg = _pg(mesh, md)
                    if dist.get_backend(pg) == "gloo":
                        dist.all_reduce(g, group=pg)
                        g.div_(mesh.shape[md])
                    else:
                        dist.all_reduce(g, op=dist.ReduceOp.AVG, group=pg)

    def step(self):
        self._k += 1
        if self._k % self._acc:
            return
        self._sync_grads()
        st = self._shard_fn
        if isinstance(st, _ShardingStage) and st.stage >= 1:
            self._sharded_step(st)
        else:
            self._inner.step()

    def _sharded_step(self, st):
        params = self._params()
        mesh = st.mesh or next((p.process_mesh for p in params if _is_dt(p)), get_mesh())
        md = mesh.dim_index(st.mesh_dim) if st.mesh_dim is not None else 0
        n, idx, pg = mesh.shape[md], _coord(mesh, md), _pg(mesh, md)
        if n == 1 or pg is None:
            return self._inner.step()
        if self._owner is None:    # greedy size-balanced ownership
            load, self._owner = [0] * n, {}
            for p in sorted(params, key=lambda q: -q.numel()):
                o = load.index(min(load))
                self._owner[id(p)] = o
                load[o] += p.numel()
        held = []
        for p in params:
            if self._owner[id(p)] != idx and _grad(p) is not None:
                held.append((p, _grad(p)))
                _set_grad(p, None)
        self._inner.step()
        ranks = mesh.ranks_along(md, env.get_rank())
        for p in params:
            dist.broadcast(_raw(p).detach(), src=ranks[self._owner[id(p)]], group=pg)
        for p, g in held:
            _set_grad(p, g)

    def clear_grad(self, set_to_zero=True):
        if self._k % self._acc == 0:
            self._inner.clear_grad(set_to_zero)

    def state_dict(self):
        return self._inner.state_dict()

    def set_state_dict(self, sd):
        return self._inner.set_state_dict(sd)


def shard_optimizer(optimizer, shard_fn=None, gradient_accumulation_steps=1):
    return _ShardOptimizer(optimizer, shard_fn, gradient_accumulation_steps)


def shard_scaler(scaler):
    return scaler


class _ShardDataLoader:
    def __init__(self, dataloader, meshes, input_keys=None, shard_dims=None, is_dataset_splitted=False):
        self.loader, self.meshes = dataloader, meshes if isinstance(meshes, (list, tuple)) else [meshes]
        self.shard_dims = shard_dims if isinstance(shard_dims, (list, tuple)) else [shard_dims] * len(self.meshes)
        self.keys, self.splitted = input_keys, is_dataset_splitted

    def __len__(self):
        return len(self.loader)

    def _place(self, t, mesh, sd):
        if not isinstance(t, torch.Tensor):
            return t
        pl = [Replicate()] * mesh.ndim
        if sd is not None:
            pl[mesh.dim_index(sd)] = Shard(0)
        return dtensor_from_local(t, mesh, pl) if self.splitted else shard_tensor(t, mesh, pl)

    def __iter__(self):
        for batch in self.loader:
            mesh, sd = self.meshes[0], self.shard_dims[0]
            if isinstance(batch, dict):
                yield {k: self._place(v, mesh, sd) for k, v in batch.items()}
            elif isinstance(batch, (list, tuple)):
                yield type(batch)(self._place(v, self.meshes[min(i, len(self.meshes) - 1)], self.shard_dims[min(i, len(self.meshes) - 1)]) for i, v in enumerate(batch))
            else:
                yield self._place(batch, mesh, sd)


def shard_dataloader(dataloader, meshes, input_keys=None, shard_dims=None, is_dataset_splitted=False):
    return _ShardDataLoader(dataloader, meshes, input_keys, shard_dims, is_dataset_splitted)


class _Cfg(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class Strategy:
    """Parity: auto_parallel/strategy.py (sharding / gradient_merge / pipeline / amp / recompute / fused_passes groups)."""

    def __init__(self, config=None):
        c = config or {}
        self.sharding = _Cfg(enable=False, stage=1, degree=8, **c.get("sharding", {}))
        self.gradient_merge = _Cfg(enable=False, k_steps=1, avg=True, **c.get("gradient_merge", {}))
        self.pipeline = _Cfg(enable=False, schedule_mode="1F1B", micro_batch_size=1, accumulate_steps=1, **c.get("pipeline", {}))
        self.amp = _Cfg(enable=False, dtype="bfloat16", level="O1", **c.get("amp", {}))
        self.recompute = _Cfg(enable=False, **c.get("recompute", {}))
        self.fused_passes = _Cfg(enable=False, fused_passes_list=[], **c.get("fused_passes", {}))
        self.full_graph = c.get("full_graph", True)


class DistModel:
    """`to_static` result: one callable that runs a whole train / eval / predict step. The step is eager (propagation rules
    above) and, on CUDA with static shapes, can be captured into a CUDA graph through paddle_b200.jit."""

    def __init__(self, layer, loader=None, loss=None, optimizer=None, strategy=None, metrics=None):
        self.network, self._loader, self._loss, self._opt, self._strategy = layer, loader, loss, optimizer, strategy or Strategy()
        self._mode = "train" if optimizer is not None and loss is not None else ("eval" if loss is not None else "predict")

    def train(self):
        self._mode = "train"
        self.network.train()

    def eval(self):
        self._mode = "eval"
        self.network.eval()

    def predict(self):
        self._mode = "predict"
        self.network.eval()

    def dist_main_program(self, mode=None):
        return None

    def state_dict(self, mode="all"):
        sd = dict(self.network.state_dict())
        if self._opt is not None and mode in ("all", "opt"):
            sd.update({f"opt.{k}": v for k, v in self._opt.state_dict().items() if isinstance(v, torch.Tensor)})
        return sd

    def set_state_dict(self, sd):
        self.network.set_state_dict({k: v for k, v in sd.items() if not k.startswith("opt.")})

    def __call__(self, *inputs):
        from ... import amp as _amp

        a = self._strategy.amp
        ctx = _amp.auto_cast(enable=bool(a.enable), level=a.level, dtype=a.dtype)
        if self._mode == "predict":
            with torch.no_grad(), ctx:
                return self.network(*inputs)
        *xs, label = inputs
        if self._mode == "eval":
            with torch.no_grad(), ctx:
                return self._loss(self.network(*xs), label)
        with ctx:
            loss = self._loss(self.network(*xs), label)
        loss.backward()
        self._opt.step()
        self._opt.clear_grad()
        return loss


def to_static(layer, loader=None, loss=None, optimizer=None, strategy=None, input_spec=None):
    if optimizer is not None and not isinstance(optimizer, _ShardOptimizer):
        st = (strategy or Strategy()).sharding
        fn = {1: ShardingStage1, 2: ShardingStage2, 3: ShardingStage3}[int(st.stage)]() if st.enable else None
        optimizer = _ShardOptimizer(optimizer, fn)
    return DistModel(layer, loader, loss, optimizer, strategy)
