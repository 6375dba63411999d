"""paddle.decomposition: rewrite composite ops of a static Program into primitive ops.

Parity: python/paddle/decomposition/decomp.py (decompose, register_decomp) + paddle/fluid/primitive/composite/composite.h (the
composite -> primitive rules: softmax, log_softmax, gelu, silu, layer_norm, rms_norm, mean, ...).

A rule is an ordinary function written with primitive tensor ops.  `decompose` finds the nodes of the recorded tape
(`static.Program.nodes`) whose op has a rule, re-records the rule on the node's example inputs into a scratch Program, and splices
the resulting primitive nodes in place of the composite node (value ids remapped, the composite's outputs keep their ids, so fetch
targets and downstream consumers are untouched).  Use: custom backends that only implement primitives, higher-order autodiff
(`incubate.autograd` differentiates primitives), and pattern passes that want one canonical form.
"""
from __future__ import annotations

import math

import torch

_REG = {}
_PRIMITIVES = {"add", "sub", "mul", "div", "neg", "exp", "log", "sqrt", "rsqrt", "pow", "abs", "erf", "tanh", "sin", "cos", "maximum", "minimum",
               "where", "sum", "max", "amax", "matmul", "reshape", "transpose", "expand", "cat", "cast", "to", "full_like", "ones_like", "zeros_like",
               "unsqueeze", "squeeze", "ge", "gt", "le", "lt", "eq", "clamp"}


def register_decomp(op_type):
    def deco(fn):
        names = op_type if isinstance(op_type, (list, tuple)) else [op_type]
        for n in names:
            _REG[n] = fn
        return fn

    return deco


def get_decomp_rule(op_type):
    return _REG.get(op_type)


def has_decomp(op_type):
    return op_type in _REG


def _t(x):
    return x.as_subclass(torch.Tensor) if isinstance(x, torch.Tensor) and type(x) is not torch.Tensor else x


# ---------------------------------------------------------------------------------------------------------------- rules
def _axis(kwargs, args, pos, default=-1):
    for k in ("dim", "axis"):
        if k in kwargs and kwargs[k] is not None:
            return kwargs[k]
    return args[pos] if len(args) > pos and args[pos] is not None else default


@register_decomp("softmax")
def _softmax(x, *args, **kwargs):
    d = _axis(kwargs, args, 0)
    dt = kwargs.get("dtype")
    x = x if dt is None else x.to(dt)
    m = torch.amax(x, dim=d, keepdim=True)
    e = torch.exp(x - m)
    return e / torch.sum(e, dim=d, keepdim=True)


@register_decomp("log_softmax")
def _log_softmax(x, *args, **kwargs):
    d = _axis(kwargs, args, 0)
    m = torch.amax(x, dim=d, keepdim=True)
    s = x - m
    return s - torch.log(torch.sum(torch.exp(s), dim=d, keepdim=True))


@register_decomp("sigmoid")
def _sigmoid(x):
    return 1.0 / (1.0 + torch.exp(-x))


@register_decomp(["silu", "swish"])
def _silu(x, *a, **k):
    return x / (1.0 + torch.exp(-x))


@register_decomp("gelu")
def _gelu(x, approximate=False, *a, **k):
    if approximate in (True, "tanh"):
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x * x * x)))
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


@register_decomp("relu")
def _relu(x, *a, **k):
    return torch.maximum(x, torch.zeros_like(x))


@register_decomp("relu6")
def _relu6(x, *a, **k):
    return torch.minimum(torch.maximum(x, torch.zeros_like(x)), torch.full_like(x, 6.0))


@register_decomp("leaky_relu")
def _leaky_relu(x, negative_slope=0.01, *a, **k):
    return torch.where(x >= 0, x, x * negative_slope)


@register_decomp("hardswish")
def _hardswish(x, *a, **k):
    return x * torch.minimum(torch.maximum(x + 3.0, torch.zeros_like(x)), torch.full_like(x, 6.0)) / 6.0


@register_decomp("softplus")
def _softplus(x, beta=1.0, threshold=20.0, *a, **k):
    return torch.where(x * beta > threshold, x, torch.log(1.0 + torch.exp(x * beta)) / beta)


def _count(x, dims=None):
    """Number of reduced elements as a RUN-TIME value (a recorded op): batch dimensions of a static Program are only examples."""
    import paddle_b200 as paddle

    if dims is None:
        return paddle.numel(x).astype(x.dtype)
    shp = paddle.shape(x)
    n = None
    for i in dims:
        n = shp[i] if n is None else n * shp[i]
    return n.astype(x.dtype)


@register_decomp("mean")
def _mean(x, *args, **kwargs):
    d = _axis(kwargs, args, 0, default=None)
    keep = bool(kwargs.get("keepdim", kwargs.get("keepdims", args[1] if len(args) > 1 else False)))
    if d is None or (isinstance(d, (list, tuple)) and len(d) == 0):
        return torch.sum(x) / _count(x)
    dims = [d] if isinstance(d, int) else list(d)
    return torch.sum(x, dim=dims, keepdim=keep) / _count(x, dims)


@register_decomp("layer_norm")
def _layer_norm(x, normalized_shape, weight=None, bias=None, eps=1e-5, *a, **k):
    eps = k.get("epsilon", eps)
    nd = len(normalized_shape) if isinstance(normalized_shape, (list, tuple, torch.Size)) else 1
    dims = list(range(x.dim() - nd, x.dim()))
    n = 1
    for i in dims:
        n *= x.shape[i]
    mu = torch.sum(x, dim=dims, keepdim=True) / n
    c = x - mu
    var = torch.sum(c * c, dim=dims, keepdim=True) / n
    y = c * torch.rsqrt(var + eps)
    if weight is not None:
        y = y * weight
    if bias is not None:
        y = y + bias
    return y


@register_decomp("rms_norm")
def _rms_norm(x, weight=None, eps=1e-6, *a, **k):
    if isinstance(weight, (list, tuple, torch.Size)):          # torch.nn.functional.rms_norm(x, normalized_shape, weight, eps)
        weight, eps = (a[0] if a else k.get("weight")), (a[1] if len(a) > 1 else k.get("eps", 1e-6)) or 1e-6
    eps = k.get("epsilon", eps)
    ms = torch.sum(x * x, dim=-1, keepdim=True) / x.shape[-1]
    y = x * torch.rsqrt(ms + eps)
    return y if weight is None else y * weight


@register_decomp("swiglu")
def _swiglu(x, y=None, *a, **k):
    if y is None:
        h = x.shape[-1] // 2
        x, y = x[..., :h], x[..., h:]
    return x / (1.0 + torch.exp(-x)) * y


@register_decomp("mse_loss")
def _mse(x, y, reduction="mean", *a, **k):
    d = (x - y) * (x - y)
    return d if reduction == "none" else (torch.sum(d) if reduction == "sum" else torch.sum(d) / _count(d))


@register_decomp("l1_loss")
def _l1(x, y, reduction="mean", *a, **k):
    d = torch.abs(x - y)
    return d if reduction == "none" else (torch.sum(d) if reduction == "sum" else torch.sum(d) / _count(d))


@register_decomp("addmm")
def _addmm(inp, x, y, beta=1.0, alpha=1.0, *a, **k):
    return inp * beta + torch.matmul(x, y) * alpha


@register_decomp("square")
def _square(x, *a, **k):
    return x * x


@register_decomp("reciprocal")
def _reciprocal(x, *a, **k):
    return 1.0 / x


@register_decomp("stack")
def _stack(xs, dim=0, *a, **k):
    d = k.get("axis", dim)
    return torch.cat([torch.unsqueeze(t, d) for t in xs], dim=d)


@register_decomp("dropout")
def _dropout(x, p=0.5, training=True, *a, **k):
    if k.get("axis") is not None or k.get("mode", "upscale_in_train") != "upscale_in_train":
        raise NotImplementedError
    if not training or p == 0.0:
        return x * 1.0
    keep = (torch.rand_like(x) >= p).to(x.dtype)
    return x * keep / (1.0 - p)


# ---------------------------------------------------------------------------------------------------------------- the pass
def _fname(fn):
    return (getattr(fn, "__name__", None) or str(fn)).strip("_")


def decompose(program, src_vars=None, blacklist=frozenset(), whitelist=frozenset()):
    """Replace every composite node of `program` (a static.Program) that has a registered rule by primitive nodes, in place.
    blacklist / whitelist filter op names. Returns `src_vars` (their value ids are preserved) or the program."""
    from . import static as S

    if not hasattr(program, "nodes"):
        return src_vars if src_vars is not None else program
    new_nodes, n_done = [], 0
    for node in program.nodes:
        name = _fname(node.fn)
        rule = _REG.get(name)
        if rule is None or node.kind != "op" or name in blacklist or (whitelist and name not in whitelist) or len(node.outs) != 1:
            new_nodes.append(node)
            continue
        try:
            sub_nodes = _expand(program, node, rule, S)
        except Exception:  # noqa: BLE001  (rule does not cover this call form: keep the composite op)
            sub_nodes = None
        if not sub_nodes:
            new_nodes.append(node)
            continue
        new_nodes.extend(sub_nodes)
        n_done += 1
    program.nodes = new_nodes
    program.__dict__["_decomposed"] = program.__dict__.get("_decomposed", 0) + n_done
    return src_vars if src_vars is not None else program


def _expand(program, node, rule, S):
    inputs = {}       # main vid -> example tensor

    def decode(x):
        if isinstance(x, S._Ref):
            t = program._keep[x.vid]
            inputs[x.vid] = t
            return t
        if isinstance(x, (list, tuple)):
            return type(x)(decode(i) for i in x)
        if isinstance(x, dict):
            return {k: decode(v) for k, v in x.items()}
        return x

    ex_args, ex_kwargs = decode(node.args), decode(node.kwargs)
    ex_kwargs = {k: v for k, v in ex_kwargs.items() if k != "name"}
    sub = S.Program()
    in_map = {}
    for vid, t in inputs.items():
        in_map[sub._new_vid(t)] = vid
    with S.program_guard(sub):
        out = rule(*ex_args, **ex_kwargs)
    if not isinstance(out, torch.Tensor) or id(out) not in sub._vids or not sub.nodes:
        return None
    expect = program._keep[node.outs[0]]
    if tuple(out.shape) != tuple(expect.shape) or out.dtype != expect.dtype:
        return None
    out_sub = sub._vids[id(out)]
    remap = dict(in_map)
    remap[out_sub] = node.outs[0]

    def vid_of(v):
        if v not in remap:                       # an intermediate of the rule: give it a value id of the main program
            remap[v] = program._next
            program._next += 1
            program._keep.append(sub._keep[v])
        return remap[v]

    def rewrite(x):
        if isinstance(x, S._Ref):
            return S._Ref(vid_of(x.vid))
        if isinstance(x, (list, tuple)):
            return type(x)(rewrite(i) for i in x)
        if isinstance(x, dict):
            return {k: rewrite(v) for k, v in x.items()}
        return x

    res = []
    for n in sub.nodes:
        res.append(S._Node(n.fn, rewrite(n.args), rewrite(n.kwargs), [vid_of(v) for v in n.outs], n.kind))
    return res
