"""Program passes over the recorded op tape. Parity (role): paddle/pir/include/pass + paddle/fluid/pir/transforms/
(dead_code_elimination_pass, constant_folding_pass, common_subexpression_elimination_pass, fused gemm epilogue pass) and
python/paddle/distributed/passes/pass_base.py (PassManager / new_pass / register_pass).

A pass takes a `static.Program` (list of `_Node(fn, args, kwargs, outs, kind)`, values referenced by `_Ref(vid)`) and
rewrites `program.nodes` in place. Fetch targets are protected through `keep` (vids) — `Executor.run` passes its
fetch_list — and nodes of kind "train" (backward + optimizer update) are roots."""
from __future__ import annotations

import torch

_REGISTRY = {}


def register_pass(name):
    def deco(fn):
        _REGISTRY[name] = fn
        fn.pass_name = name
        return fn
    return deco


def _refs(x, out=None):
    from . import _Ref

    out = [] if out is None else out
    if isinstance(x, _Ref):
        out.append(x.vid)
    elif isinstance(x, (list, tuple)):
        for i in x:
            _refs(i, out)
    elif isinstance(x, dict):
        for v in x.values():
            _refs(v, out)
    return out


def _substitute(x, mapping):
    from . import _Ref

    if isinstance(x, _Ref):
        return _Ref(mapping.get(x.vid, x.vid))
    if isinstance(x, (list, tuple)):
        return type(x)(_substitute(i, mapping) for i in x)
    if isinstance(x, dict):
        return {k: _substitute(v, mapping) for k, v in x.items()}
    return x


def _fname(fn):
    return getattr(fn, "__name__", None) or str(fn)


_IMPURE = ("dropout", "rand", "randn", "randint", "bernoulli", "normal", "uniform", "multinomial", "randperm", "poisson", "exponential")


def _is_pure(node):
    n = _fname(node.fn)
    return node.kind == "op" and not n.endswith("_") and not any(k in n for k in _IMPURE) and n not in ("copy", "set", "__setitem__")


@register_pass("dead_code_elimination")
def dead_code_elimination(program, keep=()):
    """Drop ops whose results reach neither a fetch target nor a training node."""
    live = set(keep)
    kept = []
    for node in reversed(program.nodes):
        if node.kind != "op" or not node.outs or not _is_pure(node) or any(v in live for v in node.outs):
            kept.append(node)
            live.update(_refs(node.args))
            live.update(_refs(node.kwargs))
    removed = len(program.nodes) - len(kept)
    program.nodes = kept[::-1]
    return removed


def _key_of(x):
    from . import _Ref

    if isinstance(x, _Ref):
        return ("ref", x.vid)
    if isinstance(x, torch.Tensor):
        return ("tensor", id(x))
    if isinstance(x, (list, tuple)):
        return (type(x).__name__,) + tuple(_key_of(i) for i in x)
    if isinstance(x, dict):
        return ("dict",) + tuple((k, _key_of(v)) for k, v in sorted(x.items()))
    try:
        hash(x)
        return x
    except TypeError:
        return ("obj", id(x))


@register_pass("common_subexpression_elimination")
def common_subexpression_elimination(program, keep=()):
    """Two pure ops with the same callable and the same operands compute the same value: keep the first."""
    seen, mapping, out = {}, {}, []
    for node in program.nodes:
        node.args, node.kwargs = _substitute(node.args, mapping), _substitute(node.kwargs, mapping)
        if _is_pure(node) and node.outs:
            key = (node.fn, _key_of(node.args), _key_of(node.kwargs))
            prev = seen.get(key)
            if prev is not None and len(prev.outs) == len(node.outs) and not any(v in keep for v in node.outs):
                mapping.update(dict(zip(node.outs, prev.outs)))
                continue
            seen[key] = node
        out.append(node)
    removed = len(program.nodes) - len(out)
    program.nodes = out
    return removed


@register_pass("constant_folding")
def constant_folding(program, keep=()):
    """Ops whose operands are all literals / non-trainable constants are evaluated once at compile time; their consumers then
    read the materialised tensor by reference."""
    from . import _Ref

    const = {}     # vid -> tensor
    out = []

    def lower(x):
        if isinstance(x, _Ref):
            return const.get(x.vid, x)
        if isinstance(x, (list, tuple)):
            return type(x)(lower(i) for i in x)
        if isinstance(x, dict):
            return {k: lower(v) for k, v in x.items()}
        return x

    def has_ref_or_param(x):
        if isinstance(x, _Ref):
            return True
        if isinstance(x, torch.Tensor):
            return x.requires_grad or bool(getattr(x, "__dict__", {}).get("_pd_persistable", False))
        if isinstance(x, (list, tuple)):
            return any(has_ref_or_param(i) for i in x)
        if isinstance(x, dict):
            return any(has_ref_or_param(v) for v in x.values())
        return False

    folded = 0
    for node in program.nodes:
        node.args, node.kwargs = lower(node.args), lower(node.kwargs)
        if _is_pure(node) and node.outs and not has_ref_or_param(node.args) and not has_ref_or_param(node.kwargs) \
                and not any(v in keep for v in node.outs):
            with torch.no_grad():
                res = node.fn(*node.args, **node.kwargs)
            flat = [res] if isinstance(res, torch.Tensor) else [r for r in (res if isinstance(res, (list, tuple)) else []) if isinstance(r, torch.Tensor)]
            if len(flat) == len(node.outs):
                const.update(dict(zip(node.outs, flat)))
                folded += 1
                continue
        out.append(node)
    program.nodes = out
    return folded


@register_pass("fuse_gemm_epilogue")
def fuse_gemm_epilogue(program, keep=()):
    """matmul(x, W) followed by `+ bias` (sole consumer) -> one `linear` node, i.e. the tcgen05 GEMM with the bias added in its
    epilogue instead of a second pass over the output. Parity: fused_gemm_epilogue_pass."""
    from ..nn.functional.common import linear

    uses = {}
    for node in program.nodes:
        for v in _refs(node.args) + _refs(node.kwargs):
            uses[v] = uses.get(v, 0) + 1
    producer = {v: n for n in program.nodes for v in n.outs}
    drop, fused = set(), 0
    from . import _Ref

    for node in program.nodes:
        if _fname(node.fn).strip("_") not in ("add", "radd") or len(node.args) != 2 or node.kwargs:
            continue
        a, b = node.args
        for mm_ref, bias in ((a, b), (b, a)):
            if not isinstance(mm_ref, _Ref) or uses.get(mm_ref.vid, 0) != 1 or mm_ref.vid in keep:
                continue
            mm = producer.get(mm_ref.vid)
            if mm is None or _fname(mm.fn) not in ("matmul", "mm") or len(mm.args) != 2 or mm.kwargs or id(mm) in drop:
                continue
            w = mm.args[1]
            bias_ok = isinstance(bias, torch.Tensor) and bias.dim() == 1 and isinstance(w, torch.Tensor) and w.dim() == 2 and bias.shape[0] == w.shape[1]
            if not bias_ok:
                continue
            node.fn, node.args = linear, (mm.args[0], w, bias)
            drop.add(id(mm))
            fused += 1
            break
    program.nodes = [n for n in program.nodes if id(n) not in drop]
    return fused


class PassManager:
    """Parity: distributed/passes/pass_base.py:PassManager."""

    def __init__(self, passes=None):
        self.passes = list(passes or ["constant_folding", "common_subexpression_elimination", "fuse_gemm_epilogue", "dead_code_elimination"])
        self.stats = {}

    def apply(self, program, keep=()):
        keep = set(keep)
        for p in self.passes:
            fn = _REGISTRY[p] if isinstance(p, str) else p
            self.stats[getattr(fn, "pass_name", str(fn))] = fn(program, keep)
        return program


def new_pass(name, attrs=None):
    fn = _REGISTRY[name]

    class _P:
        def apply(self, main_programs, startup_programs=None, context=None):
            for prog in (main_programs if isinstance(main_programs, (list, tuple)) else [main_programs]):
                fn(prog, set((attrs or {}).get("keep", ())))
    return _P()


def apply_pass(program, name, keep=()):
    return _REGISTRY[name](program, set(keep))
