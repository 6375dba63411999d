"""Reference evaluator of a fusion group with plain torch ops (differentiable): the third reading of a Spec next to the CUDA and host code
generators.  Used as the autograd fall-back of a fused kernel whose backward cannot be generated, and by the tests as an oracle."""
from __future__ import annotations

import torch

from .expr import Node
from .runtime import _TORCH_DT

_UN = {"neg": torch.neg, "exp": torch.exp, "log": torch.log, "sqrt": torch.sqrt, "rsqrt": torch.rsqrt, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "relu": torch.relu,
       "abs": torch.abs, "erf": torch.erf, "square": torch.square, "reciprocal": torch.reciprocal, "sin": torch.sin, "cos": torch.cos, "floor": torch.floor,
       "ceil": torch.ceil, "log1p": torch.log1p, "expm1": torch.expm1, "logical_not": torch.logical_not, "exp2": torch.exp2, "round": torch.round, "sign": torch.sign}
_BI = {"add": torch.add, "sub": torch.sub, "mul": torch.mul, "div": torch.div, "maximum": torch.maximum, "minimum": torch.minimum, "pow": torch.pow, "gt": torch.gt,
       "lt": torch.lt, "ge": torch.ge, "le": torch.le, "eq": torch.eq, "ne": torch.ne, "logical_and": torch.logical_and, "logical_or": torch.logical_or,
       "floordiv": lambda a, b: torch.div(a, b, rounding_mode="floor"), "fmod": torch.fmod}
_WIDE = {"float16": torch.float32, "bfloat16": torch.float32}


def evaluate(spec, tensors, widen=True, finished=True):
    """Outputs of the group for the given input tensors.  `widen`: compute half types in fp32 and round once at the end (what the generated
    kernels do); False rounds after every op like the unfused program."""
    vals = {}

    def cdt(n):
        t = _TORCH_DT[n.dtype]
        return _WIDE.get(n.dtype, t) if widen else t

    def arg(a, like):
        if isinstance(a, Node):
            return vals[a.id]
        return a

    it = iter(tensors)
    for n in spec.nodes:
        if n.kind == "in":
            t = next(it)
            t = t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t
            if n.attrs.get("view") and tuple(t.shape) != n.shape:
                t = t.reshape(n.shape)
            vals[n.id] = t.to(cdt(n)) if t.dtype != cdt(n) else t
            continue
        if n.kind == "creduce":
            r = vals[n.args[0].id].sum(dim=n.attrs["axes"], keepdim=n.attrs["keepdim"])
            if n.attrs.get("scale") is not None:
                r = r * n.attrs["scale"]
            vals[n.id] = r.to(cdt(n))
            continue
        if n.kind == "reduce":
            x = vals[n.args[0].id]
            keep = len(n.shape) == len(n.args[0].shape)
            vals[n.id] = {"sum": lambda: x.sum(-1, keepdim=keep), "max": lambda: x.amax(-1, keepdim=keep), "min": lambda: x.amin(-1, keepdim=keep)}[n.op]().to(cdt(n))
            continue
        a = [arg(v, n) for v in n.args]
        if n.op == "cast":
            r = a[0].to(cdt(n))
        elif n.op == "where":
            c = a[0] if a[0].dtype == torch.bool else a[0] != 0
            dev = c.device
            r = torch.where(c, *(v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=cdt(n), device=dev) for v in a[1:]))
        elif n.op in _UN:
            r = _UN[n.op](a[0])
        else:
            t0 = next(v for v in a if isinstance(v, torch.Tensor))
            a = [v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=(cdt(n) if n.dtype != "bool" else (torch.float32 if isinstance(v, float) else t0.dtype)),
                                                                  device=t0.device) for v in a]
            r = _BI[n.op](*a)
        if r.dtype != cdt(n):
            r = r.to(cdt(n))
        if tuple(r.shape) != n.shape:
            r = r.expand(n.shape) if r.dim() <= len(n.shape) else r.reshape(n.shape)
        vals[n.id] = r
    outs = []
    for n in spec.outputs:
        v = vals[n.id]
        outs.append(v.to(_TORCH_DT[n.dtype]) if v.dtype != _TORCH_DT[n.dtype] else v)
    if finished:
        from .runtime import finish

        outs = finish(spec, outs)
    return outs


__all__ = ["evaluate"]
