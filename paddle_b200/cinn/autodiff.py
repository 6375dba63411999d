"""Reverse-mode differentiation of a fusion group: forward Spec -> backward Spec (one more generated kernel).

The backward group reads the forward inputs and the output gradients, recomputes the forward values it needs (cheaper than saving them: the
chain is bandwidth-bound) and produces, per differentiable input, the gradient on the group's domain (or per row).  Inputs that were
broadcast into the group get their gradient summed back to their own shape outside the kernel (`sum_to_size`).  A broadcast along the row
(per-row values feeding column work) becomes a reduction inside the kernel, so softmax / normalisation backward is still ONE kernel.
Role parity: the primitive-op VJP rules CINN differentiates decomposed programs with (paddle/fluid/primitive/rule/vjp)."""
from __future__ import annotations

import math

from .codegen import Spec
from .expr import FLOATS, Node, Unsupported

_NO_GRAD = {"floor", "ceil", "round", "sign", "gt", "lt", "ge", "le", "eq", "ne", "logical_and", "logical_or", "logical_not", "floordiv"}


class _Builder:
    def __init__(self, full):
        self.full = tuple(full)
        self.nodes = []

    def node(self, kind, op, args, shape, dtype, space, attrs=None):
        n = Node(kind, op, args, shape, dtype, attrs, space=space)
        n.id = len(self.nodes)
        self.nodes.append(n)
        return n

    def like(self, op, args, ref, dtype=None):
        """Elementwise node in the space / shape of `ref`."""
        return self.node("ew", op, args, ref.shape, dtype or ref.dtype, ref.space)


class BackwardPlan:
    """How to assemble input gradients from the backward kernel's results."""

    def __init__(self):
        self.inputs = []          # per backward-kernel operand: ("in", forward input index) | ("grad", forward output index)
        self.parts = {}           # forward input index -> list of ("out", k) | ("grad", k): tensors to sum_to_size and add


def backward_spec(spec, need_grad):
    """-> (Spec of the backward group or None when every gradient is a pass-through, BackwardPlan)."""
    b = _Builder(spec.full)
    fwd = {}                                        # forward node id -> node in the backward group
    in_index = {n.id: k for k, n in enumerate(spec.inputs)}
    for n in spec.nodes:
        if n.kind == "in":
            fwd[n.id] = b.node("in", "load", [], n.shape, n.dtype, "in", {"src": ("in", in_index[n.id]), "view": n.attrs.get("view", False)})
        elif n.kind == "creduce":
            continue                                   # results of column reductions are never read inside a group
        else:
            fwd[n.id] = b.node(n.kind, n.op, [fwd[a.id] if isinstance(a, Node) else a for a in n.args], n.shape, n.dtype, n.space)
    adj = {}                                        # forward node id -> [contribution nodes] (in that node's space)
    parts = {}                                      # forward input node id -> {(space, shape): [contribution nodes]}
    for k, o in enumerate(spec.outputs):
        if o.kind == "creduce" and o.dtype in FLOATS:
            # d sum_K(x) / dx: the gradient, read under the keepdim shape so that it broadcasts over the reduced axes
            keep = tuple(1 if i in o.attrs["axes"] else e for i, e in enumerate(spec.full))
            g = b.node("in", "load", [], keep, o.dtype, "in", {"src": ("grad", k), "view": True})
            x = o.args[0]
            c = b.node("ew", "mul", [g, float(o.attrs.get("scale") or 1.0)], spec.full, x.dtype, "full")
            if x.kind == "in":
                if need_grad[in_index[x.id]] and x.dtype in FLOATS:
                    parts.setdefault(x.id, {}).setdefault(("full", x.shape), []).append(c)
            elif x.dtype in FLOATS:
                adj.setdefault(x.id, []).append(c)
            continue
        if o.dtype in FLOATS:
            g = b.node("in", "load", [], o.shape, o.dtype, "in", {"src": ("grad", k)})
            if o.space == "row":                     # bring a per-row gradient into the row space before column work reads it
                g = b.node("ew", "add", [g, 0.0], o.shape, o.dtype, "row")
            adj.setdefault(o.id, []).append(g)

    def total(contribs, ref):
        """Sum of the contributions, as a value in the space of `ref`."""
        t = contribs[0]
        for c in contribs[1:]:
            t = b.like("add", [t, c], ref)
        if ref.space == "full" and t.kind != "in" and t.space == "row":
            t = b.like("add", [t, 0.0], ref)
        return t

    def is_float(x):
        return isinstance(x, Node) and x.dtype in FLOATS

    def route(x, c, n):
        """Contribution `c` (in the space of forward node n) to forward argument x."""
        if not is_float(x):
            return
        if x.kind == "in":
            if need_grad[in_index[x.id]]:
                parts.setdefault(x.id, {}).setdefault((n.space, n.shape), []).append(c)
            return
        if n.space == "full" and x.space == "row":              # a per-row value was broadcast along the row: sum the row
            c = b.node("reduce", "sum", [c], x.shape, x.dtype, "row")
        adj.setdefault(x.id, []).append(c)

    for n in reversed(spec.nodes):
        if n.kind in ("in", "creduce") or n.id not in adj or n.dtype not in FLOATS:
            continue
        a = total(adj[n.id], n)
        me = fwd[n.id]
        xs = n.args
        X = [fwd[v.id] if isinstance(v, Node) else v for v in xs]
        L = lambda op, args, dtype=None: b.like(op, args, n, dtype)      # noqa: E731
        if n.kind == "reduce":
            x = xs[0]
            if n.op == "sum":
                c = a
            else:                                               # max / min: the gradient is shared by the tied extrema
                mask = b.node("ew", "cast", [b.node("ew", "eq", [X[0], me], x.shape, "bool", "full")], x.shape, n.dtype, "full")
                cnt = b.node("reduce", "sum", [mask], n.shape, n.dtype, "row")
                c = b.node("ew", "mul", [mask, b.like("div", [a, cnt], n)], x.shape, n.dtype, "full")
            if x.kind == "in":
                if need_grad[in_index[x.id]] and x.dtype in FLOATS:
                    if c.space != "full" or c.kind == "in":
                        c = b.node("ew", "add", [c, 0.0], x.shape, n.dtype, "full")
                    parts.setdefault(x.id, {}).setdefault(("full", x.shape), []).append(c)
            elif x.dtype in FLOATS:
                adj.setdefault(x.id, []).append(c)
            continue
        op = n.op
        if op in _NO_GRAD:
            continue
        if op == "add":
            route(xs[0], a, n)
            if len(xs) > 1:
                route(xs[1], a, n)
        elif op == "sub":
            route(xs[0], a, n)
            if is_float(xs[1]):
                route(xs[1], L("neg", [a]), n)
        elif op == "mul":
            if is_float(xs[0]):
                route(xs[0], L("mul", [a, X[1]]), n)
            if is_float(xs[1]):
                route(xs[1], L("mul", [a, X[0]]), n)
        elif op == "div":
            if is_float(xs[0]):
                route(xs[0], L("div", [a, X[1]]), n)
            if is_float(xs[1]):
                route(xs[1], L("neg", [L("div", [L("mul", [a, me]), X[1]])]), n)
        elif op in ("maximum", "minimum"):
            first = "gt" if op == "maximum" else "lt"
            second = "lt" if op == "maximum" else "gt"
            tie = L("where", [L("eq", [X[0], X[1]], "bool"), 0.5, 0.0])
            if is_float(xs[0]):
                route(xs[0], L("mul", [a, L("where", [L(first, [X[0], X[1]], "bool"), 1.0, tie])]), n)
            if is_float(xs[1]):
                route(xs[1], L("mul", [a, L("where", [L(second, [X[0], X[1]], "bool"), 1.0, tie])]), n)
        elif op == "pow":
            if is_float(xs[0]):
                if isinstance(X[1], Node):
                    d = L("mul", [X[1], L("pow", [X[0], L("sub", [X[1], 1.0])])])
                else:
                    d = L("mul", [L("pow", [X[0], float(X[1]) - 1.0]), float(X[1])])
                route(xs[0], L("mul", [a, d]), n)
            if is_float(xs[1]):
                lg = L("log", [X[0]]) if isinstance(X[0], Node) else math.log(X[0])
                route(xs[1], L("mul", [L("mul", [a, me]), lg]), n)
        elif op == "neg":
            route(xs[0], L("neg", [a]), n)
        elif op == "exp":
            route(xs[0], L("mul", [a, me]), n)
        elif op == "exp2":
            route(xs[0], L("mul", [L("mul", [a, me]), math.log(2.0)]), n)
        elif op == "expm1":
            route(xs[0], L("mul", [a, L("add", [me, 1.0])]), n)
        elif op == "log":
            route(xs[0], L("div", [a, X[0]]), n)
        elif op == "log1p":
            route(xs[0], L("div", [a, L("add", [X[0], 1.0])]), n)
        elif op == "sqrt":
            route(xs[0], L("div", [L("mul", [a, 0.5]), me]), n)
        elif op == "rsqrt":
            route(xs[0], L("mul", [L("mul", [a, -0.5]), L("mul", [L("mul", [me, me]), me])]), n)
        elif op == "tanh":
            route(xs[0], L("mul", [a, L("sub", [1.0, L("mul", [me, me])])]), n)
        elif op == "sigmoid":
            route(xs[0], L("mul", [a, L("mul", [me, L("sub", [1.0, me])])]), n)
        elif op == "relu":
            route(xs[0], L("where", [L("gt", [X[0], 0.0], "bool"), a, 0.0]), n)
        elif op == "abs":
            route(xs[0], L("mul", [a, L("sign", [X[0]])]), n)
        elif op == "erf":
            route(xs[0], L("mul", [L("mul", [a, 2.0 / math.sqrt(math.pi)]), L("exp", [L("neg", [L("mul", [X[0], X[0]])])])]), n)
        elif op == "square":
            route(xs[0], L("mul", [L("mul", [a, 2.0]), X[0]]), n)
        elif op == "reciprocal":
            route(xs[0], L("neg", [L("mul", [a, L("mul", [me, me])])]), n)
        elif op == "sin":
            route(xs[0], L("mul", [a, L("cos", [X[0]])]), n)
        elif op == "cos":
            route(xs[0], L("neg", [L("mul", [a, L("sin", [X[0]])])]), n)
        elif op == "where":
            if is_float(xs[1]):
                route(xs[1], L("where", [X[0], a, 0.0]), n)
            if is_float(xs[2]):
                route(xs[2], L("where", [X[0], 0.0, a]), n)
        elif op == "cast":
            if is_float(xs[0]):
                route(xs[0], L("cast", [a], xs[0].dtype), n)
        elif op == "fmod":
            route(xs[0], a, n)
            if is_float(xs[1]):
                raise Unsupported("fmod divisor gradient")
        else:
            raise Unsupported(f"no derivative rule for {op}")

    # ---- outputs: one tensor per (input, space) that received contributions ---------------------------------------------------------------
    plan = BackwardPlan()
    outputs, out_index = [], {}
    for k, inp in enumerate(spec.inputs):
        if not need_grad[k] or inp.dtype not in FLOATS:
            continue
        plan.parts[k] = []
        for (space, shape), cs in parts.get(inp.id, {}).items():
            if len(cs) == 1 and cs[0].kind == "in":
                plan.parts[k].append(cs[0].attrs["src"])                    # the gradient passes through unchanged
                continue
            broadcast = tuple(inp.shape) != tuple(shape)
            dt = inp.dtype if not broadcast or inp.dtype == "float64" else "float32"     # partial sums of a broadcast input stay wide
            t = cs[0]
            for c in cs[1:]:
                t = b.node("ew", "add", [t, c], shape, dt, space)
            if t.kind == "in" or t.dtype != dt:
                t = b.node("ew", "cast", [t], shape, dt, space)
            if id(t) not in out_index:                                     # one tensor may be the gradient of several inputs
                out_index[id(t)] = len(outputs)
                outputs.append(t)
            plan.parts[k].append(("out", out_index[id(t)]))
    if not outputs:
        return None, plan
    live, stack = set(), list(outputs)
    while stack:
        n = stack.pop()
        if id(n) in live:
            continue
        live.add(id(n))
        stack.extend(v for v in n.args if isinstance(v, Node))
    nodes = [n for n in b.nodes if id(n) in live]
    for k, n in enumerate(nodes):
        n.id = k
    ins = [n for n in nodes if n.kind == "in"]
    plan.inputs = [n.attrs["src"] for n in ins]
    return Spec(spec.name + "_bwd", spec.full, nodes, ins, outputs), plan
