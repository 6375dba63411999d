"""Tensor-expression nodes of a fusion group and the front end that builds them from recorded ops.

A group computes over ONE iteration domain `S` (the full shape; `rows = prod(S[:-1])`, `cols = S[-1]`).  Its nodes are
  * `in`      - a tensor that comes from outside the group (any shape that broadcasts against the op that reads it),
  * `ew`      - an elementwise primitive over `S` ("full") or over the per-row shape `S[:-1] + (1,)` / `S[:-1]` ("row"),
  * `reduce`  - sum / max / min over the last axis of a full node; the result is a row value.
Composite ops (softmax, log_softmax, mean, silu, ...) are decomposed into these primitives on the way in, which is what lets a chain like
`softmax(x * w + b)` become one kernel.  Role parity: CINN's op lowering to primitive compute bodies + the decomposer
(paddle/cinn/hlir/dialect/operator/transforms, paddle/fluid/primitive/composite)."""
from __future__ import annotations

import math

import torch


class Unsupported(Exception):
    pass


_DT_NAME = {torch.float32: "float32", torch.float64: "float64", torch.float16: "float16", torch.bfloat16: "bfloat16", torch.int64: "int64", torch.int32: "int32",
            torch.bool: "bool", torch.uint8: "uint8", torch.int8: "int8", torch.int16: "int16"}
FLOATS = ("float16", "bfloat16", "float32", "float64")
SUPPORTED_DTYPES = FLOATS + ("int32", "int64", "bool", "uint8")


def compute_type(dtype):
    """C type a value of this dtype is computed in (half types are widened: one rounding at the store, like every fused kernel here)."""
    return {"float16": "float", "bfloat16": "float", "float32": "float", "float64": "double", "int32": "int", "int64": "long long", "bool": "bool",
            "uint8": "int"}[dtype]


class Node:
    __slots__ = ("id", "kind", "op", "args", "shape", "dtype", "attrs", "space", "level", "value_id")

    def __init__(self, kind, op, args, shape, dtype, attrs=None, space="full"):
        self.id = -1
        self.kind, self.op, self.args, self.shape, self.dtype, self.attrs, self.space = kind, op, list(args), tuple(int(d) for d in shape), dtype, attrs or {}, space
        self.level = 0
        self.value_id = None          # IR value this node stands for (None for the inner nodes of a decomposition)

    def __repr__(self):
        return f"n{self.id}:{self.kind}:{self.op}{list(self.shape)}:{self.dtype}"


# ---- primitive table: name -> arity ------------------------------------------------------------------------------------------------------
UNARY = {"neg", "exp", "log", "sqrt", "rsqrt", "tanh", "sigmoid", "relu", "abs", "erf", "square", "reciprocal", "sin", "cos", "floor", "ceil", "log1p", "expm1",
         "logical_not", "exp2", "round", "sign"}
BINARY = {"add", "sub", "mul", "div", "maximum", "minimum", "pow", "gt", "lt", "ge", "le", "eq", "ne", "logical_and", "logical_or", "floordiv", "fmod"}
COMPARE = {"gt", "lt", "ge", "le", "eq", "ne"}
REDUCE = {"sum", "max", "min"}


class GroupBuilder:
    """Accumulates the nodes of one group.  `full` is the domain S."""

    def __init__(self, full):
        self.full = tuple(int(d) for d in full)
        self.nodes = []
        self.inputs = {}          # (ir value id, shape it is read as) -> input node
        self.by_value = {}        # ir value id -> node that computes it inside the group

    # -- shape classes
    def row_shapes(self):
        s = self.full
        return (s[:-1] + (1,), s[:-1]) if len(s) >= 1 else ()

    def space_of(self, shape):
        shape = tuple(shape)
        if shape == self.full:
            return "full"
        if len(self.full) >= 1 and shape in self.row_shapes():
            return "row"
        return None

    def _add(self, n):
        n.id = len(self.nodes)
        self.nodes.append(n)
        return n

    def input(self, value_id, shape, dtype, view=None):
        """Input node of an outside value.  `view`: read the (contiguous) tensor under another shape with the same number of elements -
        how per-channel statistics [C] line up with [N, C, H, W]."""
        if value_id in self.by_value and view is None:
            return self.by_value[value_id]
        shape = tuple(int(d) for d in (view if view is not None else shape))
        key = (value_id, shape)
        if key not in self.inputs:
            if dtype not in SUPPORTED_DTYPES:
                raise Unsupported(f"input dtype {dtype}")
            n = self._add(Node("in", "load", [], shape, dtype, {"value": value_id, "view": view is not None}, space="in"))
            self.inputs[key] = n
        return self.inputs[key]

    def view(self, node, shape):
        """The same outside tensor read under another shape."""
        if not isinstance(node, Node) or node.kind != "in":
            raise Unsupported("only values from outside the group can be re-viewed")
        n = 1
        for d in shape:
            n *= int(d)
        m = 1
        for d in node.shape:
            m *= d
        if n != m:
            raise Unsupported("view changes the number of elements")
        return self.input(node.attrs["value"], None, node.dtype, view=shape)

    def ew(self, op, args, shape, dtype):
        space = self.space_of(shape)
        if space is None:
            raise Unsupported(f"shape {list(shape)} is neither the domain {list(self.full)} nor its per-row shape")
        if space == "row" and any(n.kind == "creduce" for n in self.nodes):
            raise Unsupported("row values in a column-reduction group")
        for a in args:
            if isinstance(a, Node) and (a.attrs.get("post") or a.attrs.get("sealed")):
                raise Unsupported("a value that is finished outside the kernel cannot be read inside the group")
            if isinstance(a, Node) and a.kind != "in":
                if space == "full" and not (a.space == "full" or (a.space == "row" and a.shape == self.full[:-1] + (1,))):
                    raise Unsupported("a non-keepdim row value does not broadcast per row")
                if space == "row" and not (a.space == "row" and a.shape == tuple(shape)):
                    raise Unsupported("row op over values of another shape")
            if isinstance(a, Node) and a.kind == "in":
                _check_broadcast(a.shape, shape)
        if dtype not in SUPPORTED_DTYPES:
            raise Unsupported(f"dtype {dtype}")
        return self._add(Node("ew", op, args, shape, dtype, space=space))

    def reduce(self, op, x, keepdim, dtype):
        if not isinstance(x, Node) or x.shape != self.full or (x.kind != "in" and x.space != "full"):
            raise Unsupported("reduction input is not a full-domain value")
        if any(n.kind == "creduce" for n in self.nodes):
            raise Unsupported("row and column reductions in one group")
        shape = self.full[:-1] + ((1,) if keepdim else ())
        return self._add(Node("reduce", op, [x], shape, dtype, space="row"))

    def creduce(self, x, axes, keepdim, dtype, scale=None):
        """Sum over a contiguous run of axes that does NOT include the last one ("column" reduction: bias gradients, batch statistics):
        the domain is read as [A, K, B] with the run collapsed to K.  The result leaves the group (nothing inside may read it)."""
        if not isinstance(x, Node) or x.shape != self.full or (x.kind != "in" and x.space != "full"):
            raise Unsupported("reduction input is not a full-domain value")
        if any(n.kind == "reduce" or (n.kind == "ew" and n.space == "row") for n in self.nodes):
            raise Unsupported("row and column reductions in one group")
        nd = len(self.full)
        axes = sorted(set(int(a) % nd for a in axes))
        if not axes or axes != list(range(axes[0], axes[-1] + 1)) or axes[-1] >= nd - 1:
            raise Unsupported("reduction axes are not a contiguous run before the last axis")
        A = K = B = 1
        for i, e in enumerate(self.full):
            if i < axes[0]:
                A *= e
            elif i <= axes[-1]:
                K *= e
            else:
                B *= e
        akb = (A, K, B)
        for n in self.nodes:
            if n.kind == "creduce" and n.attrs["akb"] != akb:
                raise Unsupported("column reductions over different axes in one group")
        shape = tuple(1 if i in axes else e for i, e in enumerate(self.full)) if keepdim else tuple(e for i, e in enumerate(self.full) if i not in axes)
        return self._add(Node("creduce", "sum", [x], shape, dtype, {"akb": akb, "sealed": True, "scale": scale, "axes": tuple(axes), "keepdim": bool(keepdim)}, space="col"))


def _check_broadcast(src, dst):
    src, dst = tuple(src), tuple(dst)
    if len(src) > len(dst):
        raise Unsupported(f"cannot broadcast {list(src)} to {list(dst)}")
    for a, b in zip(reversed(src), reversed(dst)):
        if a != 1 and a != b:
            raise Unsupported(f"cannot broadcast {list(src)} to {list(dst)}")


# ---- front end: a recorded call -> nodes ---------------------------------------------------------------------------------------------
def _is_num(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool) or isinstance(v, bool)


class Frontend:
    """`lower(name, args, kwargs, out_shape, out_dtype)` where tensor arguments are already Nodes (or python numbers).  Returns the node that
    holds the op's result.  Raises Unsupported for anything that is not a fusible primitive / composite."""

    def __init__(self, gb):
        self.g = gb

    def lower(self, name, args, kwargs, shape, dtype):
        fn = getattr(self, "op_" + name, None)
        if fn is None:
            if name in UNARY:
                return self._unary(name, args, kwargs, shape, dtype)
            if name in BINARY:
                return self._binary(name, args, kwargs, shape, dtype)
            raise Unsupported(f"op {name}")
        return fn(args, kwargs, shape, dtype)

    # -- helpers
    def _unary(self, name, args, kwargs, shape, dtype):
        if len(args) != 1 or any(v is not None for k, v in kwargs.items() if k not in ("out",)) or not isinstance(args[0], Node):
            raise Unsupported(f"{name} signature")
        return self.g.ew(name, [args[0]], shape, dtype)

    def _binary(self, name, args, kwargs, shape, dtype, swap=False):
        extra = {k: v for k, v in kwargs.items() if v is not None and k not in ("out",)}
        if len(args) != 2 or extra:
            raise Unsupported(f"{name} signature")
        a, b = (args[1], args[0]) if swap else args
        if not all(isinstance(v, Node) or _is_num(v) for v in (a, b)) or not any(isinstance(v, Node) for v in (a, b)):
            raise Unsupported(f"{name} operands")
        return self.g.ew(name, [a, b], shape, dtype)

    def _scaled(self, name, args, kwargs, shape, dtype):
        alpha = kwargs.get("alpha", 1)
        rest = {k: v for k, v in kwargs.items() if k != "alpha"}
        if alpha not in (1, 1.0):
            if not _is_num(alpha) or len(args) != 2:
                raise Unsupported("alpha")
            b = args[1] * alpha if _is_num(args[1]) else self.g.ew("mul", [args[1], alpha], _shape_of(args[1]), dtype)
            args = [args[0], b]
        return self._binary(name, args, rest, shape, dtype)

    # -- arithmetic with torch spellings
    def op_add(self, a, k, s, d):
        return self._scaled("add", a, k, s, d)

    def op_sub(self, a, k, s, d):
        return self._scaled("sub", a, k, s, d)

    op_subtract = op_sub

    def op_radd(self, a, k, s, d):
        return self._binary("add", a, k, s, d, swap=True)

    def op_rsub(self, a, k, s, d):
        return self._binary("sub", a, k, s, d, swap=True)

    def op_rmul(self, a, k, s, d):
        return self._binary("mul", a, k, s, d, swap=True)

    def op_multiply(self, a, k, s, d):
        return self._binary("mul", a, k, s, d)

    def op_rtruediv(self, a, k, s, d):
        return self._binary("div", a, k, s, d, swap=True)

    def op_rdiv(self, a, k, s, d):
        return self._binary("div", a, k, s, d, swap=True)

    def op_rpow(self, a, k, s, d):
        return self._binary("pow", a, k, s, d, swap=True)

    def op_div(self, a, k, s, d):
        mode = k.get("rounding_mode")
        k = {x: v for x, v in k.items() if x != "rounding_mode"}
        if mode is None:
            if d not in FLOATS:
                raise Unsupported("integer true division")
            return self._binary("div", a, k, s, d)
        if mode == "floor":
            return self._binary("floordiv", a, k, s, d)
        raise Unsupported("div rounding mode")

    op_truediv = op_true_divide = op_divide = op_div

    def op_floor_divide(self, a, k, s, d):
        return self._binary("floordiv", a, k, s, d)

    op_floordiv = op_floor_divide

    def op_max(self, a, k, s, d):
        if len(a) == 2 and isinstance(a[1], Node) and not k:
            return self._binary("maximum", a, k, s, d)
        raise Unsupported("max with indices")

    def op_min(self, a, k, s, d):
        if len(a) == 2 and isinstance(a[1], Node) and not k:
            return self._binary("minimum", a, k, s, d)
        raise Unsupported("min with indices")

    def op_greater_than(self, a, k, s, d):
        return self._binary("gt", a, k, s, d)

    def op_less_than(self, a, k, s, d):
        return self._binary("lt", a, k, s, d)

    def op_greater_equal(self, a, k, s, d):
        return self._binary("ge", a, k, s, d)

    def op_less_equal(self, a, k, s, d):
        return self._binary("le", a, k, s, d)

    def op_equal(self, a, k, s, d):
        return self._binary("eq", a, k, s, d)

    def op_not_equal(self, a, k, s, d):
        return self._binary("ne", a, k, s, d)

    def op_negative(self, a, k, s, d):
        return self._unary("neg", a, k, s, d)

    def op_absolute(self, a, k, s, d):
        return self._unary("abs", a, k, s, d)

    # -- activations
    def op_silu(self, a, k, s, d):
        if len(a) != 1 or k.get("inplace"):
            raise Unsupported("silu signature")
        return self.g.ew("mul", [a[0], self.g.ew("sigmoid", [a[0]], s, d)], s, d)

    def op_swiglu(self, a, k, s, d):
        if len(a) != 2 or not all(isinstance(v, Node) for v in a):
            raise Unsupported("swiglu signature")
        return self.g.ew("mul", [self.g.ew("mul", [a[0], self.g.ew("sigmoid", [a[0]], s, d)], s, d), a[1]], s, d)

    def op_relu(self, a, k, s, d):
        if len(a) != 1 or k.get("inplace"):
            raise Unsupported("relu signature")
        return self.g.ew("relu", [a[0]], s, d)

    def op_gelu(self, a, k, s, d):
        approx = k.get("approximate", a[1] if len(a) > 1 else "none")
        x = a[0]
        if approx in ("none", False):
            t = self.g.ew("erf", [self.g.ew("mul", [x, 1.0 / math.sqrt(2.0)], s, d)], s, d)
            return self.g.ew("mul", [self.g.ew("mul", [x, 0.5], s, d), self.g.ew("add", [t, 1.0], s, d)], s, d)
        if approx in ("tanh", True):
            x3 = self.g.ew("mul", [self.g.ew("square", [x], s, d), x], s, d)
            inner = self.g.ew("mul", [self.g.ew("add", [x, self.g.ew("mul", [x3, 0.044715], s, d)], s, d), math.sqrt(2.0 / math.pi)], s, d)
            return self.g.ew("mul", [self.g.ew("mul", [x, 0.5], s, d), self.g.ew("add", [self.g.ew("tanh", [inner], s, d), 1.0], s, d)], s, d)
        raise Unsupported("gelu mode")

    def op_relu6(self, a, k, s, d):
        if k.get("inplace"):
            raise Unsupported("in-place activation")
        return self.g.ew("minimum", [self.g.ew("maximum", [a[0], 0.0], s, d), 6.0], s, d)

    def op_hardtanh(self, a, k, s, d):
        lo, hi = k.get("min_val", a[1] if len(a) > 1 else -1.0), k.get("max_val", a[2] if len(a) > 2 else 1.0)
        if k.get("inplace") or not (_is_num(lo) and _is_num(hi)):
            raise Unsupported("hardtanh signature")
        return self.g.ew("minimum", [self.g.ew("maximum", [a[0], float(lo)], s, d), float(hi)], s, d)

    def op_elu(self, a, k, s, d):
        alpha = k.get("alpha", a[1] if len(a) > 1 else 1.0)
        if k.get("inplace") or not _is_num(alpha):
            raise Unsupported("elu signature")
        x = a[0]
        neg = self.g.ew("mul", [self.g.ew("expm1", [x], s, d), float(alpha)], s, d)
        return self.g.ew("where", [self.g.ew("gt", [x, 0.0], s, "bool"), x, neg], s, d)

    def op_selu(self, a, k, s, d):
        if k.get("inplace") or len(a) != 1:
            raise Unsupported("selu signature")
        alpha, scale = 1.6732632423543772, 1.0507009873554805
        x = a[0]
        neg = self.g.ew("mul", [self.g.ew("expm1", [x], s, d), alpha], s, d)
        return self.g.ew("mul", [self.g.ew("where", [self.g.ew("gt", [x, 0.0], s, "bool"), x, neg], s, d), scale], s, d)

    def _softplus1(self, x, s, d):
        sp = self.g.ew("log1p", [self.g.ew("exp", [x], s, d)], s, d)
        return self.g.ew("where", [self.g.ew("gt", [x, 20.0], s, "bool"), x, sp], s, d)

    def op_mish(self, a, k, s, d):
        if k.get("inplace") or len(a) != 1:
            raise Unsupported("mish signature")
        return self.g.ew("mul", [a[0], self.g.ew("tanh", [self._softplus1(a[0], s, d)], s, d)], s, d)

    def op_log_sigmoid(self, a, k, s, d):
        if len(a) != 1 or k:
            raise Unsupported("log_sigmoid signature")
        x = a[0]
        t = self.g.ew("log1p", [self.g.ew("exp", [self.g.ew("neg", [self.g.ew("abs", [x], s, d)], s, d)], s, d)], s, d)
        return self.g.ew("sub", [self.g.ew("minimum", [x, 0.0], s, d), t], s, d)

    op_logsigmoid = op_log_sigmoid

    def op_tanhshrink(self, a, k, s, d):
        if len(a) != 1 or k:
            raise Unsupported("tanhshrink signature")
        return self.g.ew("sub", [a[0], self.g.ew("tanh", [a[0]], s, d)], s, d)

    def op_softsign(self, a, k, s, d):
        if len(a) != 1 or k:
            raise Unsupported("softsign signature")
        return self.g.ew("div", [a[0], self.g.ew("add", [self.g.ew("abs", [a[0]], s, d), 1.0], s, d)], s, d)

    # -- normalisations over the last axis (a lone norm stays on its hand-written kernel: groups need two recorded ops)
    def _affine(self, y, w, b, s, d):
        if w is not None:
            if not isinstance(w, Node):
                raise Unsupported("norm weight")
            y = self.g.ew("mul", [y, w], s, d)
        if b is not None:
            if not isinstance(b, Node):
                raise Unsupported("norm bias")
            y = self.g.ew("add", [y, b], s, d)
        return y

    def op_layer_norm(self, a, k, s, d):
        x = a[0]
        shape = k.get("normalized_shape", a[1] if len(a) > 1 else None)
        w, b = k.get("weight", a[2] if len(a) > 2 else None), k.get("bias", a[3] if len(a) > 3 else None)
        eps = k.get("eps", k.get("epsilon", a[4] if len(a) > 4 else 1e-5))
        if isinstance(shape, int):
            shape = [shape]
        if not isinstance(x, Node) or x.shape != self.g.full or d not in FLOATS or list(shape or []) != [self.g.full[-1]] or not _is_num(eps):
            raise Unsupported("layer_norm over more than the last axis")
        rs, n = self.g.full[:-1] + (1,), self.g.full[-1]
        mu = self.g.ew("mul", [self.g.reduce("sum", x, True, d), 1.0 / n], rs, d)
        xc = self.g.ew("sub", [x, mu], s, d)
        var = self.g.ew("mul", [self.g.reduce("sum", self.g.ew("mul", [xc, xc], s, d), True, d), 1.0 / n], rs, d)
        inv = self.g.ew("rsqrt", [self.g.ew("add", [var, float(eps)], rs, d)], rs, d)
        return self._affine(self.g.ew("mul", [xc, inv], s, d), w, b, s, d)

    def op_rms_norm(self, a, k, s, d):
        x = a[0]
        w = k.get("weight", a[1] if len(a) > 1 else None)
        eps = k.get("eps", k.get("epsilon", a[2] if len(a) > 2 else 1e-6))
        b = k.get("bias", a[3] if len(a) > 3 else None)
        res = k.get("residual", a[4] if len(a) > 4 else None)
        if not isinstance(x, Node) or x.shape != self.g.full or d not in FLOATS or res is not None or not _is_num(eps) or len(a) > 5:
            raise Unsupported("rms_norm signature")
        rs, n = self.g.full[:-1] + (1,), self.g.full[-1]
        ms = self.g.ew("mul", [self.g.reduce("sum", self.g.ew("mul", [x, x], s, d), True, d), 1.0 / n], rs, d)
        inv = self.g.ew("rsqrt", [self.g.ew("add", [ms, float(eps)], rs, d)], rs, d)
        return self._affine(self.g.ew("mul", [x, inv], s, d), w, b, s, d)

    def op_batch_norm(self, a, k, s, d):
        """Inference form only: y = (x - mean_c) * rsqrt(var_c + eps) * w_c + b_c with the channel on axis 1 (axis -1 for 2-D inputs)."""
        x = a[0]
        mean, var = k.get("running_mean", a[1] if len(a) > 1 else None), k.get("running_var", a[2] if len(a) > 2 else None)
        w, b = k.get("weight", a[3] if len(a) > 3 else None), k.get("bias", a[4] if len(a) > 4 else None)
        training = k.get("training", a[5] if len(a) > 5 else False)
        eps = k.get("eps", k.get("epsilon", a[7] if len(a) > 7 else 1e-5))
        if training or not isinstance(x, Node) or not isinstance(mean, Node) or not isinstance(var, Node) or len(s) < 2 or d not in FLOATS or not _is_num(eps):
            raise Unsupported("batch_norm in training mode / without running statistics")
        if k.get("data_format", "NCHW") not in ("NCHW", "NCL", "NCDHW", "NC"):
            raise Unsupported("batch_norm layout")
        cshape = (s[1],) + (1,) * (len(s) - 2)

        def per_channel(t):
            if t is None:
                return None
            if not isinstance(t, Node) or t.shape != (s[1],):
                raise Unsupported("batch_norm statistics shape")
            return self.g.view(t, cshape) if len(s) > 2 else t

        mean, var, w, b = (per_channel(t) for t in (mean, var, w, b))
        inv = self.g.ew("rsqrt", [self.g.ew("add", [var, float(eps)], s, d)], s, d)        # broadcast inputs: evaluated per element, read once per row
        return self._affine(self.g.ew("mul", [self.g.ew("sub", [x, mean], s, d), inv], s, d), w, b, s, d)

    def op_softplus(self, a, k, s, d):
        beta, thr = k.get("beta", a[1] if len(a) > 1 else 1.0), k.get("threshold", a[2] if len(a) > 2 else 20.0)
        x = a[0]
        bx = x if beta == 1.0 else self.g.ew("mul", [x, float(beta)], s, d)
        sp = self.g.ew("log1p", [self.g.ew("exp", [bx], s, d)], s, d)
        if beta != 1.0:
            sp = self.g.ew("mul", [sp, 1.0 / float(beta)], s, d)
        return self.g.ew("where", [self.g.ew("gt", [bx, float(thr)], s, "bool"), x, sp], s, d)

    def op_leaky_relu(self, a, k, s, d):
        slope = k.get("negative_slope", a[1] if len(a) > 1 else 0.01)
        if k.get("inplace") or not _is_num(slope):
            raise Unsupported("leaky_relu signature")
        x = a[0]
        return self.g.ew("where", [self.g.ew("gt", [x, 0.0], s, "bool"), x, self.g.ew("mul", [x, float(slope)], s, d)], s, d)

    def op_hardswish(self, a, k, s, d):
        if len(a) != 1 or k.get("inplace"):
            raise Unsupported("hardswish signature")
        x = a[0]
        r6 = self.g.ew("minimum", [self.g.ew("maximum", [self.g.ew("add", [x, 3.0], s, d), 0.0], s, d), 6.0], s, d)
        return self.g.ew("mul", [self.g.ew("mul", [x, r6], s, d), 1.0 / 6.0], s, d)

    # -- selection / clamp / cast
    def op_where(self, a, k, s, d):
        if len(a) != 3 or k or not isinstance(a[0], Node):
            raise Unsupported("where signature")
        return self.g.ew("where", list(a), s, d)

    def op_clamp(self, a, k, s, d):
        lo = k.get("min", a[1] if len(a) > 1 else None)
        hi = k.get("max", a[2] if len(a) > 2 else None)
        x = a[0]
        if lo is not None:
            x = self.g.ew("maximum", [x, lo], s, d)
        if hi is not None:
            x = self.g.ew("minimum", [x, hi], s, d)
        if x is a[0]:
            raise Unsupported("clamp without bounds")
        return x

    op_clip = op_clamp

    def op_pow(self, a, k, s, d):
        if len(a) != 2 or k:
            raise Unsupported("pow signature")
        if _is_num(a[1]) and a[1] == 2:
            return self.g.ew("square", [a[0]], s, d)
        if _is_num(a[1]) and a[1] == 0.5:
            return self.g.ew("sqrt", [a[0]], s, d)
        return self._binary("pow", a, k, s, d)

    def _cast(self, x, s, d):
        if not isinstance(x, Node):
            raise Unsupported("cast of a non-tensor")
        return self.g.ew("cast", [x], s, d)

    def op_to(self, a, k, s, d):
        others = [v for v in list(a[1:]) + list(k.values()) if v is not None and not isinstance(v, (torch.dtype, bool))]
        if others:
            raise Unsupported("to(device / tensor)")
        return self._cast(a[0], s, d)

    def op_type(self, a, k, s, d):
        return self._cast(a[0], s, d)

    op_float = op_half = op_bfloat16 = op_double = op_type

    def op_scale(self, a, k, s, d):
        sc, bias, after = k.get("scale", a[1] if len(a) > 1 else 1.0), k.get("bias", a[2] if len(a) > 2 else 0.0), k.get("bias_after_scale", a[3] if len(a) > 3 else True)
        if not (_is_num(sc) and _is_num(bias)):
            raise Unsupported("tensor scale")
        x = a[0]
        if after:
            return self.g.ew("add", [self.g.ew("mul", [x, float(sc)], s, d), float(bias)], s, d)
        return self.g.ew("mul", [self.g.ew("add", [x, float(bias)], s, d), float(sc)], s, d)

    # -- reductions over the last axis
    def _last_axis(self, x, dim):
        nd = len(x.shape)
        if isinstance(dim, (list, tuple)) and len(dim) == 1:
            dim = dim[0]
        if not isinstance(dim, int) or isinstance(dim, bool) or nd == 0 or dim % nd != nd - 1:
            raise Unsupported("reduction is not over the last axis")

    def _reduce(self, op, a, k, s, d):
        x = a[0]
        dim = k.get("dim", k.get("axis", a[1] if len(a) > 1 else None))
        keep = k.get("keepdim", k.get("keepdims", a[2] if len(a) > 2 else False))
        if k.get("dtype") is not None and _DT_NAME.get(k["dtype"], None) != d:
            raise Unsupported("reduction dtype")
        if not isinstance(x, Node):
            raise Unsupported("reduction input")
        if x.dtype not in FLOATS:
            raise Unsupported("integer reduction")
        nd = len(x.shape)
        everything = dim is None or (isinstance(dim, (list, tuple)) and sorted(i % nd for i in dim) == list(range(nd)) and nd > 1)
        if everything and nd >= 1:
            # reduction to a scalar (the tail of a loss): the kernel reduces every row, the handful of per-row partials is finished by one
            # small library reduction on the way out (`post`); nothing inside the group may read the value
            if keep or tuple(s) != () or x.shape != self.g.full:
                raise Unsupported("full reduction with keepdim / outside the domain")
            n = self.g.reduce(op, x, False, d)
            count = 1
            for e in x.shape:
                count *= e
            n.attrs = {"post": op}
            return n, count
        dims = [dim] if isinstance(dim, int) and not isinstance(dim, bool) else list(dim) if isinstance(dim, (list, tuple)) else None
        if dims and all(isinstance(i, int) for i in dims) and (nd - 1) not in [i % nd for i in dims]:
            if op != "sum":
                raise Unsupported("only sums reduce over leading axes")
            if x.shape != self.g.full:
                raise Unsupported("reduction outside the domain")
            count = 1
            for i in set(i % nd for i in dims):
                count *= x.shape[i]
            return self.g.creduce(x, dims, bool(keep), d), count
        self._last_axis(x, dim)
        return self.g.reduce(op, x, bool(keep), d), x.shape[-1]

    def op_sum(self, a, k, s, d):
        return self._reduce("sum", a, k, s, d)[0]

    def op_amax(self, a, k, s, d):
        return self._reduce("max", a, k, s, d)[0]

    def op_amin(self, a, k, s, d):
        return self._reduce("min", a, k, s, d)[0]

    def op_mean(self, a, k, s, d):
        r, n = self._reduce("sum", a, k, s, d)
        if r.attrs.get("post"):
            r.attrs = {"post": "sum", "scale": 1.0 / n}
            return r
        if r.kind == "creduce":
            r.attrs["scale"] = 1.0 / n
            return r
        return self.g.ew("mul", [r, 1.0 / n], r.shape, d)

    # -- composites over the last axis
    def _softmax_parts(self, a, k, s, d):
        x = a[0]
        dim = k.get("dim", k.get("axis", a[1] if len(a) > 1 else None))
        if k.get("dtype") is not None and _DT_NAME.get(k["dtype"]) != d:
            raise Unsupported("softmax dtype")
        self._last_axis(x, dim)
        if x.shape != self.g.full or d not in FLOATS:
            raise Unsupported("softmax domain")
        rs = self.g.full[:-1] + (1,)
        m = self.g.reduce("max", x, True, d)
        z = self.g.ew("sub", [x, m], s, d)
        e = self.g.ew("exp", [z], s, d)
        return z, e, self.g.reduce("sum", e, True, d), rs

    def op_softmax(self, a, k, s, d):
        z, e, den, rs = self._softmax_parts(a, k, s, d)
        return self.g.ew("mul", [e, self.g.ew("reciprocal", [den], rs, d)], s, d)

    def op_log_softmax(self, a, k, s, d):
        z, e, den, rs = self._softmax_parts(a, k, s, d)
        return self.g.ew("sub", [z, self.g.ew("log", [den], rs, d)], s, d)


def _shape_of(x):
    return x.shape if isinstance(x, Node) else ()


def dtype_name(t):
    return _DT_NAME.get(t)
