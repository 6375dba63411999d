"""Self-contained ONNX writer: traced Program -> ModelProto bytes, no `onnx` / `paddle2onnx` package needed.
Parity (role): python/paddle/onnx/export.py + the paddle2onnx converter the reference delegates to.

The protobuf wire format is emitted by hand (ModelProto / GraphProto / NodeProto / AttributeProto / TensorProto / ValueInfoProto are a
dozen fields).  The graph comes from `jit._trace_program`: one node per recorded functional op, constants (weights) become
initializers."""
from __future__ import annotations

import struct

import numpy as np
import torch

# ---- protobuf wire helpers ------------------------------------------------------------------------------------------------


def _varint(n):
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _key(field, wire):
    return _varint((field << 3) | wire)


def _f_varint(field, v):
    return _key(field, 0) + _varint(int(v))


def _f_bytes(field, b):
    b = b.encode() if isinstance(b, str) else bytes(b)
    return _key(field, 2) + _varint(len(b)) + b


def _f_float(field, v):
    return _key(field, 5) + struct.pack("<f", float(v))


_DTYPES = {torch.float32: 1, torch.uint8: 2, torch.int8: 3, torch.int16: 5, torch.int32: 6, torch.int64: 7, torch.bool: 9, torch.float16: 10,
           torch.float64: 11, torch.bfloat16: 16}


def _tensor_proto(name, t):
    t = t.detach().cpu().contiguous()
    raw = t.view(torch.int16).numpy().tobytes() if t.dtype == torch.bfloat16 else t.numpy().tobytes()
    out = b"".join(_f_varint(1, d) for d in t.shape) + _f_varint(2, _DTYPES[t.dtype]) + _f_bytes(8, name) + _f_bytes(9, raw)
    return out


def _attr(name, v):
    body = _f_bytes(1, name)
    if isinstance(v, bool) or isinstance(v, (int, np.integer)):
        body += _f_varint(3, int(v)) + _f_varint(20, 2)
    elif isinstance(v, float):
        body += _f_float(2, v) + _f_varint(20, 1)
    elif isinstance(v, str):
        body += _f_bytes(4, v) + _f_varint(20, 3)
    elif isinstance(v, torch.Tensor):
        body += _f_bytes(5, _tensor_proto("", v)) + _f_varint(20, 4)
    elif isinstance(v, (list, tuple)) and all(isinstance(i, (int, np.integer)) for i in v):
        body += b"".join(_f_varint(8, int(i)) for i in v) + _f_varint(20, 7)
    elif isinstance(v, (list, tuple)):
        body += b"".join(_f_float(7, float(i)) for i in v) + _f_varint(20, 6)
    else:
        raise TypeError(f"onnx attribute {name}: unsupported value {v!r}")
    return body


def _value_info(name, dtype, shape):
    dims = b""
    for i, d in enumerate(shape):
        dims += _f_bytes(1, _f_varint(1, d) if (d is not None and d >= 0) else _f_bytes(2, f"dyn_{i}"))
    tensor_type = _f_varint(1, _DTYPES[dtype]) + _f_bytes(2, dims)
    return _f_bytes(1, name) + _f_bytes(2, _f_bytes(1, tensor_type))


# ---- graph builder ----------------------------------------------------------------------------------------------------------------


class _Graph:
    def __init__(self):
        self.nodes, self.inits, self._n = [], [], 0

    def fresh(self, hint="t"):
        self._n += 1
        return f"{hint}_{self._n}"

    def const(self, value, dtype=None):
        t = value if isinstance(value, torch.Tensor) else torch.as_tensor(value, dtype=dtype)
        t = t.as_subclass(torch.Tensor)
        name = self.fresh("c")
        self.inits.append(_tensor_proto(name, t))
        return name

    def node(self, op, inputs, outputs=None, **attrs):
        outputs = outputs or [self.fresh(op.lower())]
        body = b"".join(_f_bytes(1, i) for i in inputs) + b"".join(_f_bytes(2, o) for o in outputs) + _f_bytes(3, self.fresh("n")) + _f_bytes(4, op)
        body += b"".join(_f_bytes(5, _attr(k, v)) for k, v in attrs.items() if v is not None)
        self.nodes.append(body)
        return outputs[0]


def _pair(v, n=2):
    return [int(v)] * n if isinstance(v, (int, np.integer)) else [int(i) for i in v]


def _norm_axis(a, rank):
    return a if a >= 0 else a + rank


class _Ctx:
    def __init__(self, g, names, ranks):
        self.g, self.names, self.ranks = g, names, ranks

    def val(self, a, dtype=None):
        from .static import _Ref

        if isinstance(a, _Ref):
            return self.names[a.vid]
        if isinstance(a, torch.Tensor):
            return self.g.const(a)
        return self.g.const(a, dtype=dtype or (torch.float32 if isinstance(a, float) else torch.int64))

    def rank(self, a):
        from .static import _Ref

        return self.ranks.get(a.vid) if isinstance(a, _Ref) else (a.dim() if isinstance(a, torch.Tensor) else 0)


def _binary(op):
    def f(c, a, k):
        # python scalars become float32 constants (the recorded graphs are floating point; integer arithmetic keeps tensors on both sides)
        return c.g.node(op, [c.val(a[0], torch.float32), c.val(a[1], torch.float32)])
    return f


def _unary(op, **attrs):
    return lambda c, a, k: c.g.node(op, [c.val(a[0])], **attrs)


def _conv2d(c, a, k):
    x, w = a[0], a[1]
    b = a[2] if len(a) > 2 else k.get("bias")
    stride, pad, dil, groups = (list(a[3:7]) + [1, 0, 1, 1][len(a[3:7]):]) if len(a) > 3 else (k.get("stride", 1), k.get("padding", 0), k.get("dilation", 1), k.get("groups", 1))
    p = _pair(pad)
    ins = [c.val(x), c.val(w)] + ([c.val(b)] if b is not None else [])
    return c.g.node("Conv", ins, strides=_pair(stride), pads=p + p, dilations=_pair(dil), group=int(groups))


def _batch_norm(c, a, k):
    x, mean, var = a[0], a[1], a[2]
    w, b = k.get("weight", a[3] if len(a) > 3 else None), k.get("bias", a[4] if len(a) > 4 else None)
    if k.get("training", a[5] if len(a) > 5 else False):
        raise ValueError("onnx export: batch_norm must be in eval mode (layer.eval())")
    ch = mean.shape[0]
    w = w if w is not None else torch.ones(ch)
    b = b if b is not None else torch.zeros(ch)
    return c.g.node("BatchNormalization", [c.val(x), c.val(w), c.val(b), c.val(mean), c.val(var)], epsilon=float(k.get("eps", 1e-5)))


def _pool(op):
    def f(c, a, k):
        ks = _pair(a[1] if len(a) > 1 else k["kernel_size"])
        st = k.get("stride", a[2] if len(a) > 2 else None)
        st = ks if st in (None, []) else _pair(st)
        p = _pair(k.get("padding", a[3] if len(a) > 3 else 0))
        attrs = dict(kernel_shape=ks, strides=st, pads=p + p, ceil_mode=int(bool(k.get("ceil_mode", False))))
        if op == "AveragePool":
            attrs["count_include_pad"] = int(bool(k.get("count_include_pad", True)))
        return c.g.node(op, [c.val(a[0])], **attrs)
    return f


def _adaptive_avg(c, a, k):
    out = a[1] if len(a) > 1 else k["output_size"]
    if _pair(out) != [1, 1]:
        raise ValueError("onnx export: adaptive_avg_pool2d is only exportable for output_size 1 (GlobalAveragePool)")
    return c.g.node("GlobalAveragePool", [c.val(a[0])])


def _flatten(c, a, k):
    start = a[1] if len(a) > 1 else k.get("start_dim", 0)
    end = a[2] if len(a) > 2 else k.get("end_dim", -1)
    rank = c.rank(a[0])
    if end in (-1, (rank or 0) - 1) and start >= 0:
        if start == 1:
            return c.g.node("Flatten", [c.val(a[0])], axis=1)
        shape = [0] * start + [-1]
        return c.g.node("Reshape", [c.val(a[0]), c.g.const(shape, torch.int64)])
    raise ValueError("onnx export: flatten with end_dim != -1 is not supported")


def _linear(c, a, k):
    x, w = a[0], a[1]            # paddle layout: w is [in, out]
    b = a[2] if len(a) > 2 else k.get("bias")
    y = c.g.node("MatMul", [c.val(x), c.val(w)])
    return c.g.node("Add", [y, c.val(b)]) if b is not None else y


def _matmul(c, a, k):
    return c.g.node("MatMul", [c.val(a[0]), c.val(a[1])])


def _gelu(c, a, k):
    x = c.val(a[0])
    if k.get("approximate", "none") == "tanh":
        inner = c.g.node("Mul", [c.g.node("Add", [x, c.g.node("Mul", [c.g.node("Pow", [x, c.g.const(3.0)]), c.g.const(0.044715)])]), c.g.const(0.7978845608028654)])
        return c.g.node("Mul", [c.g.node("Mul", [x, c.g.const(0.5)]), c.g.node("Add", [c.g.node("Tanh", [inner]), c.g.const(1.0)])])
    erf = c.g.node("Erf", [c.g.node("Div", [x, c.g.const(1.4142135623730951)])])
    return c.g.node("Mul", [c.g.node("Mul", [x, c.g.const(0.5)]), c.g.node("Add", [erf, c.g.const(1.0)])])


def _silu(c, a, k):
    x = c.val(a[0])
    return c.g.node("Mul", [x, c.g.node("Sigmoid", [x])])


def _softmax(op):
    def f(c, a, k):
        dim = k.get("dim", a[1] if len(a) > 1 else -1)
        return c.g.node(op, [c.val(a[0])], axis=int(dim if dim is not None else -1))
    return f


def _layer_norm(c, a, k):
    x, shape = a[0], a[1]
    w = a[2] if len(a) > 2 else k.get("weight")
    b = a[3] if len(a) > 3 else k.get("bias")
    eps = a[4] if len(a) > 4 else k.get("eps", 1e-5)
    n = len(shape)
    w = w if w is not None else torch.ones(list(shape))
    ins = [c.val(x), c.val(w)] + ([c.val(b)] if b is not None else [])
    return c.g.node("LayerNormalization", ins, axis=-n, epsilon=float(eps))


def _reshape(c, a, k):
    shape = a[1] if len(a) == 2 and isinstance(a[1], (list, tuple)) else list(a[1:])
    return c.g.node("Reshape", [c.val(a[0]), c.g.const([int(s) for s in shape], torch.int64)])


def _permute(c, a, k):
    perm = a[1] if len(a) == 2 and isinstance(a[1], (list, tuple)) else list(a[1:])
    return c.g.node("Transpose", [c.val(a[0])], perm=[int(p) for p in perm])


def _transpose(c, a, k):
    rank = c.rank(a[0])
    d0, d1 = _norm_axis(int(a[1]), rank), _norm_axis(int(a[2]), rank)
    perm = list(range(rank))
    perm[d0], perm[d1] = perm[d1], perm[d0]
    return c.g.node("Transpose", [c.val(a[0])], perm=perm)


def _cat(c, a, k):
    dim = a[1] if len(a) > 1 else k.get("dim", 0)
    return c.g.node("Concat", [c.val(t) for t in a[0]], axis=int(dim))


def _embedding(c, a, k):
    ids, w = a[0], a[1]
    return c.g.node("Gather", [c.val(w), c.val(ids)], axis=0)


def _reduce(op):
    def f(c, a, k):
        dim = a[1] if len(a) > 1 else k.get("dim")
        keep = int(bool(k.get("keepdim", a[2] if len(a) > 2 else False)))
        if dim is None:
            return c.g.node(op, [c.val(a[0])], keepdims=keep)
        axes = _pair(dim, 1)
        if op == "ReduceSum":      # axes are an input since opset 13
            return c.g.node(op, [c.val(a[0]), c.g.const(axes, torch.int64)], keepdims=keep)
        return c.g.node(op, [c.val(a[0])], axes=axes, keepdims=keep)
    return f


def _clip(c, a, k):
    lo = k.get("min", a[1] if len(a) > 1 else None)
    hi = k.get("max", a[2] if len(a) > 2 else None)
    ins = [c.val(a[0]), c.g.const(float(lo)) if lo is not None else "", c.g.const(float(hi)) if hi is not None else ""]
    return c.g.node("Clip", ins)


def _unsqueeze(op):
    return lambda c, a, k: c.g.node(op, [c.val(a[0]), c.g.const([int(a[1] if len(a) > 1 else k["dim"])], torch.int64)])


_CONVERTERS = {
    "conv2d": _conv2d, "batch_norm": _batch_norm, "relu": _unary("Relu"), "sigmoid": _unary("Sigmoid"), "tanh": _unary("Tanh"), "gelu": _gelu,
    "silu": _silu, "softmax": _softmax("Softmax"), "log_softmax": _softmax("LogSoftmax"), "max_pool2d": _pool("MaxPool"), "avg_pool2d": _pool("AveragePool"),
    "adaptive_avg_pool2d": _adaptive_avg, "flatten": _flatten, "linear": _linear, "matmul": _matmul, "mm": _matmul, "bmm": _matmul,
    "add": _binary("Add"), "sub": _binary("Sub"), "subtract": _binary("Sub"), "mul": _binary("Mul"), "multiply": _binary("Mul"), "div": _binary("Div"),
    "divide": _binary("Div"), "true_divide": _binary("Div"), "pow": _binary("Pow"), "layer_norm": _layer_norm, "reshape": _reshape, "view": _reshape,
    "permute": _permute, "transpose": _transpose, "cat": _cat, "concat": _cat, "dropout": _unary("Identity"), "embedding": _embedding,
    "mean": _reduce("ReduceMean"), "sum": _reduce("ReduceSum"), "exp": _unary("Exp"), "log": _unary("Log"), "sqrt": _unary("Sqrt"), "neg": _unary("Neg"),
    "abs": _unary("Abs"), "erf": _unary("Erf"), "clamp": _clip, "clip": _clip, "hardtanh": lambda c, a, k: _clip(c, [a[0], k.get("min_val", -1.0), k.get("max_val", 1.0)], {}),
    "leaky_relu": lambda c, a, k: c.g.node("LeakyRelu", [c.val(a[0])], alpha=float(a[1] if len(a) > 1 else k.get("negative_slope", 0.01))),
    "elu": lambda c, a, k: c.g.node("Elu", [c.val(a[0])], alpha=float(a[1] if len(a) > 1 else k.get("alpha", 1.0))),
    "long": _unary("Cast", to=7), "int": _unary("Cast", to=6), "float": _unary("Cast", to=1), "double": _unary("Cast", to=11), "half": _unary("Cast", to=10),
    "bfloat16": _unary("Cast", to=16), "bool": _unary("Cast", to=9),
    "unsqueeze": _unsqueeze("Unsqueeze"), "squeeze": _unsqueeze("Squeeze"), "contiguous": _unary("Identity"), "clone": _unary("Identity"),
}


def supported_ops():
    return sorted(_CONVERTERS)


def export_program(blob, path, opset_version=17, producer="paddle_b200"):
    """`blob` = jit._trace_program(...) result. Writes `path` (ModelProto bytes) and returns the list of ONNX op types emitted."""
    from .static import passes as SP

    prog = blob["program"]
    g = _Graph()
    names, ranks = {}, {}
    inputs = b""
    keep = getattr(prog, "_keep", None) or []
    for name, vid in prog.placeholders.items():
        names[vid] = name
        t = keep[vid] if vid < len(keep) and isinstance(keep[vid], torch.Tensor) else None
        shape = list(torch.Tensor.size(t)) if t is not None else []
        decl = getattr(t, "_decl_shape", None) or shape
        ranks[vid] = len(shape)
        inputs += _f_bytes(11, _value_info(name, t.dtype if t is not None else torch.float32, [d if (d is not None and d >= 0) else -1 for d in decl]))
    emitted = []
    for n in prog.nodes:
        op = SP._fname(n.fn)
        conv = _CONVERTERS.get(op)
        if conv is None:
            raise NotImplementedError(f"onnx export: no converter for op '{op}' (supported: {', '.join(supported_ops())})")
        before = len(g.nodes)
        out_name = conv(_Ctx(g, names, ranks), list(n.args), dict(n.kwargs))
        emitted.extend([None] * (len(g.nodes) - before))
        vid = n.outs[0]
        names[vid] = out_name
        t = keep[vid] if vid < len(keep) and isinstance(keep[vid], torch.Tensor) else None
        ranks[vid] = t.dim() if t is not None else None
    outputs = b""
    for vid in blob["fetch_vids"]:
        t = keep[vid] if vid < len(keep) and isinstance(keep[vid], torch.Tensor) else None
        outputs += _f_bytes(12, _value_info(names[vid], t.dtype if t is not None else torch.float32, list(torch.Tensor.size(t)) if t is not None else []))
    graph = b"".join(_f_bytes(1, nb) for nb in g.nodes) + _f_bytes(2, "paddle_b200_graph") + b"".join(_f_bytes(5, i) for i in g.inits) + inputs + outputs
    model = _f_varint(1, 8) + _f_bytes(2, producer) + _f_bytes(3, "0.1") + _f_bytes(7, graph) + _f_bytes(8, _f_bytes(1, "") + _f_varint(2, int(opset_version)))
    with open(path, "wb") as f:
        f.write(model)
    return len(g.nodes)
