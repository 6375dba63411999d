"""paddle.pir: the SSA program IR and its pass manager.

Parity: paddle/pir (Program / Operation / Value, IrPrinter, PassManager, pattern rewriter), paddle/fluid/pir/transforms (dead_code_elimination,
common_subexpression_elimination, constant_folding, identity_op_clean, fused gemm-epilogue / fused add-norm patterns, inplace pass) and
python/paddle/pir (`paddle.pir.Program`, `paddle.pir.PassManager`, `translate_to_pir`).

The IR, the printer / parser, the passes and the declarative pattern rewriter are native (csrc/runtime/ir.cpp, bound as `_C.IrProgram` /
`_C.IrPassManager`).  This module is the front end: it translates a recorded `static.Program` (the tape of torch-level calls) into the IR,
runs a pass pipeline and lowers the result back into a Program the Executor replays - fewer nodes (DCE / CSE / identity removal /
constant folding) and fused kernels (`matmul + add -> fused_linear`, `silu * mul -> swiglu`, ...) instead of the recorded sequence.
"""
from __future__ import annotations

import torch

from .._build import load as _load

__all__ = ["Program", "PassManager", "translate_to_pir", "optimize", "default_passes", "DEFAULT_PATTERNS", "core_available", "Value", "Operation", "Block", "global_block", "parse",
           "register_op_impl", "lower"]


def _native():
    m = _load()
    if m is None or not hasattr(m, "IrProgram"):
        raise RuntimeError("paddle_b200.pir needs the native extension (paddle_b200._C): build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    return m


def core_available():
    m = _load()
    return m is not None and hasattr(m, "IrProgram")


def Program():
    """An empty IR program (`_C.IrProgram`): add_input / add_param / add_op / set_outputs / verify / ops / str() / parse()."""
    return _native().IrProgram()


def parse(text):
    return _native().IrProgram.parse(text)


# ------------------------------------------------------------------------------------------------ object views (paddle.pir Python API)
class Value:
    """View of one SSA value of a program (`paddle.pir.Value`): shape / dtype / defining op / users, replace_all_uses_with."""

    __slots__ = ("_p", "id")

    def __init__(self, program, vid):
        self._p, self.id = program, int(vid)

    def _info(self):
        return self._p.value_info(self.id)

    @property
    def shape(self):
        return list(self._info()["type"][1])

    @property
    def dtype(self):
        return self._info()["type"][0]

    @property
    def name(self):
        return self._info()["name"] or f"%{self.id}"

    def is_block_argument(self):
        return self._info()["def_op"] < 0

    def get_defining_op(self):
        d = self._info()["def_op"]
        return None if d < 0 else Operation(self._p, d)

    def all_used_ops(self):
        return [Operation(self._p, o["id"]) for o in self._p.ops() if self.id in o["operands"]]

    def use_count(self):
        return self._p.use_counts()[self.id]

    def use_empty(self):
        return self.use_count() == 0

    def replace_all_uses_with(self, other):
        self._p.replace_all_uses(self.id, other.id if isinstance(other, Value) else int(other))

    def __eq__(self, o):
        return isinstance(o, Value) and o._p is self._p and o.id == self.id

    def __hash__(self):
        return hash((id(self._p), self.id))

    def __repr__(self):
        t = self._info()["type"]
        return f"Value(%{self.id}: tensor<{'x'.join(str(d) for d in t[1])}{'x' if t[1] else ''}{t[0]}>)"


class Operation:
    """View of one operation (`paddle.pir.Operation`)."""

    __slots__ = ("_p", "id")

    def __init__(self, program, op_id):
        self._p, self.id = program, int(op_id)

    def _rec(self):
        for o in self._p.ops():
            if o["id"] == self.id:
                return o
        raise RuntimeError(f"operation {self.id} was erased")

    def name(self):
        return self._rec()["name"]

    def attrs(self):
        return dict(self._rec()["attrs"])

    def num_operands(self):
        return len(self._rec()["operands"])

    def num_results(self):
        return len(self._rec()["results"])

    def operands_source(self):
        return [Value(self._p, v) for v in self._rec()["operands"]]

    def operand_source(self, i):
        return Value(self._p, self._rec()["operands"][i])

    def results(self):
        return [Value(self._p, v) for v in self._rec()["results"]]

    def result(self, i):
        return Value(self._p, self._rec()["results"][i])

    def num_regions(self):
        return self._rec()["num_regions"]

    def blocks(self):
        return [Block(self._p.region(self.id, k)) for k in range(self.num_regions())]

    def erase(self):
        self._p.erase_op(self.id)

    def __repr__(self):
        r = self._rec()
        return f"Operation({r['name']}, operands={r['operands']}, results={r['results']})"


class Block:
    """`program.global_block()`: the ordered operations and the block arguments (inputs / parameters)."""

    def __init__(self, program):
        self.program = program

    @property
    def ops(self):
        return [Operation(self.program, o["id"]) for o in self.program.ops()]

    def args(self):
        return [Value(self.program, a[0]) for a in self.program.args()]

    def kwargs(self):
        return {a[2]: Value(self.program, a[0]) for a in self.program.args()}

    def __len__(self):
        return self.program.num_ops()

    def __iter__(self):
        return iter(self.ops)


def global_block(program):
    """Block view of an IR program (the native object has no Python subclass: `pir.global_block(p)` plays `p.global_block()`)."""
    return Block(program)


# (name, source ops, result ops): op = (op name, [input symbols], [output symbols], {attribute constraints / "$sym.attr" copies})
DEFAULT_PATTERNS = [
    ("fuse_matmul_add", [("matmul", ["x", "w"], ["t"]), ("add", ["t", "b"], ["y"])], [("fused_linear", ["x", "w", "b"], ["y"])]),
    ("fuse_linear_act_gelu", [("fused_linear", ["x", "w", "b"], ["t"]), ("gelu", ["t"], ["y"])], [("fused_linear", ["x", "w", "b"], ["y"], {"activation": "gelu"})]),
    ("fuse_linear_act_relu", [("fused_linear", ["x", "w", "b"], ["t"]), ("relu", ["t"], ["y"])], [("fused_linear", ["x", "w", "b"], ["y"], {"activation": "relu"})]),
    ("fuse_swiglu", [("silu", ["a"], ["t"]), ("mul", ["t", "b"], ["y"])], [("swiglu", ["a", "b"], ["y"])]),
    ("fuse_add_rms_norm", [("add", ["x", "r"], ["h"]), ("rms_norm", ["h", "g"], ["y"])], [("fused_add_rms_norm", ["x", "r", "g"], ["y"], {"eps": "$y.eps"})]),
]


def default_passes():
    # dead code goes first: a dead consumer would otherwise count as a second use and block a fusion
    return ["identity_elim", "cse", "constant_fold", "dce"] + [p[0] for p in DEFAULT_PATTERNS] + ["dce", "inplace"]


class PassManager:
    """Named passes in order.  Built in: dce, cse, identity_elim, constant_fold, inplace, compact; every registered rewrite pattern is a
    pass of its own name.  `run(program)` returns one record per pass ({pass, ops_before, ops_after, changed})."""

    def __init__(self, passes=None, patterns=None, opt_level=2):
        self._pm = _native().IrPassManager([])
        self._names = []
        for name, src, res in (DEFAULT_PATTERNS if patterns is None else patterns):
            self._pm.add_pattern(name, list(src), list(res))
        for p in (default_passes() if passes is None else passes):
            self.add_pass(p)
        self._consts = []
        self._pm.set_folder(self._fold)

    def add_pass(self, name):
        self._pm.add_pass(name)
        self._names.append(name)
        return self

    def add_pattern(self, name, source, result, as_pass=True):
        self._pm.add_pattern(name, list(source), list(result))
        if as_pass:
            self.add_pass(name)
        return self

    def passes(self):
        return list(self._names)

    def enable_ir_printing(self, on=True):
        self._pm.enable_ir_printing(on)

    # constants live on the Python side (tensors); the IR refers to them by index
    def add_constant(self, value):
        self._consts.append(value)
        return len(self._consts) - 1

    def constant(self, idx):
        return self._consts[idx]

    def _fold(self, op, const_ids, attrs):
        fn = _FOLDABLE.get(op)
        if fn is None:
            return None
        try:
            out = fn(*[self._consts[i] for i in const_ids], attrs)
        except Exception:  # noqa: BLE001  (a fold that fails leaves the op in place)
            return None
        return self.add_constant(out)

    def run(self, program):
        return list(self._pm.run(program))


def _t(x):
    return x.as_subclass(torch.Tensor) if isinstance(x, torch.Tensor) and type(x) is not torch.Tensor else x


_FOLDABLE = {
    "add": lambda a, b, at: _t(a) + _t(b), "sub": lambda a, b, at: _t(a) - _t(b), "mul": lambda a, b, at: _t(a) * _t(b), "div": lambda a, b, at: _t(a) / _t(b),
    "neg": lambda a, at: -_t(a), "exp": lambda a, at: torch.exp(_t(a)), "sqrt": lambda a, at: torch.sqrt(_t(a)), "rsqrt": lambda a, at: torch.rsqrt(_t(a)),
    "matmul": lambda a, b, at: torch.matmul(_t(a), _t(b)), "transpose": lambda a, at: _t(a).permute(*at["perm"]) if "perm" in at else None,
}


# ------------------------------------------------------------------------------------------------ static.Program <-> IR
class _Slot:
    __slots__ = ("k",)

    def __init__(self, k):
        self.k = k


def _fname(fn):
    return (getattr(fn, "__name__", None) or str(fn)).strip("_")


_DT = {torch.float32: "float32", torch.float64: "float64", torch.float16: "float16", torch.bfloat16: "bfloat16", torch.int64: "int64", torch.int32: "int32",
       torch.int16: "int16", torch.int8: "int8", torch.uint8: "uint8", torch.bool: "bool", torch.complex64: "complex64", torch.complex128: "complex128"}


def _ty(t):
    return (_DT.get(t.dtype, str(t.dtype).replace("torch.", "")), [int(d) for d in t.shape])


class Translation:
    """Result of `translate_to_pir`: the IR program plus the tables that tie it back to the recorded one."""

    def __init__(self, ir, templates, vid_of, params, source):
        self.ir, self.templates, self.vid_of, self.params, self.source = ir, templates, vid_of, params, source


def translate_to_pir(program, fetch_vids=None):
    """Record -> IR.  Program values become IR values (types from the recorded example tensors), tensors captured by reference (parameters,
    constants) become `param` block arguments, every other argument becomes an attribute (so CSE keys on it)."""
    from ..static import _Ref

    ir = Program()
    vid_of, ir_of = {}, {}           # ir value -> program vid, program vid -> ir value
    params, param_ir = {}, {}        # ir value -> tensor, id(tensor) -> ir value
    for name, vid in program.placeholders.items():
        v = ir.add_input(name, *_ty(program._keep[vid]))
        vid_of[v], ir_of[vid] = vid, v
    templates = {}
    for idx, node in enumerate(program.nodes):
        operands = []

        def enc(x):
            if isinstance(x, _Ref):
                if x.vid not in ir_of:           # produced outside the tape (e.g. created by a control node): opaque input
                    v = ir.add_input(f"v{x.vid}", *_ty(program._keep[x.vid]))
                    vid_of[v], ir_of[x.vid] = x.vid, v
                operands.append(ir_of[x.vid])
                return _Slot(len(operands) - 1)
            if isinstance(x, torch.Tensor):
                k = id(x)
                if k not in param_ir:
                    v = ir.add_param(getattr(x, "name", None) or f"p{len(param_ir)}", *_ty(x))
                    param_ir[k], params[v] = v, x
                operands.append(param_ir[k])
                return _Slot(len(operands) - 1)
            if isinstance(x, (list, tuple)):
                return type(x)(enc(i) for i in x)
            if isinstance(x, dict):
                return {k: enc(v) for k, v in x.items()}
            return x

        if node.kind != "op":
            refs = []

            def collect(x):
                if isinstance(x, _Ref):
                    refs.append(x.vid)
                elif isinstance(x, (list, tuple)):
                    for i in x:
                        collect(i)
                elif isinstance(x, dict):
                    for i in x.values():
                        collect(i)

            collect(node.args)
            collect(node.kwargs)
            res = ir.add_op("pd_op.train_step" if node.kind == "train" else "pd_op.py_func", [ir_of[v] for v in refs if v in ir_of], {"side_effect": True, "node": idx},
                            [_ty(program._keep[v]) for v in node.outs])
            for v, r in zip(node.outs, res):
                vid_of[r], ir_of[v] = v, r
            templates[_last_op_id(ir)] = ("node", node, [ir_of[v] for v in refs if v in ir_of])
            continue
        t_args, t_kwargs = enc(node.args), enc(node.kwargs)
        attrs = {}
        flat_attr(attrs, "a", t_args)
        flat_attr(attrs, "k", t_kwargs)
        name = _fname(node.fn)
        if name in ("transpose", "permute"):
            perm = _perm_of(name, node, program)
            if perm is not None:
                attrs["perm"] = perm
        if name == "rms_norm" and "eps" not in attrs:
            e = node.kwargs.get("eps", node.kwargs.get("epsilon"))
            if isinstance(e, float):
                attrs["eps"] = e
        res = ir.add_op("pd_op." + name, operands, attrs, [_ty(program._keep[v]) for v in node.outs])
        for v, r in zip(node.outs, res):
            vid_of[r], ir_of[v] = v, r
        templates[_last_op_id(ir)] = ("call", node.fn, t_args, t_kwargs, node.kind)
    outs = [ir_of[v] for v in (fetch_vids if fetch_vids is not None else _default_fetch(program)) if v in ir_of]
    ir.set_outputs(outs)
    ir.verify()
    return Translation(ir, templates, vid_of, params, program)


def _last_op_id(ir):
    return ir.ops()[-1]["id"]


def _default_fetch(program):
    used = set()
    from ..static import _Ref

    def collect(x):
        if isinstance(x, _Ref):
            used.add(x.vid)
        elif isinstance(x, (list, tuple)):
            for i in x:
                collect(i)
        elif isinstance(x, dict):
            for i in x.values():
                collect(i)

    for n in program.nodes:
        collect(n.args)
        collect(n.kwargs)
    return [v for n in program.nodes for v in n.outs if v not in used]     # values nobody consumes are the program's results


def flat_attr(attrs, prefix, obj):
    """Non-tensor arguments as attributes (CSE must distinguish `sum(x, 0)` from `sum(x, 1)`); anything exotic travels as repr()."""
    items = enumerate(obj) if isinstance(obj, (list, tuple)) else obj.items()
    for k, v in items:
        if isinstance(v, _Slot):
            continue
        key = f"{prefix}{k}"
        if isinstance(v, bool) or isinstance(v, (int, float, str)):
            attrs[key] = v
        elif isinstance(v, (list, tuple)) and v and all(isinstance(i, int) and not isinstance(i, bool) for i in v):
            attrs[key] = list(v)
        elif isinstance(v, (list, tuple, dict)) and _has_slot(v):
            attrs[key] = "slots:" + _shape_repr(v)
        else:
            attrs[key] = repr(v)


def _has_slot(v):
    if isinstance(v, _Slot):
        return True
    if isinstance(v, (list, tuple)):
        return any(_has_slot(i) for i in v)
    if isinstance(v, dict):
        return any(_has_slot(i) for i in v.values())
    return False


def _shape_repr(v):
    if isinstance(v, _Slot):
        return f"%{v.k}"
    if isinstance(v, (list, tuple)):
        return "[" + ",".join(_shape_repr(i) for i in v) + "]"
    if isinstance(v, dict):
        return "{" + ",".join(f"{k}:{_shape_repr(i)}" for k, i in v.items()) + "}"
    return repr(v)


def _perm_of(name, node, program):
    from ..static import _Ref

    nd = None
    for a in node.args:
        if isinstance(a, _Ref):
            nd = program._keep[a.vid].dim()
            break
    if nd is None:
        return None
    rest = [a for a in node.args[1:]]
    if name == "transpose" and len(rest) == 2 and all(isinstance(i, int) for i in rest):     # torch.transpose(x, d0, d1)
        p = list(range(nd))
        d0, d1 = rest[0] % nd, rest[1] % nd
        p[d0], p[d1] = p[d1], p[d0]
        return p
    perm = node.kwargs.get("perm", node.kwargs.get("dims"))
    if perm is None and rest:
        perm = rest[0] if isinstance(rest[0], (list, tuple)) else rest
    if perm is not None and len(perm) == nd and all(isinstance(i, int) for i in perm):
        return [int(i) % nd for i in perm]
    return None


def _identity(x):
    return x


# implementations of the ops only rewrite patterns create
def _fused_linear(x, w, b, activation=None):
    from ..nn import functional as F

    y = F.linear(x, w, b)
    if activation == "gelu":
        return F.gelu(y)
    if activation == "relu":
        return F.relu(y)
    return y


def _swiglu(a, b):
    from ..incubate.nn.functional import swiglu

    return swiglu(a, b)


def _fused_add_rms_norm(x, r, g, eps=1e-6):
    from ..incubate.nn.functional import fused_rms_norm

    h = x + r
    out = fused_rms_norm(h, g, None, eps, h.dim() - 1)
    return out[0] if isinstance(out, (tuple, list)) else out


_IMPL = {"fused_linear": _fused_linear, "swiglu": _swiglu, "fused_add_rms_norm": _fused_add_rms_norm}


def register_op_impl(name, fn):
    """Python implementation of an op that a custom rewrite pattern introduces (called with the operands, attributes as keywords)."""
    _IMPL[name] = fn


def lower(tr, pm=None):
    """IR -> a new `static.Program` for the Executor.  Untouched ops are rebuilt from their recorded call (operands re-wired), ops created by
    patterns call their registered implementation, folded constants become captured tensors."""
    from ..static import Program as SProgram, _Node, _Ref

    src, ir = tr.source, tr.ir
    out = SProgram()
    out.placeholders = dict(src.placeholders)
    out._name2vid = dict(src._name2vid)
    out._keep = list(src._keep)
    out._next = src._next
    out.random_seed = src.random_seed
    vid_of = dict(tr.vid_of)
    repl = dict(ir.replacements())

    def origin_vid(v):
        # a value a rewrite created: give it the program slot of the value whose uses it took over (fetch targets keep working)
        for old, new in repl.items():
            if new == v and old in vid_of:
                return vid_of[old]
        return None

    def vid(v):
        if v not in vid_of:
            o = origin_vid(v)
            if o is None:
                o = out._next
                out._next += 1
                out._keep.append(None)
            vid_of[v] = o
        return vid_of[v]

    def fill(t, operands):
        if isinstance(t, _Slot):
            v = operands[t.k]
            return tr.params[v] if v in tr.params else _Ref(vid(v))
        if isinstance(t, (list, tuple)):
            return type(t)(fill(i, operands) for i in t)
        if isinstance(t, dict):
            return {k: fill(i, operands) for k, i in t.items()}
        return t

    for op in ir.ops():
        tpl = tr.templates.get(op["id"])
        operands = op["operands"]
        if tpl is not None and tpl[0] == "node":
            # an opaque node (run-time control flow) reads the value table by the ORIGINAL slots: when a rewrite moved one of its inputs to
            # another slot (CSE / a fusion), bind the old slot to the survivor first
            node = tpl[1]
            for old_ir, cur_ir in zip(tpl[2], operands):
                if old_ir != cur_ir and old_ir in tr.vid_of:
                    out.nodes.append(_Node(_identity, (_Ref(vid(cur_ir)),), {}, [tr.vid_of[old_ir]]))
            out.nodes.append(node)
            continue
        if tpl is not None:
            _, fn, t_args, t_kwargs, kind = tpl
            out.nodes.append(_Node(fn, fill(t_args, operands), fill(t_kwargs, operands), [vid(r) for r in op["results"]], kind))
            continue
        name = op["name"].split(".", 1)[-1]
        if name == "constant":
            const = pm.constant(op["attrs"]["const_id"]) if pm is not None else None
            out.nodes.append(_Node((lambda c=const: c), (), {}, [vid(r) for r in op["results"]]))
            continue
        impl = _IMPL.get(name)
        if impl is None:
            raise NotImplementedError(f"pir.lower: op '{name}' was introduced by a rewrite but has no implementation (pir.register_op_impl)")
        args = tuple(tr.params[v] if v in tr.params else _Ref(vid(v)) for v in operands)
        kw = {k: v for k, v in op["attrs"].items() if k not in ("inplace",)}
        out.nodes.append(_Node(impl, args, kw, [vid(r) for r in op["results"]]))
    # fetch targets that were merged away (CSE) or replaced (patterns) point at their survivor
    alias = {}
    for old, new in repl.items():
        if old in tr.vid_of:
            v = new
            while v in repl and repl[v] != v:
                v = repl[v]
            alias[tr.vid_of[old]] = vid(v)
    out._fetch_alias = {k: alias.get(v, v) for k, v in src._fetch_alias.items()}
    out._name2vid = {k: alias.get(v, v) for k, v in out._name2vid.items()}
    out.__dict__["_pir_alias"] = alias
    return out


def conv_bn_fuse(tr):
    """Inference: `batch_norm(conv(x, W, b), mean, var, gamma, beta)` with constant statistics becomes `conv(x, W', b')`,
    W' = W * gamma / sqrt(var + eps) per output channel, b' = (b - mean) * gamma / sqrt(var + eps) + beta.  Works on the Translation (it creates
    new parameter values).  Parity: conv_bn_fuse_pass of the reference's inference pipeline (paddle/fluid/framework/ir/conv_bn_fuse_pass.cc).
    Returns the number of fused pairs."""
    ir = tr.ir
    fused = 0
    uses = ir.use_counts()
    by_result = {}
    for op in ir.ops():
        for r in op["results"]:
            by_result[r] = op
    for bn in ir.ops():
        if bn["name"] != "pd_op.batch_norm":
            continue
        tb = tr.templates.get(bn["id"])
        if tb is None or tb[0] != "call":
            continue
        args, kw = list(tb[2]), dict(tb[3])

        def arg(i, key, default=None):
            return kw[key] if key in kw else (args[i] if len(args) > i else default)

        xs, mean_s, var_s, w_s, b_s = arg(0, "input"), arg(1, "running_mean"), arg(2, "running_var"), arg(3, "weight"), arg(4, "bias")
        training, eps = arg(5, "training", False), arg(7, "eps", 1e-5)
        if training or not all(isinstance(v, _Slot) for v in (xs, mean_s, var_s)) or not isinstance(eps, float):
            continue
        val = lambda slot: bn["operands"][slot.k]      # noqa: E731
        conv = by_result.get(val(xs))
        if conv is None or conv["name"] not in ("pd_op.conv2d", "pd_op.conv1d", "pd_op.conv3d") or uses[val(xs)] != 1 or val(xs) in ir.outputs():
            continue
        tc = tr.templates.get(conv["id"])
        if tc is None or tc[0] != "call" or tc[3] or len(tc[2]) < 2 or not isinstance(tc[2][0], _Slot) or not isinstance(tc[2][1], _Slot):
            continue
        cargs = list(tc[2])
        wv = conv["operands"][cargs[1].k]
        bias_slot = cargs[2] if len(cargs) > 2 else None
        if wv not in tr.params or (isinstance(bias_slot, _Slot) and conv["operands"][bias_slot.k] not in tr.params) or (bias_slot is not None and not isinstance(bias_slot, _Slot)):
            continue
        stats = [tr.params.get(val(s_)) if isinstance(s_, _Slot) else None for s_ in (mean_s, var_s, w_s, b_s)]
        if stats[0] is None or stats[1] is None or (isinstance(w_s, _Slot) and stats[2] is None) or (isinstance(b_s, _Slot) and stats[3] is None):
            continue
        W = _t(tr.params[wv]).detach()
        mean, var = _t(stats[0]).detach().to(W.dtype), _t(stats[1]).detach().to(W.dtype)
        gamma = _t(stats[2]).detach().to(W.dtype) if stats[2] is not None else torch.ones_like(mean)
        beta = _t(stats[3]).detach().to(W.dtype) if stats[3] is not None else torch.zeros_like(mean)
        b0 = _t(tr.params[conv["operands"][bias_slot.k]]).detach() if isinstance(bias_slot, _Slot) else torch.zeros_like(mean)
        scale = gamma * torch.rsqrt(var + eps)
        W2 = (W * scale.reshape(-1, *([1] * (W.dim() - 1)))).contiguous()
        b2 = ((b0 - mean) * scale + beta).contiguous()
        wn = ir.add_param(f"conv_bn_w{fused}", *_ty(W2))
        bnew = ir.add_param(f"conv_bn_b{fused}", *_ty(b2))
        tr.params[wn], tr.params[bnew] = W2, b2
        ir.set_insertion_point_after(bn["id"])
        new = ir.add_op(conv["name"], [conv["operands"][cargs[0].k], wn, bnew], dict(conv["attrs"]), [ir.value_type(bn["results"][0])])
        ir.reset_insertion_point()
        rest = tuple(cargs[3:])
        tr.templates[_op_of(ir, new[0])] = ("call", tc[1], (_Slot(0), _Slot(1), _Slot(2)) + rest, {}, "op")
        ir.replace_all_uses(bn["results"][0], new[0])
        ir.erase_op(bn["id"])
        ir.erase_op(conv["id"])
        uses = ir.use_counts()
        fused += 1
    if fused:
        ir.verify()
    return fused


def _op_of(ir, value):
    return ir.value_info(value)["def_op"]


TRANSLATION_PASSES = {"conv_bn_fuse": conv_bn_fuse}


def optimize(program, fetch_list=None, passes=None, patterns=None, return_report=False, cinn=None):
    """Run a pass pipeline over a recorded static Program and return the optimised Program (same feeds, same fetch targets).
    `cinn`: True / dict of `cinn.fuse` options - after the passes, fusible elementwise / reduction chains become generated kernels
    (default: FLAGS_use_cinn)."""
    fetch_vids = None
    if fetch_list is not None:
        fetch_vids = []
        for f in fetch_list:
            v = program._name2vid.get(f) if isinstance(f, str) else program._fetch_alias.get(id(f))
            if v is not None:
                fetch_vids.append(v)
        for n in program.nodes:          # side-effect nodes keep their inputs alive through is_pure(); nothing else to add
            pass
    tr = translate_to_pir(program, fetch_vids)
    early = []
    if passes is not None:                       # passes that need the Translation (they create parameters) run first; the rest is native
        for name in [p for p in passes if p in TRANSLATION_PASSES]:
            before = tr.ir.num_ops()
            early.append({"pass": name, "ops_before": before, "changed": TRANSLATION_PASSES[name](tr), "ops_after": tr.ir.num_ops()})
        passes = [p for p in passes if p not in TRANSLATION_PASSES]
    pm = PassManager(passes, patterns)
    report = early + list(pm.run(tr.ir))
    if cinn is None:
        from ..framework.flags import flag

        cinn = bool(flag("FLAGS_use_cinn", False))
    if cinn:
        from ..cinn import fuse as _cinn_fuse

        before = tr.ir.num_ops()
        fr = _cinn_fuse(tr, **(cinn if isinstance(cinn, dict) else {}))
        report = list(report) + [{"pass": "cinn_fusion", "ops_before": before, "ops_after": tr.ir.num_ops(), "changed": len(fr.groups)}]
    new = lower(tr, pm)
    new.__dict__["_pir_report"] = report
    if cinn:
        new.__dict__["_cinn_report"] = fr
    return (new, report) if return_report else new
