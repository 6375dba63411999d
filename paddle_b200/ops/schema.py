"""Operator schema registry + infer-meta (shape / dtype inference without running a kernel).

Parity: paddle/phi/ops/yaml/ops.yaml (op : / args : / output : / infer_meta : / kernel : / backward : entries), the generated C++ API
and paddle/phi/infermeta/* (L5 of SURVEY.md).  The reference generates its op library FROM the YAML; here the op library is the Python
functions of `paddle_b200.ops`, so the schema is derived from them (signature introspection, `dump_yaml`) and kept next to the code
as `ops.yaml`; `load_yaml` + `validate` check the file against the live functions, which is what keeps a declared contract (argument
names, order, attribute types and defaults, backward pairing) from drifting.

Infer-meta is generic: every op is evaluated on *meta tensors* (torch `device="meta"`: shapes, strides and dtypes propagate, no
storage, no kernel), which gives InferMeta for the whole op library at once and is what `static` uses to answer `Variable.shape`
questions for ops whose output shape is not trivially the input's.  Hand-written rules can override (`register_infer_meta`) for
ops whose meta evaluation is data dependent (nonzero, masked_select, unique -> dynamic dims = -1).
"""
from __future__ import annotations

import inspect
import os
import re
from dataclasses import dataclass, field

import torch

from ..common.ddim import DDim

__all__ = ["ArgSpec", "OpSchema", "MetaTensor", "REGISTRY", "build_registry", "get", "infer_meta", "register_infer_meta", "dump_yaml", "load_yaml",
           "validate", "parse_signature"]

_TENSOR_ARGS = {"x", "y", "input", "label", "index", "weight", "bias", "mask", "condition", "other", "tensor", "src", "value", "values", "indices",
                "updates", "ids", "logits", "labels", "mat1", "mat2", "vec", "a", "b", "q", "k", "v", "query", "key", "positions", "sorted_sequence",
                "boundaries", "repeats", "rois", "boxes", "scores", "grad", "out_grad", "start", "end_t", "step_t", "prob", "arr", "tau", "pivots",
                "lu_data", "source", "offsets", "weights", "min_t", "max_t", "x1", "x2", "input1", "input2", "target", "anchor", "positive", "negative"}
_TENSOR_LIST_ARGS = {"inputs", "xs", "tensors", "args_list", "operands"}


@dataclass
class ArgSpec:
    name: str
    type: str                  # Tensor | Tensor[] | int | float | bool | str | int[] | float[] | Scalar | DataType | any
    default: object = inspect.Parameter.empty
    optional: bool = False

    def render(self):
        s = f"{self.type} {self.name}"
        if self.default is not inspect.Parameter.empty:
            d = self.default
            s += "=" + ("none" if d is None else ("true" if d is True else "false" if d is False else (f'"{d}"' if isinstance(d, str) else
                        ("{" + ", ".join(map(str, d)) + "}" if isinstance(d, (list, tuple)) else str(d)))))
        return s


@dataclass
class OpSchema:
    name: str
    args: list = field(default_factory=list)
    outputs: list = field(default_factory=lambda: ["Tensor(out)"])
    module: str = ""
    backward: str | None = None
    inplace_of: str | None = None
    infer_meta: str = "generic"
    func: object = None

    @property
    def tensor_args(self):
        return [a for a in self.args if a.type.startswith("Tensor")]

    @property
    def attrs(self):
        return [a for a in self.args if not a.type.startswith("Tensor")]

    def signature(self):
        return f"{self.name}({', '.join(a.render() for a in self.args)}) -> {', '.join(self.outputs)}"


class MetaTensor:
    """Shape + dtype carrier (phi::MetaTensor). dims use -1 for sizes only known at run time."""

    __slots__ = ("dims", "dtype")

    def __init__(self, dims, dtype=torch.float32):
        self.dims, self.dtype = DDim(dims), dtype

    @classmethod
    def from_tensor(cls, t):
        return cls(list(t.shape), t.dtype)

    def to_meta(self, dynamic_size=2):
        return torch.empty([int(d) if d >= 0 else dynamic_size for d in self.dims], dtype=self.dtype, device="meta")

    def __repr__(self):
        return f"MetaTensor({list(self.dims)}, {str(self.dtype).split('.')[-1]})"

    def __eq__(self, o):
        return isinstance(o, MetaTensor) and list(self.dims) == list(o.dims) and self.dtype == o.dtype


REGISTRY: dict[str, OpSchema] = {}
_CUSTOM_META = {}


def _arg_type(name, default, annotation):
    if name in _TENSOR_LIST_ARGS:
        return "Tensor[]"
    if name in _TENSOR_ARGS:
        return "Tensor"
    if name in ("dtype", "out_dtype"):
        return "DataType"
    if name in ("shape", "axes", "perm", "dims", "sizes", "repeat_times", "strides", "paddings", "dilations", "kernel_size", "output_size", "starts", "ends"):
        return "int[]"
    if name in ("axis", "dim"):
        return "int[]" if isinstance(default, (list, tuple)) else "int"
    if isinstance(default, bool):
        return "bool"
    if isinstance(default, int):
        return "int"
    if isinstance(default, float):
        return "float"
    if isinstance(default, str):
        return "str"
    if isinstance(default, (list, tuple)):
        return "float[]" if default and isinstance(default[0], float) else "int[]"
    return "Scalar" if name in ("min", "max", "scale", "alpha", "beta", "p", "fill_value", "epsilon", "eps", "rtol", "atol", "threshold", "factor") else "any"


def _schema_of(name, fn, module):
    sig = inspect.signature(fn)
    args = []
    for pn, p in sig.parameters.items():
        if pn in ("name", "out") or p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            continue
        default = p.default
        typ = _arg_type(pn, None if default is inspect.Parameter.empty else default, p.annotation)
        first_positional = not args and default is inspect.Parameter.empty
        if first_positional and typ == "any":
            typ = "Tensor"
        args.append(ArgSpec(pn, typ, default, optional=default is None))
    s = OpSchema(name=name, args=args, module=module, func=fn)
    if name.endswith("_") and name[:-1] in REGISTRY:
        s.inplace_of = name[:-1]
    s.backward = None if (name.startswith(("is_", "arg", "equal", "not_", "less", "greater", "logical", "bitwise", "all", "any", "shape", "numel", "rank"))
                          or s.inplace_of) else f"{name}_grad"
    return s


def build_registry(force=False):
    """Derive a schema for every public op of paddle_b200.ops.* (the live op library)."""
    if REGISTRY and not force:
        return REGISTRY
    from . import creation, linalg, logic, manipulation, math, random, search, stat

    REGISTRY.clear()
    for mod in (math, manipulation, logic, search, stat, linalg, creation, random):
        short = mod.__name__.rsplit(".", 1)[-1]
        for name in sorted(getattr(mod, "__all__", [])):
            fn = getattr(mod, name, None)
            if fn is None or not callable(fn) or inspect.isclass(fn):
                continue
            try:
                REGISTRY[name] = _schema_of(name, fn, short)
            except (TypeError, ValueError):
                continue
    return REGISTRY


def get(name):
    build_registry()
    return REGISTRY[name]


def register_infer_meta(name):
    def deco(fn):
        _CUSTOM_META[name] = fn
        return fn
    return deco


def _to_meta(a, dyn=2):
    if isinstance(a, MetaTensor):
        return a.to_meta(dyn)
    if isinstance(a, torch.Tensor):
        return torch.empty_like(a, device="meta") if a.device.type != "meta" else a
    if isinstance(a, (list, tuple)):
        return type(a)(_to_meta(b, dyn) for b in a)
    return a


def _from_meta(r):
    if isinstance(r, torch.Tensor):
        return MetaTensor(list(r.shape), r.dtype)
    if isinstance(r, (list, tuple)):
        return type(r)(_from_meta(b) for b in r)
    return r


def infer_meta(name, *args, **attrs):
    """Output MetaTensor(s) of op `name` for MetaTensor / Tensor inputs (no kernel runs, no memory is touched)."""
    build_registry()
    if name in _CUSTOM_META:
        return _CUSTOM_META[name](*args, **attrs)
    s = REGISTRY[name]

    def has_dyn(a):
        return (isinstance(a, MetaTensor) and a.dims.is_dynamic()) or (isinstance(a, (list, tuple)) and any(has_dyn(b) for b in a))

    def run(stand_in):
        return _from_meta(s.func(*[_to_meta(a, stand_in) for a in args], **{k: _to_meta(v, stand_in) for k, v in attrs.items()}))

    res = run(2)
    if any(has_dyn(a) for a in list(args) + list(attrs.values())):
        # dynamic (-1) input dims: evaluate with two different stand-in sizes; every output dim that changes depends on them -> -1
        other = run(3)

        def merge(a, b):
            if isinstance(a, MetaTensor) and isinstance(b, MetaTensor) and len(a.dims) == len(b.dims):
                return MetaTensor([x if x == y else -1 for x, y in zip(a.dims, b.dims)], a.dtype)
            if isinstance(a, (list, tuple)):
                return type(a)(merge(x, y) for x, y in zip(a, b))
            return a

        res = merge(res, other)
    return res


@register_infer_meta("nonzero")
def _nonzero_meta(x, as_tuple=False):
    nd = len(x.dims) if isinstance(x, MetaTensor) else x.dim()
    return MetaTensor([-1, nd], torch.int64)


@register_infer_meta("masked_select")
def _masked_select_meta(x, mask):
    return MetaTensor([-1], x.dtype)


@register_infer_meta("unique")
def _unique_meta(x, *a, **k):
    return MetaTensor([-1], x.dtype)


# ---------------------------------------------------------------------------------------------------------------- YAML
_YAML_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ops.yaml")


def dump_yaml(path=None):
    build_registry()
    lines = ["# Operator schema of paddle_b200.ops - generated by `python -m paddle_b200.ops.schema --dump` from the live functions.",
             "# Same entry layout as the reference's paddle/phi/ops/yaml/ops.yaml; `tests/test_op_schema_cpu.py` validates it against the code.", ""]
    for name in sorted(REGISTRY):
        s = REGISTRY[name]
        lines.append(f"- op : {name}")
        lines.append(f"  args : ({', '.join(a.render() for a in s.args)})")
        lines.append(f"  output : {', '.join(s.outputs)}")
        lines.append("  infer_meta :")
        lines.append(f"    func : {'custom' if name in _CUSTOM_META else 'MetaEval'}")
        lines.append("  kernel :")
        lines.append(f"    func : {s.module}.{name}")
        if s.inplace_of:
            lines.append(f"  inplace : (x -> out)  # in-place form of {s.inplace_of}")
        if s.backward:
            lines.append(f"  backward : {s.backward}")
        lines.append("")
    text = "\n".join(lines)
    with open(path or _YAML_PATH, "w") as f:
        f.write(text)
    return text


_ARG_RE = re.compile(r"\s*([A-Za-z\[\]]+)\s+(\w+)\s*(?:=\s*(.+))?\s*$")


def _split_args(body):
    out, depth, cur = [], 0, ""
    for ch in body:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def _parse_default(typ, text):
    t = text.strip()
    if t == "none":
        return None
    if t in ("true", "false"):
        return t == "true"
    if t.startswith('"'):
        return t.strip('"')
    if t.startswith("{"):
        inner = t.strip("{}").strip()
        vals = [v.strip() for v in inner.split(",")] if inner else []
        return [float(v) if typ == "float[]" else int(v) for v in vals]
    try:
        return int(t)
    except ValueError:
        try:
            return float(t)
        except ValueError:
            return t


def parse_signature(sig):
    """'matmul(Tensor x, Tensor y, bool transpose_x=false) -> Tensor(out)' -> OpSchema (the args : / output : syntax of the YAML)."""
    head, _, outs = sig.strip().partition("->")
    m = re.match(r"\s*(\w+)\s*\((.*)\)\s*$", head.strip(), re.S)
    if not m:
        raise ValueError(f"cannot parse op signature: {sig!r}")
    name, body, outs = m.group(1), m.group(2), outs.strip() or None
    args = []
    for part in _split_args(body):
        if not part.strip():
            continue
        am = _ARG_RE.match(part)
        if not am:
            raise ValueError(f"cannot parse argument {part!r} of op {name}")
        typ, an, dv = am.group(1), am.group(2), am.group(3)
        default = inspect.Parameter.empty if dv is None else _parse_default(typ, dv)
        args.append(ArgSpec(an, typ, default, optional=default is None))
    return OpSchema(name=name, args=args, outputs=[o.strip() for o in (outs or "Tensor(out)").split(",")])


def load_yaml(path=None):
    """Parse ops.yaml into {name: OpSchema}."""
    out, cur = {}, None
    with open(path or _YAML_PATH) as f:
        for raw in f:
            line = raw.rstrip("\n")
            if line.startswith("- op :"):
                cur = {"name": line.split(":", 1)[1].strip()}
                out[cur["name"]] = cur
            elif cur is not None and line.strip().startswith("args :"):
                cur["args"] = line.split(":", 1)[1].strip()
            elif cur is not None and line.strip().startswith("output :"):
                cur["output"] = line.split(":", 1)[1].strip()
            elif cur is not None and line.strip().startswith("backward :"):
                cur["backward"] = line.split(":", 1)[1].strip()
            elif cur is not None and line.strip().startswith("func :") and "kernel" in cur.get("_last", ""):
                cur["kernel"] = line.split(":", 1)[1].strip()
            if cur is not None:
                cur["_last"] = line.strip().split(":")[0].strip() if line.strip() else cur.get("_last", "")
    res = {}
    for name, d in out.items():
        s = parse_signature(f"{name}{d.get('args', '()')} -> {d.get('output', 'Tensor(out)')}")
        s.backward = d.get("backward")
        s.module = (d.get("kernel") or ".").split(".")[0]
        res[name] = s
    return res


def validate(path=None):
    """Differences between ops.yaml and the live op library: [(op, problem)] (empty = in sync)."""
    build_registry()
    declared = load_yaml(path)
    problems = []
    for name, s in REGISTRY.items():
        d = declared.get(name)
        if d is None:
            problems.append((name, "missing from ops.yaml"))
            continue
        live = [(a.name, a.type) for a in s.args]
        decl = [(a.name, a.type) for a in d.args]
        if live != decl:
            problems.append((name, f"args differ: yaml {decl} vs code {live}"))
            continue
        for a, b in zip(s.args, d.args):
            da = None if a.default is inspect.Parameter.empty else a.default
            db = None if b.default is inspect.Parameter.empty else b.default
            if isinstance(da, tuple):
                da = list(da)
            if da != db and not (isinstance(da, float) and isinstance(db, (int, float)) and abs(da - db) < 1e-12) and not callable(da) \
                    and isinstance(da, (int, float, bool, str, list, type(None))):
                problems.append((name, f"default of {a.name}: yaml {db!r} vs code {da!r}"))
    for name in declared:
        if name not in REGISTRY:
            problems.append((name, "declared in ops.yaml but not implemented"))
    return problems


if __name__ == "__main__":
    import sys

    if "--dump" in sys.argv:
        dump_yaml()
        print(f"wrote {_YAML_PATH} ({len(REGISTRY)} ops)")
    else:
        for p in validate():
            print(p)
