"""Fusion-group formation over the SSA program (`pir.translate_to_pir`) and the rewrite that replaces each group by one generated kernel.

The walk is a single pass in program order.  An op joins the OPEN group of one of its producers when its result lives on that group's
domain (or on its per-row shape); otherwise it opens a group of its own.  A group is closed the moment a value it computes is read by an op
that is not a member: the fused op is placed where the group's last member stood, so every outside reader must come after it, and every
outside value a member reads was defined before that member - hence before the fused op.  No cycle can form.
Role parity: CINN's op fusion / fusion merge passes (paddle/cinn/hlir/framework/pir/op_lowering..., paddle/cinn/operator_fusion)."""
from __future__ import annotations

import hashlib

import torch

from . import codegen
from .expr import Frontend, GroupBuilder, Node, Unsupported
from .runtime import CompileError, FusedKernel

_PATTERN_OPS = {"swiglu"}                                                   # ops DRR patterns introduce that lower to primitives
_ROW_OPS = {"sum", "mean", "amax", "amin", "softmax", "log_softmax", "layer_norm", "rms_norm"}       # their domain is the INPUT's shape


class Group:
    def __init__(self, gid, full):
        self.id, self.gb = gid, GroupBuilder(full)
        self.fe = Frontend(self.gb)
        self.op_ids, self.open, self.dead, self.pos = [], True, False, None

    def snapshot(self):
        return len(self.gb.nodes), dict(self.gb.inputs), dict(self.gb.by_value)

    def restore(self, snap):
        n, ins, byv = snap
        del self.gb.nodes[n:]
        self.gb.inputs, self.gb.by_value = ins, byv

    def absorb(self, other):
        """Merge another OPEN group on the same domain into this one (both are still unread from outside, so running them as one kernel at the
        later position is legal)."""
        remap = {}
        for n in other.gb.nodes:
            if n.kind == "in":
                v = (n.attrs["value"], n.shape)
                if v in self.gb.inputs:
                    remap[id(n)] = self.gb.inputs[v]
                    continue
                self.gb.inputs[v] = n
            else:
                n.args = [remap.get(id(a), a) if isinstance(a, Node) else a for a in n.args]
            n.id = len(self.gb.nodes)
            self.gb.nodes.append(n)
        for v, n in other.gb.by_value.items():
            self.gb.by_value[v] = remap.get(id(n), n)
        self.op_ids = sorted(set(self.op_ids) | set(other.op_ids), key=lambda o: self.pos[o])
        other.open, other.dead = False, True


class FusionResult:
    def __init__(self):
        self.groups = []          # dicts: name, ops (names), inputs, outputs, kind, kernel
        self.rejected = []        # (op name, reason) of candidates that could not be lowered

    def __repr__(self):
        return f"FusionResult({len(self.groups)} kernels, {sum(len(g['ops']) for g in self.groups)} ops fused)"


def _slot_type():
    from ..pir import _Slot

    return _Slot


def _decode(t, operand_node, Slot):
    if isinstance(t, Slot):
        return operand_node(t.k)
    if isinstance(t, (list, tuple)):
        out = [_decode(i, operand_node, Slot) for i in t]
        if any(isinstance(i, Node) for i in out):
            raise Unsupported("tensor list argument")
        return type(t)(out)
    if isinstance(t, dict):
        raise Unsupported("dict argument")
    if isinstance(t, torch.Tensor):
        raise Unsupported("captured tensor outside the operand list")
    return t


def fuse(tr, min_ops=2, targets=("cuda",), precompile=False):
    """Form groups over `tr.ir`, generate their kernels and rewrite the program in place.  Returns a FusionResult.
    `targets`: which sources must generate without error for a group to be accepted; `precompile` builds them now (else at first launch)."""
    from .. import pir

    Slot = _slot_type()
    ir = tr.ir
    ops = ir.ops()
    group_of = {}                                  # value id -> Group that computes it
    groups, res = [], FusionResult()

    def vtype(v):
        dt, shape = ir.value_type(v)
        return dt, tuple(int(d) for d in shape)

    def close_producers(op, keep=None):
        for v in op["operands"]:
            g = group_of.get(v)
            if g is not None and g is not keep:
                g.open = False

    pos = {op["id"]: k for k, op in enumerate(ops)}
    for op in ops:
        tpl = tr.templates.get(op["id"])
        name = op["name"].split(".", 1)[-1]
        cand = tpl is not None and tpl[0] == "call" and tpl[4] == "op" and len(op["results"]) == 1 and op["num_regions"] == 0
        if tpl is None and name in _PATTERN_OPS and len(op["results"]) == 1:        # created by a rewrite pattern: operands are its arguments
            tpl = ("call", None, tuple(Slot(k) for k in range(len(op["operands"]))), {}, "op")
            cand = True
        if cand:
            raw_name = getattr(tpl[1], "__name__", "") or ""
            if raw_name.endswith("_") and not raw_name.endswith("__"):        # in-place spelling: it mutates its first operand
                cand = False
        if not cand:
            close_producers(op)
            continue
        r = op["results"][0]
        rdt, rshape = vtype(r)
        numel = 1
        for d in rshape:
            numel *= d
        tensor_operands = op["operands"]
        if not tensor_operands or numel == 0:
            close_producers(op)
            continue
        domain = vtype(tensor_operands[0])[1] if name in _ROW_OPS else rshape
        # candidate groups: open groups of producers on which this op's result fits
        chosen = None
        for v in tensor_operands:
            g = group_of.get(v)
            if g is not None and g.open and (g.gb.full == domain if name in _ROW_OPS else g.gb.space_of(rshape) is not None):
                if chosen is None:
                    chosen = g
                elif g is not chosen and g.gb.full == chosen.gb.full and _compatible(g, chosen):       # two open groups feed this op: one kernel
                    chosen.absorb(g)
                    for val, owner in list(group_of.items()):
                        if owner is g:
                            group_of[val] = chosen
        sideways = None
        if chosen is None:
            # horizontal fusion: no producer is in a group, but an open group on this domain already reads one of this op's operands - joining
            # it saves a pass over that tensor (dbeta = dy.sum(0) next to dgamma = (dy * xhat).sum(0))
            wanted = set(tensor_operands)
            for g in reversed(groups):
                if g.open and not g.dead and (g.gb.full == domain if name in _ROW_OPS else g.gb.full == rshape) \
                        and any(v in wanted for (v, _shape) in g.gb.inputs):
                    sideways = g
                    break
        node = fresh = None
        for cand_group in ([chosen, None] if chosen is not None else [sideways, None] if sideways is not None else [None]):
            fresh = cand_group is None
            grp = Group(len(groups), domain) if fresh else cand_group
            if fresh:
                grp.pos = pos
            snap = grp.snapshot()

            def operand_node(k, g=grp, op=op):
                v = op["operands"][k]
                if v in g.gb.by_value:
                    return g.gb.by_value[v]
                dt, shape = vtype(v)
                return g.gb.input(v, shape, dt)

            try:
                args = [_decode(a, operand_node, Slot) for a in tpl[2]]
                kwargs = {k: _decode(a, operand_node, Slot) for k, a in tpl[3].items()}
                kwargs = {k: v for k, v in kwargs.items() if not (k == "name" and (v is None or isinstance(v, str)))}
                node = grp.fe.lower(name, args, kwargs, rshape, rdt)
                if not isinstance(node, Node) or node.kind == "in":
                    raise Unsupported("op lowered to no computation")
                if node.attrs.get("post") and rshape == () and node.dtype == rdt:
                    pass                                  # per-row partials, finished to the recorded scalar on the way out
                elif node.shape != rshape or node.dtype != rdt:
                    raise Unsupported(f"lowered type {node.dtype}{list(node.shape)} differs from the recorded {rdt}{list(rshape)}")
                chosen = grp
                break
            except Unsupported as e:
                grp.restore(snap)
                node = None
                if cand_group is None:                   # a join that does not fit falls back to a group of its own; only that failure is final
                    res.rejected.append((name, str(e)))
        if node is None:
            close_producers(op)
            continue
        node.value_id = r
        chosen.gb.by_value[r] = node
        chosen.op_ids.append(op["id"])
        group_of[r] = chosen
        if fresh:
            groups.append(chosen)
        close_producers(op, keep=chosen)

    # ---- outputs of every group: values read outside it, or returned by the program
    groups = [g for g in groups if not g.dead]
    member = {}
    for g in groups:
        for oid in g.op_ids:
            member[oid] = g
    outside = {g.id: [] for g in groups}
    for op in ops:
        for v in op["operands"]:
            g = group_of.get(v)
            if g is not None and member.get(op["id"]) is not g and v not in outside[g.id]:
                outside[g.id].append(v)
    for v in ir.outputs():
        g = group_of.get(v)
        if g is not None and v not in outside[g.id]:
            outside[g.id].append(v)

    op_name = {op["id"]: op["name"].split(".", 1)[-1] for op in ops}
    renamed = {}
    for g in groups:
        outs = outside[g.id]
        if len(g.op_ids) < min_ops or not outs:
            continue
        out_nodes = [g.gb.by_value[v] for v in outs]
        nodes = _prune(g.gb.nodes, out_nodes)
        in_nodes = [n for n in nodes if n.kind == "in"]
        if not in_nodes:
            continue
        for k, n in enumerate(nodes):
            n.id = k
        sig = hashlib.sha1(repr([(n.kind, n.op, [a.id if isinstance(a, Node) else a for a in n.args], n.shape, n.dtype,
                                  sorted((k, v) for k, v in n.attrs.items() if k != "value")) for n in nodes]
                                + [o.id for o in out_nodes]).encode()).hexdigest()[:12]
        kname = f"cinn_fused_{sig}"
        known = pir._IMPL.get(kname)
        if isinstance(known, FusedKernel):                 # the same computation on the same types was compiled before (another layer, another program)
            kernel, spec = known, known.spec
        else:
            spec = codegen.Spec(kname, g.gb.full, nodes, in_nodes, out_nodes)
            kernel = FusedKernel(spec)
        try:
            for t in targets:
                kernel.source(t)
                if precompile:
                    kernel.build(t)
        except (Unsupported, CompileError) as e:
            res.rejected.append((kname, str(e)))
            continue
        # rewrite: one op where the last member stood
        ir.set_insertion_point_after(g.op_ids[-1])
        new_vals = ir.add_op("cinn_op." + kname, [renamed.get(n.attrs["value"], n.attrs["value"]) for n in in_nodes], {"kernel": kname}, [ir.value_type(v) for v in outs])
        ir.reset_insertion_point()
        for old, new in zip(outs, new_vals):
            ir.replace_all_uses(old, new)
            renamed[old] = new                 # later groups recorded the old value as their input
        for oid in reversed(g.op_ids):
            ir.erase_op(oid)
        pir.register_op_impl(kname, kernel)
        res.groups.append({"name": kname, "ops": [op_name[o] for o in g.op_ids], "inputs": len(in_nodes), "outputs": len(out_nodes),
                           "kind": "column" if spec.col else "reduce" if spec.has_reduce else "elementwise", "kernel": kernel, "domain": list(g.gb.full)})
    ir.verify()
    return res


def _compatible(a, b):
    """May two groups on the same domain share a kernel?  Not when one reduces along rows and the other along columns (different schedules), nor
    when their column reductions collapse different axes."""
    def kinds(g):
        row = any(n.kind == "reduce" or (n.kind == "ew" and n.space == "row") for n in g.gb.nodes)
        col = {n.attrs["akb"] for n in g.gb.nodes if n.kind == "creduce"}
        return row, col

    (ra, ca), (rb, cb) = kinds(a), kinds(b)
    if (ra and cb) or (rb and ca):
        return False
    return len(ca | cb) <= 1


def _prune(nodes, outputs):
    """Nodes the outputs depend on (dead branches of a decomposition / of ops whose value nobody reads are dropped), still topological."""
    live, stack = set(), list(outputs)
    while stack:
        n = stack.pop()
        if id(n) in live:
            continue
        live.add(id(n))
        stack.extend(a for a in n.args if isinstance(a, Node))
    return [n for n in nodes if id(n) in live]
