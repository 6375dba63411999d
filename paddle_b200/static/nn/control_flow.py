"""Data-dependent control flow for recorded programs. Parity: python/paddle/static/nn/control_flow.py (cond / case / switch_case /
while_loop over Variables whose values are only known at run time).

A recorded program cannot branch in Python on a placeholder value, so:
  * `cond` records BOTH branches and selects the results with `where(pred, true_out, false_out)` (branches must be side-effect free
    and return the same structure — the reference has the same structural requirement); `case` / `switch_case` are folds of `cond`;
  * `while_loop` records the condition and the body once as sub-tapes and emits one control node that re-executes them at run time
    until the condition is false, rebinding the loop variables after every iteration.
In dygraph (or when the predicate is a plain Python / constant value) the ordinary eager semantics apply."""
from __future__ import annotations

import torch

from ...tensor import Tensor


def _prog_of(*values):
    from ... import static

    prog = static._recording[0]
    if prog is None:
        return None
    return prog if any(isinstance(v, torch.Tensor) and prog._touches_program(v) for v in values) else None


def _truth(p):
    return bool(p.item()) if isinstance(p, torch.Tensor) else bool(p)


def _select(pred, a, b):
    if a is None and b is None:
        return None
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        if list(a.shape) != list(b.shape):
            raise ValueError(f"cond: the branches return tensors of different shapes {list(a.shape)} vs {list(b.shape)}")
        p = pred.reshape([1] * a.dim()) if a.dim() else pred.reshape([])
        return torch.where(p.astype("bool") if hasattr(p, "astype") else p.bool(), a, b)
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)) and len(a) == len(b):
        return type(a)(_select(pred, x, y) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict) and a.keys() == b.keys():
        return {k: _select(pred, a[k], b[k]) for k in a}
    if not isinstance(a, torch.Tensor) and a == b:
        return a
    raise TypeError("cond: true_fn and false_fn must return the same structure of tensors")


def cond(pred, true_fn=None, false_fn=None, name=None, return_names=None):
    if _prog_of(pred) is None:
        if _truth(pred):
            return true_fn() if true_fn is not None else None
        return false_fn() if false_fn is not None else None
    t = true_fn() if true_fn is not None else None
    f = false_fn() if false_fn is not None else None
    if (t is None) != (f is None):
        raise ValueError("cond in a static program: both branches must return values (or neither); side effects cannot be recorded conditionally")
    return _select(pred, t, f)


def case(pred_fn_pairs, default=None, name=None):
    pairs = list(pred_fn_pairs)
    if _prog_of(*[p for p, _ in pairs]) is None:
        for pred, fn in pairs:
            if _truth(pred):
                return fn()
        return default() if default is not None else pairs[-1][1]()
    result = default() if default is not None else pairs[-1][1]()
    for pred, fn in reversed(pairs):
        result = _select(pred, fn(), result)
    return result


def switch_case(branch_index, branch_fns, default=None, name=None):
    fns = dict(branch_fns) if not isinstance(branch_fns, dict) else dict(branch_fns)
    if isinstance(branch_fns, (list, tuple)) and branch_fns and not isinstance(branch_fns[0], (list, tuple)):
        fns = dict(enumerate(branch_fns))
    if _prog_of(branch_index) is None:
        i = int(branch_index.item() if isinstance(branch_index, torch.Tensor) else branch_index)
        if i in fns:
            return fns[i]()
        return default() if default is not None else fns[max(fns)]()
    return case([(branch_index == k, fn) for k, fn in sorted(fns.items())], default=default if default is not None else fns[max(fns)])


def _exec(nodes, env):
    from ... import static

    def decode(x):
        if isinstance(x, static._Ref):
            return env[x.vid]
        if isinstance(x, (list, tuple)):
            return type(x)(decode(i) for i in x)
        if isinstance(x, dict):
            return {k: decode(v) for k, v in x.items()}
        return x

    for n in nodes:
        if n.kind != "op":
            n.fn(env)
            continue
        out = n.fn(*decode(n.args), **decode(n.kwargs))
        flat = []

        def fl(o):
            if isinstance(o, torch.Tensor):
                flat.append(o)
            elif isinstance(o, (list, tuple)):
                for i in o:
                    fl(i)

        fl(out)
        for vid, t in zip(n.outs, flat):
            env[vid] = t


def while_loop(cond, body, loop_vars, is_test=False, name=None):
    from ... import static

    vs = list(loop_vars)
    probe = None
    prog = static._recording[0]
    n0 = len(prog.nodes) if prog is not None else 0
    if prog is not None:
        probe = cond(*vs)
    if prog is None or not (isinstance(probe, torch.Tensor) and (prog._touches_program(probe) or any(isinstance(v, torch.Tensor) and prog._touches_program(v) for v in vs))):
        if prog is not None:
            del prog.nodes[n0:]            # the probe was recorded for nothing: constant loop, unroll eagerly at record time
        while _truth(cond(*vs)):
            out = body(*vs)
            vs = list(out) if isinstance(out, (list, tuple)) else [out]
        return vs
    # ---- run-time loop: sub-tapes for the condition and the body ------------------------------------------------------------
    loop_vids = []
    init_consts = {}
    for v in vs:
        if not isinstance(v, torch.Tensor):
            raise TypeError("while_loop in a static program: loop_vars must be tensors")
        if id(v) not in prog._vids:
            init_consts[prog._new_vid(v)] = v          # a constant used as loop state: give it a slot, seed it when the loop starts
        loop_vids.append(prog._vids[id(v)])
    del prog.nodes[n0:]
    c = cond(*vs)
    cond_nodes = prog.nodes[n0:]
    del prog.nodes[n0:]
    c_vid = prog._vids[id(c)]
    out = body(*vs)
    outs = list(out) if isinstance(out, (list, tuple)) else [out]
    if len(outs) != len(vs):
        raise ValueError("while_loop: body must return as many values as there are loop_vars")
    body_nodes = prog.nodes[n0:]
    del prog.nodes[n0:]
    out_vids = []
    for o in outs:
        if id(o) not in prog._vids:
            init_consts[prog._new_vid(o)] = o
        out_vids.append(prog._vids[id(o)])
    with torch._C.DisableTorchFunction():       # plain copies (no subclass / mode dispatch): they must not be recorded as ops of the program
        results = [torch.Tensor.clone(v).as_subclass(Tensor) for v in vs]
    res_vids = [prog._new_vid(r) for r in results]
    from ..passes import _refs

    refs = sorted({r for nodes in (cond_nodes, body_nodes) for n in nodes for r in _refs(n.args) + _refs(n.kwargs)} | set(loop_vids))

    def run(env, _max=10 ** 7):
        for vid, t in init_consts.items():
            env.setdefault(vid, t)
        it = 0
        while True:
            _exec(cond_nodes, env)
            if not _truth(env[c_vid]):
                break
            _exec(body_nodes, env)
            new = [env[v] for v in out_vids]
            for lv, t in zip(loop_vids, new):
                env[lv] = t
            it += 1
            if it > _max:
                raise RuntimeError("while_loop: iteration limit reached")
        for rv, lv in zip(res_vids, loop_vids):
            env[rv] = env[lv]

    prog.nodes.append(static._Node(run, [static._Ref(r) for r in refs], {}, res_vids, kind="control"))
    return results
