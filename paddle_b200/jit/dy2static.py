"""AST dynamic-to-static conversion of tensor-dependent control flow.

Parity: python/paddle/jit/dy2static/program_translator.py:1759 (ProgramTranslator), transformers/ifelse_transformer.py,
loop_transformer.py, logical_transformer.py, convert_operators.py (convert_ifelse / convert_while_loop / convert_logical_*).

B200 design: `to_static` captures a function into a CUDA graph, so a Python `if tensor:` would need a host read in the middle of the
capture.  The transformer rewrites

    if cond:  A            ->   def __t(): nonlocal v...; A; return v...
    else:     B                 def __f(): nonlocal v...; B; return v...
                                v... = _jst.convert_ifelse(cond, __t, __f, get_state, set_state, names)

and at run time `convert_ifelse` keeps plain Python semantics for Python conditions, while a TENSOR condition runs BOTH branches
from the same starting state and merges every modified variable with a device-side select (`where(cond, a, b)`), which is capture
safe.  `while` / `for range(tensor)` loops become `convert_while_loop` (Python loop for Python conditions; a tensor condition is read
back once per iteration - loops whose trip count depends on data cannot live inside one CUDA graph, the call then stays eager).
`and` / `or` / `not` on tensors become logical ops.  Statements the merge cannot express (return / break / continue / yield inside a
converted branch) leave that `if` untouched.
"""
from __future__ import annotations

import ast
import functools
import inspect
import textwrap

import torch

__all__ = ["convert_to_static", "get_code", "ProgramTranslator", "convert_ifelse", "convert_while_loop", "convert_logical_and",
           "convert_logical_or", "convert_logical_not", "UNDEFINED"]


class _Undefined:
    """Value of a name that a branch did not assign.  After a tensor-dependent `if` a name assigned in ONE branch only stays undefined - which is
    fine as long as the code does not use it (a branch-local temporary); any use fails with the name in the message."""

    def __init__(self, name=None):
        self._name = name

    def __repr__(self):
        return "UNDEFINED"

    def __bool__(self):
        return False

    def _fail(self, *a, **k):
        raise NameError(f"dy2static: variable '{self._name}' is assigned in only one branch of a tensor-dependent `if` and used after it; give it a value "
                        "before the `if`")

    __getattr__ = __call__ = __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __getitem__ = __iter__ = __neg__ = \
        __matmul__ = __rmatmul__ = __lt__ = __gt__ = __le__ = __ge__ = __float__ = __int__ = __len__ = _fail


UNDEFINED = _Undefined()


def _is_tensor(x):
    return isinstance(x, torch.Tensor)


def _select(pred, a, b, name):
    if isinstance(a, _Undefined) or isinstance(b, _Undefined):
        return _Undefined(name)              # a branch-local temporary: legal unless the code uses it after the `if`
    if _is_tensor(a) or _is_tensor(b):
        ta = a if _is_tensor(a) else torch.as_tensor(a, device=b.device, dtype=b.dtype)
        tb = b if _is_tensor(b) else torch.as_tensor(b, device=a.device, dtype=a.dtype)
        if ta.dtype != tb.dtype:
            dt = torch.promote_types(ta.dtype, tb.dtype)
            ta, tb = ta.to(dt), tb.to(dt)
        p = pred.reshape([1] * max(ta.dim(), tb.dim())) if pred.numel() == 1 else pred
        out = torch.where(p.to(torch.bool), ta, tb)
        if _is_tensor(a) and type(a) is not torch.Tensor and not isinstance(out, type(a)):
            return out.as_subclass(type(a))
        return out                       # unchanged object: a program being recorded tracks its values by identity
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)) and len(a) == len(b):
        return type(a)(_select(pred, x, y, name) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict) and a.keys() == b.keys():
        return {k: _select(pred, a[k], b[k], f"{name}[{k!r}]") for k in a}
    if a is b or a == b:
        return a
    raise ValueError(f"dy2static: non-tensor variable '{name}' takes different values ({a!r} / {b!r}) in the branches of a tensor-dependent `if`")


def convert_ifelse(pred, true_fn, false_fn, get_state, set_state, names):
    """`if pred: true_fn() else: false_fn()` - Python semantics for Python predicates, run-both-and-select for tensor predicates."""
    if not _is_tensor(pred):
        return true_fn() if pred else false_fn()
    if pred.numel() != 1:
        raise ValueError("dy2static: the condition of an `if` must be a scalar tensor")
    saved = get_state()
    out_t = true_fn()
    state_t = get_state()
    set_state(saved)
    out_f = false_fn()
    state_f = get_state()
    merged = tuple(_select(pred, a, b, n) for a, b, n in zip(state_t, state_f, names))
    set_state(merged)
    return merged if len(names) != 1 else merged


def convert_while_loop(cond_fn, body_fn, get_state, set_state, names):
    """`while cond: body`.  A tensor condition is evaluated on the host each iteration (data-dependent trip count)."""
    while True:
        c = cond_fn()
        if _is_tensor(c):
            c = bool(c.item())
        if not c:
            break
        body_fn()
    return get_state()


def convert_logical_and(x_fn, y_fn):
    x = x_fn()
    if not _is_tensor(x):
        return x and y_fn()
    y = y_fn()
    return torch.logical_and(x, y if _is_tensor(y) else torch.as_tensor(bool(y), device=x.device))


def convert_logical_or(x_fn, y_fn):
    x = x_fn()
    if not _is_tensor(x):
        return x or y_fn()
    y = y_fn()
    return torch.logical_or(x, y if _is_tensor(y) else torch.as_tensor(bool(y), device=x.device))


def convert_logical_not(x):
    return torch.logical_not(x) if _is_tensor(x) else (not x)


def convert_range_bound(n):
    """`for i in range(n)` with a tensor bound: the trip count is read once."""
    return int(n.item()) if _is_tensor(n) else n


# ---------------------------------------------------------------------------------------------------------------- AST transform
class _Names(ast.NodeVisitor):
    """Names bound by a block (assign / augassign / for targets / with-as), not descending into nested function or class bodies."""

    def __init__(self):
        self.stores = []

    def _add(self, n):
        if n not in self.stores:
            self.stores.append(n)

    def visit_Name(self, node):
        if isinstance(node.ctx, (ast.Store, ast.Del)):
            self._add(node.id)

    def visit_FunctionDef(self, node):
        self._add(node.name)

    visit_AsyncFunctionDef = visit_FunctionDef

    def visit_ClassDef(self, node):
        self._add(node.name)

    def visit_Lambda(self, node):
        pass

    def visit_ListComp(self, node):
        pass

    visit_SetComp = visit_DictComp = visit_GeneratorExp = visit_ListComp


def _stores(stmts):
    v = _Names()
    for s in stmts:
        v.visit(s)
    return [n for n in v.stores if not n.startswith("__jst_")]


class _HasFlow(ast.NodeVisitor):
    def __init__(self):
        self.found = False

    def visit_Return(self, node):
        self.found = True

    visit_Break = visit_Continue = visit_Yield = visit_YieldFrom = visit_Global = visit_Nonlocal = visit_Return

    def visit_FunctionDef(self, node):
        pass

    visit_AsyncFunctionDef = visit_Lambda = visit_ClassDef = visit_FunctionDef


def _has_flow(stmts, loops_ok=False):
    v = _HasFlow()
    for s in stmts:
        if loops_ok and isinstance(s, (ast.For, ast.While)):
            # break / continue inside a nested loop belong to that loop; returns still escape
            r = _HasReturn()
            r.visit(s)
            if r.found:
                return True
            continue
        v.visit(s)
    return v.found


class _HasReturn(_HasFlow):
    def visit_Break(self, node):
        pass

    visit_Continue = visit_Break


def _tuple(names, ctx):
    return ast.Tuple(elts=[ast.Name(id=n, ctx=ctx()) for n in names], ctx=ctx())


def _state_fns(names, uid):
    """def __jst_get_k(): return (a, b) ; def __jst_set_k(v): nonlocal a, b; (a, b) = v"""
    get = ast.FunctionDef(name=f"__jst_get_{uid}", args=ast.arguments(posonlyargs=[], args=[], kwonlyargs=[], kw_defaults=[], defaults=[]),
                          body=[ast.Return(value=_tuple(names, ast.Load))], decorator_list=[], type_params=[])
    body = []
    if names:
        body.append(ast.Nonlocal(names=list(names)))
        body.append(ast.Assign(targets=[_tuple(names, ast.Store)], value=ast.Name(id="__jst_v", ctx=ast.Load())))
    else:
        body.append(ast.Pass())
    setf = ast.FunctionDef(name=f"__jst_set_{uid}", args=ast.arguments(posonlyargs=[], args=[ast.arg(arg="__jst_v")], kwonlyargs=[], kw_defaults=[], defaults=[]),
                           body=body, decorator_list=[], type_params=[])
    return get, setf


def _branch_fn(name, names, body):
    stmts = ([ast.Nonlocal(names=list(names))] if names else []) + (list(body) or [ast.Pass()]) + [ast.Return(value=_tuple(names, ast.Load))]
    return ast.FunctionDef(name=name, args=ast.arguments(posonlyargs=[], args=[], kwonlyargs=[], kw_defaults=[], defaults=[]), body=stmts,
                           decorator_list=[], type_params=[])


def _jst(attr):
    return ast.Attribute(value=ast.Name(id="__jst", ctx=ast.Load()), attr=attr, ctx=ast.Load())


class _Transformer(ast.NodeTransformer):
    def __init__(self):
        self.uid = 0
        self.func_depth = 0
        self.known = [set()]       # names certainly bound before the current statement, per function scope

    def _next(self):
        self.uid += 1
        return self.uid

    # nested defs get their own scope bookkeeping
    def visit_FunctionDef(self, node):
        self.func_depth += 1
        bound = {a.arg for a in node.args.args + node.args.kwonlyargs + node.args.posonlyargs}
        if node.args.vararg:
            bound.add(node.args.vararg.arg)
        if node.args.kwarg:
            bound.add(node.args.kwarg.arg)
        self.known.append(bound)
        node.body = self._block(node.body)
        self.known.pop()
        self.func_depth -= 1
        return node

    def _block(self, stmts):
        out = []
        for s in stmts:
            r = self.visit(s)
            if isinstance(r, list):
                out.extend(r)
            elif r is not None:
                out.append(r)
            self.known[-1].update(_stores([s]))
        return out

    def visit_BoolOp(self, node):
        self.generic_visit(node)
        fn = "convert_logical_and" if isinstance(node.op, ast.And) else "convert_logical_or"
        expr = node.values[-1]
        for v in reversed(node.values[:-1]):
            lam = lambda e: ast.Lambda(args=ast.arguments(posonlyargs=[], args=[], kwonlyargs=[], kw_defaults=[], defaults=[]), body=e)  # noqa: E731
            expr = ast.Call(func=_jst(fn), args=[lam(v), lam(expr)], keywords=[])
        return expr

    def visit_UnaryOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Not):
            return ast.Call(func=_jst("convert_logical_not"), args=[node.operand], keywords=[])
        return node

    def visit_If(self, node):
        node.test = self.visit(node.test)
        before = set(self.known[-1])
        node.body = self._block(node.body)
        self.known[-1] = set(before)
        node.orelse = self._block(node.orelse)
        self.known[-1] = set(before)
        if self.func_depth == 0 or _has_flow(node.body, loops_ok=True) or _has_flow(node.orelse, loops_ok=True):
            return node
        names = _stores(node.body + node.orelse)
        k = self._next()
        pre = [ast.Assign(targets=[ast.Name(id=n, ctx=ast.Store())], value=_jst("UNDEFINED")) for n in names if n not in before]
        get, setf = _state_fns(names, k)
        tfn = _branch_fn(f"__jst_true_{k}", names, node.body)
        ffn = _branch_fn(f"__jst_false_{k}", names, node.orelse)
        call = ast.Call(func=_jst("convert_ifelse"),
                        args=[node.test, ast.Name(id=tfn.name, ctx=ast.Load()), ast.Name(id=ffn.name, ctx=ast.Load()), ast.Name(id=get.name, ctx=ast.Load()),
                              ast.Name(id=setf.name, ctx=ast.Load()), ast.Tuple(elts=[ast.Constant(value=n) for n in names], ctx=ast.Load())], keywords=[])
        assign = ast.Assign(targets=[_tuple(names, ast.Store)], value=call) if names else ast.Expr(value=call)
        return pre + [get, setf, tfn, ffn, assign]

    def visit_While(self, node):
        node.test = self.visit(node.test)
        before = set(self.known[-1])
        node.body = self._block(node.body)
        self.known[-1] = set(before)
        if self.func_depth == 0 or node.orelse or _has_flow(node.body):
            return node
        names = _stores(node.body)
        k = self._next()
        pre = [ast.Assign(targets=[ast.Name(id=n, ctx=ast.Store())], value=_jst("UNDEFINED")) for n in names if n not in before]
        get, setf = _state_fns(names, k)
        cond = ast.FunctionDef(name=f"__jst_cond_{k}", args=ast.arguments(posonlyargs=[], args=[], kwonlyargs=[], kw_defaults=[], defaults=[]),
                               body=[ast.Return(value=node.test)], decorator_list=[], type_params=[])
        body = _branch_fn(f"__jst_body_{k}", names, node.body)
        call = ast.Call(func=_jst("convert_while_loop"),
                        args=[ast.Name(id=cond.name, ctx=ast.Load()), ast.Name(id=body.name, ctx=ast.Load()), ast.Name(id=get.name, ctx=ast.Load()),
                              ast.Name(id=setf.name, ctx=ast.Load()), ast.Tuple(elts=[ast.Constant(value=n) for n in names], ctx=ast.Load())], keywords=[])
        assign = ast.Assign(targets=[_tuple(names, ast.Store)], value=call) if names else ast.Expr(value=call)
        return pre + [get, setf, cond, body, assign]

    def visit_For(self, node):
        # for i in range(<maybe tensor>): the bound is converted; the loop itself stays a Python loop (static trip count once read)
        self.generic_visit(node)
        it = node.iter
        if isinstance(it, ast.Call) and isinstance(it.func, ast.Name) and it.func.id == "range":
            it.args = [ast.Call(func=_jst("convert_range_bound"), args=[a], keywords=[]) for a in it.args]
        return node


def _contains_return(stmts):
    class V(ast.NodeVisitor):
        found = False

        def visit_Return(self, node):
            self.found = True

        def visit_FunctionDef(self, node):          # a nested function's returns are its own
            pass

        visit_AsyncFunctionDef = visit_Lambda = visit_FunctionDef

    v = V()
    for s in stmts:
        v.visit(s)
    return v.found


def _normalize_early_returns(stmts, counter):
    """`if c: ...; return A` followed by `...; return B`  ->  `if c: ...; r = A  else: ...; r = B` + `return r`.
    Only the top-level shape: the `if` body ends with its single `return`, there is no `else` (or one that ends with its single `return`), and the
    statements after it end the function.  The rewritten `if` has no `return` inside, so the converter can treat a tensor condition
    (run both branches, select)."""
    for i, st in enumerate(stmts):
        if not isinstance(st, ast.If) or not st.body or not isinstance(st.body[-1], ast.Return) or _contains_return(st.body[:-1]):
            continue
        if _contains_return(stmts[:i]):
            return stmts
        rest = stmts[i + 1:]
        if st.orelse:
            if rest or not isinstance(st.orelse[-1], ast.Return) or _contains_return(st.orelse[:-1]):
                return stmts
            other = list(st.orelse)
        else:
            other = _normalize_early_returns(list(rest), counter)
            if not other or not isinstance(other[-1], ast.Return) or _contains_return(other[:-1]):
                return stmts
        counter[0] += 1
        name = f"_jst_ret_{counter[0]}"

        def assign(ret):
            return ast.Assign(targets=[ast.Name(id=name, ctx=ast.Store())], value=ret.value if ret.value is not None else ast.Constant(value=None))

        new_if = ast.If(test=st.test, body=st.body[:-1] + [assign(st.body[-1])], orelse=other[:-1] + [assign(other[-1])])
        return stmts[:i] + [new_if, ast.Return(value=ast.Name(id=name, ctx=ast.Load()))]
    return stmts


def _transform_source(fn):
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = next(n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef)))
    fdef.decorator_list = []          # the decorator (to_static itself) must not run again
    fdef.body = _normalize_early_returns(list(fdef.body), [0])
    tr = _Transformer()
    tree = tr.visit(tree)
    ast.fix_missing_locations(tree)
    return tree, fdef.name, tr.uid


def get_code(fn):
    """Transformed source of `fn` (what ProgramTranslator.get_code returns in the reference)."""
    fn = getattr(fn, "__func__", fn)
    tree, _, _ = _transform_source(fn)
    return ast.unparse(tree)


def convert_to_static(fn):
    """Returns a function equivalent to `fn` whose tensor-dependent control flow is expressed with the _jst runtime converters.
    Falls back to `fn` itself when the source is unavailable or nothing needed converting."""
    if getattr(fn, "__jst_converted__", False) or getattr(fn, "_not_to_static", False):
        return fn
    bound_self = getattr(fn, "__self__", None)
    raw = getattr(fn, "__func__", fn)
    try:
        tree, name, n_converted = _transform_source(raw)
    except (OSError, TypeError, SyntaxError, StopIteration, IndentationError):
        return fn
    if n_converted == 0 and "convert_logical" not in ast.dump(tree):
        return fn
    import sys

    this = sys.modules[__name__]
    glb = dict(raw.__globals__)
    glb["__jst"] = this
    # closure variables of the original become globals of the rebuilt function (read-only view, like the reference's converter)
    if raw.__closure__:
        for nm, cell in zip(raw.__code__.co_freevars, raw.__closure__):
            try:
                glb[nm] = cell.cell_contents
            except ValueError:
                pass
    try:
        code = compile(tree, filename=f"<dy2static {raw.__qualname__}>", mode="exec")
        loc = {}
        exec(code, glb, loc)
        new = loc[name]
    except (SyntaxError, KeyError):      # e.g. a branch assigns a name declared `global`: keep the original function
        return fn
    new = functools.wraps(raw)(new)
    new.__jst_converted__ = True
    new.__jst_source__ = ast.unparse(tree)
    if raw.__defaults__:
        new.__defaults__ = raw.__defaults__
    if raw.__kwdefaults__:
        new.__kwdefaults__ = dict(raw.__kwdefaults__)
    return new.__get__(bound_self, type(bound_self)) if bound_self is not None else new


class ProgramTranslator:
    """Parity shim of paddle.jit.dy2static.program_translator.ProgramTranslator (singleton, enable / get_code / get_func)."""

    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
            cls._inst.enable_to_static = True
        return cls._inst

    @classmethod
    def get_instance(cls):
        return cls()

    def enable(self, flag):
        from . import enable_to_static

        self.enable_to_static = bool(flag)
        enable_to_static(flag)

    def get_code(self, dygraph_func):
        return get_code(dygraph_func)

    def get_func(self, dygraph_func):
        return convert_to_static(dygraph_func)

    def get_output(self, dygraph_func, *args, **kwargs):
        return convert_to_static(dygraph_func)(*args, **kwargs)
