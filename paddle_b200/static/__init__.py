"""paddle.static. Parity: python/paddle/static/__init__.py, python/paddle/base/{framework,executor}.py.

Design: a ``Program`` is a recorded op tape, not a protobuf ProgramDesc.  Under ``program_guard`` every tensor op that
touches a program Variable is executed once on placeholder data (shape inference by example) and recorded as a node
``(callable, argument refs) -> outputs``.  ``Executor.run(program, feed, fetch_list)`` binds the feed tensors to the
placeholders and replays the tape through the native scheduler (csrc/runtime/graph_exec.cpp GraphExecutor: topological
order from data dependencies); ``optimizer.minimize(loss)`` inside the guard appends a backward+update node, so training
programs work.  Parameters created by ``static.nn`` / ``create_parameter`` live in the global scope across runs.
"""
from __future__ import annotations

import contextlib
import os
import pickle

import numpy as np
import torch
from torch.overrides import TorchFunctionMode

from ..framework import dtype as _dt
from ..framework.io import load as _pload
from ..framework.io import save as _psave
from ..framework.place import CPUPlace, CUDAPlace  # noqa: F401
from ..nn.layer import ParamAttr  # noqa: F401
from ..tensor import Parameter, Tensor, to_tensor
from .input import InputSpec  # noqa: F401

_static_mode = [False]


def enable_static():
    _static_mode[0] = True


def disable_static(place=None):
    _static_mode[0] = False


def in_dynamic_mode():
    return not _static_mode[0]


def in_static_mode():
    return _static_mode[0]


class Variable(Tensor):
    """A program variable: a Tensor carrying its symbolic declaration (name, shape with -1)."""

    @property
    def desc_shape(self):
        return self.__dict__.get("_pd_desc_shape", list(self.size()))

    @property
    def shape(self):
        """Declared shape: dynamic dimensions read -1 (their run-time extent is `paddle.shape(x)[i]`), as in the reference's static mode.
        Python code that bakes `x.shape[0]` of the placeholder into the program would silently fix the batch size."""
        return list(self.desc_shape)


class _Node:
    __slots__ = ("fn", "args", "kwargs", "outs", "kind")

    def __init__(self, fn, args, kwargs, outs, kind="op"):
        self.fn, self.args, self.kwargs, self.outs, self.kind = fn, args, kwargs, outs, kind

    # Saved programs pickle their nodes.  Tensor method descriptors / builtins recorded through __torch_function__ (e.g. `Tensor.pow` from
    # `x ** 2`) are not picklable by reference, so they travel by name and are looked up again on load.
    def __getstate__(self):
        fn = self.fn
        try:
            pickle.dumps(fn)
            enc = ("obj", fn)
        except Exception:  # noqa: BLE001
            inner = getattr(fn, "__wrapped__", None)         # a recordable wrapper made at call time: at run time its body is all that matters
            if inner is not None:
                try:
                    pickle.dumps(inner)
                    return (("obj", inner), self.args, self.kwargs, self.outs, self.kind)
                except Exception:  # noqa: BLE001
                    pass
            name = getattr(fn, "__name__", None)
            objclass = getattr(fn, "__objclass__", None)
            owner = None
            if name and objclass is not None and issubclass(torch.Tensor, objclass) and hasattr(torch.Tensor, name):
                owner = "tensor"                 # method descriptor of Tensor / TensorBase (x.mean(), ...)
            elif name:
                for cand in (name, f"__{name}__", f"__r{name}__", f"__i{name}__"):      # python wrappers such as Tensor.__pow__ (`x ** 2`)
                    if getattr(torch.Tensor, cand, None) is fn:
                        owner, name = "tensor", cand
                        break
            if owner is None and name and getattr(torch, name, None) is fn:
                owner = "torch"
            elif owner is None and name and getattr(torch.nn.functional, name, None) is fn:
                owner = "functional"
            if owner is None:
                raise pickle.PicklingError(f"program node '{name or fn}' cannot be saved: its function is neither importable nor a torch op")
            enc = (owner, name)
        return (enc, self.args, self.kwargs, self.outs, self.kind)

    def __setstate__(self, state):
        enc, self.args, self.kwargs, self.outs, self.kind = state
        kind, v = enc
        self.fn = v if kind == "obj" else getattr({"tensor": torch.Tensor, "torch": torch, "functional": torch.nn.functional}[kind], v)


class _Ref:
    __slots__ = ("vid",)

    def __init__(self, vid):
        self.vid = vid


class Program:
    def __init__(self):
        self.nodes = []
        self.placeholders = {}       # name -> vid
        self._vids = {}              # id(tensor) -> vid  (only valid while recording)
        self._keep = []              # keep recorded example tensors alive so ids stay unique
        self._next = 0
        self.random_seed = 0
        self._fetch_alias = {}       # id(example tensor) -> vid (for fetch_list lookups after recording)
        self._name2vid = {}
        try:
            _live_programs.add(self)     # global_scope().find_var(name) looks parameters up through the live programs
        except NameError:                # module still initialising
            pass

    # ---- recording -----------------------------------------------------------------------------------------
    def _new_vid(self, t, name=None):
        vid = self._next
        self._next += 1
        self._vids[id(t)] = vid
        self._fetch_alias[id(t)] = vid
        self._keep.append(t)
        if name:
            self._name2vid[name] = vid
        return vid

    def _encode(self, x):
        if isinstance(x, torch.Tensor):
            vid = self._vids.get(id(x))
            return _Ref(vid) if vid is not None else x   # unknown tensors are constants / parameters (by reference)
        if isinstance(x, (list, tuple)):
            return type(x)(self._encode(i) for i in x)
        if isinstance(x, dict):
            return {k: self._encode(v) for k, v in x.items()}
        return x

    def _touches_program(self, x):
        if isinstance(x, torch.Tensor):
            return id(x) in self._vids
        if isinstance(x, (list, tuple)):
            return any(self._touches_program(i) for i in x)
        if isinstance(x, dict):
            return any(self._touches_program(v) for v in x.values())
        return False

    def _record(self, fn, args, kwargs, out):
        outs = []

        def reg(o):
            if isinstance(o, torch.Tensor):
                if id(o) not in self._vids:
                    self._new_vid(o)
                outs.append(self._vids[id(o)])
            elif isinstance(o, (list, tuple)):
                for i in o:
                    reg(i)

        enc_a, enc_k = self._encode(args), self._encode(kwargs)
        reg(out)
        self.nodes.append(_Node(fn, enc_a, enc_k, outs))

    # ---- api ------------------------------------------------------------------------------------------------------
    def global_block(self):
        return self

    @staticmethod
    def _eval_node(n):
        """The inference twin of a node: ops with a `training` switch (dropout family, batch / instance norm, rrelu) get it turned off."""
        import inspect

        if n.kind != "op":
            return n
        if n.kwargs.get("training") is True or n.kwargs.get("train") is True:
            kw = dict(n.kwargs)
            kw["training" if "training" in kw else "train"] = False
            return _Node(n.fn, n.args, kw, n.outs, n.kind)
        try:
            params = list(inspect.signature(n.fn).parameters)
        except (TypeError, ValueError):
            return n
        if "training" in params:
            i = params.index("training")
            if i < len(n.args) and n.args[i] is True:
                args = list(n.args)
                args[i] = False
                return _Node(n.fn, type(n.args)(args), n.kwargs, n.outs, n.kind)
            if i >= len(n.args) and "training" not in n.kwargs:
                default = inspect.signature(n.fn).parameters["training"].default
                if default is True:
                    return _Node(n.fn, n.args, dict(n.kwargs, training=False), n.outs, n.kind)
        return n

    def clone(self, for_test=False):
        p = Program()
        p.nodes = [self._eval_node(n) if for_test else n for n in self.nodes if not (for_test and n.kind == "train")]
        p.placeholders = dict(self.placeholders)
        p._fetch_alias, p._name2vid, p._next, p._keep = dict(self._fetch_alias), dict(self._name2vid), self._next, list(self._keep)
        return p

    def all_parameters(self):
        seen, out = set(), []

        def walk(x):
            if isinstance(x, Parameter) and id(x) not in seen:
                seen.add(id(x))
                out.append(x)
            elif isinstance(x, (list, tuple)):
                for i in x:
                    walk(i)
            elif isinstance(x, dict):
                for v in x.values():
                    walk(v)

        for n in self.nodes:
            walk(n.args)
            walk(n.kwargs)
            if n.kind == "train":
                for p in n.fn.__self__._parameter_list if hasattr(n.fn, "__self__") else []:
                    walk(p)
        return out

    def list_vars(self):
        return list(self._keep)

    def var(self, name):
        return self._keep[self._name2vid[name]] if name in self._name2vid else None

    @property
    def num_blocks(self):
        return 1

    # ---- introspection (reference: Program.global_block().ops, op.type / input_arg_names / output_arg_names, Program.to_string) ------------
    @property
    def idx(self):
        return 0

    @property
    def blocks(self):
        return [self]

    def block(self, index=0):
        return self

    def _value_name(self, vid):
        t = self._keep[vid] if 0 <= vid < len(self._keep) else None
        return getattr(t, "name", None) or f"tmp_{vid}"

    @property
    def ops(self):
        prog = self

        class _OpView:
            def __init__(self, idx, node):
                self.idx, self._n = idx, node
                fn = node.fn
                self.type = {"train": "backward_and_update", "control": "control_flow"}.get(node.kind) or (getattr(fn, "__name__", None) or str(fn)).strip("_")

            def _refs(self):
                out = set()
                _ValueGC._refs(self._n.args, out)
                _ValueGC._refs(self._n.kwargs, out)
                return sorted(out)

            @property
            def input_arg_names(self):
                return [prog._value_name(v) for v in self._refs()]

            @property
            def output_arg_names(self):
                return [prog._value_name(v) for v in self._n.outs]

            def attr(self, name):
                return self._n.kwargs.get(name)

            def all_attrs(self):
                return {k: v for k, v in self._n.kwargs.items() if not isinstance(v, (_Ref, torch.Tensor))}

            def __repr__(self):
                return f"{{Out={self.output_arg_names}}} = {self.type}(inputs={self.input_arg_names})"

        return [_OpView(i, n) for i, n in enumerate(self.nodes)]

    def to_string(self, throw_on_error=False, with_details=False):
        lines = [f"{{ // block 0  ({len(self.nodes)} ops, feeds: {list(self.placeholders)})"]
        for name, vid in self.placeholders.items():
            t = self._keep[vid]
            lines.append(f"    var {name} : shape{getattr(t, 'desc_shape', list(t.shape))} dtype({str(t.dtype).replace('torch.', '')})")
        for p in self.all_parameters():
            lines.append(f"    persist trainable param {p.name} : shape{list(p.shape)} dtype({str(p.dtype).replace('torch.', '')})")
        lines += [f"    {op!r}" for op in self.ops]
        return "\n".join(lines + ["}"])

    def __str__(self):
        return self.to_string()

    def __repr__(self):
        return f"Program(nodes={len(self.nodes)}, feeds={list(self.placeholders)})"


_main = [Program()]
_startup = [Program()]
from ..framework.recording import current as _recording  # noqa: E402  (shared with the kernel wrappers' @recordable)


def default_main_program():
    return _main[0]


def default_startup_program():
    return _startup[0]


class _Recorder(TorchFunctionMode):
    def __init__(self, program):
        super().__init__()
        self.program = program

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        out = func(*args, **kwargs)
        p = self.program
        if p._touches_program(args) or p._touches_program(kwargs):
            name = getattr(func, "__name__", "")
            if name not in ("__get__", "size", "dim", "numel", "stride", "is_contiguous", "data_ptr", "__len__", "is_floating_point", "is_complex", "element_size", "_is_view", "storage_offset"):
                p._record(func, args, kwargs, out)
        return out


@contextlib.contextmanager
def program_guard(main_program, startup_program=None):
    prev_m, prev_s, prev_r = _main[0], _startup[0], _recording[0]
    _main[0] = main_program
    if startup_program is not None:
        _startup[0] = startup_program
    rec = _Recorder(main_program)
    _recording[0] = main_program
    try:
        with rec:
            yield
    finally:
        _main[0], _startup[0], _recording[0] = prev_m, prev_s, prev_r


def data(name, shape, dtype="float32", lod_level=0):
    prog = _main[0]
    ex_shape = [1 if (s is None or s < 0) else int(s) for s in shape]
    d = _dt.convert_dtype(dtype) or torch.float32
    with torch._C.DisableTorchFunction():
        t = torch.zeros(ex_shape, dtype=d)
    v = t.as_subclass(Variable)
    v.name = name
    v.__dict__["_pd_desc_shape"] = [-1 if (s is None or s < 0) else int(s) for s in shape]
    prog.placeholders[name] = prog._new_vid(v, name)
    return v


def create_parameter(shape, dtype, name=None, attr=None, is_bias=False, default_initializer=None):
    from ..nn.layer import _make_parameter

    return _make_parameter(list(shape), _dt.convert_dtype(dtype), attr=attr, is_bias=is_bias, default_initializer=default_initializer, name=name)


def create_global_var(shape, value, dtype, persistable=False, force_cpu=False, name=None):
    t = torch.full(list(shape), value, dtype=_dt.convert_dtype(dtype)).as_subclass(Tensor)
    t.persistable = persistable
    if name:
        t.name = name
    if persistable and name:
        _global_vars.append(t)
    return t


def _minimize_node(optimizer, loss, program):
    """Appended by Optimizer.minimize in static mode: replayed as backward + update."""
    vid = program._vids.get(id(loss), program._fetch_alias.get(id(loss)))

    def train_step(env):
        l = env[vid]
        optimizer.clear_grad()
        l.backward()
        optimizer.step()

    n = _Node(train_step, (), {}, [], kind="train")
    n.args = (optimizer,)
    if vid is not None:
        n.kwargs = {"loss": _Ref(vid)}          # declared read: keeps the loss alive (and re-bound to this slot) when the program goes through IR passes
    program.nodes.append(n)


def append_backward(loss, parameter_list=None, no_grad_set=None, callbacks=None):
    prog = _main[0]
    vid = prog._vids.get(id(loss))
    params = list(parameter_list) if parameter_list is not None else prog.all_parameters()

    def bwd(env):
        for p in params:
            p.clear_grad()
        env[vid].backward(retain_graph=True)

    n = _Node(bwd, (), {}, [], kind="train")
    if vid is not None:
        n.kwargs = {"loss": _Ref(vid)}
    prog.nodes.append(n)
    return [(p, p) for p in params]


def gradients(targets, inputs, target_gradients=None, no_grad_set=None):
    from ..autograd import grad

    return grad(targets, inputs, target_gradients, allow_unused=True)


class _ValueGC:
    """Garbage collection of the executor's value table (the interpreter GC of the reference, new_executor/garbage_collector): every value
    carries the number of nodes that still have to read it; when the last reader has run the entry is dropped, so the tensor's memory goes back to
    the allocator while the program is still running.  Counting (rather than last-use positions) stays correct under any dependency-respecting
    execution order.  Programs with training / control nodes read the table through closures: no collection there.  FLAGS_eager_delete_tensor_gb
    < 0 turns it off."""

    def __init__(self, readers, keep, reads):
        self.readers, self.keep, self.reads = readers, keep, reads
        self.stats = {"freed": 0, "peak_live": 0}

    @staticmethod
    def _refs(x, out):
        if isinstance(x, _Ref):
            out.add(x.vid)
        elif isinstance(x, (list, tuple)):
            for i in x:
                _ValueGC._refs(i, out)
        elif isinstance(x, dict):
            for i in x.values():
                _ValueGC._refs(i, out)

    @staticmethod
    def plan(program, fetch_list, fed):
        if any(n.kind != "op" for n in program.nodes):
            return None
        keep = set(fed)
        for f in fetch_list or []:
            vid = program._name2vid.get(f) if isinstance(f, str) else program._fetch_alias.get(id(f))
            if vid is not None:
                keep.add(vid)
        keep.update(program.__dict__.get("_pir_alias", {}).values())
        readers, reads = {}, {}
        for n in program.nodes:
            r = set()
            _ValueGC._refs(n.args, r)
            _ValueGC._refs(n.kwargs, r)
            reads[id(n)] = r
            for v in r:
                readers[v] = readers.get(v, 0) + 1
        return _ValueGC(readers, keep, reads)

    def after(self, node, env):
        self.stats["peak_live"] = max(self.stats["peak_live"], len(env))
        for v in self.reads.get(id(node), ()):
            c = self.readers.get(v, 0) - 1
            self.readers[v] = c
            if c == 0 and v not in self.keep and v in env:
                del env[v]
                self.stats["freed"] += 1
        for v in node.outs:                              # produced but never read and not fetched
            if self.readers.get(v, 0) == 0 and v not in self.keep and v in env:
                del env[v]
                self.stats["freed"] += 1


class Executor:
    def __init__(self, place=None):
        self.place = place

    def run(self, program=None, feed=None, fetch_list=None, feed_var_name="feed", fetch_var_name="fetch", scope=None, return_numpy=True,
            use_program_cache=False, use_prune=False):
        program = program if program is not None else _main[0]
        if isinstance(program, CompiledProgram):
            program = program._program
        if not program.nodes and not fetch_list:
            return []   # startup program: parameters are initialised at creation time
        from ..framework.flags import flag

        feed = feed or {}
        if isinstance(feed, (list, tuple)):                          # one dict per place (DataLoader(return_list=False) on several places): this process runs the first
            feed = feed[0] if feed else {}
        on_run = program.__dict__.get("_cinn_on_run")                 # set by inference.Config.enable_cinn(): specialise per feed signature
        if program.nodes and (on_run or (flag("FLAGS_enable_pir_api", False) and not program.__dict__.get("_pir_report"))):
            if on_run or flag("FLAGS_use_cinn", False):
                hit = self._specialised(program, fetch_list, feed, return_numpy, passes=[] if on_run else None)
                if isinstance(hit, tuple):                              # first run with this signature: the generic program's results
                    return hit[1]
                program = hit
            else:
                program = self._pir_optimized(program, fetch_list)
        return self._execute(program, feed, fetch_list, return_numpy)[0]

    def _execute(self, program, feed, fetch_list, return_numpy, keep_env=False):
        """One pass over the program's nodes.  Returns (fetched values, value table when keep_env)."""
        from ..framework.flags import flag

        env = {}
        dev = None
        for name, vid in program.placeholders.items():
            if name in feed:
                v = feed[name]
                t = v if isinstance(v, torch.Tensor) else to_tensor(np.asarray(v))
                env[vid] = t
                dev = t.device

        def decode(x):
            if isinstance(x, _Ref):
                if x.vid not in env:
                    raise KeyError(f"program variable #{x.vid} is not available: is a feed missing? feeds={list(program.placeholders)}")
                return env[x.vid]
            if isinstance(x, (list, tuple)):
                return type(x)(decode(i) for i in x)
            if isinstance(x, dict):
                return {k: decode(v) for k, v in x.items()}
            return x

        def run_node(n):
            if n.kind in ("train", "control"):      # backward + update / run-time control flow: the node works on the value table
                n.fn(env)
                return
            try:
                out = n.fn(*decode(n.args), **decode(n.kwargs))
            except RuntimeError as e:
                if "shape" in str(e) and any(-1 in getattr(program._keep[v], "desc_shape", []) for v in program.placeholders.values()):
                    name = (getattr(n.fn, "__name__", None) or str(n.fn)).strip("_")
                    raise RuntimeError(f"op '{name}' of the program failed on this feed: {e}.  The program declares dynamic dims (-1); it was recorded with extent 1 "
                                       "for them, and python code that read a tensor's size while the program was built froze that 1 into an argument.  Use "
                                       "reshape([-1, ...]) / reshape_as / paddle.shape(x) for run-time extents, or declare the dims with their real sizes.") from e
                raise
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
            if gc is not None:                          # values nobody will read any more leave the table right away
                gc.after(n, env)

        gc = _ValueGC.plan(program, fetch_list, set(env)) if (flag("FLAGS_eager_delete_tensor_gb", 0.0) >= 0 and not keep_env) else None
        self._schedule(program, run_node)
        self.last_gc_stats = gc.stats if gc is not None else None
        outs = []
        for f in fetch_list or []:
            if isinstance(f, str):
                vid = program._name2vid[f]
            else:
                vid = program._fetch_alias.get(id(f))
                if vid is None:
                    if isinstance(f, torch.Tensor) and program.placeholders and not getattr(f, "persistable", False) and not isinstance(f, torch.nn.Parameter) \
                            and not getattr(f, "is_parameter", False):
                        import warnings

                        warnings.warn("Executor.run: a fetch target is not a value of this program (it was computed outside the recorded ops, e.g. on raw tensors); "
                                      "its trace-time value is returned and does not depend on the feeds", stacklevel=3)
                    outs.append(f.numpy() if return_numpy and isinstance(f, torch.Tensor) else f)
                    continue
            v = env[vid]
            outs.append(v.detach().cpu().as_subclass(Tensor).numpy() if return_numpy else v)
        return outs, (env if keep_env else None)

    def _specialised(self, program, fetch_list, feed, return_numpy, passes=None):
        """Generated kernels are compiled for concrete shapes, a program may declare dynamic ones (-1: its recorded example uses extent 1).  So the
        program is specialised per FEED SIGNATURE: the first run with a signature executes the generic program (its results are returned) and keeps
        every value; their actual types go into the IR, the passes + kernel fusion run on that, and later runs with the same signature use the
        result.  Kernels whose source does not depend on the leading extents share one compiled object across signatures (cinn/codegen.py).
        Returns the specialised program, or (None, outputs) on the first run."""
        import copy

        from .. import pir

        if not pir.core_available() or any(n.kind not in ("op", "control", "train") for n in program.nodes) \
                or any(n.kind == "train" and not n.kwargs for n in program.nodes):
            return program
        sig = tuple(sorted((name, tuple(v.shape), str(getattr(v, "dtype", ""))) for name, v in feed.items() if name in program.placeholders))
        key = (tuple(f if isinstance(f, str) else id(f) for f in (fetch_list or [])), sig, len(program.nodes))
        cache = program.__dict__.setdefault("_cinn_cache", {})
        if key in cache:
            return cache[key]
        outs, env = self._execute(program, feed, fetch_list, return_numpy, keep_env=True)
        try:
            shaped = copy.copy(program)
            shaped.__dict__ = dict(program.__dict__)
            shaped.__dict__.pop("_pir_report", None)
            shaped._keep = [env.get(vid, t) for vid, t in enumerate(program._keep)]
            cache[key] = pir.optimize(shaped, fetch_list=fetch_list, passes=passes, cinn=True)
        except Exception:  # noqa: BLE001  (an op the translator cannot encode, a kernel that cannot be generated: run as recorded)
            import os

            if os.environ.get("B200_JIT_DEBUG"):
                raise
            cache[key] = program
        return None, outs

    @staticmethod
    def _pir_optimized(program, fetch_list):
        """FLAGS_enable_pir_api: the program runs through the native IR pass pipeline (paddle_b200.pir: DCE / CSE / identity removal / constant
        folding / fusion patterns) once per fetch set; the lowered program is cached on the source program."""
        from .. import pir

        if not pir.core_available() or any(n.kind not in ("op", "control", "train") for n in program.nodes):
            return program
        if any(n.kind == "train" and not n.kwargs for n in program.nodes):
            return program          # a training node that does not declare what it reads (older pickles): keep the recorded form
        key = tuple(f if isinstance(f, str) else id(f) for f in (fetch_list or []))
        cache = program.__dict__.setdefault("_pir_cache", {})
        if key not in cache or cache[key][0] != len(program.nodes):
            try:
                cache[key] = (len(program.nodes), pir.optimize(program, fetch_list=fetch_list, cinn=False))      # shape-independent passes only
            except Exception:  # noqa: BLE001  (an op the translator cannot encode: run as recorded)
                cache[key] = (len(program.nodes), program)
        return cache[key][1]

    def train_from_dataset(self, program=None, dataset=None, scope=None, thread=0, debug=False, fetch_list=None, fetch_info=None, print_period=100,
                           fetch_handler=None):
        """Run `program` once per batch of a slot dataset (distributed.InMemoryDataset / QueueDataset): each batch dict feeds the
        placeholders named after the dataset's `use_var` entries. Parity: base/executor.py:Executor.train_from_dataset."""
        return self._run_from_dataset(program, dataset, fetch_list, fetch_info, print_period, debug)

    def infer_from_dataset(self, program=None, dataset=None, scope=None, thread=0, debug=False, fetch_list=None, fetch_info=None, print_period=100,
                           fetch_handler=None):
        program = program if program is not None else _main[0]
        prog = program._program if isinstance(program, CompiledProgram) else program
        return self._run_from_dataset(prog.clone(for_test=True), dataset, fetch_list, fetch_info, print_period, debug)

    def _run_from_dataset(self, program, dataset, fetch_list, fetch_info, print_period, debug):
        if dataset is None:
            raise RuntimeError("dataset is need and should be initialized")
        program = program if program is not None else _main[0]
        results = []
        for step, batch in enumerate(dataset):
            if isinstance(batch, dict):
                feed = {}
                for k, v in batch.items():
                    if isinstance(v, tuple):            # ragged slot: (values, lod) -> LoD tensor of packed rows
                        t = to_tensor(np.asarray(v[0]).reshape(-1, 1))
                        t.set_lod([[int(o) for o in v[1]]])
                        feed[k] = t
                    else:
                        feed[k] = v
            else:
                names = list((program._program if isinstance(program, CompiledProgram) else program).placeholders)
                feed = {names[0]: batch} if names else {}
            out = self.run(program, feed=feed, fetch_list=fetch_list)
            if fetch_list:
                results.append(out)
                if debug or (print_period and step % print_period == 0 and fetch_info):
                    print(" ".join(f"{n}: {np.asarray(o).reshape(-1)[:4]}" for n, o in zip(fetch_info or [], out)))
        return results

    def flush(self):
        pass

    def close(self):
        pass

    @staticmethod
    def _schedule(program, run_node):
        """Replay in dependency order through the native GraphExecutor when available (falls back to tape order)."""
        from .._build import load

        m = load()
        nodes = program.nodes
        if m is None or not hasattr(m, "GraphExecutor") or len(nodes) < 2:
            for n in nodes:
                run_node(n)
            return
        ge = m.GraphExecutor()
        producer = {}
        last_train = -1
        for i, n in enumerate(nodes):
            deps = set()

            def collect(x):
                if isinstance(x, _Ref) and x.vid in producer:
                    deps.add(producer[x.vid])
                elif isinstance(x, (list, tuple)):
                    for j in x:
                        collect(j)
                elif isinstance(x, dict):
                    for v in x.values():
                        collect(v)

            collect(n.args)
            collect(n.kwargs)
            if n.kind == "train":   # side effects: order after everything recorded before it
                deps.update(range(i))
            elif last_train >= 0:
                deps.add(last_train)
            if i > 0 and not n.outs:
                deps.add(i - 1)
            ge.add_node((lambda nn=n: run_node(nn)), sorted(deps), 0, 0)
            for vid in n.outs:
                producer[vid] = i
            if n.kind == "train":
                last_train = i
        ge.run(-1)

    def close(self):
        pass


class CompiledProgram:
    def __init__(self, program_or_graph, build_strategy=None):
        self._program = program_or_graph
        self._build_strategy = build_strategy
        if getattr(build_strategy, "build_cinn_pass", False) and hasattr(program_or_graph, "nodes"):
            # BuildStrategy.build_cinn_pass: pointwise / reduction chains run as generated kernels, specialised per feed signature by the Executor
            program_or_graph.__dict__["_cinn_on_run"] = True

    def with_data_parallel(self, *a, **k):
        return self


class BuildStrategy:
    def __init__(self):
        self.fuse_all_reduce_ops = self.fuse_elewise_add_act_ops = self.fuse_bn_act_ops = self.enable_inplace = False
        self.memory_optimize = self.build_cinn_pass = False


class ExecutionStrategy:
    def __init__(self):
        self.num_threads, self.num_iteration_per_drop_scope = 1, 10


class _TensorView:
    """What `scope.find_var(name).get_tensor()` hands out (the LoDTensor handle of the reference): reads and writes go to the live tensor, so
    `np.array(t)` sees the current parameter value and `t.set(array, place)` changes what the next Executor.run computes with."""

    def __init__(self, holder):
        self._h = holder

    def _t(self):
        t = self._h._tensor
        if t is None:
            raise RuntimeError(f"variable '{self._h.name}' holds no tensor yet")
        return t

    def __array__(self, dtype=None, copy=None):
        a = self._t().detach().cpu().as_subclass(torch.Tensor)
        a = a.float().numpy() if a.dtype == torch.bfloat16 else a.numpy()
        return a.astype(dtype) if dtype is not None else a

    def set(self, value, place=None):
        src = value if isinstance(value, torch.Tensor) else torch.as_tensor(np.asarray(value))
        t = self._h._tensor
        if t is None:
            self._h._tensor = src.clone().as_subclass(Tensor)
            return
        if tuple(src.shape) != tuple(t.shape):
            raise ValueError(f"set: shape {list(src.shape)} does not match variable '{self._h.name}' {list(t.shape)}")
        with torch.no_grad():
            t.as_subclass(torch.Tensor).copy_(src.to(t.dtype))

    def shape(self):
        return list(self._t().shape)

    def _dtype(self):
        return self._t().dtype

    def _place(self):
        return self._t().place if hasattr(self._t(), "place") else CPUPlace()

    def lod(self):
        return []

    def recursive_sequence_lengths(self):
        return []

    def set_lod(self, lod):
        pass

    def set_recursive_sequence_lengths(self, lens):
        pass

    def _is_initialized(self):
        return self._h._tensor is not None


class _ScopeVar:
    def __init__(self, name, tensor=None):
        self.name, self._tensor = name, tensor

    def get_tensor(self):
        return _TensorView(self)

    def get_value(self):
        return self._tensor

    def set_value(self, value):
        _TensorView(self).set(value)


class _Scope:
    """Name -> variable table (reference: paddle/fluid/framework/scope.h, `paddle.static.global_scope()`).  Persistable variables - the parameters
    and global vars of every live Program - are visible in the global scope by name; `var(name)` creates a local one; child scopes chain up."""

    def __init__(self, parent=None):
        self.vars = {}
        self._parent = parent
        self._kids = []

    def var(self, name=None):
        name = name or f"_generated_var_{len(self.vars)}"
        if name not in self.vars:
            self.vars[name] = _ScopeVar(name)
        return self.vars[name]

    def find_var(self, name):
        v = self.vars.get(name)
        if v is not None:
            return v
        if self._parent is not None:
            return self._parent.find_var(name)
        if self is _root_scope:
            t = _find_persistable(name)
            if t is not None:
                v = self.vars[name] = _ScopeVar(name, t)
                return v
        return None

    def erase(self, names):
        for n in names:
            self.vars.pop(n, None)

    def local_var_names(self):
        return list(self.vars)

    def new_scope(self):
        s = _Scope(parent=self)
        self._kids.append(s)
        return s

    def drop_kids(self):
        self._kids.clear()


def _find_persistable(name):
    """A parameter / persistable global var of any live Program (or of the default programs), by name."""
    for prog in list(_live_programs):
        for p in prog.all_parameters():
            if getattr(p, "name", None) == name:
                return p
    for t in list(_global_vars):
        if getattr(t, "name", None) == name:
            return t
    return None


import weakref as _weakref  # noqa: E402

_live_programs = _weakref.WeakSet()
_global_vars = []
_root_scope = _Scope()
_scope = _root_scope
Scope = _Scope


def global_scope():
    return _scope


@contextlib.contextmanager
def scope_guard(scope):
    global _scope
    prev = _scope
    _scope = scope
    try:
        yield
    finally:
        _scope = prev


@contextlib.contextmanager
def name_scope(prefix=None):
    yield


@contextlib.contextmanager
def device_guard(device=None):
    yield


def cpu_places(device_count=None):
    return [CPUPlace()] * (device_count or 1)


def cuda_places(device_ids=None):
    ids = device_ids if device_ids is not None else range(max(1, torch.cuda.device_count()))
    return [CUDAPlace(i) for i in ids]


def xpu_places(device_ids=None):
    return []


def save(program, model_path, protocol=4, **configs):
    params = {p.name: p for p in program.all_parameters()}
    _psave(params, model_path + ".pdparams", protocol)
    with open(model_path + ".pdmodel", "wb") as f:
        pickle.dump({"feeds": list(program.placeholders), "n_nodes": len(program.nodes)}, f)


def load(program, model_path, executor=None, var_list=None):
    sd = _pload(model_path + ".pdparams")
    for p in program.all_parameters():
        if p.name in sd:
            p.set_value(sd[p.name])


def load_program_state(model_path, var_list=None):
    return _pload(model_path + ".pdparams", return_numpy=True)


def set_program_state(program, state_dict):
    for p in program.all_parameters():
        if p.name in state_dict:
            p.set_value(state_dict[p.name])


def save_inference_model(path_prefix, feed_vars, fetch_vars, executor, program=None, **kwargs):
    program = program or _main[0]
    feed_vars = feed_vars if isinstance(feed_vars, (list, tuple)) else [feed_vars]
    fetch_vars = fetch_vars if isinstance(fetch_vars, (list, tuple)) else [fetch_vars]
    d = os.path.dirname(path_prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    blob = _inference_blob(feed_vars, fetch_vars, program)
    with open(path_prefix + ".pdmodel", "wb") as f:
        pickle.dump(blob, f)


def load_inference_model(path_prefix, executor, **kwargs):
    with open(path_prefix + ".pdmodel", "rb") as f:
        blob = pickle.load(f)
    prog = blob["program"]
    prog._fetch_alias = {id(t): vid for vid, t in enumerate(prog._keep)}   # object ids changed across the pickle round trip
    fetch = [prog._keep[v] for v in blob["fetch_vids"]]
    return prog, blob["feeds"], fetch


def normalize_program(program, feed_vars, fetch_vars, **kwargs):
    return program.clone(for_test=True)


def _inference_blob(feed_vars, fetch_vars, program):
    feed_vars = feed_vars if isinstance(feed_vars, (list, tuple)) else [feed_vars]
    fetch_vars = fetch_vars if isinstance(fetch_vars, (list, tuple)) else [fetch_vars]
    infer = program.clone(for_test=True)
    fetch_vids = [program._fetch_alias[id(v)] for v in fetch_vars]
    from . import passes as _passes

    _passes.dead_code_elimination(infer, keep=set(fetch_vids))      # prune to what the fetch targets need (drops the loss sub-graph)
    return {"program": infer, "feeds": [v.name for v in feed_vars], "fetch_vids": fetch_vids}


def serialize_program(feed_vars, fetch_vars, **kwargs):
    """Bytes of the pruned inference program (same content as the `.pdmodel` written by save_inference_model)."""
    return pickle.dumps(_inference_blob(feed_vars, fetch_vars, kwargs.get("program") or _main[0]))


def serialize_persistables(feed_vars, fetch_vars, executor, **kwargs):
    return pickle.dumps({p.name: p.numpy() for p in (kwargs.get("program") or _main[0]).all_parameters()})


def deserialize_program(data):
    """Program from serialize_program bytes; the feed names / fetch variables ride along as `_feed_names` / `_fetch_vars`."""
    blob = pickle.loads(data)
    prog = blob["program"]
    prog._fetch_alias = {id(t): vid for vid, t in enumerate(prog._keep)}
    prog._feed_names, prog._fetch_vars = blob["feeds"], [prog._keep[v] for v in blob["fetch_vids"]]
    return prog


def deserialize_persistables(program, data, executor):
    set_program_state(program, pickle.loads(data))


def save_to_file(path, content):
    with open(path, "wb") as f:
        f.write(content)


def load_from_file(path):
    with open(path, "rb") as f:
        return f.read()


class _PrintOp:
    """The body of a `Print` node (a picklable object, so programs with prints can be saved)."""

    __name__ = "print"

    def __init__(self, first_n, message, summarize, show_type, show_shape):
        self.first_n, self.message, self.summarize, self.show_type, self.show_shape, self.n = first_n, message, summarize, show_type, show_shape, 0

    def __call__(self, t):
        from ..framework import recording

        if recording._inside[0] == 0 and (self.first_n < 0 or self.n < self.first_n):      # _inside > 0: the recorder is only taking the example output
            self.n += 1
            r = t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else t
            head = [self.message or ""]
            if self.show_shape and isinstance(r, torch.Tensor):
                head.append(f"shape={list(r.shape)}")
            if self.show_type and isinstance(r, torch.Tensor):
                head.append(f"dtype={str(r.dtype).replace('torch.', '')}")
            flat = r.detach().reshape(-1)[: max(int(self.summarize), 0) if self.summarize and self.summarize > 0 else None] if isinstance(r, torch.Tensor) else r
            print(" ".join(h for h in head if h), "data:", flat.tolist() if isinstance(flat, torch.Tensor) else flat)
        return t * 1 if isinstance(t, torch.Tensor) else t


def Print(input, first_n=-1, message=None, summarize=20, print_tensor_name=True, print_tensor_type=True, print_tensor_shape=True,
          print_tensor_layout=True, print_tensor_lod=True, print_phase="both"):
    """Prints the tensor when the op RUNS (in a program: at every Executor.run, up to `first_n` times), not while the program is being built."""
    from ..framework import recording

    return recording.recordable(_PrintOp(first_n, message, summarize, print_tensor_type, print_tensor_shape))(input)


def py_func(func, x, out, backward_func=None, skip_vars_in_backward_input=None):
    """Runs `func` on the values of `x` when the op runs.  In a program the call is ONE recorded node (its example outputs come from calling `func`
    once on the placeholders while building); the returned variables are the op's results (`out` only declares their types in the reference)."""
    from ..framework import recording

    xs = list(x) if isinstance(x, (list, tuple)) else [x]

    def _py_func_op(*vals):
        res = func(*[v.as_subclass(Tensor) if isinstance(v, torch.Tensor) and not isinstance(v, Tensor) else v for v in vals])
        conv = lambda r: r if isinstance(r, torch.Tensor) else to_tensor(np.asarray(r))      # noqa: E731
        if isinstance(res, (list, tuple)):
            return type(res)(conv(r) for r in res)
        return None if res is None else conv(res)

    _py_func_op.__name__ = "py_func"
    return recording.recordable(_py_func_op)(*xs)


def accuracy(input, label, k=1, correct=None, total=None):
    from ..metric import accuracy as acc

    return acc(input, label, k)


def auc(input, label, curve="ROC", num_thresholds=4095, topk=1, slide_steps=1, ins_tag_weight=None):
    """Area under the curve, accumulated over the runs of the op (the statistics live with the op, like the reference's stat variables).
    Returns (global auc, batch auc, [statistics])."""
    from ..framework import recording
    from ..metric import Auc

    total = Auc(curve, num_thresholds)

    def _auc_op(pred, lbl):
        if recording._inside[0] > 0:                      # building: types only, the placeholders carry no data
            z = to_tensor(np.asarray(0.0, dtype=np.float32))
            return z, z * 1
        batch = Auc(curve, num_thresholds)
        batch.update(pred, lbl)
        total.update(pred, lbl)
        return to_tensor(np.asarray(total.accumulate(), dtype=np.float32)), to_tensor(np.asarray(batch.accumulate(), dtype=np.float32))

    _auc_op.__name__ = "auc"
    g, b = recording.recordable(_auc_op)(input, label)
    return g, b, [total]


def ctr_metric_bundle(input, label, ins_tag_weight=None):
    """Running CTR statistics of a batch: (sqrerr, abserr, prob, q, pos_num, ins_num), each a [1] tensor that the caller
    accumulates (the reference keeps them in persistable variables). Parity: python/paddle/static/nn/metric.py:ctr_metric_bundle."""
    import torch

    p = input.as_subclass(torch.Tensor).reshape(-1).float()
    y = label.as_subclass(torch.Tensor).reshape(-1).float()
    w = torch.ones_like(p) if ins_tag_weight is None else ins_tag_weight.as_subclass(torch.Tensor).reshape(-1).float()
    outs = (((p - y) ** 2 * w).sum(), ((p - y).abs() * w).sum(), (p * w).sum(), (p * w).sum(), (y * w).sum(), w.sum())
    return tuple(o.reshape(1).as_subclass(Tensor) for o in outs)


class WeightNormParamAttr(ParamAttr):
    def __init__(self, dim=None, **kw):
        super().__init__(**kw)
        self.dim = dim


class ExponentialMovingAverage:
    """EMA_t = decay * EMA_{t-1} + (1 - decay) * theta_t, EMA_0 = 0; `apply()` swaps in the bias-corrected EMA_t / (1 - decay^t).
    In a program `update()` appends an op, so the averages move with EVERY Executor.run of the training program (call it after
    `optimizer.minimize`); in dynamic mode every call is one update.  `thres_steps` (a step-count variable / number) caps the decay at
    (1 + steps) / (10 + steps).  Parity: python/paddle/static/nn/... ExponentialMovingAverage."""

    def __init__(self, decay=0.999, thres_steps=None, name=None):
        self._decay, self._thres, self._shadow, self._backup, self._params = float(decay), thres_steps, {}, {}, None
        self._step, self._decay_pow = 0, 1.0

    def _one_update(self, env=None):
        d = self._decay
        if self._thres is not None:
            t = self._thres
            steps = float(t.as_subclass(torch.Tensor).reshape(-1)[0]) if isinstance(t, torch.Tensor) else float(t)
            d = min(d, (1.0 + steps) / (10.0 + steps))
        self._step += 1
        self._decay_pow *= d
        with torch.no_grad():
            for p in self._params:
                s = self._shadow.get(p.name)
                if s is None:
                    s = self._shadow[p.name] = torch.zeros_like(p.detach().as_subclass(torch.Tensor), dtype=torch.float32)
                s.mul_(d).add_(p.detach().as_subclass(torch.Tensor).float(), alpha=1.0 - d)

    def update(self, parameters=None):
        prog = _recording[0]
        params = list(parameters) if parameters is not None else (self._params or (prog or _main[0]).all_parameters())
        self._params = [p for p in params if not getattr(p, "stop_gradient", False)] or list(params)
        if prog is not None:                               # building a program: the update is an op of it
            prog.nodes.append(_Node(self._one_update, (), {}, [], kind="control"))
        else:
            self._one_update()

    @contextlib.contextmanager
    def apply(self, executor=None, need_restore=True):
        corr = 1.0 - self._decay_pow
        with torch.no_grad():
            for p in self._params or []:
                self._backup[p.name] = p.detach().clone()
                if p.name in self._shadow and corr > 0:
                    p.as_subclass(torch.Tensor).copy_((self._shadow[p.name] / corr).to(p.dtype))
        try:
            yield
        finally:
            if need_restore:
                self.restore()

    def restore(self, executor=None):
        with torch.no_grad():
            for p in self._params or []:
                if p.name in self._backup:
                    p.as_subclass(torch.Tensor).copy_(self._backup[p.name])


class IpuStrategy:
    def __init__(self):
        raise RuntimeError("IPU is not supported")


IpuCompiledProgram = IpuStrategy


def ipu_shard_guard(*a, **k):
    raise RuntimeError("IPU is not supported")


def set_ipu_shard(*a, **k):
    raise RuntimeError("IPU is not supported")


from . import nn  # noqa: E402,F401
