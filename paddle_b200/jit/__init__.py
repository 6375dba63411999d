"""paddle.jit. Parity: python/paddle/jit/api.py (to_static, save, load, not_to_static, enable_to_static, ignore_module),
translated_layer.py (TranslatedLayer).

B200 design ("CUDA streams and graphs instead of a tracing compiler"): by default ``to_static`` does not trace into an IR.  A
static function keeps running the eager kernels, but once its inputs have a stable signature (shapes/dtypes) the whole
call is captured into a CUDA graph and replayed with one launch; a guard cache keyed by the signature holds one graph
per shape.  ``backend="CINN"`` selects the other mode: trace -> IR passes -> generated kernels (paddle_b200.cinn), verified
against the eager function on the first call of every signature.  ``jit.save`` writes the parameters (.pdiparams, same pickle format as paddle.save) plus a .pdmodel file that
pickles the Layer's class path + constructor spec + InputSpec, so ``jit.load`` rebuilds a ``TranslatedLayer``.
"""
from __future__ import annotations

import functools
import importlib
import os
import pickle

import torch

from ..framework.io import load as _load
from ..framework.io import save as _save
from ..nn.layer import Layer
from ..static.input import InputSpec
from ..tensor import Tensor

_enabled = [True]
_ignored_modules = []


def enable_to_static(enable_to_static_bool):
    _enabled[0] = bool(enable_to_static_bool)


def ignore_module(modules):
    _ignored_modules.extend(modules if isinstance(modules, (list, tuple)) else [modules])


def not_to_static(func=None):
    if func is None:
        return not_to_static
    func._not_to_static = True
    return func


def _sig(args, kwargs):
    out = []
    for a in list(args) + [kwargs[k] for k in sorted(kwargs)]:
        if isinstance(a, torch.Tensor):
            out.append((tuple(a.shape), a.dtype, a.device.type, a.requires_grad))
        elif isinstance(a, (int, float, bool, str, type(None))):
            out.append(a)
        else:
            out.append(id(a))
    return tuple(out)


class StaticFunction:
    """Callable produced by to_static. Eager warm-up, then CUDA-graph replay for inference-mode CUDA calls."""

    def __init__(self, fn, layer=None, input_spec=None, build_strategy=None, backend=None, full_graph=False):
        self._dygraph_fn = fn
        try:      # AST conversion of tensor-dependent control flow (jit/dy2static.py): `if tensor:` becomes run-both + device select
            from .dy2static import convert_to_static

            fn = convert_to_static(fn)
        except Exception:  # noqa: BLE001  (source unavailable, exotic syntax): the function runs as written
            fn = self._dygraph_fn
        self._fn, self._layer, self._input_spec = fn, layer, input_spec
        self._backend = backend.upper() if isinstance(backend, str) else None
        if self._backend is None and getattr(build_strategy, "build_cinn_pass", False):       # BuildStrategy().build_cinn_pass = True (older spelling)
            self._backend = "CINN"
        self._cinn = {}                 # input signature -> (executor, program, feed names, fetch targets, single result?) or None
        self._graphs = {}
        self._train_graphs = {}
        self._warm = {}
        self._capture_after = 2
        functools.update_wrapper(self, self._dygraph_fn)

    @property
    def dygraph_function(self):
        return self._dygraph_fn

    @property
    def code(self):
        """Source after the dy2static AST transform (reference: StaticFunction.code)."""
        from .dy2static import get_code

        return get_code(self._dygraph_fn)

    def concrete_program_specify_input_spec(self, *a, **k):
        return None

    def rollback(self):
        return self._dygraph_fn

    # ---- training capture: forward AND backward as CUDA graphs ------------------------------------------------------------------------
    def _train_params(self):
        if self._layer is None:
            return []
        return [p for p in self._layer.parameters() if not p.stop_gradient]     # the leaf Parameters themselves (an alias would receive no gradient)

    def _can_train_graph(self, args, kwargs):
        from ..framework.flags import flag

        if kwargs or not flag("FLAGS_b200_to_static_train_graph", True) or not torch.is_grad_enabled():
            return False
        ts = [a for a in args if isinstance(a, torch.Tensor)]
        if not ts or len(ts) != len(args) or not all(t.is_cuda for t in ts):
            return False
        return bool(self._train_params()) or any(t.requires_grad for t in ts)

    def _capture_train(self, args):
        """torch.cuda.make_graphed_callables over (inputs, *parameters): the parameters are passed as explicit (unused) arguments so that
        the captured backward graph produces their gradients too - the role dy2static's program + run_program backward play in the
        reference (python/paddle/jit/dy2static/pir_partial_program.py)."""
        params = self._train_params()
        n_in = len(args)
        fn = self._fn

        def pure(*flat):
            out = fn(*[a.as_subclass(Tensor) if isinstance(a, torch.Tensor) and not isinstance(a, Tensor) else a for a in flat[:n_in]])
            return out.as_subclass(torch.Tensor) if isinstance(out, torch.Tensor) and type(out) is not torch.Tensor else out

        sample = tuple((a.as_subclass(torch.Tensor) if type(a) is not torch.Tensor else a).detach().clone().requires_grad_(a.requires_grad) for a in args)
        try:
            graphed = torch.cuda.make_graphed_callables(pure, sample + tuple(params), num_warmup_iters=2, allow_unused_input=True)
        except Exception as e:  # noqa: BLE001  capture-unsafe body: stay eager for this signature
            _capture_failed(e)
            return None
        return graphed, params

    def _can_graph(self, args, kwargs):
        if torch.is_grad_enabled() and any(isinstance(a, torch.Tensor) and a.requires_grad for a in args):
            return False
        if self._layer is not None and self._layer.training and torch.is_grad_enabled():
            return False
        ts = [a for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor)]
        return bool(ts) and all(t.is_cuda for t in ts)

    # ---- backend="CINN": trace -> IR passes -> fused generated kernels (paddle_b200.cinn); inference and training ---------------------------
    def _cinn_entry(self, args):
        from .. import cinn, static

        prog = static.Program()
        fwd = None
        if self._layer is not None:
            fwd = self._layer.__dict__.get("forward")
            if isinstance(fwd, StaticFunction):
                self._layer.__dict__.pop("forward", None)
        try:
            with static.program_guard(prog), torch.no_grad():
                ins, call = [], []
                for i, a in enumerate(args):
                    if isinstance(a, torch.Tensor):
                        v = static.data(f"x{i}", list(a.shape), str(a.dtype).replace("torch.", ""))
                        ins.append(v)
                        call.append(v)
                    else:
                        call.append(a)
                out = self._fn(*call)
        finally:
            if isinstance(fwd, StaticFunction):
                self._layer.forward = fwd
        outs = list(out) if isinstance(out, (list, tuple)) else [out]
        if not outs or not all(isinstance(o, torch.Tensor) and id(o) in prog._fetch_alias for o in outs):
            return None                                           # results that are not values of the program: run as written
        training = torch.is_grad_enabled() and self._layer is not None and self._layer.training
        infer = prog if training else prog.clone(for_test=True)        # a training trace keeps dropout etc. live; fused kernels are differentiable
        new, report = cinn.compile_program(infer, outs)
        return static.Executor(), new, [v.name for v in ins], outs, not isinstance(out, (list, tuple)), report

    def _run_cinn_entry(self, entry, args):
        exe, prog, feeds, fetch, single, _ = entry
        res = exe.run(prog, feed=dict(zip(feeds, [a for a in args if isinstance(a, torch.Tensor)])), fetch_list=fetch, return_numpy=False)
        res = [o.as_subclass(Tensor) if isinstance(o, torch.Tensor) else o for o in res]
        return res[0] if single else res

    def _call_cinn(self, args):
        key = _sig(args, {}) + (bool(torch.is_grad_enabled() and self._layer is not None and self._layer.training),)
        if key not in self._cinn:
            # First call with this signature: trace + compile, then CHECK the compiled program against the function itself on the real
            # arguments (same RNG state).  A forward that leaves the recorded tensor type (`as_subclass(torch.Tensor)` fast paths, .numpy(),
            # python branching on values) bakes trace-time constants into the program; such a function keeps running as written.
            try:
                entry = self._cinn_entry(args)
            except Exception as e:  # noqa: BLE001  (untraceable function): eager
                if os.environ.get("B200_JIT_DEBUG"):
                    raise
                entry, self._cinn_error = None, e
            cpu_rng = torch.get_rng_state()
            cuda_rng = torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None
            ref = self._fn(*args)
            if entry is not None:
                now_cpu = torch.get_rng_state()
                torch.set_rng_state(cpu_rng)
                if cuda_rng is not None:
                    now_cuda = torch.cuda.get_rng_state_all()
                    torch.cuda.set_rng_state_all(cuda_rng)
                try:
                    with torch.no_grad():
                        got = self._run_cinn_entry(entry, args)
                    if not _same_tree(ref, got):
                        entry, self._cinn_error = None, RuntimeError("the traced program does not reproduce the function (it leaves the recorded tensor type); running eagerly")
                except Exception as e:  # noqa: BLE001
                    if os.environ.get("B200_JIT_DEBUG"):
                        raise
                    entry, self._cinn_error = None, e
                finally:
                    torch.set_rng_state(now_cpu)
                    if cuda_rng is not None:
                        torch.cuda.set_rng_state_all(now_cuda)
            self._cinn[key] = entry
            return ref
        entry = self._cinn[key]
        if entry is None:
            return self._fn(*args)
        return self._run_cinn_entry(entry, args)

    def cinn_report(self, *args):
        """FusionResult of the program compiled for these arguments (None when it was not compiled)."""
        for training in (False, True):
            e = self._cinn.get(_sig(args, {}) + (training,))
            if e is not None:
                return e[5]
        return None

    def __call__(self, *args, **kwargs):
        if not _enabled[0] or getattr(self._dygraph_fn, "_not_to_static", False):
            return self._dygraph_fn(*args, **kwargs)
        if self._backend == "CINN" and not kwargs:
            return self._call_cinn(args)
        if self._can_train_graph(args, kwargs):
            key = ("train",) + _sig(args, kwargs)
            entry = self._train_graphs.get(key, False)
            if entry is False:
                n = self._warm.get(key, 0) + 1
                self._warm[key] = n
                if n <= self._capture_after:
                    return self._fn(*args, **kwargs)
                entry = self._train_graphs[key] = self._capture_train(args)
            if entry is not None:
                graphed, params = entry
                out = graphed(*[a.as_subclass(torch.Tensor) if type(a) is not torch.Tensor else a for a in args], *params)
                return out.as_subclass(Tensor) if isinstance(out, torch.Tensor) else out
            return self._fn(*args, **kwargs)
        if not self._can_graph(args, kwargs) or torch.is_grad_enabled():
            return self._fn(*args, **kwargs)
        key = _sig(args, kwargs)
        entry = self._graphs.get(key)
        if entry is None:
            n = self._warm.get(key, 0) + 1
            self._warm[key] = n
            if n <= self._capture_after:
                return self._fn(*args, **kwargs)
            entry = self._capture(args, kwargs)
            self._graphs[key] = entry
            if entry is None:
                return self._fn(*args, **kwargs)
        if entry is None:
            return self._fn(*args, **kwargs)
        graph, static_in, static_out = entry
        for s, a in zip(static_in, [a for a in list(args) + [kwargs[k] for k in sorted(kwargs)] if isinstance(a, torch.Tensor)]):
            s.copy_(a)
        graph.replay()
        return _clone_tree(static_out)

    def _capture(self, args, kwargs):
        try:
            static_args = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
            static_kwargs = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in kwargs.items()}
            static_in = [a for a in static_args + [static_kwargs[k] for k in sorted(static_kwargs)] if isinstance(a, torch.Tensor)]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._fn(*static_args, **static_kwargs)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._fn(*static_args, **static_kwargs)
            return g, static_in, out
        except Exception as e:  # noqa: BLE001  capture-unsafe function (host sync, dynamic shapes...): stay eager for this signature
            _capture_failed(e)
            return None


def _capture_failed(exc):
    """A capture that died half way leaves the device RNG registered as "capturing" (every later random op then raises "Offset increment
    outside graph capture"): drain the device and run one trivial capture to completion, which resets the generator.  B200_JIT_DEBUG=1 prints
    why the capture failed."""
    import os

    if os.environ.get("B200_JIT_DEBUG"):
        import traceback

        traceback.print_exception(type(exc), exc, exc.__traceback__)
    try:
        torch.cuda.synchronize()
    except Exception:  # noqa: BLE001
        pass
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                torch.zeros(1, device="cuda")
        torch.cuda.synchronize()
    except Exception:  # noqa: BLE001
        pass


def _clone_tree(o):
    if isinstance(o, torch.Tensor):
        return o.clone()
    if isinstance(o, (list, tuple)):
        return type(o)(_clone_tree(i) for i in o)
    if isinstance(o, dict):
        return {k: _clone_tree(v) for k, v in o.items()}
    return o


def _same_tree(a, b):
    """Same structure, and tensors equal within the tolerance of a re-associated / once-rounded computation."""
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        if a.shape != b.shape or a.dtype != b.dtype:
            return False
        if not a.is_floating_point():
            return bool(torch.equal(a, b))
        tol = {torch.float16: 2e-2, torch.bfloat16: 5e-2}.get(a.dtype, 1e-3)
        x, y = a.detach().float(), b.detach().float()
        return bool(torch.allclose(x, y, rtol=tol, atol=tol * max(1.0, float(x.abs().max()) if x.numel() else 1.0), equal_nan=True))
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(_same_tree(x, y) for x, y in zip(a, b))
    return isinstance(a, torch.Tensor) == isinstance(b, torch.Tensor)


def _needs_grad(sf, args):
    if any(isinstance(a, torch.Tensor) and a.requires_grad for a in args):
        return True
    return sf._layer is not None and sf._layer.training and any(not p.stop_gradient for p in sf._layer.parameters())


def to_static(function=None, input_spec=None, build_strategy=None, backend=None, **kwargs):
    def decorate(fn):
        if isinstance(fn, Layer):
            layer = fn
            sf = StaticFunction(layer.forward, layer, input_spec, build_strategy, backend)
            layer.forward = sf
            layer._input_spec = input_spec
            return layer
        return StaticFunction(fn, None, input_spec, build_strategy, backend)

    if function is not None:
        return decorate(function)
    return decorate


def _layer_spec(layer):
    cls = type(layer)
    spec = {"module": cls.__module__, "qualname": cls.__qualname__, "init_args": getattr(layer, "_init_args", None)}
    return spec


def save(layer, path, input_spec=None, **configs):
    """jit.save(layer, path): path.pdmodel (structure spec) + path.pdiparams (weights)."""
    if not isinstance(layer, Layer):
        raise TypeError("jit.save expects a Layer (functions are captured at call time by to_static)")
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    _save(layer.state_dict(), path + ".pdiparams")
    spec = _layer_spec(layer)
    spec["input_spec"] = [(list(s.shape), str(s.dtype), s.name) if isinstance(s, InputSpec) else None for s in (input_spec or getattr(layer, "_input_spec", None) or [])]
    try:
        fwd = layer.forward
        if isinstance(fwd, StaticFunction):
            layer.__dict__.pop("forward", None)
        blob = pickle.dumps(layer)
        if isinstance(fwd, StaticFunction):
            layer.forward = fwd
        spec["pickled_layer"] = blob
    except Exception:
        spec["pickled_layer"] = None
    if spec["pickled_layer"] is None and spec["input_spec"] and all(s is not None for s in spec["input_spec"]):
        # classes that cannot be re-imported (defined in a function / __main__): store the traced forward as a static Program.
        # The saved artifact is the inference program: trace in eval mode (dropout off, batch norm on running statistics).
        was_training = bool(getattr(layer, "training", False))
        layer.eval()
        try:
            traced = _trace_program(layer, spec["input_spec"])
            _check_trace(layer, traced, spec["input_spec"])
            spec["program"] = pickle.dumps(traced)
        except Exception as e:  # noqa: BLE001
            spec["program"] = None
            trace_error = e
        finally:
            if was_training:
                layer.train()
    if spec["pickled_layer"] is None and not spec.get("program"):
        cls = type(layer)
        importable = "<locals>" not in cls.__qualname__ and cls.__module__ != "__main__"
        if not importable:
            raise RuntimeError(f"jit.save: {cls.__qualname__} can neither be pickled nor re-imported, and tracing it failed"
                               + (f" ({type(trace_error).__name__}: {trace_error})" if "trace_error" in locals() else " (pass input_spec so that the forward can be traced)"))
    with open(path + ".pdmodel", "wb") as f:
        pickle.dump(spec, f)


def _check_trace(layer, traced, input_spec):
    """Run the traced program and the layer on random inputs of the traced sizes: a forward that leaves the recorded tensor type (raw-tensor
    fast paths, .numpy(), python branches on values) bakes trace-time constants into the program, and a saved model must not be silently wrong."""
    from .. import static

    g = torch.Generator().manual_seed(0)
    ins = []
    for shape, dtype, _ in input_spec:
        dt = getattr(torch, dtype.replace("paddle.", "").replace("torch.", ""), torch.float32)
        dims = [1 if (d is None or d < 0) else int(d) for d in shape]
        t = torch.randn(dims, generator=g) if dt.is_floating_point else torch.randint(0, 2, dims, generator=g)
        ins.append(t.to(dt).as_subclass(Tensor))
    dev = next((p.device for p in layer.parameters()), torch.device("cpu"))
    ins = [t.to(dev) for t in ins]
    with torch.no_grad():
        ref = layer(*ins)
        prog = traced["program"]
        prog._fetch_alias = {id(t): vid for vid, t in enumerate(prog._keep)}
        got = static.Executor().run(prog, feed=dict(zip(traced["feeds"], ins)), fetch_list=[prog._keep[v] for v in traced["fetch_vids"]], return_numpy=False)
    ref = list(ref) if isinstance(ref, (list, tuple)) else [ref]
    if not _same_tree(ref, list(got)):
        raise RuntimeError("the traced program does not reproduce the layer's forward (it computes on raw tensors / python values the tracer cannot see)")


def _trace_program(layer, input_spec):
    from .. import static

    prog = static.Program()
    fwd = layer.__dict__.get("forward")
    if isinstance(fwd, StaticFunction):
        layer.__dict__.pop("forward", None)
    try:
        with static.program_guard(prog), torch.no_grad():
            ins = [static.data(name or f"x{i}", shape, dtype.replace("paddle.", "").replace("torch.", "")) for i, (shape, dtype, name) in enumerate(input_spec)]
            out = layer(*ins)
    finally:
        if isinstance(fwd, StaticFunction):
            layer.forward = fwd
    outs = list(out) if isinstance(out, (list, tuple)) else [out]
    if not all(isinstance(o, torch.Tensor) and id(o) in prog._fetch_alias for o in outs):
        raise RuntimeError("the forward's results are not values of the traced program (it computes them on raw tensors / python values the tracer cannot see)")
    infer = prog.clone(for_test=True)
    return {"program": infer, "feeds": [v.name for v in ins], "fetch_vids": [prog._fetch_alias[id(o)] for o in outs], "single": not isinstance(out, (list, tuple))}


class _ProgramLayer(Layer):
    """Runs a traced Program (jit.load of a model whose Python class is not importable)."""

    def __init__(self, blob):
        super().__init__()
        from .. import static

        self._blob = blob
        self._exe = static.Executor()
        prog = blob["program"]
        prog._fetch_alias = {id(t): vid for vid, t in enumerate(prog._keep)}
        self._fetch = [prog._keep[v] for v in blob["fetch_vids"]]

    def forward(self, *args):
        outs = self._exe.run(self._blob["program"], feed=dict(zip(self._blob["feeds"], args)), fetch_list=self._fetch, return_numpy=False)
        outs = [o.as_subclass(Tensor) if isinstance(o, torch.Tensor) else o for o in outs]
        return outs[0] if self._blob["single"] else outs


class TranslatedLayer(Layer):
    """Layer rebuilt by jit.load. Parity: python/paddle/jit/translated_layer.py."""

    def __init__(self, inner, spec):
        super().__init__()
        self._inner = inner
        self._spec = spec
        self._static = StaticFunction(inner.forward, inner)

    def forward(self, *args, **kwargs):
        return self._static(*args, **kwargs)

    def program(self, method_name="forward"):
        return self._spec


def load(path, **configs):
    with open(path + ".pdmodel", "rb") as f:
        spec = pickle.load(f)
    layer = None
    if spec.get("pickled_layer"):
        layer = pickle.loads(spec["pickled_layer"])
    elif spec.get("program"):
        blob = spec["program"]
        if isinstance(blob, (bytes, bytearray)):         # jit.save of a traced layer: the blob travels pickled inside the spec
            blob = pickle.loads(blob)
        else:                                            # static.save_inference_model: the .pdmodel IS the blob
            blob = dict(spec)
            blob.setdefault("single", len(blob.get("fetch_vids", [])) == 1)
        layer = _ProgramLayer(blob)
        layer.eval()
        return layer
    else:
        mod = importlib.import_module(spec["module"])
        cls = mod
        for part in spec["qualname"].split("."):
            cls = getattr(cls, part)
        args = spec.get("init_args") or ((), {})
        layer = cls(*args[0], **args[1])
    layer.set_state_dict(_load(path + ".pdiparams"))
    layer.eval()
    return TranslatedLayer(layer, spec)


def set_code_level(level=100, also_to_stdout=False):
    pass


def set_verbosity(level=0, also_to_stdout=False):
    pass


def marker_unified(*a, **k):
    def deco(fn):
        return fn

    return deco


__all__ = ["to_static", "save", "load", "TranslatedLayer", "not_to_static", "enable_to_static", "ignore_module", "set_code_level", "set_verbosity"]


from .train_step import CapturedTrainStep, capture_train_step  # noqa: F401,E402
