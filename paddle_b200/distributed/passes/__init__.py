"""paddle.distributed.passes: program rewrites for distributed training. Parity: python/paddle/distributed/passes/
(pass_base.py new_pass / PassManager / PassContext; auto_parallel_amp / _fp16 / _bf16, auto_parallel_recompute,
auto_parallel_gradient_merge, auto_parallel_sharding, fuse_all_reduce, fuse_gemm_epilogue, fuse_optimizer ...).

Programs here are recorded op tapes (static/passes.py); a distributed pass rewrites `program.nodes`:
 * amp passes wrap white-list GEMM-like nodes so they run under `amp.auto_cast` with the requested dtype,
 * recompute wraps checkpoint segments (nodes between two `checkpoints` vids) so their intermediates are dropped from the
   executor's value table after use,
 * gradient merge / sharding / fused all-reduce are runtime properties of the arena optimizer and the DataParallel reducer, so
   these passes record their configuration on the program (`program._dist_attrs`) where `Executor`/`fleet` pick it up."""
from __future__ import annotations

from ...static import passes as _sp
from ...static.passes import PassManager as _ProgramPassManager
from ...static.passes import register_pass


class PassContext:
    """Carries attributes between passes. Parity: pass_base.py:PassContext."""

    def __init__(self):
        self._attrs, self._applied = {}, []

    def set_attr(self, k, v):
        self._attrs[k] = v

    def get_attr(self, k, default=None):
        return self._attrs.get(k, default)

    @property
    def passes(self):
        return list(self._applied)


_GEMM_LIKE = ("matmul", "linear", "mm", "bmm", "conv1d", "conv2d", "conv3d", "addmm", "einsum")


def _amp_pass(dtype):
    def run(program, keep=(), attrs=None):
        from ...amp import auto_cast

        attrs = attrs or {}
        white = set(attrs.get("custom_white_list", ())) | set(_GEMM_LIKE)
        black = set(attrs.get("custom_black_list", ()))
        dt = attrs.get("dtype", dtype)
        n = 0
        for node in program.nodes:
            name = _sp._fname(node.fn)
            if node.kind != "op" or name in black or name not in white or getattr(node.fn, "_amp_wrapped", False):
                continue
            inner = node.fn

            def wrapped(*a, _f=inner, **k):
                with auto_cast(True, level="O1", dtype=dt):
                    return _f(*a, **k)
            wrapped.__name__ = name
            wrapped._amp_wrapped = True
            node.fn = wrapped
            n += 1
        program.__dict__.setdefault("_dist_attrs", {})["amp"] = {"dtype": dt, "wrapped": n}
        return n
    return run


for _n, _d in (("auto_parallel_amp", "float16"), ("auto_parallel_fp16", "float16"), ("auto_parallel_bf16", "bfloat16")):
    register_pass(_n)(_amp_pass(_d))


def _attr_pass(key):
    def run(program, keep=(), attrs=None):
        program.__dict__.setdefault("_dist_attrs", {})[key] = dict(attrs or {})
        return 1
    return run


for _n, _k in (("auto_parallel_gradient_merge_pass", "gradient_merge"), ("auto_parallel_gradient_merge", "gradient_merge"),
               ("auto_parallel_sharding", "sharding"), ("fuse_all_reduce", "fuse_all_reduce"), ("fuse_optimizer", "fuse_optimizer"),
               ("auto_parallel_recompute", "recompute"), ("auto_parallel_grad_clip", "grad_clip"), ("pipeline_scheduler_1F1B", "pipeline"),
               ("pipeline_scheduler_FThenB", "pipeline"), ("pipeline_scheduler_VPP", "pipeline"),
               ("auto_parallel_sequence_parallel_optimization", "sequence_parallel"), ("allreduce_matmul_grad_overlapping", "overlap")):
    register_pass(_n)(_attr_pass(_k))


class _Pass:
    def __init__(self, name, attrs):
        self.name, self._attrs = name, dict(attrs or {})
        self._fn = _sp._REGISTRY[name]

    def set_attr(self, k, v):
        self._attrs[k] = v
        return self

    def get_attr(self, k, default=None):
        return self._attrs.get(k, default)

    def apply(self, main_programs, startup_programs=None, context=None):
        import inspect

        progs = main_programs if isinstance(main_programs, (list, tuple)) else [main_programs]
        takes_attrs = "attrs" in inspect.signature(self._fn).parameters
        for p in progs:
            keep = set(self._attrs.get("keep", ()))
            self._fn(p, keep, self._attrs) if takes_attrs else self._fn(p, keep)
        if context is not None:
            context._applied.append(self)
        return context


def new_pass(name, pass_attrs=None):
    if name not in _sp._REGISTRY:
        raise ValueError(f"unknown pass {name!r}; registered: {sorted(_sp._REGISTRY)}")
    return _Pass(name, pass_attrs)


class PassManager:
    """Ordered list of passes applied to (lists of) programs. Parity: pass_base.py:PassManager."""

    def __init__(self, passes=None, context=None, auto_solve_conflict=True):
        self._passes = [p if isinstance(p, _Pass) else new_pass(p) for p in (passes or [])]
        self._context = context or PassContext()

    def append(self, p):
        self._passes.append(p if isinstance(p, _Pass) else new_pass(p))

    def apply(self, main_programs, startup_programs=None):
        for p in self._passes:
            p.apply(main_programs, startup_programs, self._context)
        return self._context

    @property
    def context(self):
        return self._context

    @property
    def names(self):
        return [p.name for p in self._passes]


__all__ = ["new_pass", "PassManager", "PassContext", "register_pass"]
