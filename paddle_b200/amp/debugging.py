"""Numerics debugging. Parity: python/paddle/amp/debugging.py, paddle/fluid/eager/nan_inf_utils.cc."""
from __future__ import annotations

import contextlib
from collections import defaultdict
from enum import Enum

import torch

from ..framework.flags import set_flags


class DebugMode(Enum):
    CHECK_NAN_INF_AND_ABORT = 0
    CHECK_NAN_INF = 1
    CHECK_ALL_FOR_OVERFLOW = 2
    CHECK_ALL = 3
    CHECK_ALL_AND_ABORT = 4
    DUMP_ALL = 5


class TensorCheckerConfig:
    def __init__(self, enable, debug_mode=DebugMode.CHECK_NAN_INF_AND_ABORT, output_dir=None, checked_op_list=None,
                 skipped_op_list=None, debug_step=None, stack_height_limit=1):
        self.enable, self.debug_mode, self.output_dir = enable, debug_mode, output_dir
        self.checked_op_list, self.skipped_op_list = checked_op_list, skipped_op_list
        self.debug_step, self.stack_height_limit = debug_step, stack_height_limit


def check_numerics(tensor, op_type="", var_name="", debug_mode=DebugMode.CHECK_NAN_INF_AND_ABORT):
    t = tensor.as_subclass(torch.Tensor) if isinstance(tensor, torch.Tensor) else torch.as_tensor(tensor)
    tf = t.float()
    n_nan, n_inf = int(torch.isnan(tf).sum()), int(torch.isinf(tf).sum())
    stats = torch.tensor([n_nan, n_inf, int((tf == 0).sum())], dtype=torch.int64)
    finite = tf[torch.isfinite(tf)]
    vals = torch.tensor([finite.max().item() if finite.numel() else 0.0, finite.min().item() if finite.numel() else 0.0,
                         finite.mean().item() if finite.numel() else 0.0])
    if (n_nan or n_inf) and debug_mode in (DebugMode.CHECK_NAN_INF_AND_ABORT, DebugMode.CHECK_ALL_AND_ABORT):
        raise RuntimeError(f"[check_numerics] op={op_type} var={var_name}: {n_nan} nan, {n_inf} inf")
    from ..tensor import Tensor

    return stats.as_subclass(Tensor), vals.as_subclass(Tensor)


_checker = {"hooks": [], "enabled": False}


def enable_tensor_checker(checker_config):
    set_flags({"FLAGS_check_nan_inf": bool(checker_config.enable)})
    _checker["enabled"] = bool(checker_config.enable)
    if checker_config.enable:
        torch.autograd.set_detect_anomaly(True, check_nan=True)
        if checker_config.output_dir and checker_config.debug_mode in (DebugMode.DUMP_ALL, DebugMode.CHECK_ALL, DebugMode.CHECK_ALL_AND_ABORT,
                                                                       DebugMode.CHECK_ALL_FOR_OVERFLOW):
            from .accuracy_compare import TensorDumpMode

            mode = TensorDumpMode(checker_config.output_dir, checker_config.checked_op_list, checker_config.skipped_op_list,
                                  abort_on_nonfinite=checker_config.debug_mode == DebugMode.CHECK_ALL_AND_ABORT)
            mode.__enter__()
            _checker["dump"] = mode


def disable_tensor_checker():
    set_flags({"FLAGS_check_nan_inf": False})
    _checker["enabled"] = False
    torch.autograd.set_detect_anomaly(False)
    mode = _checker.pop("dump", None)
    if mode is not None:
        mode.__exit__(None, None, None)
        mode.close()


_op_stats = defaultdict(lambda: [0, 0, 0, 0])  # fp16, bf16, fp32, other


class _StatsMode(torch.overrides.TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        t = out[0] if isinstance(out, (tuple, list)) and out else out
        if isinstance(t, torch.Tensor):
            name = getattr(func, "__name__", str(func))
            idx = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}.get(t.dtype, 3)
            _op_stats[name][idx] += 1
        return out


_mode = [None]


def enable_operator_stats_collection():
    _op_stats.clear()
    _mode[0] = _StatsMode()
    _mode[0].__enter__()


def disable_operator_stats_collection():
    if _mode[0] is not None:
        _mode[0].__exit__(None, None, None)
        _mode[0] = None
    print("<{:-^120}>".format(" op list "))
    print("<{:-^40}".format(" Op Name ") + "|{:-^17}|{:-^17}|{:-^17}|{:-^17}>".format(" FP16 Calls ", " BF16 Calls ", " FP32 Calls ", " Other Calls "))
    for k, v in sorted(_op_stats.items()):
        print("  {:<40}|  {:<15}|  {:<15}|  {:<15}|  {:<15}".format(k, *v))
    print("<{:-^120}>".format(f" op count: {len(_op_stats)} "))


@contextlib.contextmanager
def collect_operator_stats():
    enable_operator_stats_collection()
    try:
        yield
    finally:
        disable_operator_stats_collection()


def compare_accuracy(dump_path, another_dump_path, output_filename, loss_scale=1, dump_all_tensors=False):
    """Compare the per-op tensor logs of two runs (TensorCheckerConfig(output_dir=..., debug_mode=DebugMode.DUMP_ALL))."""
    from .accuracy_compare import compare_accuracy as _cmp

    return _cmp(dump_path, another_dump_path, output_filename, loss_scale, dump_all_tensors)


def check_layer_numerics(func):
    """Decorator for Layer.forward: checks inputs and outputs for NaN / Inf. Parity: amp/debugging.py:check_layer_numerics."""
    import functools

    import torch

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        def chk(t, what):
            if isinstance(t, torch.Tensor) and t.is_floating_point() and not bool(torch.isfinite(t).all()):
                raise RuntimeError(f"{type(self).__name__}: non-finite values in {what}")
        for a in args:
            chk(a, "input")
        out = func(self, *args, **kwargs)
        for o in (out if isinstance(out, (list, tuple)) else [out]):
            chk(o, "output")
        return out

    return wrapper
