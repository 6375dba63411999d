"""paddle.incubate.jit.inference: decorator that serves a Layer's forward (or a function) through the inference path.
Parity: python/paddle/incubate/jit/inference_decorator.py (there: export to a static model + Predictor with TensorRT options).
Here: no_grad + eval mode + the CUDA-graph replay of `paddle.jit.to_static`, optional precision cast."""
from __future__ import annotations

import functools

import torch


def inference(function=None, cache_static_model=False, save_model_dir=None, memory_pool_init_size_mb=1000, precision_mode="float32", switch_ir_optim=True,
              switch_ir_debug=False, enable_cinn=False, with_trt=False, trt_precision_mode="float32", trt_use_static=False, collect_shape=False,
              enable_new_ir=True, exp_enable_use_cutlass=False, delete_pass_lists=None, skip_prune_program=False):
    from .. import jit
    from ..nn.layer import Layer

    def decorate(fn):
        if isinstance(fn, Layer):
            fn.eval()
            if precision_mode in ("float16", "bfloat16"):
                fn.to(dtype=precision_mode)
            fn.forward = decorate(fn.forward)
            return fn
        static = jit.to_static(fn)

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            with torch.no_grad():
                return static(*args, **kwargs)

        wrapper._inference_options = {"precision_mode": precision_mode, "with_trt": with_trt, "save_model_dir": save_model_dir}
        return wrapper

    return decorate(function) if function is not None else decorate
