"""paddle.static.amp: mixed precision for static programs. Parity: python/paddle/static/amp/{decorator,fp16_lists,fp16_utils}.py.

`decorate(optimizer, ...)` returns an optimizer whose `minimize(loss)` (a) rewrites the white-list ops of the recorded program to run
under auto_cast in the requested dtype (the `auto_parallel_fp16 / bf16` program pass) and (b) applies (dynamic) loss scaling around the
recorded backward + update."""
from __future__ import annotations

from ..amp import GradScaler, auto_cast  # noqa: F401
from ..amp.auto_cast import fp32_guard as fp16_guard  # noqa: F401


class AutoMixedPrecisionLists:
    """White / black op lists. Parity: static/amp/fp16_lists.py:AutoMixedPrecisionLists."""

    def __init__(self, custom_white_list=None, custom_black_list=None, custom_black_varnames=None, dtype="float16"):
        self.white_list, self.black_list = set(custom_white_list or ()), set(custom_black_list or ())
        self.black_varnames, self.dtype = set(custom_black_varnames or ()), dtype
        both = self.white_list & self.black_list
        if both:
            raise ValueError(f"ops in both the custom white and black list: {sorted(both)}")


CustomOpLists = AutoMixedPrecisionLists


class OptimizerWithMixedPrecision:
    def __init__(self, optimizer, amp_lists=None, level="O1", dtype="float16", init_loss_scaling=2.0 ** 15, incr_every_n_steps=1000,
                 decr_every_n_nan_or_inf=2, incr_ratio=2.0, decr_ratio=0.8, use_dynamic_loss_scaling=None, use_amp_guard=False, use_promote=False):
        self._optimizer, self._lists, self._level, self._dtype = optimizer, amp_lists or AutoMixedPrecisionLists(dtype=dtype), level, dtype
        dyn = (dtype == "float16") if use_dynamic_loss_scaling is None else use_dynamic_loss_scaling
        self._scaler = GradScaler(enable=dtype == "float16", init_loss_scaling=init_loss_scaling, incr_ratio=incr_ratio, decr_ratio=decr_ratio,
                                  incr_every_n_steps=incr_every_n_steps, decr_every_n_nan_or_inf=decr_every_n_nan_or_inf, use_dynamic_loss_scaling=dyn)

    def __getattr__(self, name):
        return getattr(self._optimizer, name)

    def get_loss_scaling(self):
        return self._scaler.get_loss_scaling()

    def amp_init(self, place=None, scope=None, test_program=None, use_fp16_test=False):
        """O2: cast the parameters once to the low-precision dtype (fp32 master weights are kept by multi_precision optimizers)."""
        if self._level == "O2":
            for p in self._optimizer._parameter_list or []:
                p.set_value(p.astype(self._dtype)) if hasattr(p, "astype") else None

    def minimize(self, loss, startup_program=None, parameters=None, no_grad_set=None):
        from .. import static
        from ..distributed.passes import new_pass

        prog = static._recording[0]
        if prog is not None:
            new_pass("auto_parallel_bf16" if self._dtype == "bfloat16" else "auto_parallel_fp16",
                     {"custom_white_list": list(self._lists.white_list), "custom_black_list": list(self._lists.black_list), "dtype": self._dtype}).apply([prog])
        return self._optimizer.minimize(loss, startup_program, parameters, no_grad_set)

    def backward(self, loss, **kw):
        return self._optimizer.backward(self._scaler.scale(loss), **kw)

    def apply_gradients(self, params_grads):
        return self._optimizer.apply_gradients(params_grads)


def decorate(optimizer, amp_lists=None, level="O1", dtype="float16", master_weight=None, master_grad=False, init_loss_scaling=2.0 ** 15, incr_every_n_steps=1000,
             decr_every_n_nan_or_inf=2, incr_ratio=2.0, decr_ratio=0.8, use_dynamic_loss_scaling=None, use_amp_guard=False, use_promote=False):
    if master_weight and hasattr(optimizer, "_multi_precision"):
        optimizer._multi_precision = True
    return OptimizerWithMixedPrecision(optimizer, amp_lists, level, dtype, init_loss_scaling, incr_every_n_steps, decr_every_n_nan_or_inf, incr_ratio, decr_ratio,
                                       use_dynamic_loss_scaling, use_amp_guard, use_promote)


def cast_model_to_fp16(program, amp_lists=None, use_fp16_guard=True, dest_type="float16"):
    from ..distributed.passes import new_pass

    new_pass("auto_parallel_bf16" if dest_type == "bfloat16" else "auto_parallel_fp16", {"dtype": dest_type}).apply([program])
    return set()


def cast_parameters_to_fp16(place, program, scope=None, to_fp16_var_names=None, dest_type="float16"):
    for p in program.all_parameters():
        if to_fp16_var_names is None or p.name in to_fp16_var_names:
            p.set_value(p.astype(dest_type))


class bf16:  # namespace paddle.static.amp.bf16
    AutoMixedPrecisionListsBF16 = AutoMixedPrecisionLists

    @staticmethod
    def decorate_bf16(optimizer, amp_lists=None, use_pure_bf16=False, use_bf16_guard=None):
        return decorate(optimizer, amp_lists, level="O2" if use_pure_bf16 else "O1", dtype="bfloat16")

    @staticmethod
    def bf16_guard():
        return auto_cast(True, dtype="bfloat16")
