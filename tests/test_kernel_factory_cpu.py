"""KernelFactory (kernels/registry.py), paddle._C_ops, core._get_all_register_op_kernels."""
import pytest
import torch

import paddle_b200 as paddle
from paddle_b200.kernels.registry import KernelFactory, KernelKey, dispatch, register_kernel


def test_table_and_selection():
    f = KernelFactory.instance()
    assert {"matmul", "flash_attn", "rms_norm", "layer_norm", "swiglu", "weight_only_linear", "mx_gemm", "cross_entropy_with_softmax"} <= set(f.ops())
    assert f.has_kernel("rms_norm", KernelKey("GPU", "bfloat16")) and not f.has_kernel("rms_norm", KernelKey("CPU", "float32"))
    assert f.has_kernel("flash_attn", KernelKey("GPU", "float16")) and not f.has_kernel("flash_attn", KernelKey("GPU", "float32"))
    for ks in f.kernels().values():                      # every registered target imports
        for k in ks:
            assert callable(k.fn)
    x, w = paddle.randn([4, 8]), paddle.ones([8])
    k = f.select("rms_norm", x, w, 1e-6)                 # CPU tensor: no CPU key -> the ANY (reference) kernel
    assert k.backend == "ANY"
    ref = x / torch.sqrt((x * x).mean(-1, keepdim=True) + 1e-6)
    assert torch.allclose(dispatch("rms_norm", x, w, 1e-6).as_subclass(torch.Tensor), ref.as_subclass(torch.Tensor), atol=1e-6)
    with pytest.raises(KeyError):
        f.select("no_such_op", x)
    listing = paddle.base.core._get_all_register_op_kernels()
    assert "(GPU, ANY, bfloat16)" in listing["matmul"] and "abs" in paddle.base.core.get_all_op_names()


def test_user_kernels_join_the_table_and_predicates_rank():
    calls = []

    @register_kernel("my_scale", backend="CPU", dtypes=("float32",), priority=5, predicate=lambda x, s: s == 2.0)
    def double(x, s):
        calls.append("double")
        return x + x

    @register_kernel("my_scale", backend="ANY", dtypes=("*",))
    def generic(x, s):
        calls.append("generic")
        return x * s

    x = paddle.ones([3])
    assert float(dispatch("my_scale", x, 2.0).sum()) == 6.0 and float(dispatch("my_scale", x, 3.0).sum()) == 9.0
    assert calls == ["double", "generic"]
    assert float(dispatch("my_scale", x.astype("float64"), 2.0).sum()) == 6.0 and calls[-1] == "generic"        # dtype not in the CPU kernel's set
    assert float(paddle._C_ops.my_scale(x, 2.0).sum()) == 6.0


def test_c_ops_positional_calls():
    C = paddle._C_ops
    x, y = paddle.randn([4, 8]), paddle.randn([8, 3])
    assert torch.allclose(C.matmul(x, y).as_subclass(torch.Tensor), (x @ y).as_subclass(torch.Tensor), atol=1e-6)
    assert torch.equal(C.abs(x).as_subclass(torch.Tensor), x.abs().as_subclass(torch.Tensor))
    assert C.softmax(x, -1).shape == [4, 8] and C.concat([x, x], 0).shape == [8, 8]
    assert C.rms_norm(x, paddle.ones([8]), 1e-6).shape == [4, 8]
    with pytest.raises(AttributeError):
        C.this_op_does_not_exist
