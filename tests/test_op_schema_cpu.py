"""Operator schema (ops.yaml) and infer-meta. Parity: the reference's YAML-driven op definitions + phi infermeta unit tests
(test/cpp/phi/core/test_meta_fn_utils.cc, paddle/phi/ops/yaml)."""
import inspect

import numpy as np
import pytest
import torch

import paddle_b200 as paddle
from paddle_b200.common import DDim
from paddle_b200.ops import schema as S

M = S.MetaTensor


def test_yaml_matches_live_ops():
    problems = S.validate()
    assert not problems, problems[:10]
    declared = S.load_yaml()
    assert len(declared) >= 400
    mm = declared["matmul"]
    assert [a.name for a in mm.args] == ["x", "y", "transpose_x", "transpose_y"] and mm.args[2].default is False and mm.backward == "matmul_grad"
    assert S.parse_signature("clip(Tensor x, Scalar min=none, Scalar max=none) -> Tensor(out)").args[1].default is None


def test_infer_meta_shapes_and_dtypes():
    assert S.infer_meta("matmul", M([4, 8, 16]), M([16, 32])) == M([4, 8, 32])
    assert S.infer_meta("matmul", M([8, 16], torch.bfloat16), M([32, 16], torch.bfloat16), transpose_y=True) == M([8, 32], torch.bfloat16)
    assert S.infer_meta("concat", [M([2, 3]), M([5, 3])], axis=0) == M([7, 3])
    assert S.infer_meta("transpose", M([2, 3, 4]), [2, 0, 1]) == M([4, 2, 3])
    assert S.infer_meta("sum", M([-1, 3, 4]), axis=1) == M([-1, 4])                 # dynamic batch stays dynamic
    assert S.infer_meta("reshape", M([-1, 6]), [-1, 2, 3]) == M([-1, 2, 3])
    assert S.infer_meta("argmax", M([5, 7]), axis=1) == M([5], torch.int64)
    v, i = S.infer_meta("topk", M([5, 10]), 3)
    assert v == M([5, 3]) and i == M([5, 3], torch.int64)
    assert S.infer_meta("nonzero", M([3, 4])) == M([-1, 2], torch.int64)            # data dependent
    assert S.infer_meta("cast", M([2, 2]), "float16").dtype == torch.float16
    with pytest.raises(Exception):
        S.infer_meta("matmul", M([4, 8]), M([9, 3]))                                 # shape errors surface at infer time, before any kernel


def test_infer_meta_agrees_with_execution_over_the_unary_library():
    S.build_registry()
    checked = 0
    for name, s in S.REGISTRY.items():
        if name in S._CUSTOM_META or not s.args or s.args[0].type != "Tensor" or len(s.tensor_args) != 1:
            continue
        if not all(a.default is not inspect.Parameter.empty for a in s.args[1:]):
            continue
        try:
            r = S.infer_meta(name, M([4, 6]))
        except Exception:  # noqa: BLE001  (needs integer / complex input, random op, ...)
            continue
        try:
            real = s.func(paddle.ones([4, 6]))
        except Exception:  # noqa: BLE001
            continue
        if isinstance(r, M) and isinstance(real, torch.Tensor):
            assert list(r.dims) == list(real.shape) and r.dtype == real.dtype, (name, r, tuple(real.shape), real.dtype)
            checked += 1
    assert checked >= 120, checked


def test_ddim_and_errors():
    from paddle_b200 import common as C

    d = C.make_ddim([2, 3, 4])
    assert C.product(d) == 24 and C.flatten_to_2d(d, 1) == [2, 12] and list(C.stride(d)) == [12, 4, 1] and C.slice_ddim(d, 1, 3) == [3, 4]
    assert d.reshape([-1, 4]) == [6, 4] and DDim([2, -1]).is_dynamic() and C.product([2, -1]) == -1
    with pytest.raises(ValueError):
        C.enforce_eq(1, 2)
    with pytest.raises(IndexError):
        C.enforce(False, C.OutOfRangeError, "index %d out of range", 5)
    with pytest.raises(NotImplementedError):
        C.throw(C.UnimplementedError, "nope")
    assert issubclass(C.InvalidArgumentError, C.EnforceNotMet) and C.ResourceExhaustedError.code == C.ErrorCode.RESOURCE_EXHAUSTED
