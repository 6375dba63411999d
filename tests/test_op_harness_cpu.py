"""Ops checked through the OpTest harness (forward vs numpy, analytic vs numeric gradients) + incubate.autograd."""
import numpy as np
import pytest

import paddle_b200 as paddle
from op_test import OpTest

rng = np.random.RandomState(0)


def _softmax(a):
    e = np.exp(a - a.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


CASES = {
    "tanh": (paddle.tanh, np.tanh, [rng.randn(3, 4)]),
    "matmul": (paddle.matmul, np.matmul, [rng.randn(3, 4), rng.randn(4, 2)]),
    "softmax": (lambda x: paddle.nn.functional.softmax(x) * paddle.to_tensor(np.arange(4.0)), lambda a: _softmax(a) * np.arange(4.0), [rng.randn(3, 4)]),
    "layer_norm": (lambda x: paddle.nn.functional.layer_norm(x, [4]) * paddle.to_tensor(np.arange(1.0, 5.0)),
                   lambda a: (a - a.mean(-1, keepdims=True)) / np.sqrt(a.var(-1, keepdims=True) + 1e-5) * np.arange(1.0, 5.0), [rng.randn(3, 4)]),
    "logsumexp": (lambda x: paddle.logsumexp(x, axis=1), lambda a: np.log(np.exp(a).sum(1)), [rng.randn(3, 4)]),
    "gelu": (paddle.nn.functional.gelu, lambda a: 0.5 * a * (1 + np.vectorize(__import__("math").erf)(a / np.sqrt(2))), [rng.randn(2, 5)]),
    "div_bcast": (paddle.divide, np.divide, [rng.randn(3, 4), rng.rand(4) + 0.5]),
    "cumsum": (lambda x: paddle.cumsum(x, axis=1), lambda a: np.cumsum(a, 1), [rng.randn(2, 5)]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_op(name):
    op, ref, inputs = CASES[name]

    class T(OpTest):
        pass

    T.op, T.ref, T.inputs = staticmethod(op), staticmethod(ref), inputs
    t = T()
    t.rtol, t.atol = 1e-6, 1e-8
    t.check_output()
    t.check_grad()


def test_incubate_autograd_functional():
    from paddle_b200.incubate import autograd as IA

    x = paddle.to_tensor(rng.randn(3).astype("float32"))
    f = lambda t: t * t * t  # noqa: E731
    y, g = IA.vjp(f, x)
    np.testing.assert_allclose(g.numpy(), 3 * x.numpy() ** 2, rtol=1e-5)
    y, t = IA.jvp(f, x, paddle.ones([3]))
    np.testing.assert_allclose(t.numpy(), 3 * x.numpy() ** 2, rtol=1e-5)
    J = IA.Jacobian(f, x)
    np.testing.assert_allclose(J[:].numpy(), np.diag(3 * x.numpy() ** 2), rtol=1e-5)
    H = IA.Hessian(lambda t: (t * t * t).sum(), x)
    np.testing.assert_allclose(H[:].numpy(), np.diag(6 * x.numpy()), rtol=1e-5)
    xb = paddle.to_tensor(rng.randn(2, 3).astype("float32"))
    Jb = IA.Jacobian(lambda t: t * 2.0, xb, is_batched=True)
    assert Jb.shape == [2, 3, 3]
    np.testing.assert_allclose(Jb[0].numpy(), 2 * np.eye(3), rtol=1e-6)


@pytest.mark.parametrize("name", ["tanh", "matmul", "softmax", "logsumexp", "gelu", "div_bcast", "cumsum"])
def test_op_dtype_place_matrix(name):
    """Every case over fp64 / fp32 / bf16 / fp16 on every available place (reference: OpTest's per-dtype, per-place checks)."""
    from op_test import OpTestMatrix

    op, ref, inputs = CASES[name]

    class T(OpTestMatrix):
        pass

    T.op, T.ref, T.inputs = staticmethod(op), staticmethod(ref), inputs
    ran = T().check_all()
    assert len(ran) >= 4
