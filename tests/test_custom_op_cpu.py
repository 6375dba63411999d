"""Custom C++ op through paddle.utils.cpp_extension.load (JIT build with the in-image toolchain) + indexing semantics vs numpy.
Parity: test/custom_op/test_custom_relu_op_jit.py, test/indexing/."""
import numpy as np
import pytest

import paddle_b200 as paddle


@pytest.mark.timeout(600)
def test_custom_cpp_op_jit(tmp_path):
    src = tmp_path / "my_relu.cc"
    src.write_text('''
#include <torch/extension.h>
torch::Tensor my_leaky(const torch::Tensor& x, double slope) { return torch::where(x > 0, x, x * slope); }
std::vector<torch::Tensor> my_leaky_grad(const torch::Tensor& x, const torch::Tensor& gy, double slope) {
  return {torch::where(x > 0, gy, gy * slope)};
}
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("my_leaky", &my_leaky);
  m.def("my_leaky_grad", &my_leaky_grad);
}
''')
    from paddle_b200.utils import cpp_extension

    mod = cpp_extension.load(name="my_leaky_ext", sources=[str(src)], build_directory=str(tmp_path / "build"), verbose=False)
    x = paddle.to_tensor(np.array([-2.0, -0.5, 0.0, 3.0], "float32"))
    y = mod.my_leaky(x, 0.1)
    assert isinstance(y, paddle.Tensor)
    np.testing.assert_allclose(y.numpy(), [-0.2, -0.05, 0.0, 3.0], rtol=1e-6)
    (g,) = mod.my_leaky_grad(x, paddle.ones([4]), 0.1)
    np.testing.assert_allclose(g.numpy(), [0.1, 0.1, 0.1, 1.0], rtol=1e-6)

    class Leaky(paddle.autograd.PyLayer):
        @staticmethod
        def forward(ctx, v):
            ctx.save_for_backward(v)
            return mod.my_leaky(v, 0.1)

        @staticmethod
        def backward(ctx, gy):
            (v,) = ctx.saved_tensor()
            return mod.my_leaky_grad(v, gy, 0.1)[0]

    x.stop_gradient = False
    Leaky.apply(x).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), [0.1, 0.1, 0.1, 1.0], rtol=1e-6)


def test_indexing_semantics_match_numpy():
    a = np.arange(2 * 3 * 4, dtype="float32").reshape(2, 3, 4)
    x = paddle.to_tensor(a)
    idx = np.array([2, 0])
    cases = [
        (slice(None), 1), (Ellipsis, 2), (0, slice(1, None), slice(None, None, 2)), (None, 1, ..., None), (slice(None), idx), (idx[1:], slice(None), idx),
        (a[..., 0] > 10,), (slice(None), slice(None), np.array([True, False, True, False])), (np.array([[0, 1], [1, 0]]), 1), (-1, -1, -1), (slice(None, None, -1),),
        (1, [0, 2], [1, 3]),
    ]
    for c in cases:
        c = c if isinstance(c, tuple) else (c,)
        pc = tuple(paddle.to_tensor(i) if isinstance(i, np.ndarray) else i for i in c)
        got = x[pc] if len(pc) > 1 else x[pc[0]]
        ref = a[c] if len(c) > 1 else a[c[0]]
        np.testing.assert_allclose(np.asarray(got.numpy()), ref, err_msg=str(c))
    # setitem: scalar, broadcast row, tensor value, boolean mask, index arrays, step slices
    b = a.copy()
    y = paddle.to_tensor(a.copy())
    for key, val in [((0, 1), 5.0), ((slice(None), 0), np.array([1.0, 2.0, 3.0, 4.0], "float32")), ((a > 20,), -1.0), ((slice(None), idx, 0), np.array([[7.0, 8.0]], "float32")),
                     ((1, slice(None, None, 2), slice(1, None, 2)), 0.5), ((Ellipsis, -1), np.zeros((2, 3), "float32"))]:
        pk = tuple(paddle.to_tensor(i) if isinstance(i, np.ndarray) else i for i in key)
        pv = paddle.to_tensor(val) if isinstance(val, np.ndarray) else val
        y[pk if len(pk) > 1 else pk[0]] = pv
        b[key if len(key) > 1 else key[0]] = val
        np.testing.assert_allclose(y.numpy(), b, err_msg=str(key))
    # gradient through a gather-style index and an in-place masked write
    z = paddle.to_tensor(a.copy(), stop_gradient=False)
    (z[:, idx] * 2).sum().backward()
    ref = np.zeros_like(a)
    ref[:, idx] = 2
    np.testing.assert_allclose(z.grad.numpy(), ref)
    w = paddle.to_tensor(a.copy(), stop_gradient=False)
    v = w * 1.0
    v[0] = 0.0
    v.sum().backward()
    ref = np.ones_like(a)
    ref[0] = 0
    np.testing.assert_allclose(w.grad.numpy(), ref)
