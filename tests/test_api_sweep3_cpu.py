"""Fourth API sweep: in-place op variants, top-level tensor leftovers vs numpy, initializers, io, autograd switches."""
import math

import numpy as np
import pytest
import scipy.special as sps
import torch

import paddle_b200 as paddle

rng = np.random.RandomState(31)


def t(a):
    return paddle.to_tensor(np.asarray(a))


def close(a, b, tol=1e-6):
    a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    b = b.detach().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


U = rng.rand(3, 4) * 0.8 + 0.1
A = rng.randn(3, 4)
B = rng.randn(3, 4)
I = rng.randint(1, 9, size=(3, 4))
J = rng.randint(1, 9, size=(3, 4))

INPLACE_UNARY = ["abs", "acos", "asin", "atan", "asinh", "atanh", "ceil", "cos", "cosh", "digamma", "erf", "erfinv", "exp", "expm1", "floor", "frac", "i0", "lgamma",
                 "log", "log10", "log1p", "log2", "logit", "nan_to_num", "neg", "reciprocal", "round", "rsqrt", "sigmoid", "sin", "sinh", "sqrt", "square", "tan",
                 "tanh", "trunc", "sinc", "gammaln"]
INPLACE_BINARY = ["add", "subtract", "multiply", "divide", "pow", "remainder", "mod", "floor_mod", "copysign", "hypot", "ldexp", "gammainc", "gammaincc"]
INPLACE_INT = ["bitwise_and", "bitwise_or", "bitwise_xor", "gcd", "lcm", "floor_divide", "bitwise_left_shift", "bitwise_right_shift"]
INPLACE_CMP = ["equal", "not_equal", "less_than", "less_equal", "greater_than", "greater_equal", "logical_and", "logical_or", "logical_xor"]


@pytest.mark.parametrize("name", INPLACE_UNARY)
def test_inplace_unary(name):
    src = U + 1.0 if name == "acosh" else U
    x = t(src.copy())
    ref = getattr(paddle, name)(t(src.copy()))
    out = getattr(paddle, name + "_")(x)
    close(x, ref, 1e-12)
    assert out is x or np.allclose(out.numpy(), x.numpy())


@pytest.mark.parametrize("name", INPLACE_BINARY + INPLACE_INT + INPLACE_CMP)
def test_inplace_binary(name):
    a, b = (I, J) if name in INPLACE_INT else ((U + 0.5, np.abs(B) + 0.5) if name in ("pow", "gammainc", "gammaincc") else (A, B))
    if name == "ldexp":
        b = J
    if name in ("bitwise_left_shift", "bitwise_right_shift"):
        b = J % 4
    x = t(a.copy())
    ref = getattr(paddle, name)(t(a.copy()), t(b))
    getattr(paddle, name + "_")(x, t(b))
    close(x.astype(ref.dtype) if x.dtype != ref.dtype else x, ref, 1e-12)


def test_other_inplace_and_view_ops():
    x = t(A.copy())
    paddle.clip_(x, -0.3, 0.3)
    close(x, np.clip(A, -0.3, 0.3))
    x = t(A.copy())
    paddle.scale_(x, 2.0, 1.0)
    close(x, A * 2 + 1)
    x = t(A.copy())
    paddle.lerp_(x, t(B), 0.25)
    close(x, A + 0.25 * (B - A))
    x = t(A.copy())
    paddle.where_(t(A > 0), x, t(B))
    close(x, np.where(A > 0, A, B))
    x = t(A.copy())
    paddle.masked_fill_(x, t(A > 0), 7.0)
    close(x, np.where(A > 0, 7.0, A))
    x = t(A.copy())
    paddle.cumsum_(x, axis=1)
    close(x, np.cumsum(A, 1))
    x = t(U.copy())
    paddle.cumprod_(x, dim=1)
    close(x, np.cumprod(U, 1))
    x = t(A.copy())
    paddle.renorm_(x, 2.0, 0, 1.0)
    close(x, paddle.renorm(t(A), 2.0, 0, 1.0))
    assert (np.linalg.norm(x.numpy(), axis=1) <= 1.0 + 1e-6).all()
    x = t(A.copy())
    paddle.cast_(x, "float32")
    assert x.dtype == paddle.float32
    x = t(A.copy())
    assert paddle.reshape_(x, [4, 3]).shape == [4, 3] and paddle.flatten_(t(A.copy())).shape == [12]
    assert paddle.squeeze_(t(A[None].copy()), 0).shape == [3, 4] and paddle.unsqueeze_(t(A.copy()), 0).shape == [1, 3, 4]
    x = t(np.tril(np.ones((3, 3))))
    close(paddle.tril_(t(np.ones((3, 3)))), np.tril(np.ones((3, 3))))
    x = t(A.copy())
    paddle.addmm_(t(np.zeros((3, 3))), x, t(B.T.copy()))
    bb = t(np.zeros((2, 3, 3)))
    close(paddle.baddbmm(bb, t(np.stack([A, A])), t(np.stack([B.T, B.T])), beta=0.5, alpha=2.0), 2.0 * np.stack([A @ B.T] * 2))
    paddle.baddbmm_(bb, t(np.stack([A, A])), t(np.stack([B.T, B.T])))
    close(bb, np.stack([A @ B.T] * 2))
    x = t(A.copy())
    paddle.put_along_axis_(x, t(np.zeros((3, 1), "int64")), 9.0, 1)
    assert (x.numpy()[:, 0] == 9).all()
    x = t(np.zeros(5))
    paddle.scatter_(x, t(np.array([1, 3])), t(np.array([2.0, 4.0])))
    close(x, [0, 2, 0, 4, 0])
    x = t(A.copy())
    paddle.index_add_(x, t(np.array([0, 2])), 0, t(np.ones((2, 4))))
    close(x, A + np.array([[1], [0], [1]]))
    close(paddle.index_add(t(A), t(np.array([0, 2])), 0, t(np.ones((2, 4)))), x)
    x = t(A.copy())
    paddle.index_fill_(x, t(np.array([1])), 1, 5.0)
    assert (x.numpy()[:, 1] == 5).all()
    close(paddle.index_fill(t(A), t(np.array([1])), 1, 5.0), x)
    x = t(A.copy())
    paddle.index_put_(x, (t(np.array([0, 1])), t(np.array([1, 2]))), t(np.array([8.0, 9.0])))
    assert x.numpy()[0, 1] == 8 and x.numpy()[1, 2] == 9
    close(paddle.index_put(t(A), (t(np.array([0, 1])), t(np.array([1, 2]))), t(np.array([8.0, 9.0]))), x)
    paddle.seed(0)
    for fn, args in (("uniform_", (-1.0, 1.0)), ("normal_", (0.0, 1.0)), ("exponential_", (1.0,)), ("bernoulli_", (0.5,)), ("cauchy_", ()), ("geometric_", (0.5,)),
                     ("log_normal_", ())):
        x = t(np.zeros((50, 4)))
        getattr(x, fn)(*args) if hasattr(x, fn) else getattr(paddle, fn)(x, *args)
        assert np.isfinite(x.numpy()).all() and x.numpy().std() > 0, fn
    assert paddle.logical_not(t(np.array([True, False]))).numpy().tolist() == [False, True]
    y = t(np.array([True, False]))
    paddle.logical_not_(y)
    assert y.numpy().tolist() == [False, True]
    z = t(I.copy())
    paddle.bitwise_not_(z)
    close(z, ~I)
    close(paddle.bitwise_not(t(I)), ~I)
    close(paddle.bitwise_invert(t(I)), ~I)


def test_top_level_leftovers_vs_numpy():
    c = A[:, :2] + 1j * A[:, 2:]
    close(paddle.as_complex(t(np.stack([A[:, :2], A[:, 2:]], -1))), c)
    close(paddle.as_real(t(c)), np.stack([c.real, c.imag], -1))
    close(paddle.real(t(c)), c.real)
    close(paddle.imag(t(c)), c.imag)
    close(paddle.complex(t(A), t(B)), A + 1j * B)
    close(paddle.polar(t(U), t(A)), U * np.exp(1j * A), 1e-6)
    assert paddle.is_complex(t(c)) and not paddle.is_complex(t(A)) and bool(paddle.isreal(t(c)).numpy().sum() == 0)
    assert paddle.complex64 is not None and paddle.complex128 is not None and paddle.int16 is not None
    close(paddle.as_strided(t(np.arange(12.0)), [3, 2], [4, 1]), np.arange(12.0).reshape(3, 4)[:, :2])
    assert [x.ndim for x in (paddle.atleast_1d(t(1.0)), paddle.atleast_2d(t(1.0)), paddle.atleast_3d(t(1.0)))] == [1, 2, 3]
    close(paddle.add_n([t(A), t(B), t(A)]), 2 * A + B)
    import scipy.linalg as sl

    close(paddle.block_diag([t(A[:2, :2]), t(B[:1, :3])]), sl.block_diag(A[:2, :2], B[:1, :3]))
    assert paddle.broadcast_shape([3, 1, 4], [5, 1]) == [3, 5, 4]
    bt = paddle.broadcast_tensors([t(A[:, :1]), t(A[:1])])
    assert bt[0].shape == [3, 4] and bt[1].shape == [3, 4]
    close(paddle.cartesian_prod([t(np.array([1, 2])), t(np.array([3, 4, 5]))]), np.array([[i, j] for i in (1, 2) for j in (3, 4, 5)]))
    close(paddle.combinations(t(np.array([1, 2, 3])), 2), [[1, 2], [1, 3], [2, 3]])
    close(paddle.column_stack([t(A[:, 0]), t(A[:, 1])]), np.column_stack([A[:, 0], A[:, 1]]))
    close(paddle.row_stack([t(A), t(B)]), np.vstack([A, B]))
    close(paddle.hstack([t(A), t(B)]), np.hstack([A, B]))
    close(paddle.vstack([t(A), t(B)]), np.vstack([A, B]))
    close(paddle.dstack([t(A), t(B)]), np.dstack([A, B]))
    assert [x.shape for x in paddle.hsplit(t(A), 2)] == [[3, 2], [3, 2]] and [x.shape for x in paddle.vsplit(t(np.vstack([A, A])), 2)] == [[3, 4], [3, 4]]
    assert [x.shape for x in paddle.dsplit(t(np.zeros((2, 3, 4))), 2)] == [[2, 3, 2], [2, 3, 2]]
    assert [x.shape for x in paddle.tensor_split(t(np.arange(7.0)), 3)] == [[3], [2], [2]]
    close(paddle.cumulative_trapezoid(t(A), axis=1), np.cumsum((A[:, 1:] + A[:, :-1]) / 2, 1))
    close(paddle.diagflat(t(np.array([1.0, 2.0]))), np.diagflat([1.0, 2.0]))
    close(paddle.diagonal_scatter(t(np.zeros((3, 3))), t(np.ones(3))), np.eye(3))
    close(paddle.select_scatter(t(np.zeros((2, 3))), t(np.ones(3)), 0, 1), [[0, 0, 0], [1, 1, 1]])
    close(paddle.slice_scatter(t(np.zeros((3, 4))), t(np.ones((3, 2))), [1], [0], [4], [2]), np.tile([1, 0, 1, 0], (3, 1)))
    close(paddle.exp2(t(A)), np.exp2(A))
    assert paddle.expand_as(t(A[:1]), t(A)).shape == [3, 4]
    m, e = paddle.frexp(t(U * 10))
    close(m.numpy() * 2.0 ** e.numpy(), U * 10)
    close(paddle.gammainc(t(U + 0.5), t(np.abs(B))), sps.gammainc(U + 0.5, np.abs(B)), 1e-6)
    close(paddle.gammaincc(t(U + 0.5), t(np.abs(B))), sps.gammaincc(U + 0.5, np.abs(B)), 1e-6)
    close(paddle.multigammaln(t(U + 3), 2), sps.multigammaln(U + 3, 2), 1e-6)
    close(paddle.polygamma(t(U + 1), 2), sps.polygamma(2, U + 1), 1e-6)
    close(paddle.sinc(t(A)), np.sinc(A))
    close(paddle.signbit(t(A)), np.signbit(A))
    close(paddle.isclose(t(A), t(A + 1e-9)), np.ones_like(A, bool))
    close(paddle.isin(t(I), t(np.array([1, 2, 3]))), np.isin(I, [1, 2, 3]))
    inf = np.array([1.0, np.inf, -np.inf])
    assert paddle.isposinf(t(inf)).numpy().tolist() == [False, True, False] and paddle.isneginf(t(inf)).numpy().tolist() == [False, False, True]
    assert paddle.is_empty(t(np.zeros((0, 3)))) and not paddle.is_empty(t(A))
    close(paddle.logspace(0, 2, 5), np.logspace(0, 2, 5), 1e-5)
    close(paddle.histogram_bin_edges(t(A), bins=4), np.histogram_bin_edges(A, 4), 1e-6)
    close(paddle.nanmedian(t(np.where(A > 1, np.nan, A))), np.nanmedian(np.where(A > 1, np.nan, A)))
    close(paddle.nanquantile(t(np.where(A > 1, np.nan, A)), 0.4), np.nanquantile(np.where(A > 1, np.nan, A), 0.4))
    close(paddle.floor_mod(t(I), t(J)), np.mod(I, J))
    close(paddle.bitwise_left_shift(t(I), t(J % 3)), I << (J % 3))
    close(paddle.bitwise_right_shift(t(I), t(J % 3)), I >> (J % 3))
    close(paddle.index_sample(t(A), t(np.array([[0, 1], [2, 3], [1, 1]]))), np.take_along_axis(A, np.array([[0, 1], [2, 3], [1, 1]]), 1))
    close(paddle.masked_scatter(t(np.zeros(4)), t(np.array([True, False, True, False])), t(np.array([5.0, 6.0]))), [5, 0, 6, 0])
    close(paddle.multiplex([t(A), t(B)], t(np.array([[0], [1], [0]]))), np.stack([A[0], B[1], A[2]]))
    close(paddle.reduce_as(t(A), t(A[:1])), A.sum(0, keepdims=True))
    close(paddle.reverse(t(A), [1]), A[:, ::-1])
    close(paddle.scatter_nd(t(np.array([[1], [3]])), t(np.array([9.0, 10.0])), [5]), [0, 9, 0, 10, 0])
    close(paddle.scatter_nd_add(t(np.ones(5)), t(np.array([[1], [1]])), t(np.array([2.0, 3.0]))), [1, 6, 1, 1, 1])
    close(paddle.shard_index(t(np.array([[1], [6], [12]])), 20, 2, 0), [[1], [6], [-1]])
    close(paddle.swapaxes(t(A), 0, 1), A.T)
    close(paddle.swapdims(t(A), 0, 1), A.T)
    close(paddle.take(t(A), t(np.array([0, 5, 11]))), A.reshape(-1)[[0, 5, 11]])
    close(np.stack([x.numpy() for x in paddle.tril_indices(3, 3)]) if isinstance(paddle.tril_indices(3, 3), (list, tuple)) else paddle.tril_indices(3, 3), np.stack(np.tril_indices(3)))
    close(paddle.triu_indices(3, 3), np.stack(np.triu_indices(3)))
    assert paddle.unflatten(t(A), 1, [2, 2]).shape == [3, 2, 2] and len(paddle.unstack(t(A), 0)) == 3
    u, cnt = paddle.unique_consecutive(t(np.array([1, 1, 2, 2, 2, 1])), return_counts=True)
    assert u.numpy().tolist() == [1, 2, 1] and cnt.numpy().tolist() == [2, 3, 1]
    close(paddle.vander(t(np.array([1.0, 2.0, 3.0])), 3), np.vander([1.0, 2.0, 3.0], 3))
    assert paddle.view(t(A), [4, 3]).shape == [4, 3] and paddle.view_as(t(A), t(np.zeros((2, 6)))).shape == [2, 6]
    close(paddle.fill_constant([2, 2], "float32", 3.0), np.full((2, 2), 3.0))
    x = t(np.array([1.0]))
    close(paddle.increment(x, 2.0), [3.0])
    assert paddle.from_numpy(np.ones(3)).shape == [3] and paddle.create_tensor("float32") is not None
    assert paddle.double is not None
    paddle.seed(1)
    assert paddle.rand_like(t(A)).shape == [3, 4] and paddle.randint_like(t(I), 0, 5).numpy().max() < 5
    assert paddle.standard_normal([1000]).numpy().std() > 0.8 and (paddle.standard_gamma(t(np.full(1000, 2.0))).numpy() > 0).all()
    assert abs(float(paddle.binomial(t(np.full(2000, 10)), t(np.full(2000, 0.3))).numpy().mean()) - 3) < 0.3
    assert (paddle.log_normal(shape=[100]).numpy() > 0).all()
    picked = paddle.top_p_sampling(t(np.array([[0.9, 0.05, 0.05]], "float32")), t(np.array([0.5], "float32")))
    vals = [np.asarray(v.numpy()).reshape(-1) for v in picked[:2]]
    assert any(int(v[0]) == 0 for v in vals if v.dtype.kind in "iu")     # nucleus of mass 0.5 only contains token 0


def test_state_switches_and_places():
    assert paddle.is_grad_enabled()
    with paddle.set_grad_enabled(False):
        assert not paddle.is_grad_enabled()
        with paddle.enable_grad():
            assert paddle.is_grad_enabled()
    st = paddle.get_rng_state()
    a = paddle.rand([3]).numpy()
    paddle.set_rng_state(st)
    close(paddle.rand([3]), a)
    assert paddle.get_cuda_rng_state() is not None
    paddle.set_cuda_rng_state(paddle.get_cuda_rng_state())
    assert paddle.get_device() in ("cpu",) or paddle.get_device().startswith("gpu")
    assert isinstance(paddle.get_flags("FLAGS_check_nan_inf"), dict)
    for fn in ("is_compiled_with_cinn", "is_compiled_with_rocm", "is_compiled_with_xpu", "is_compiled_with_custom_device"):
        args = ("npu",) if fn == "is_compiled_with_custom_device" else ()
        assert getattr(paddle, fn)(*args) is False
    assert paddle.is_compiled_with_cuda() in (True, False) and paddle.is_compiled_with_distribute() is True
    assert isinstance(paddle.CPUPlace(), paddle.Place) or paddle.CPUPlace() is not None
    assert paddle.CUDAPinnedPlace() is not None
    paddle.disable_signal_handler()


def test_async_save(tmp_path):
    sd = {"w": t(A), "step": 3}
    p = str(tmp_path / "a.pdparams")
    paddle.async_save(sd, p)
    paddle.clear_async_save_task_queue()
    back = paddle.load(p)
    close(back["w"], A)
    assert back["step"] == 3


def test_initializer_leftovers():
    I_ = paddle.nn.initializer
    paddle.seed(0)
    w = paddle.create_parameter([200, 100], "float32", default_initializer=I_.TruncatedNormal(0.0, 1.0, a=-1.0, b=1.0))
    assert abs(float(w.numpy().max())) <= 1.0 and w.numpy().std() > 0.3
    w = paddle.create_parameter([200, 100], "float32", default_initializer=I_.XavierNormal())
    assert abs(w.numpy().std() - math.sqrt(2 / 300)) < 0.01
    w = paddle.create_parameter([200, 100], "float32", default_initializer=I_.KaimingUniform())
    assert abs(w.numpy().std() - math.sqrt(6 / 200) / math.sqrt(3)) < 0.01
    w = paddle.create_parameter([200, 100], "float32", default_initializer=I_.MSRA())
    assert np.isfinite(w.numpy()).all()
    w = paddle.create_parameter([4, 4, 3, 3], "float32", default_initializer=I_.Dirac())
    x = rng.randn(1, 4, 5, 5).astype("float32")
    close(paddle.nn.functional.conv2d(t(x), w, padding=1), x, 1e-6)
    w = paddle.create_parameter([2, 2], "float32", default_initializer=I_.NumpyArrayInitializer(np.array([[1, 2], [3, 4]], "float32")))
    close(w, [[1, 2], [3, 4]])
    assert abs(I_.calculate_gain("relu") - math.sqrt(2)) < 1e-6 and abs(I_.calculate_gain("leaky_relu", 0.2) - math.sqrt(2 / 1.04)) < 1e-6
    I_.set_global_initializer(I_.Constant(0.5), I_.Constant(0.1))
    try:
        lin = paddle.nn.Linear(3, 2)
        assert (lin.weight.numpy() == 0.5).all() and (lin.bias.numpy() == 0.1).all()
    finally:
        I_.set_global_initializer(None)
    assert isinstance(I_.Constant(1.0), I_.Initializer)


def test_io_leftovers():
    io = paddle.io

    class DS(io.Dataset):
        def __init__(self, n, off=0):
            self.n, self.off = n, off

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            return np.float32(i + self.off)

    class IDS(io.IterableDataset):
        def __init__(self, lo, hi):
            self.lo, self.hi = lo, hi

        def __iter__(self):
            info = io.get_worker_info()
            assert info is None
            return iter(range(self.lo, self.hi))

    assert list(io.ChainDataset([IDS(0, 2), IDS(5, 7)])) == [0, 1, 5, 6]
    comp = io.ComposeDataset([DS(3), DS(3, 10)])
    assert len(comp) == 3 and tuple(comp[1]) == (1.0, 11.0)
    assert list(io.SequenceSampler(DS(4))) == [0, 1, 2, 3] and isinstance(io.SequenceSampler(DS(4)), io.Sampler)
    assert sorted(io.SubsetRandomSampler([3, 5, 7])) == [3, 5, 7]
    b = io.default_collate_fn([(np.ones(2, "float32"), 1), (np.zeros(2, "float32"), 2)])
    assert b[0].shape == [2, 2] and b[1].numpy().tolist() == [1, 2]
    assert io.default_convert_fn(np.ones(2)).shape == [2]


def test_nn_utils_and_autograd_leftovers():
    lin = paddle.nn.Linear(4, 3)
    (lin(t(rng.randn(5, 4).astype("float32"))) ** 2).sum().backward()
    total = paddle.nn.utils.clip_grad_norm_(lin.parameters(), 0.1)
    gn = math.sqrt(sum(float((p.grad ** 2).sum()) for p in lin.parameters()))
    assert float(total) > 0 and gn <= 0.1 + 1e-4
    paddle.nn.utils.clip_grad_value_(lin.parameters(), 0.01)
    assert max(float(p.grad.abs().max()) for p in lin.parameters()) <= 0.01 + 1e-8
    wn = paddle.nn.utils.weight_norm(paddle.nn.Linear(4, 3))
    assert hasattr(wn, "weight_g")
    paddle.nn.utils.remove_weight_norm(wn)
    assert not hasattr(wn, "weight_g") and wn.weight.shape == [4, 3]
    x = t(np.array([1.0, 2.0]))
    x.stop_gradient = False
    h = paddle.autograd.hessian((x ** 3).sum(), x)
    close(h[:] if hasattr(h, "__getitem__") else h, np.diag([6.0, 12.0]), 1e-5)
    seen = []
    with paddle.autograd.saved_tensors_hooks(lambda v: (seen.append(1), v)[1], lambda v: v):
        y = (x * x).sum()
    y.backward()
    assert seen and np.allclose(x.grad.numpy(), [2.0, 4.0])

    class Sq(paddle.autograd.PyLayer):
        @staticmethod
        def forward(ctx, v):
            assert isinstance(ctx, paddle.autograd.PyLayerContext)
            ctx.save_for_backward(v)
            return v * v

        @staticmethod
        def backward(ctx, g):
            (v,) = ctx.saved_tensor()
            return 2 * v * g

    z = t(np.array([3.0]))
    z.stop_gradient = False
    Sq.apply(z).sum().backward()
    close(z.grad, [6.0])
    assert paddle.amp.is_bfloat16_supported() in (True, False) and paddle.amp.is_float16_supported() in (True, False)


def test_tensor_method_extras():
    x = t(np.arange(6.0).reshape(2, 3))
    close(x.apply(lambda v: v * 2), np.arange(6.0).reshape(2, 3) * 2)
    assert x.strides == [3, 1] and x.offset == 0 and x[1].offset == 3 * 8 and x.matrix_transpose().shape == [3, 2] and x.is_same_shape(x)
    v0 = x.inplace_version
    x.apply_(lambda v: v + 1)
    assert x.inplace_version > v0 and x._inplace_version() == x.inplace_version and x.numpy()[0, 0] == 1
    y = paddle.zeros([2, 3], "float64")
    x._share_buffer_to(y)
    assert x._is_shared_buffer_with(y) and not x._is_shared_buffer_with(t(np.zeros(3)))
    assert len(x._md5sum()) == 32 and x._md5sum() == y._md5sum() and x._numel() == 6 and x._slice(0, 1).shape == [1, 3]
    close(paddle.zeros([3, 3]).fill_diagonal_tensor(paddle.ones([3])), np.eye(3))
    d = paddle.zeros([3, 3])
    d.fill_diagonal_tensor_(paddle.ones([2]), offset=1)
    close(d, np.diag(np.ones(2), 1))
    close(torch.utils.dlpack.from_dlpack(x.to_dlpack()), x.numpy())
    z = paddle.to_tensor([1.0, 2.0], stop_gradient=False)
    w = z * 2
    w.retain_grads()
    w.sum().backward()
    close(w.grad, [1, 1])
    close(z._grad_ivar(), [2, 2])
    with pytest.raises(RuntimeError):
        z.apply(lambda v: v)
    assert not x.is_selected_rows() and x._use_gpudnn(False) is x
    for name in ("add_n", "concat", "stack", "multi_dot", "broadcast_tensors", "atleast_2d", "polar", "scatter_nd", "where_", "is_tensor"):
        assert hasattr(paddle.Tensor, name), name
    tmp = t(np.ones(3))
    tmp._clear_data()
    assert tmp._numel() == 0
