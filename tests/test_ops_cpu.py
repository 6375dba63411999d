"""Tensor-op API vs numpy. Parity: test/legacy_test/test_*_op.py (OpTest numpy references), here one table-driven file."""
import numpy as np
import pytest

import paddle_b200 as paddle

rng = np.random.RandomState(0)
A = rng.rand(3, 4).astype("float32") + 0.1
B = rng.rand(3, 4).astype("float32") + 0.1
M = rng.rand(4, 5).astype("float32")
I = rng.randint(0, 10, (3, 4)).astype("int64")


def T(x):
    return paddle.to_tensor(x)


UNARY = [("exp", np.exp), ("log", np.log), ("sqrt", np.sqrt), ("abs", np.abs), ("sin", np.sin), ("cos", np.cos), ("tanh", np.tanh),
         ("floor", np.floor), ("ceil", np.ceil), ("square", np.square), ("reciprocal", np.reciprocal), ("sign", np.sign),
         ("log1p", np.log1p), ("expm1", np.expm1), ("rsqrt", lambda a: 1 / np.sqrt(a)), ("neg", np.negative), ("atan", np.arctan),
         ("asin", lambda a: np.arcsin(np.clip(a, -1, 1))), ("sinh", np.sinh), ("cosh", np.cosh), ("log2", np.log2), ("log10", np.log10),
         ("trunc", np.trunc), ("round", np.round)]


@pytest.mark.parametrize("name,ref", UNARY)
def test_unary(name, ref):
    x = np.clip(A, 0.1, 0.95) if name == "asin" else A
    np.testing.assert_allclose(getattr(paddle, name)(T(x)).numpy(), ref(x), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(getattr(T(x), name)().numpy(), ref(x), rtol=1e-5, atol=1e-6)


BINARY = [("add", np.add), ("subtract", np.subtract), ("multiply", np.multiply), ("divide", np.divide), ("maximum", np.maximum),
          ("minimum", np.minimum), ("pow", np.power), ("atan2", np.arctan2), ("fmax", np.fmax), ("fmin", np.fmin)]


@pytest.mark.parametrize("name,ref", BINARY)
def test_binary(name, ref):
    np.testing.assert_allclose(getattr(paddle, name)(T(A), T(B)).numpy(), ref(A, B), rtol=1e-5)
    np.testing.assert_allclose(getattr(paddle, name)(T(A), T(B[0])).numpy(), ref(A, B[0]), rtol=1e-5)   # broadcast


def test_int_ops_and_logic():
    J = rng.randint(1, 5, (3, 4)).astype("int64")
    np.testing.assert_array_equal(paddle.floor_divide(T(I), T(J)).numpy(), I // J)
    np.testing.assert_array_equal(paddle.remainder(T(I), T(J)).numpy(), I % J)
    np.testing.assert_array_equal(paddle.bitwise_and(T(I), T(J)).numpy(), I & J)
    np.testing.assert_array_equal(paddle.bitwise_xor(T(I), T(J)).numpy(), I ^ J)
    np.testing.assert_array_equal((T(A) > T(B)).numpy(), A > B)
    np.testing.assert_array_equal(paddle.logical_and(T(A > 0.5), T(B > 0.5)).numpy(), (A > 0.5) & (B > 0.5))
    np.testing.assert_array_equal(paddle.equal(T(I), T(I)).numpy(), np.ones_like(I, bool))
    assert bool(paddle.equal_all(T(I), T(I))) and bool(paddle.allclose(T(A), T(A + 1e-9)))
    np.testing.assert_array_equal(paddle.isnan(T(np.array([1.0, np.nan]))).numpy(), [False, True])
    np.testing.assert_array_equal(paddle.where(T(A > 0.5), T(A), T(B)).numpy(), np.where(A > 0.5, A, B))


def test_reductions():
    for name, ref in [("sum", np.sum), ("mean", np.mean), ("max", np.max), ("min", np.min), ("prod", np.prod)]:
        np.testing.assert_allclose(getattr(paddle, name)(T(A)).numpy(), ref(A), rtol=1e-5)
        np.testing.assert_allclose(getattr(paddle, name)(T(A), axis=1).numpy(), ref(A, axis=1), rtol=1e-5)
        np.testing.assert_allclose(getattr(paddle, name)(T(A), axis=[0, 1], keepdim=True).numpy(), ref(A, axis=(0, 1), keepdims=True), rtol=1e-5)
    np.testing.assert_allclose(paddle.std(T(A), axis=0).numpy(), A.std(0, ddof=1), rtol=1e-5)
    np.testing.assert_allclose(paddle.var(T(A), unbiased=False).numpy(), A.var(), rtol=1e-5)
    np.testing.assert_allclose(paddle.median(T(A), axis=1).numpy(), np.median(A, 1), rtol=1e-6)
    np.testing.assert_allclose(paddle.logsumexp(T(A), axis=1).numpy(), np.log(np.exp(A).sum(1)), rtol=1e-5)
    np.testing.assert_allclose(paddle.cumsum(T(A), axis=1).numpy(), np.cumsum(A, 1), rtol=1e-5)
    np.testing.assert_allclose(paddle.cumprod(T(A), dim=0).numpy(), np.cumprod(A, 0), rtol=1e-5)
    np.testing.assert_array_equal(paddle.argmax(T(A), axis=1).numpy(), A.argmax(1))
    np.testing.assert_array_equal(paddle.argmin(T(A)).numpy(), A.argmin())
    assert bool(paddle.all(T(A > 0))) and not bool(paddle.any(T(A > 5)))
    np.testing.assert_allclose(paddle.quantile(T(A), 0.3, axis=1).numpy(), np.quantile(A, 0.3, axis=1), rtol=1e-5)
    np.testing.assert_array_equal(paddle.count_nonzero(T(I)).numpy(), np.count_nonzero(I))


def test_creation():
    assert paddle.zeros([2, 3]).shape == [2, 3] and paddle.ones([2], dtype="int32").dtype == paddle.int32
    np.testing.assert_array_equal(paddle.arange(2, 10, 3).numpy(), np.arange(2, 10, 3))
    np.testing.assert_allclose(paddle.linspace(0, 1, 5).numpy(), np.linspace(0, 1, 5), rtol=1e-6)
    np.testing.assert_array_equal(paddle.eye(3, 4).numpy(), np.eye(3, 4, dtype="float32"))
    np.testing.assert_array_equal(paddle.full([2, 2], 7, dtype="int64").numpy(), np.full((2, 2), 7))
    np.testing.assert_array_equal(paddle.tril(T(A)).numpy(), np.tril(A))
    np.testing.assert_array_equal(paddle.triu(T(A), 1).numpy(), np.triu(A, 1))
    np.testing.assert_array_equal(paddle.diag(T(A[0])).numpy(), np.diag(A[0]))
    np.testing.assert_array_equal(paddle.zeros_like(T(A)).numpy(), np.zeros_like(A))
    a, b = paddle.meshgrid(paddle.arange(3), paddle.arange(2))
    assert a.shape == [3, 2] and b.shape == [3, 2]
    assert paddle.empty([3, 1]).shape == [3, 1] and paddle.to_tensor([1, 2]).dtype == paddle.int64 and paddle.to_tensor([1.0]).dtype == paddle.float32
    np.testing.assert_array_equal(paddle.assign(T(A)).numpy(), A)
    x = paddle.clone(T(A))
    assert x.numpy() is not A and paddle.numel(x).item() == 12 and paddle.rank(x).item() == 2


def test_manipulation():
    x = T(A)
    np.testing.assert_array_equal(paddle.reshape(x, [4, 3]).numpy(), A.reshape(4, 3))
    np.testing.assert_array_equal(paddle.reshape(x, [-1, 2]).numpy(), A.reshape(-1, 2))
    np.testing.assert_array_equal(paddle.transpose(x, [1, 0]).numpy(), A.T)
    np.testing.assert_array_equal(paddle.concat([x, T(B)], axis=0).numpy(), np.concatenate([A, B], 0))
    np.testing.assert_array_equal(paddle.stack([x, T(B)], axis=1).numpy(), np.stack([A, B], 1))
    parts = paddle.split(x, [1, 3], axis=1)
    assert [p.shape for p in parts] == [[3, 1], [3, 3]]
    parts = paddle.split(x, 2, axis=1)
    assert len(parts) == 2 and len(paddle.chunk(x, 3, axis=0)) == 3 and len(paddle.unbind(x, 0)) == 3
    np.testing.assert_array_equal(paddle.squeeze(T(A[None]), 0).numpy(), A)
    np.testing.assert_array_equal(paddle.unsqueeze(x, [0, 2]).numpy(), A[None, :, None])
    np.testing.assert_array_equal(paddle.flatten(T(A[None]), 1).numpy(), A.reshape(1, -1))
    np.testing.assert_array_equal(paddle.flip(x, [0]).numpy(), A[::-1])
    np.testing.assert_array_equal(paddle.roll(x, 1, 1).numpy(), np.roll(A, 1, 1))
    np.testing.assert_array_equal(paddle.tile(x, [2, 1]).numpy(), np.tile(A, (2, 1)))
    np.testing.assert_array_equal(paddle.expand(T(A[:1]), [3, 4]).numpy(), np.broadcast_to(A[:1], (3, 4)))
    np.testing.assert_array_equal(paddle.broadcast_to(T(A[0]), [2, 4]).numpy(), np.broadcast_to(A[0], (2, 4)))
    np.testing.assert_array_equal(paddle.gather(x, T(np.array([2, 0])), axis=0).numpy(), A[[2, 0]])
    np.testing.assert_array_equal(paddle.index_select(x, T(np.array([3, 1])), axis=1).numpy(), A[:, [3, 1]])
    np.testing.assert_array_equal(paddle.gather_nd(x, T(np.array([[0, 1], [2, 3]]))).numpy(), A[[0, 2], [1, 3]])
    np.testing.assert_array_equal(paddle.take_along_axis(x, T(I % 4), 1).numpy(), np.take_along_axis(A, I % 4, 1))
    np.testing.assert_array_equal(paddle.masked_select(x, T(A > 0.5)).numpy(), A[A > 0.5])
    np.testing.assert_array_equal(paddle.slice(x, [0, 1], [1, 0], [3, 2]).numpy(), A[1:3, 0:2])
    np.testing.assert_array_equal(paddle.strided_slice(x, [1], [0], [4], [2]).numpy(), A[:, 0:4:2])
    np.testing.assert_array_equal(paddle.cast(x, "int32").numpy(), A.astype("int32"))
    np.testing.assert_array_equal(paddle.moveaxis(T(A[None]), 0, 2).numpy(), np.moveaxis(A[None], 0, 2))
    np.testing.assert_array_equal(paddle.repeat_interleave(x, 2, 0).numpy(), np.repeat(A, 2, 0))
    s = paddle.scatter(paddle.zeros([4, 2]), T(np.array([1, 3])), T(np.ones((2, 2), "float32")))
    np.testing.assert_array_equal(s.numpy(), np.array([[0, 0], [1, 1], [0, 0], [1, 1]], "float32"))
    u = paddle.put_along_axis(paddle.zeros([2, 3]), T(np.array([[1], [2]])), 5.0, 1)
    assert u.numpy()[0, 1] == 5 and u.numpy()[1, 2] == 5
    p = paddle.nn.functional.pad(T(A[None, None]), [1, 1, 2, 0])
    assert p.shape == [1, 1, 5, 6]
    assert paddle.shape(x).numpy().tolist() == [3, 4] and x.shape == [3, 4] and x.ndim == 2
    y = x[1:, ::2]
    np.testing.assert_array_equal(y.numpy(), A[1:, ::2])
    z = paddle.zeros([3, 4])
    z[1] = 2.0
    z[:, 0] = T(np.arange(3, dtype="float32"))
    assert z.numpy()[1, 1] == 2 and z.numpy()[2, 0] == 2


def test_search_sort():
    v, i = paddle.topk(T(A), 2, axis=1)
    np.testing.assert_allclose(v.numpy(), np.sort(A, 1)[:, ::-1][:, :2])
    np.testing.assert_array_equal(paddle.sort(T(A), axis=1, descending=True).numpy(), np.sort(A, 1)[:, ::-1])
    np.testing.assert_array_equal(paddle.argsort(T(A), axis=0).numpy(), np.argsort(A, 0))
    np.testing.assert_array_equal(paddle.nonzero(T(I > 4)).numpy(), np.stack(np.nonzero(I > 4), 1))
    u = paddle.unique(T(I))
    np.testing.assert_array_equal(u.numpy(), np.unique(I))
    u, inv, cnt = paddle.unique(T(I), return_inverse=True, return_counts=True)
    np.testing.assert_array_equal(cnt.numpy(), np.unique(I, return_counts=True)[1])
    np.testing.assert_array_equal(paddle.searchsorted(T(np.array([1.0, 3, 5])), T(np.array([2.0, 5]))).numpy(), [1, 2])
    np.testing.assert_array_equal(paddle.bincount(T(np.array([1, 1, 3]))).numpy(), [0, 2, 0, 1])
    k, ki = paddle.kthvalue(T(A), 2, axis=1)
    np.testing.assert_allclose(k.numpy(), np.sort(A, 1)[:, 1])
    m, mi = paddle.mode(T(np.array([[1, 2, 2], [3, 3, 1]])), axis=1)
    np.testing.assert_array_equal(m.numpy(), [2, 3])
    np.testing.assert_array_equal(paddle.histogram(T(A), bins=4, min=0, max=1).numpy(), np.histogram(A, 4, (0, 1))[0])


def test_linalg():
    np.testing.assert_allclose(paddle.matmul(T(A), T(M)).numpy(), A @ M, rtol=1e-5)
    np.testing.assert_allclose(paddle.matmul(T(A), T(B), transpose_y=True).numpy(), A @ B.T, rtol=1e-5)
    np.testing.assert_allclose(paddle.bmm(T(A[None]), T(M[None])).numpy(), (A @ M)[None], rtol=1e-5)
    np.testing.assert_allclose(paddle.dot(T(A[0]), T(B[0])).numpy(), A[0] @ B[0], rtol=1e-5)
    np.testing.assert_allclose(paddle.einsum("ij,jk->ik", T(A), T(M)).numpy(), A @ M, rtol=1e-5)
    np.testing.assert_allclose(paddle.einsum("ij,ij->i", T(A), T(B)).numpy(), (A * B).sum(1), rtol=1e-5)
    S = (M.T @ M + np.eye(5, dtype="float32")).astype("float32")
    np.testing.assert_allclose(paddle.linalg.inv(T(S)).numpy() @ S, np.eye(5), atol=1e-4)
    np.testing.assert_allclose(paddle.linalg.det(T(S)).numpy(), np.linalg.det(S), rtol=1e-4)
    L = paddle.linalg.cholesky(T(S)).numpy()
    np.testing.assert_allclose(L @ L.T, S, rtol=1e-4, atol=1e-5)
    q, r = paddle.linalg.qr(T(M))
    np.testing.assert_allclose(q.numpy() @ r.numpy(), M, atol=1e-5)
    u, s, vh = paddle.linalg.svd(T(M))
    np.testing.assert_allclose(s.numpy(), np.linalg.svd(M)[1], rtol=1e-4)
    w, v = paddle.linalg.eigh(T(S))
    np.testing.assert_allclose(w.numpy(), np.linalg.eigvalsh(S), rtol=1e-4)
    np.testing.assert_allclose(paddle.linalg.norm(T(A)).numpy(), np.linalg.norm(A), rtol=1e-5)
    np.testing.assert_allclose(paddle.linalg.norm(T(A), p=1, axis=1).numpy(), np.abs(A).sum(1), rtol=1e-5)
    b = rng.rand(5, 2).astype("float32")
    np.testing.assert_allclose(paddle.linalg.solve(T(S), T(b)).numpy(), np.linalg.solve(S, b), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(paddle.linalg.pinv(T(M)).numpy(), np.linalg.pinv(M), atol=1e-4)
    assert int(paddle.linalg.matrix_rank(T(S))) == 5
    np.testing.assert_allclose(paddle.trace(T(S)).numpy(), np.trace(S), rtol=1e-5)
    np.testing.assert_allclose(paddle.cross(T(A[:, :3]), T(B[:, :3]), axis=1).numpy(), np.cross(A[:, :3], B[:, :3]), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(paddle.kron(T(A[:2, :2]), T(B[:2, :2])).numpy(), np.kron(A[:2, :2], B[:2, :2]), rtol=1e-5)
    np.testing.assert_allclose(paddle.outer(T(A[0]), T(B[0])).numpy(), np.outer(A[0], B[0]), rtol=1e-5)
    np.testing.assert_allclose(paddle.mv(T(A), T(M[:, 0])).numpy(), A @ M[:, 0], rtol=1e-5)


def test_random_and_seed():
    paddle.seed(7)
    a = paddle.rand([4, 4]).numpy()
    paddle.seed(7)
    b = paddle.rand([4, 4]).numpy()
    np.testing.assert_array_equal(a, b)
    assert paddle.randn([1000]).numpy().std() > 0.8 and paddle.randint(0, 5, [100]).numpy().max() < 5
    assert sorted(paddle.randperm(6).numpy().tolist()) == list(range(6))
    u = paddle.uniform([1000], min=-2, max=3).numpy()
    assert u.min() >= -2 and u.max() <= 3
    assert paddle.normal(1.0, 0.1, [1000]).numpy().mean() > 0.9
    assert set(paddle.bernoulli(paddle.full([100], 0.5)).numpy().tolist()) <= {0.0, 1.0}
    assert paddle.multinomial(T(np.array([0.1, 0.9], "float32")), 5, replacement=True).shape == [5]


def test_autograd_basic():
    x = paddle.to_tensor(A, stop_gradient=False)
    y = (x * x).sum()
    y.backward()
    np.testing.assert_allclose(x.grad.numpy(), 2 * A, rtol=1e-6)
    x.clear_grad()
    assert x.grad is None or float(x.grad.abs().sum()) == 0
    g, = paddle.grad((x ** 3).sum(), x, create_graph=True)
    g2, = paddle.grad(g.sum(), x)
    np.testing.assert_allclose(g2.numpy(), 6 * A, rtol=1e-5)
    with paddle.no_grad():
        z = x * 2
    assert z.stop_gradient
    d = x.detach()
    assert d.stop_gradient and not x.stop_gradient

    class Double(paddle.autograd.PyLayer):
        @staticmethod
        def forward(ctx, a):
            ctx.save_for_backward(a)
            return a * 2

        @staticmethod
        def backward(ctx, g):
            return g * 2

    x2 = paddle.to_tensor(A, stop_gradient=False)
    Double.apply(x2).sum().backward()
    np.testing.assert_allclose(x2.grad.numpy(), np.full_like(A, 2))
    xj = paddle.to_tensor(A[0], stop_gradient=False)
    j = paddle.autograd.jacobian(xj * xj, xj)
    np.testing.assert_allclose(np.asarray(j[:].numpy() if hasattr(j[:], "numpy") else j[:]), np.diag(2 * A[0]), rtol=1e-5)


def test_dtype_and_place():
    x = T(A)
    assert x.dtype == paddle.float32 and x.astype("float64").dtype == paddle.float64 and x.astype(paddle.bfloat16).dtype == paddle.bfloat16
    assert x.place.is_cpu_place() and "cpu" in str(x.place).lower()
    assert paddle.get_default_dtype() == "float32"
    paddle.set_default_dtype("float64")
    try:
        assert paddle.ones([1]).dtype == paddle.float64
    finally:
        paddle.set_default_dtype("float32")
    assert paddle.is_tensor(x) and not paddle.is_tensor(A) and paddle.is_floating_point(x) and not paddle.is_integer(x)
    assert paddle.iinfo(paddle.int32).max == 2 ** 31 - 1 and paddle.finfo(paddle.float32).eps > 0
    bf = x.astype("bfloat16").numpy()
    assert bf.dtype == np.uint16   # bf16 crosses to numpy as raw uint16, like the reference


def test_fft_signal():
    x = rng.rand(16).astype("float32")
    np.testing.assert_allclose(paddle.fft.fft(T(x)).numpy(), np.fft.fft(x), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(paddle.fft.rfft(T(x)).numpy(), np.fft.rfft(x), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(paddle.fft.irfft(paddle.fft.rfft(T(x))).numpy(), x, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(paddle.fft.fft2(T(A)).numpy(), np.fft.fft2(A), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(paddle.fft.fftshift(T(x)).numpy(), np.fft.fftshift(x))
    sig = T(rng.rand(2, 256).astype("float32"))
    spec = paddle.signal.stft(sig, n_fft=64, hop_length=16)
    rec = paddle.signal.istft(spec, n_fft=64, hop_length=16, length=256)
    np.testing.assert_allclose(rec.numpy(), sig.numpy(), atol=1e-4)


def test_remaining_toplevel_names():
    x = T(A.copy())
    paddle.t_(x)
    np.testing.assert_array_equal(x.numpy(), A.T)
    y = T(A.copy())
    paddle.transpose_(y, [1, 0])
    assert y.shape == [4, 3]
    z = T(A.copy())
    paddle.triu_(z, 1)
    np.testing.assert_array_equal(z.numpy(), np.triu(A, 1))
    np.testing.assert_allclose(paddle.sgn(T(A - 0.5)).numpy(), np.sign(A - 0.5))
    from scipy.spatial.distance import pdist as sp_pdist

    np.testing.assert_allclose(paddle.pdist(T(A)).numpy(), sp_pdist(A), rtol=1e-5)
    m = T(np.zeros((2, 2), "float32"))
    paddle.masked_scatter_(m, T(np.array([[True, False], [False, True]])), T(np.array([5.0, 7.0], "float32")))
    np.testing.assert_array_equal(m.numpy(), [[5, 0], [0, 7]])
    paddle.set_printoptions(precision=4)
    paddle.check_shape([1, 2], "op")
    with pytest.raises(TypeError):
        paddle.check_shape("x", "op")
    with paddle.LazyGuard():
        pass
    import paddle_b200.distributed as D

    assert D.ParallelMode.PIPELINE_PARALLEL == 2 and D.ReduceType.kRedSum == 0 and D.ShowClickEntry("s", "c")._to_attr() == "show_click_entry:s:c"
