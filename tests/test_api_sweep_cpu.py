"""Second API sweep: distributions vs scipy.stats, linalg / fft leftovers vs numpy, lr schedulers vs closed forms, vision
(models / transforms / ops), geometric, sparse and incubate fused functional ops vs plain compositions."""
import math

import numpy as np
import pytest
import scipy.stats as st
import torch

import paddle_b200 as paddle

rng = np.random.RandomState(11)
D = paddle.distribution


def t(a, dtype=None):
    return paddle.to_tensor(np.asarray(a, dtype=dtype) if dtype else np.asarray(a))


def close(a, b, tol=1e-4):
    a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    b = b.detach().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


XS = np.array([0.3, 0.9, 1.7], "float32")
DISTS = {
    "Normal": (lambda: D.Normal(0.5, 1.5), st.norm(0.5, 1.5), XS), "LogNormal": (lambda: D.LogNormal(0.2, 0.7), st.lognorm(0.7, scale=math.exp(0.2)), XS),
    "Uniform": (lambda: D.Uniform(0.0, 2.0), st.uniform(0, 2), XS), "Beta": (lambda: D.Beta(2.0, 3.0), st.beta(2, 3), XS[:2] * 0.5),
    "Cauchy": (lambda: D.Cauchy(0.1, 1.2), st.cauchy(0.1, 1.2), XS), "Chi2": (lambda: D.Chi2(t(3.0)), st.chi2(3), XS),
    "Exponential": (lambda: D.Exponential(t(1.5)), st.expon(scale=1 / 1.5), XS), "Gamma": (lambda: D.Gamma(t(2.0), t(1.5)), st.gamma(2.0, scale=1 / 1.5), XS),
    "Gumbel": (lambda: D.Gumbel(0.3, 1.1), st.gumbel_r(0.3, 1.1), XS), "Laplace": (lambda: D.Laplace(0.2, 0.9), st.laplace(0.2, 0.9), XS),
    "StudentT": (lambda: D.StudentT(4.0, 0.1, 1.3), st.t(4, 0.1, 1.3), XS),
}
DISCRETE = {
    "Bernoulli": (lambda: D.Bernoulli(t(0.3)), st.bernoulli(0.3), np.array([0.0, 1.0], "float32")),
    "Binomial": (lambda: D.Binomial(t(5), t(0.4)), st.binom(5, 0.4), np.array([0.0, 2.0, 5.0], "float32")),
    "Geometric": (lambda: D.Geometric(t(0.3)), st.geom(0.3, loc=-1), np.array([0.0, 2.0, 4.0], "float32")),
    "Poisson": (lambda: D.Poisson(t(2.5)), st.poisson(2.5), np.array([0.0, 2.0, 6.0], "float32")),
}


@pytest.mark.parametrize("name", sorted(DISTS))
def test_continuous_distribution(name):
    make, ref, xs = DISTS[name]
    d = make()
    close(d.log_prob(t(xs)), ref.logpdf(xs))
    close(d.prob(t(xs)), ref.pdf(xs))
    if name != "Cauchy":
        close(d.mean, ref.mean(), 1e-4)
        close(d.variance, ref.var(), 1e-3)
    close(d.entropy(), ref.entropy(), 1e-4)
    if hasattr(d, "cdf") and name not in ("Beta", "Chi2", "Gamma", "StudentT"):
        close(d.cdf(t(xs)), ref.cdf(xs))
    paddle.seed(5)
    s = d.sample([4000]).numpy()
    assert s.shape[0] == 4000 and np.isfinite(s).all()
    if name not in ("Cauchy", "StudentT", "LogNormal"):
        assert abs(s.mean() - ref.mean()) < 0.15 * max(1.0, ref.std())
    assert isinstance(d, D.Distribution)


@pytest.mark.parametrize("name", sorted(DISCRETE))
def test_discrete_distribution(name):
    make, ref, xs = DISCRETE[name]
    d = make()
    close(d.log_prob(t(xs)), ref.logpmf(xs))
    close(d.mean, ref.mean())
    close(d.variance, ref.var())
    paddle.seed(5)
    assert abs(float(d.sample([5000]).numpy().astype("float64").mean()) - ref.mean()) < 0.1 * max(1.0, ref.mean())


def test_multivariate_and_composed_distributions():
    loc, A = rng.randn(3).astype("float32"), rng.randn(3, 3).astype("float32")
    cov = (A @ A.T + np.eye(3)).astype("float32")
    x = rng.randn(3).astype("float32")
    mvn = D.MultivariateNormal(t(loc), covariance_matrix=t(cov))
    ref = st.multivariate_normal(loc, cov)
    close(mvn.log_prob(t(x)), ref.logpdf(x), 1e-3)
    close(mvn.entropy(), ref.entropy(), 1e-3)
    conc = np.array([1.5, 2.0, 3.0], "float32")
    p = np.array([0.2, 0.3, 0.5], "float32")
    close(D.Dirichlet(t(conc)).log_prob(t(p)), st.dirichlet(conc).logpdf(p), 1e-4)
    close(D.Dirichlet(t(conc)).entropy(), st.dirichlet(conc).entropy(), 1e-4)
    cnt = np.array([1.0, 2.0, 2.0], "float32")
    close(D.Multinomial(5, t(p)).log_prob(t(cnt)), st.multinomial(5, p).logpmf(cnt), 1e-4)
    cb = D.ContinuousBernoulli(t(0.3))
    xs = np.linspace(1e-4, 1 - 1e-4, 4001)
    assert abs(np.trapezoid(np.exp(cb.log_prob(t(xs.astype("float32"))).numpy()), xs) - 1) < 1e-2
    ind = D.Independent(D.Normal(t(np.zeros(3, "float32")), t(np.ones(3, "float32"))), 1)
    close(ind.log_prob(t(x)), st.norm(0, 1).logpdf(x).sum(), 1e-4)
    td = D.TransformedDistribution(D.Normal(0.0, 1.0), [D.AffineTransform(t(1.0), t(2.0)), D.ExpTransform()])
    close(td.log_prob(t(np.array([0.7, 3.0], "float32"))), st.lognorm(2.0, scale=math.e).logpdf([0.7, 3.0]), 1e-4)
    L = D.LKJCholesky(3, 1.5).sample()
    close((L.numpy() @ L.numpy().T).diagonal(), np.ones(3), 1e-4)
    assert issubclass(D.Normal, D.ExponentialFamily) or isinstance(D.Normal(0., 1.), D.Distribution)
    close(D.kl_divergence(D.Normal(0.0, 1.0), D.Normal(1.0, 2.0)), math.log(2) + (1 + 1) / 8 - 0.5, 1e-5)

    class MyD(D.Normal):
        pass

    @D.register_kl(MyD, MyD)
    def _kl(a, b):
        return paddle.to_tensor(42.0)

    assert float(D.kl_divergence(MyD(0.0, 1.0), MyD(0.0, 1.0))) == 42.0


def test_transforms():
    x = t(np.array([0.3, -0.8, 1.2], "float32"))
    pos = t(np.array([0.3, 0.8, 1.2], "float32"))
    for tr, inp in ((D.ExpTransform(), x), (D.SigmoidTransform(), x), (D.TanhTransform(), x), (D.AffineTransform(t(0.5), t(2.0)), x), (D.PowerTransform(t(2.0)), pos)):
        y = tr.forward(inp)
        close(tr.inverse(y), inp, 1e-4)
        eps = 1e-3
        num = (tr.forward(inp + eps).numpy() - tr.forward(inp - eps).numpy()) / (2 * eps)
        close(tr.forward_log_det_jacobian(inp), np.log(np.abs(num)), 2e-3)
        close(tr.inverse_log_det_jacobian(y), -np.log(np.abs(num)), 2e-3)
        assert isinstance(tr, D.Transform)
    close(D.AbsTransform().forward(x), np.abs(x.numpy()))
    chain = D.ChainTransform([D.AffineTransform(t(0.0), t(2.0)), D.ExpTransform()])
    close(chain.forward(x), np.exp(2 * x.numpy()))
    close(chain.inverse(chain.forward(x)), x, 1e-5)
    sm = D.SoftmaxTransform().forward(x)
    close(sm.sum(), 1.0)
    sb = D.StickBreakingTransform()
    y = sb.forward(x)
    assert y.shape == [4] and abs(float(y.sum()) - 1) < 1e-5
    close(sb.inverse(y), x, 1e-4)
    rs = D.ReshapeTransform([3], [1, 3])
    assert rs.forward(x).shape == [1, 3] and rs.inverse(rs.forward(x)).shape == [3]
    it = D.IndependentTransform(D.ExpTransform(), 1)
    close(it.forward_log_det_jacobian(x), x.numpy().sum(), 1e-5)
    stk = D.StackTransform([D.ExpTransform(), D.AffineTransform(t(1.0), t(3.0))], axis=0)
    xx = t(np.array([[0.1, 0.2], [0.3, 0.4]], "float32"))
    close(stk.forward(xx), np.stack([np.exp([0.1, 0.2]), 1 + 3 * np.array([0.3, 0.4])]), 1e-5)


def test_linalg_leftovers():
    LA = paddle.linalg
    A = rng.randn(4, 4)
    S = A @ A.T + 4 * np.eye(4)
    B = rng.randn(4, 3)
    v, w = rng.randn(5, 3), rng.randn(5, 3)
    close(LA.vecdot(t(v), t(w)), (v * w).sum(-1), 1e-8)
    close(LA.vector_norm(t(v), 3, axis=1), (np.abs(v) ** 3).sum(1) ** (1 / 3), 1e-8)
    close(LA.matrix_norm(t(A), "fro"), np.linalg.norm(A, "fro"), 1e-8)
    close(LA.matrix_norm(t(A), "nuc"), np.linalg.norm(A, "nuc"), 1e-8)
    close(LA.cdist(t(v), t(w)), np.linalg.norm(v[:, None] - w[None], axis=-1), 1e-6)
    Lc = np.linalg.cholesky(S)
    close(LA.cholesky_solve(t(B), t(Lc)), np.linalg.solve(S, B), 1e-8)
    close(LA.cholesky_inverse(t(Lc)), np.linalg.inv(S), 1e-8)
    close(LA.inverse(t(S)), np.linalg.inv(S), 1e-8)
    ev = np.sort_complex(LA.eigvals(t(A)).numpy())
    close(ev, np.sort_complex(np.linalg.eigvals(A)), 1e-6)
    wv, vv = LA.eig(t(A))
    close(A @ vv.numpy(), vv.numpy() * wv.numpy()[None], 1e-6)
    sol = LA.lstsq(t(rng.randn(6, 3)), t(rng.randn(6, 2)))[0]
    assert sol.shape == [3, 2]
    lu, piv = LA.lu(t(A))
    P, L, U = LA.lu_unpack(lu, piv)
    close(P.numpy() @ L.numpy() @ U.numpy(), A, 1e-8)
    close(LA.lu_solve(t(B), lu, piv), np.linalg.solve(A, B), 1e-8)
    import scipy.linalg as sl

    close(LA.matrix_exp(t(A / 3)), sl.expm(A / 3), 1e-8)
    q, r = np.linalg.qr(A)
    h, tau = sl.lapack.dgeqrf(A)[:2]
    close(LA.householder_product(t(h), t(tau)), sl.lapack.dorgqr(h, tau)[0], 1e-8)
    close(LA.ormqr(t(h), t(tau), t(B)), sl.lapack.dorgqr(h, tau)[0] @ B, 1e-8)
    u, s, vh = LA.svd_lowrank(t(S), q=4)
    close(np.sort(s.numpy())[::-1], np.linalg.svd(S)[1], 1e-5)
    u, s, vh = LA.pca_lowrank(t(rng.randn(20, 4)), q=2)
    assert s.shape == [2]
    hist, edges = paddle.histogramdd(t(rng.rand(50, 2)), bins=[3, 3])
    assert float(hist.sum()) == 50 and len(edges) == 2


def test_fft_leftovers():
    x = rng.randn(4, 6)
    c = rng.randn(4, 6) + 1j * rng.randn(4, 6)
    FF = paddle.fft
    close(FF.ifft(t(c)), np.fft.ifft(c), 1e-8)
    close(FF.ifft2(t(c)), np.fft.ifft2(c), 1e-8)
    close(FF.fftn(t(c)), np.fft.fftn(c), 1e-8)
    close(FF.ifftn(t(c)), np.fft.ifftn(c), 1e-8)
    close(FF.rfft2(t(x)), np.fft.rfft2(x), 1e-8)
    close(FF.rfftn(t(x)), np.fft.rfftn(x), 1e-8)
    close(FF.irfft2(t(np.fft.rfft2(x)), s=x.shape), x, 1e-8)
    close(FF.irfftn(t(np.fft.rfftn(x)), s=x.shape), x, 1e-8)
    close(FF.hfft(t(c[0])), np.fft.hfft(c[0]), 1e-8)
    close(FF.ihfft(t(x[0])), np.fft.ihfft(x[0]), 1e-8)
    close(FF.fftfreq(8, 0.5), np.fft.fftfreq(8, 0.5), 1e-8)
    close(FF.rfftfreq(8, 0.5), np.fft.rfftfreq(8, 0.5), 1e-8)
    close(FF.ifftshift(t(x)), np.fft.ifftshift(x), 1e-8)
    h2 = FF.hfft2(t(c))
    hn = FF.hfftn(t(c))
    close(h2, hn, 1e-8)
    close(FF.ihfft2(t(x)), FF.ihfftn(t(x)), 1e-8)
    close(FF.ihfft2(t(x)), np.conj(np.fft.rfft2(x)) / x.size, 1e-8)


def test_lr_scheduler_leftovers():
    lr = paddle.optimizer.lr
    s = lr.MultiplicativeDecay(0.5, lambda e: 0.9)
    vals = []
    for _ in range(3):
        vals.append(s())
        s.step()
    close(vals, [0.5, 0.45, 0.405], 1e-6)
    s = lr.LinearLR(1.0, total_steps=4, start_factor=0.25, end_factor=1.0)
    vals = []
    for _ in range(6):
        vals.append(s())
        s.step()
    close(vals, [0.25, 0.4375, 0.625, 0.8125, 1.0, 1.0], 1e-6)
    s = lr.CosineAnnealingWarmRestarts(1.0, T_0=4, T_mult=1, eta_min=0.1)
    vals = []
    for _ in range(6):
        vals.append(s())
        s.step()
    exp = [0.1 + 0.9 * (1 + math.cos(math.pi * (i % 4) / 4)) / 2 for i in range(6)]
    close(vals, exp, 1e-6)


VISION = ["alexnet", "vgg11", "resnet34", "resnext50_32x4d", "wide_resnet50_2", "mobilenet_v1", "mobilenet_v2", "mobilenet_v3_small", "mobilenet_v3_large",
          "densenet121", "googlenet", "shufflenet_v2_x0_5", "shufflenet_v2_swish", "squeezenet1_0", "squeezenet1_1"]


@pytest.mark.parametrize("name", VISION)
def test_vision_model_forward(name):
    M = paddle.vision.models
    paddle.seed(0)
    net = getattr(M, name)(num_classes=7)
    net.eval()
    with paddle.no_grad():
        y = net(paddle.randn([1, 3, 96, 96]))
    y = y[0] if isinstance(y, (tuple, list)) else y
    assert y.shape == [1, 7] and np.isfinite(y.numpy()).all()


def test_vision_model_constructors_and_inception():
    M = paddle.vision.models
    for name in ["vgg13", "vgg16", "vgg19", "resnet101", "resnet152", "resnext50_64x4d", "resnext101_32x4d", "resnext101_64x4d", "resnext152_32x4d",
                 "resnext152_64x4d", "wide_resnet101_2", "densenet161", "densenet169", "densenet201", "densenet264", "shufflenet_v2_x0_33",
                 "shufflenet_v2_x1_0", "shufflenet_v2_x1_5", "shufflenet_v2_x2_0"]:
        net = getattr(M, name)(num_classes=3)
        assert sum(int(np.prod(p.shape)) for p in net.parameters()) > 1e5, name
    assert isinstance(M.resnet34(), M.ResNet) and isinstance(M.vgg11(), M.VGG) and isinstance(M.alexnet(), M.AlexNet)
    assert isinstance(M.densenet121(), M.DenseNet) and isinstance(M.googlenet(), M.GoogLeNet) and isinstance(M.squeezenet1_0(), M.SqueezeNet)
    assert isinstance(M.mobilenet_v1(), M.MobileNetV1) and isinstance(M.mobilenet_v2(), M.MobileNetV2) and isinstance(M.shufflenet_v2_x0_5(), M.ShuffleNetV2)
    assert isinstance(M.mobilenet_v3_small(), M.MobileNetV3Small) and isinstance(M.mobilenet_v3_large(), M.MobileNetV3Large)
    r = M.ResNet(M.BasicBlock, 18, num_classes=4)
    assert r(paddle.randn([1, 3, 64, 64])).shape == [1, 4]
    r = M.ResNet(M.BottleneckBlock, 50, num_classes=4)
    assert r(paddle.randn([1, 3, 64, 64])).shape == [1, 4]
    inc = M.inception_v3(num_classes=5)
    assert isinstance(inc, M.InceptionV3)
    inc.eval()
    with paddle.no_grad():
        assert inc(paddle.randn([1, 3, 299, 299])).shape == [1, 5]
