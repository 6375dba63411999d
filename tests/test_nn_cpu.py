"""nn.Layer machinery, layers and functional ops vs numpy / closed forms. Parity: test/legacy_test/test_layers.py,
test_imperative_*.py, test_*_layer.py."""
import numpy as np
import pytest

import paddle_b200 as paddle
from paddle_b200 import nn
import paddle_b200.nn.functional as F

rng = np.random.RandomState(1)


def T(x, **k):
    return paddle.to_tensor(x, **k)


def test_layer_basics_and_state_dict(tmp_path):
    class Net(nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc1 = nn.Linear(4, 8)
            self.bn = nn.BatchNorm1D(8)
            self.blocks = nn.LayerList([nn.Linear(8, 8) for _ in range(2)])
            self.register_buffer("steps", paddle.zeros([1]))
            self.scale = self.create_parameter([1], default_initializer=nn.initializer.Constant(2.0))

        def forward(self, x):
            x = self.bn(self.fc1(x))
            for b in self.blocks:
                x = F.relu(b(x))
            return x * self.scale

    net = Net()
    names = [n for n, _ in net.named_parameters()]
    assert "fc1.weight" in names and "blocks.1.bias" in names and "scale" in names
    assert len(net.parameters()) == 2 + 2 + 4 + 1 and len(list(net.sublayers())) == 5
    sd = net.state_dict()
    assert "bn._mean" in sd and "steps" in sd and net.fc1.weight.shape == [4, 8]
    paddle.save(sd, str(tmp_path / "m.pdparams"))
    net2 = Net()
    net2.set_state_dict(paddle.load(str(tmp_path / "m.pdparams")))
    x = T(rng.rand(5, 4).astype("float32"))
    net.eval(), net2.eval()
    np.testing.assert_allclose(net(x).numpy(), net2(x).numpy(), rtol=1e-6)
    net.train()
    assert net.training and net.bn.training
    calls = []
    h = net.fc1.register_forward_post_hook(lambda l, i, o: calls.append(o.shape))
    net(x)
    h.remove()
    net(x)
    assert calls == [[5, 8]]
    net.to(dtype="float64")
    assert net.fc1.weight.dtype == paddle.float64
    net.apply(lambda l: None)
    assert "Linear" in repr(net)
    # paddle-style unique parameter names
    assert net.fc1.weight.name.startswith("linear_") and net.fc1.weight.name.endswith(".w_0")


def test_linear_conv_norm_numerics():
    x = rng.rand(2, 3, 8, 8).astype("float32")
    conv = nn.Conv2D(3, 4, 3, padding=1, stride=2)
    y = conv(T(x))
    assert y.shape == [2, 4, 4, 4]
    # direct reference for one output position
    w, b = conv.weight.numpy(), conv.bias.numpy()
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    ref = (xp[0, :, 2:5, 2:5] * w[1]).sum() + b[1]
    np.testing.assert_allclose(y.numpy()[0, 1, 1, 1], ref, rtol=1e-4)
    assert nn.Conv2DTranspose(3, 2, 4, stride=2, padding=1)(T(x)).shape == [2, 2, 16, 16]
    assert nn.Conv1D(3, 5, 3)(T(x[:, :, 0])).shape == [2, 5, 6] and nn.Conv3D(3, 2, 3, padding=1)(T(x[:, :, None])).shape == [2, 2, 1, 8, 8]
    ln = nn.LayerNorm(8)
    z = ln(T(x)).numpy()
    np.testing.assert_allclose(z.mean(-1), 0, atol=1e-5)
    np.testing.assert_allclose(z.std(-1), 1, atol=1e-2)
    bn = nn.BatchNorm2D(3)
    z = bn(T(x)).numpy()
    np.testing.assert_allclose(z.mean((0, 2, 3)), 0, atol=1e-5)
    np.testing.assert_allclose(bn._mean.numpy(), 0.1 * x.mean((0, 2, 3)), rtol=1e-4)   # momentum 0.9
    bn.eval()
    assert np.abs(bn(T(x)).numpy().mean()) > 0.05   # running stats now
    gn = nn.GroupNorm(1, 3)(T(x)).numpy()
    np.testing.assert_allclose(gn.reshape(2, -1).mean(1), 0, atol=1e-5)
    assert nn.InstanceNorm2D(3)(T(x)).shape == [2, 3, 8, 8]
    h = rng.rand(4, 16).astype("float32")
    rms = nn.RMSNorm(16) if hasattr(nn, "RMSNorm") else None
    if rms is not None:
        np.testing.assert_allclose(rms(T(h)).numpy(), h / np.sqrt((h ** 2).mean(-1, keepdims=True) + 1e-6), rtol=1e-4)
    lin = nn.Linear(16, 3)
    np.testing.assert_allclose(lin(T(h)).numpy(), h @ lin.weight.numpy() + lin.bias.numpy(), rtol=1e-5)
    emb = nn.Embedding(10, 4, padding_idx=0)
    e = emb(T(np.array([[0, 3]])))
    assert e.shape == [1, 2, 4] and float(e[0, 0].abs().sum()) == 0


def test_pooling_and_resize():
    x = rng.rand(1, 2, 6, 6).astype("float32")
    np.testing.assert_allclose(nn.MaxPool2D(2)(T(x)).numpy(), x.reshape(1, 2, 3, 2, 3, 2).max((3, 5)))
    np.testing.assert_allclose(nn.AvgPool2D(2)(T(x)).numpy(), x.reshape(1, 2, 3, 2, 3, 2).mean((3, 5)), rtol=1e-6)
    np.testing.assert_allclose(nn.AdaptiveAvgPool2D(1)(T(x)).numpy()[..., 0, 0], x.mean((2, 3)), rtol=1e-6)
    assert nn.AdaptiveMaxPool2D(3)(T(x)).shape == [1, 2, 3, 3] and nn.MaxPool1D(2)(T(x[:, :, 0])).shape == [1, 2, 3]
    assert F.interpolate(T(x), scale_factor=2, mode="nearest").shape == [1, 2, 12, 12]
    assert F.interpolate(T(x), size=[3, 3], mode="bilinear", align_corners=False).shape == [1, 2, 3, 3]
    assert nn.Upsample(scale_factor=2, mode="bicubic")(T(x)).shape == [1, 2, 12, 12]
    assert F.pixel_shuffle(T(rng.rand(1, 8, 3, 3).astype("float32")), 2).shape == [1, 2, 6, 6]
    assert F.unfold(T(x), 3).shape == [1, 18, 16]
    assert nn.Flatten()(T(x)).shape == [1, 72] and nn.Pad2D([1, 1, 1, 1])(T(x)).shape == [1, 2, 8, 8]


ACTS = [("relu", lambda a: np.maximum(a, 0)), ("sigmoid", lambda a: 1 / (1 + np.exp(-a))), ("tanh", np.tanh),
        ("softplus", lambda a: np.log1p(np.exp(a))), ("silu", lambda a: a / (1 + np.exp(-a))), ("relu6", lambda a: np.clip(a, 0, 6)),
        ("leaky_relu", lambda a: np.where(a > 0, a, 0.01 * a)), ("elu", lambda a: np.where(a > 0, a, np.exp(a) - 1)),
        ("hardswish", lambda a: a * np.clip(a + 3, 0, 6) / 6), ("hardsigmoid", lambda a: np.clip(a / 6 + 0.5, 0, 1)),
        ("softsign", lambda a: a / (1 + np.abs(a))), ("mish", lambda a: a * np.tanh(np.log1p(np.exp(a)))),
        ("log_sigmoid", lambda a: -np.log1p(np.exp(-a))), ("tanhshrink", lambda a: a - np.tanh(a))]


@pytest.mark.parametrize("name,ref", ACTS)
def test_activations(name, ref):
    a = (rng.rand(3, 5).astype("float32") - 0.5) * 6
    np.testing.assert_allclose(getattr(F, name)(T(a)).numpy(), ref(a), rtol=1e-4, atol=1e-5)


def test_softmax_gelu_glu_and_layers():
    a = (rng.rand(3, 6).astype("float32") - 0.5) * 4
    e = np.exp(a - a.max(-1, keepdims=True))
    np.testing.assert_allclose(F.softmax(T(a)).numpy(), e / e.sum(-1, keepdims=True), rtol=1e-5)
    np.testing.assert_allclose(F.log_softmax(T(a), axis=0).numpy(), a - np.log(np.exp(a).sum(0, keepdims=True)), rtol=1e-4, atol=1e-5)
    from scipy.special import erf

    np.testing.assert_allclose(F.gelu(T(a)).numpy(), 0.5 * a * (1 + erf(a / np.sqrt(2))), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(F.glu(T(a)).numpy(), a[:, :3] / (1 + np.exp(-a[:, 3:])), rtol=1e-4)
    np.testing.assert_allclose(F.swiglu(T(a[:, :3]), T(a[:, 3:])).numpy(), a[:, :3] / (1 + np.exp(-a[:, :3])) * a[:, 3:], rtol=1e-4, atol=1e-6)
    for cls in (nn.ReLU, nn.GELU, nn.Sigmoid, nn.Tanh, nn.Softmax, nn.LeakyReLU, nn.PReLU, nn.Silu, nn.Hardswish, nn.ELU, nn.Softplus):
        assert cls()(T(a)).shape == [3, 6]
    d = nn.Dropout(0.5)
    d.eval()
    np.testing.assert_array_equal(d(T(a)).numpy(), a)
    d.train()
    z = d(T(np.ones((100, 100), "float32"))).numpy()
    assert 0.35 < (z == 0).mean() < 0.65 and np.allclose(z[z != 0], 2.0)


def test_losses():
    logits = (rng.rand(4, 5).astype("float32") - 0.5) * 3
    lab = np.array([1, 0, 4, 2])
    lsm = logits - np.log(np.exp(logits).sum(-1, keepdims=True))
    np.testing.assert_allclose(F.cross_entropy(T(logits), T(lab)).numpy(), -lsm[np.arange(4), lab].mean(), rtol=1e-5)
    np.testing.assert_allclose(F.cross_entropy(T(logits), T(lab), reduction="none").numpy().reshape(-1), -lsm[np.arange(4), lab], rtol=1e-5)
    soft = np.eye(5, dtype="float32")[lab] * 0.9 + 0.02
    np.testing.assert_allclose(F.cross_entropy(T(logits), T(soft), soft_label=True).numpy(), -(soft * lsm).sum(-1).mean(), rtol=1e-5)
    lab_ig = np.array([1, -100, 4, 2])
    np.testing.assert_allclose(F.cross_entropy(T(logits), T(lab_ig), ignore_index=-100).numpy(), -lsm[[0, 2, 3], [1, 4, 2]].mean(), rtol=1e-5)
    a, b = rng.rand(4, 3).astype("float32"), rng.rand(4, 3).astype("float32")
    np.testing.assert_allclose(F.mse_loss(T(a), T(b)).numpy(), ((a - b) ** 2).mean(), rtol=1e-6)
    np.testing.assert_allclose(F.l1_loss(T(a), T(b), reduction="sum").numpy(), np.abs(a - b).sum(), rtol=1e-6)
    np.testing.assert_allclose(F.binary_cross_entropy(T(a), T((b > 0.5).astype("float32"))).numpy(),
                               -(np.where(b > 0.5, np.log(a), np.log(1 - a))).mean(), rtol=1e-5)
    np.testing.assert_allclose(F.binary_cross_entropy_with_logits(T(logits), T((logits > 0).astype("float32"))).numpy(),
                               np.log1p(np.exp(-np.abs(logits))).mean(), rtol=1e-5)
    d = a - b
    np.testing.assert_allclose(F.smooth_l1_loss(T(a), T(b)).numpy(), np.where(np.abs(d) < 1, 0.5 * d * d, np.abs(d) - 0.5).mean(), rtol=1e-5)
    np.testing.assert_allclose(F.kl_div(T(np.log(a)), T(b), reduction="sum").numpy(), (b * (np.log(b) - np.log(a))).sum(), rtol=1e-5)
    np.testing.assert_allclose(F.nll_loss(T(lsm), T(lab)).numpy(), -lsm[np.arange(4), lab].mean(), rtol=1e-5)
    assert F.margin_ranking_loss(T(a[:, 0]), T(b[:, 0]), T(np.ones(4, "float32"))).shape == []
    assert F.cosine_similarity(T(a), T(b)).shape == [4] and F.normalize(T(a)).shape == [4, 3]
    for cls in (nn.CrossEntropyLoss, nn.MSELoss, nn.L1Loss, nn.BCEWithLogitsLoss, nn.SmoothL1Loss, nn.KLDivLoss, nn.NLLLoss):
        assert cls is not None
    lp = F.log_softmax(T(rng.rand(6, 2, 5).astype("float32")), axis=-1)
    ctc = F.ctc_loss(lp, T(np.array([[1, 2], [3, 3]], "int32")), T(np.array([6, 6])), T(np.array([2, 2])))
    assert np.isfinite(float(ctc))


def test_rnn_and_transformer():
    x = T(rng.rand(2, 5, 4).astype("float32"))
    out, (h, c) = nn.LSTM(4, 6, num_layers=2, direction="bidirect")(x)
    assert out.shape == [2, 5, 12] and h.shape == [4, 2, 6] and c.shape == [4, 2, 6]
    out, h = nn.GRU(4, 6)(x)
    assert out.shape == [2, 5, 6] and h.shape == [1, 2, 6]
    out, h = nn.SimpleRNN(4, 3)(x)
    assert out.shape == [2, 5, 3]
    cell = nn.LSTMCell(4, 6)
    y, (h1, c1) = cell(x[:, 0])
    assert y.shape == [2, 6]
    rnn = nn.RNN(nn.GRUCell(4, 6))
    assert rnn(x)[0].shape == [2, 5, 6]
    mha = nn.MultiHeadAttention(8, 2)
    q = T(rng.rand(2, 5, 8).astype("float32"))
    assert mha(q, q, q).shape == [2, 5, 8]
    enc = nn.TransformerEncoder(nn.TransformerEncoderLayer(8, 2, 16, dropout=0.0), 2)
    assert enc(q).shape == [2, 5, 8]
    tr = nn.Transformer(8, 2, 1, 1, 16, dropout=0.0)
    assert tr(q, q).shape == [2, 5, 8]
    mask = tr.generate_square_subsequent_mask(5)
    assert mask.shape == [5, 5]
    # scaled_dot_product_attention vs explicit softmax
    qq = rng.rand(1, 4, 2, 8).astype("float32")   # [B, S, H, D]
    o = F.scaled_dot_product_attention(T(qq), T(qq), T(qq), is_causal=True).numpy()
    qh = qq.transpose(0, 2, 1, 3)
    s = qh @ qh.transpose(0, 1, 3, 2) / np.sqrt(8)
    s = np.where(np.tril(np.ones((4, 4))) > 0, s, -1e30)
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    np.testing.assert_allclose(o, (p @ qh).transpose(0, 2, 1, 3), rtol=1e-4, atol=1e-5)


def test_initializers_and_clip():
    paddle.seed(3)
    w = nn.Linear(256, 256, weight_attr=paddle.ParamAttr(initializer=nn.initializer.Normal(0.0, 0.02))).weight.numpy()
    assert abs(w.std() - 0.02) < 0.002
    w = nn.Linear(256, 128, weight_attr=paddle.ParamAttr(initializer=nn.initializer.XavierUniform())).weight.numpy()
    assert abs(w.max() - np.sqrt(6 / 384)) < 0.01
    w = nn.Linear(64, 64, weight_attr=paddle.ParamAttr(initializer=nn.initializer.KaimingNormal())).weight.numpy()
    assert abs(w.std() - np.sqrt(2 / 64)) < 0.03
    assert np.all(nn.Linear(3, 3, bias_attr=paddle.ParamAttr(initializer=nn.initializer.Constant(0.5))).bias.numpy() == 0.5)
    o = nn.Linear(16, 16, weight_attr=paddle.ParamAttr(initializer=nn.initializer.Orthogonal())).weight.numpy()
    np.testing.assert_allclose(o @ o.T, np.eye(16), atol=1e-4)
    assert nn.Linear(3, 3, bias_attr=False).bias is None
    p = T(np.ones((2, 2), "float32"), stop_gradient=False)
    (p * 10).sum().backward()
    clip = nn.ClipGradByGlobalNorm(1.0)
    (_, g), = clip([(p, p.grad)])
    np.testing.assert_allclose(np.linalg.norm(g.numpy()), 1.0, rtol=1e-5)
    (_, g), = nn.ClipGradByValue(0.5)([(p, p.grad)])
    assert g.numpy().max() == 0.5
    (_, g), = nn.ClipGradByNorm(2.0)([(p, p.grad)])
    np.testing.assert_allclose(np.linalg.norm(g.numpy()), 2.0, rtol=1e-5)


def test_containers_and_utils():
    seq = nn.Sequential(nn.Linear(2, 3), nn.ReLU(), nn.Linear(3, 1))
    assert len(seq) == 3 and isinstance(seq[1], nn.ReLU) and seq(T(np.ones((4, 2), "float32"))).shape == [4, 1]
    pl = nn.ParameterList([paddle.create_parameter([2], "float32") for _ in range(3)])
    assert len(list(pl.parameters())) == 3
    ld = nn.LayerDict({"a": nn.Linear(1, 1)})
    assert "a" in ld and len(ld) == 1
    v = nn.utils.parameters_to_vector(seq.parameters())
    assert v.shape == [2 * 3 + 3 + 3 + 1]
    nn.utils.vector_to_parameters(v * 0, seq.parameters())
    assert float(seq[0].weight.abs().sum()) == 0
    wn = nn.utils.weight_norm(nn.Linear(4, 4))
    assert any("weight_g" in n for n, _ in wn.named_parameters())
    sn = nn.utils.spectral_norm(nn.Linear(4, 4))
    assert sn(T(np.ones((1, 4), "float32"))).shape == [1, 4]


def test_recompute_matches_plain_with_dropout_and_sequential():
    """recompute / recompute_sequential: same loss and gradients as the plain forward, dropout masks replayed in the re-run."""
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200.distributed.fleet import recompute, recompute_hybrid, recompute_sequential

    def make():
        paddle.seed(5)
        return paddle.nn.Sequential(paddle.nn.Linear(8, 16), paddle.nn.GELU(), paddle.nn.Dropout(0.3), paddle.nn.Linear(16, 8), paddle.nn.Tanh(), paddle.nn.Linear(8, 4))

    x = paddle.to_tensor(np.random.RandomState(0).randn(6, 8).astype("float32"))
    results = []
    for mode in ("plain", "recompute", "sequential", "hybrid_offload"):
        net = make()
        paddle.seed(77)                           # same dropout stream in every mode
        inp = paddle.to_tensor(x.numpy())
        inp.stop_gradient = False
        if mode == "plain":
            y = net(inp)
        elif mode == "recompute":
            y = recompute(net, inp)
        elif mode == "sequential":
            y = recompute_sequential({"segments": 2}, net, inp)
        else:
            y = recompute_hybrid({"mp_group": None, "offload": True, "partition": False}, net, inp)
        loss = (y ** 2).sum()
        loss.backward()
        results.append((float(loss), inp.grad.numpy().copy(), [p.grad.numpy().copy() for p in net.parameters()]))
    for loss, dx, grads in results[1:]:
        assert abs(loss - results[0][0]) < 1e-5
        np.testing.assert_allclose(dx, results[0][1], rtol=1e-5, atol=1e-6)
        for a, b in zip(grads, results[0][2]):
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
