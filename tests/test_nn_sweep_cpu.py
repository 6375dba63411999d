"""nn / nn.functional sweep against plain PyTorch references (layers, pooling, losses, shape ops).
Parity: test/legacy_test/test_*_layer.py / test_*_loss.py families, collapsed into tables."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

import paddle_b200 as paddle

F = paddle.nn.functional
nn = paddle.nn
rng = np.random.RandomState(3)


def t(a):
    return paddle.to_tensor(np.asarray(a))


def tt(a):
    return torch.as_tensor(np.asarray(a))


def close(a, b, tol=1e-5):
    a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    b = b.detach().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


X1 = rng.randn(2, 3, 12).astype("float32")
X2 = rng.randn(2, 4, 8, 8).astype("float32")
X3 = rng.randn(1, 2, 4, 6, 6).astype("float32")
V = rng.randn(4, 6).astype("float32")
W = rng.randn(4, 6).astype("float32")
LBL = rng.randint(0, 6, size=(4,))
PM = np.where(rng.rand(4, 6) > 0.5, 1.0, -1.0).astype("float32")
PROB = (rng.rand(4, 6) * 0.8 + 0.1).astype("float32")
BIN = (rng.rand(4, 6) > 0.5).astype("float32")

STATELESS = {
    # layer factory, torch reference, input
    "AdaptiveAvgPool1D": (lambda: nn.AdaptiveAvgPool1D(4), lambda x: TF.adaptive_avg_pool1d(x, 4), X1),
    "AdaptiveAvgPool3D": (lambda: nn.AdaptiveAvgPool3D(2), lambda x: TF.adaptive_avg_pool3d(x, 2), X3),
    "AdaptiveMaxPool1D": (lambda: nn.AdaptiveMaxPool1D(4), lambda x: TF.adaptive_max_pool1d(x, 4), X1),
    "AdaptiveMaxPool3D": (lambda: nn.AdaptiveMaxPool3D(2), lambda x: TF.adaptive_max_pool3d(x, 2), X3),
    "AvgPool1D": (lambda: nn.AvgPool1D(3, 2), lambda x: TF.avg_pool1d(x, 3, 2), X1),
    "AvgPool3D": (lambda: nn.AvgPool3D(2, 2), lambda x: TF.avg_pool3d(x, 2, 2), X3),
    "MaxPool3D": (lambda: nn.MaxPool3D(2, 2), lambda x: TF.max_pool3d(x, 2, 2), X3),
    "LPPool1D": (lambda: nn.LPPool1D(2, 3, 2), lambda x: TF.lp_pool1d(x, 2, 3, 2), np.abs(X1)),
    "LPPool2D": (lambda: nn.LPPool2D(2, 2, 2), lambda x: TF.lp_pool2d(x, 2, 2, 2), np.abs(X2)),
    "CELU": (lambda: nn.CELU(1.2), lambda x: TF.celu(x, 1.2), V), "GLU": (lambda: nn.GLU(), lambda x: TF.glu(x), V),
    "Hardshrink": (lambda: nn.Hardshrink(), TF.hardshrink, V), "Hardsigmoid": (lambda: nn.Hardsigmoid(), lambda x: torch.clamp(x * 0.1666667 + 0.5, 0, 1), V),
    "Hardtanh": (lambda: nn.Hardtanh(), TF.hardtanh, V * 2), "Identity": (lambda: nn.Identity(), lambda x: x, V),
    "LogSigmoid": (lambda: nn.LogSigmoid(), TF.logsigmoid, V), "LogSoftmax": (lambda: nn.LogSoftmax(), lambda x: TF.log_softmax(x, -1), V),
    "Mish": (lambda: nn.Mish(), TF.mish, V), "ReLU6": (lambda: nn.ReLU6(), TF.relu6, V * 4), "SELU": (lambda: nn.SELU(), TF.selu, V),
    "Softshrink": (lambda: nn.Softshrink(), TF.softshrink, V), "Softsign": (lambda: nn.Softsign(), TF.softsign, V),
    "Swish": (lambda: nn.Swish(), TF.silu, V), "Tanhshrink": (lambda: nn.Tanhshrink(), TF.tanhshrink, V),
    "ThresholdedReLU": (lambda: nn.ThresholdedReLU(0.2), lambda x: torch.where(x > 0.2, x, torch.zeros_like(x)), V),
    "Softmax2D": (lambda: nn.Softmax2D(), lambda x: TF.softmax(x, -3), X2), "Maxout": (lambda: nn.Maxout(2), lambda x: x.reshape(2, 2, 2, 8, 8).max(2)[0], X2),
    "RReLU_eval": (lambda: _eval(nn.RReLU(0.1, 0.3)), lambda x: TF.rrelu(x, 0.1, 0.3, training=False), V),
    "Pad1D": (lambda: nn.Pad1D([1, 2]), lambda x: TF.pad(x, [1, 2]), X1), "Pad3D": (lambda: nn.Pad3D([1, 1, 0, 2, 1, 0]), lambda x: TF.pad(x, [1, 1, 0, 2, 1, 0]), X3),
    "PixelShuffle": (lambda: nn.PixelShuffle(2), lambda x: TF.pixel_shuffle(x, 2), X2), "PixelUnshuffle": (lambda: nn.PixelUnshuffle(2), lambda x: TF.pixel_unshuffle(x, 2), X2),
    "ChannelShuffle": (lambda: nn.ChannelShuffle(2), lambda x: TF.channel_shuffle(x, 2), X2),
    "Unfold": (lambda: nn.Unfold(3, 1, 1, 1), lambda x: TF.unfold(x, 3, 1, 1, 1), X2),
    "Fold": (lambda: nn.Fold([4, 4], 2, strides=2), lambda x: TF.fold(x, [4, 4], 2, 1, 0, 2), rng.randn(2, 12, 4).astype("float32")),
    "Unflatten": (lambda: nn.Unflatten(1, [2, 3]), lambda x: x.reshape(4, 2, 3), V),
    "CosineSimilarity": (lambda: _two(nn.CosineSimilarity(axis=1), W), lambda x: TF.cosine_similarity(x, tt(W), 1), V),
    "PairwiseDistance": (lambda: _two(nn.PairwiseDistance(), W), lambda x: TF.pairwise_distance(x, tt(W)), V),
    "LocalResponseNorm": (lambda: nn.LocalResponseNorm(3), lambda x: TF.local_response_norm(x, 3, 1e-4, 0.75, 1.0), X2),
    "InstanceNorm1D": (lambda: nn.InstanceNorm1D(3), lambda x: TF.instance_norm(x, weight=torch.ones(3), bias=torch.zeros(3)), X1),
    "InstanceNorm3D": (lambda: nn.InstanceNorm3D(2), lambda x: TF.instance_norm(x, weight=torch.ones(2), bias=torch.zeros(2)), X3),
    "Dropout2D_eval": (lambda: _eval(nn.Dropout2D(0.5)), lambda x: x, X2), "Dropout3D_eval": (lambda: _eval(nn.Dropout3D(0.5)), lambda x: x, X3),
    "AlphaDropout_eval": (lambda: _eval(nn.AlphaDropout(0.5)), lambda x: x, V), "FeatureAlphaDropout_eval": (lambda: _eval(nn.FeatureAlphaDropout(0.5)), lambda x: x, X2),
}


def _eval(layer):
    layer.eval()
    return layer


def _two(layer, other):
    return lambda x: layer(x, t(other))


@pytest.mark.parametrize("name", sorted(STATELESS))
def test_stateless_layer(name):
    make, ref, x = STATELESS[name]
    close(make()(t(x)), ref(tt(x)))


LOSSES = {
    "BCELoss": (lambda: nn.BCELoss(), lambda a, b: TF.binary_cross_entropy(a, b), PROB, BIN),
    "CosineEmbeddingLoss": (lambda: (lambda a, b: nn.CosineEmbeddingLoss(margin=0.1)(a, t(W), b)), lambda a, b: TF.cosine_embedding_loss(a, tt(W), b, margin=0.1), V, PM[:, 0]),
    "GaussianNLLLoss": (lambda: (lambda a, b: nn.GaussianNLLLoss()(a, b, t(PROB))), lambda a, b: TF.gaussian_nll_loss(a, b, tt(PROB)), V, W),
    "HingeEmbeddingLoss": (lambda: nn.HingeEmbeddingLoss(), lambda a, b: TF.hinge_embedding_loss(a, b), V, PM),
    "HuberLoss": (lambda: nn.HuberLoss(delta=0.7), lambda a, b: TF.huber_loss(a, b, delta=0.7), V, W),
    "MarginRankingLoss": (lambda: (lambda a, b: nn.MarginRankingLoss(0.2)(a, t(W), b)), lambda a, b: TF.margin_ranking_loss(a, tt(W), b, margin=0.2), V, PM),
    "MultiLabelSoftMarginLoss": (lambda: nn.MultiLabelSoftMarginLoss(), lambda a, b: TF.multilabel_soft_margin_loss(a, b), V, BIN),
    "MultiMarginLoss": (lambda: nn.MultiMarginLoss(), lambda a, b: TF.multi_margin_loss(a, b), V, LBL),
    "PoissonNLLLoss": (lambda: nn.PoissonNLLLoss(), lambda a, b: TF.poisson_nll_loss(a, b), V, np.abs(W)),
    "SoftMarginLoss": (lambda: nn.SoftMarginLoss(), lambda a, b: TF.soft_margin_loss(a, b), V, PM),
    "TripletMarginLoss": (lambda: (lambda a, b: nn.TripletMarginLoss()(a, b, t(PROB))), lambda a, b: TF.triplet_margin_loss(a, b, tt(PROB)), V, W),
    "TripletMarginWithDistanceLoss": (lambda: (lambda a, b: nn.TripletMarginWithDistanceLoss()(a, b, t(PROB))), lambda a, b: TF.triplet_margin_with_distance_loss(a, b, tt(PROB)), V, W),
    "square_error_cost": (lambda: F.square_error_cost, lambda a, b: (a - b) ** 2, V, W),
    "log_loss": (lambda: F.log_loss, lambda a, b: -b * torch.log(a + 1e-4) - (1 - b) * torch.log(1 - a + 1e-4), PROB[:, :1], BIN[:, :1]),
    "label_smooth": (lambda: (lambda a, b: F.label_smooth(a, epsilon=0.1)), lambda a, b: a * 0.9 + 0.1 / 6, BIN, BIN),
    "sigmoid_focal_loss": (lambda: (lambda a, b: F.sigmoid_focal_loss(a, b, reduction="mean")), lambda a, b: _focal(a, b), V, BIN),
    "dice_loss": (lambda: (lambda a, b: F.dice_loss(a, b)), lambda a, b: _dice(a, b), PROB, LBL[:, None]),
    "softmax_with_cross_entropy": (lambda: (lambda a, b: F.softmax_with_cross_entropy(a, b)), lambda a, b: TF.cross_entropy(a, b[:, 0], reduction="none")[:, None], V, LBL[:, None]),
}


def _focal(a, b, alpha=0.25, gamma=2.0):
    p = torch.sigmoid(a)
    ce = TF.binary_cross_entropy_with_logits(a, b, reduction="none")
    pt = p * b + (1 - p) * (1 - b)
    return (ce * (1 - pt) ** gamma * (alpha * b + (1 - alpha) * (1 - b))).mean()


def _dice(a, b, eps=1e-5):
    oh = TF.one_hot(b[:, 0], 6).float()
    inter = (a * oh).sum(1)
    return (1 - 2 * inter / (a.sum(1) + oh.sum(1) + eps)).mean()


@pytest.mark.parametrize("name", sorted(LOSSES))
def test_loss(name):
    make, ref, a, b = LOSSES[name]
    close(make()(t(a), t(b)), ref(tt(a), tt(b)))


def test_functional_conv_pool_norm_family():
    w1 = rng.randn(5, 3, 3).astype("float32")
    close(F.conv1d(t(X1), t(w1), padding=1), TF.conv1d(tt(X1), tt(w1), padding=1), 1e-4)
    wt1 = rng.randn(3, 5, 3).astype("float32")
    close(F.conv1d_transpose(t(X1), t(wt1), stride=2), TF.conv_transpose1d(tt(X1), tt(wt1), stride=2), 1e-4)
    wt2 = rng.randn(4, 5, 3, 3).astype("float32")
    close(F.conv2d_transpose(t(X2), t(wt2), stride=2, padding=1), TF.conv_transpose2d(tt(X2), tt(wt2), stride=2, padding=1), 1e-4)
    w3 = rng.randn(3, 2, 3, 3, 3).astype("float32")
    close(F.conv3d(t(X3), t(w3), padding=1), TF.conv3d(tt(X3), tt(w3), padding=1), 1e-4)
    wt3 = rng.randn(2, 3, 3, 3, 3).astype("float32")
    close(F.conv3d_transpose(t(X3), t(wt3)), TF.conv_transpose3d(tt(X3), tt(wt3)), 1e-4)
    close(nn.Conv1DTranspose(3, 5, 3)(t(X1)).shape, [2, 5, 14], 0)
    close(nn.Conv3DTranspose(2, 3, 3)(t(X3)).shape, [1, 3, 6, 8, 8], 0)
    for k in (1, 2, 3):
        x = {1: X1, 2: X2, 3: X3}[k]
        close(getattr(F, f"max_pool{k}d")(t(x), 2, 2), getattr(TF, f"max_pool{k}d")(tt(x), 2, 2))
        close(getattr(F, f"avg_pool{k}d")(t(x), 2, 2), getattr(TF, f"avg_pool{k}d")(tt(x), 2, 2))
        close(getattr(F, f"adaptive_avg_pool{k}d")(t(x), 2), getattr(TF, f"adaptive_avg_pool{k}d")(tt(x), 2))
        close(getattr(F, f"adaptive_max_pool{k}d")(t(x), 2), getattr(TF, f"adaptive_max_pool{k}d")(tt(x), 2))
        out, idx = getattr(F, f"max_pool{k}d")(t(x), 2, 2, return_mask=True)
        ro, ri = getattr(TF, f"max_pool{k}d")(tt(x), 2, 2, return_indices=True)
        close(getattr(F, f"max_unpool{k}d")(out, idx, 2, 2), getattr(TF, f"max_unpool{k}d")(ro, ri, 2, 2))
        up = getattr(nn, f"MaxUnPool{k}D")(2, 2)(out, idx)
        close(up, getattr(TF, f"max_unpool{k}d")(ro, ri, 2, 2))
    rm, rv = np.zeros(4, "float32"), np.ones(4, "float32")
    g, b = rng.rand(4).astype("float32"), rng.randn(4).astype("float32")
    close(F.batch_norm(t(X2), t(rm), t(rv), t(g), t(b), training=False), TF.batch_norm(tt(X2), tt(rm), tt(rv), tt(g), tt(b), False), 1e-4)
    close(F.group_norm(t(X2), 2, weight=t(g), bias=t(b)), TF.group_norm(tt(X2), 2, tt(g), tt(b)), 1e-4)
    close(F.instance_norm(t(X2), weight=t(g), bias=t(b)), TF.instance_norm(tt(X2), weight=tt(g), bias=tt(b)), 1e-4)
    close(F.local_response_norm(t(X2), 3), TF.local_response_norm(tt(X2), 3, 1e-4, 0.75, 1.0), 1e-5)
    close(nn.BatchNorm3D(2)(t(X3)).shape, list(X3.shape), 0)
    bn = nn.BatchNorm(4)
    bn.eval()
    close(bn(t(X2)), X2 / np.sqrt(1 + 1e-5), 1e-4)
    sbn = nn.SyncBatchNorm(4)
    close(sbn(t(X2)), TF.batch_norm(tt(X2), None, None, torch.ones(4), torch.zeros(4), True), 1e-4)
    assert isinstance(nn.SyncBatchNorm.convert_sync_batchnorm(nn.Sequential(nn.BatchNorm2D(4)))[0], nn.SyncBatchNorm)


def test_functional_misc_family():
    theta = rng.randn(2, 2, 3).astype("float32")
    grid = F.affine_grid(t(theta), [2, 4, 5, 5], align_corners=False)
    rg = TF.affine_grid(tt(theta), [2, 4, 5, 5], align_corners=False)
    close(grid, rg)
    close(F.grid_sample(t(X2), grid, align_corners=False), TF.grid_sample(tt(X2), rg, align_corners=False), 1e-4)
    close(F.diag_embed(t(V)), torch.diag_embed(tt(V)))
    close(F.sequence_mask(t(np.array([1, 3, 2])), 4), np.array([[1, 0, 0, 0], [1, 1, 1, 0], [1, 1, 0, 0]]))
    close(F.prelu(t(V), t(np.array([0.2], "float32"))), TF.prelu(tt(V), torch.tensor([0.2])))
    close(F.maxout(t(X2), 2), tt(X2).reshape(2, 2, 2, 8, 8).max(2)[0])
    close(F.pairwise_distance(t(V), t(W)), TF.pairwise_distance(tt(V), tt(W)))
    close(F.pixel_unshuffle(t(X2), 2), TF.pixel_unshuffle(tt(X2), 2))
    close(F.channel_shuffle(t(X2), 2), TF.channel_shuffle(tt(X2), 2))
    x = t(V.copy())
    for name, ref in (("relu_", TF.relu), ("elu_", TF.elu), ("leaky_relu_", lambda a: TF.leaky_relu(a, 0.01)), ("hardtanh_", TF.hardtanh), ("softmax_", lambda a: TF.softmax(a, -1))):
        y = t(V.copy())
        out = getattr(F, name)(y)
        close(out, ref(tt(V)))
        close(y, ref(tt(V)))   # in place
    paddle.seed(0)
    gs = F.gumbel_softmax(t(V), temperature=0.5, hard=True)
    assert np.allclose(gs.numpy().sum(-1), 1) and set(np.unique(gs.numpy())) <= {0.0, 1.0}
    assert F.dropout2d(t(X2), 0.5, training=False).shape == list(X2.shape) and F.dropout3d(t(X3), 0.5, training=False).shape == list(X3.shape)
    assert F.alpha_dropout(t(V), 0.5, training=False).shape == [4, 6]
    r = F.rrelu(t(V), 0.1, 0.3, training=True).numpy()
    neg = V < 0
    ratio = r[neg] / V[neg]
    assert np.allclose(r[~neg], V[~neg]) and (ratio >= 0.1 - 1e-6).all() and (ratio <= 0.3 + 1e-6).all()
    ids = np.array([[[2, 2], [6, 1]], [[3, 9], [6, 1]], [[0, 1], [9, 0]]])
    parents = np.array([[[0, 0], [1, 1]], [[1, 0], [1, 0]], [[0, 0], [0, 1]]])
    assert F.gather_tree(t(ids), t(parents)).numpy().tolist() == [[[2, 2], [1, 6]], [[3, 3], [6, 1]], [[0, 1], [9, 0]]]
    anchor, pos = rng.randn(6, 8).astype("float32"), rng.randn(6, 8).astype("float32")
    assert np.isfinite(float(F.npair_loss(t(anchor), t(pos), t(np.arange(6)))))
    lab, sampled = F.class_center_sample(t(np.array([3, 7, 3, 1])), 10, 5)
    assert len(sampled.numpy()) == 5 and {1, 3, 7} <= set(sampled.numpy().tolist()) and (sampled.numpy()[lab.numpy()] == [3, 7, 3, 1]).all()
    logits = rng.randn(4, 10).astype("float32") * 0.1
    assert np.isfinite(float(F.margin_cross_entropy(t(np.clip(logits, -1, 1)), t(np.array([1, 2, 3, 4])), reduction="mean")))


def test_rnn_cells_birnn_transformer_decoder_spectral_norm():
    cell = nn.SimpleRNNCell(6, 5)
    assert isinstance(cell, nn.RNNCellBase)
    h, _ = cell(t(V))
    ref = torch.tanh(tt(V) @ tt(cell.weight_ih.numpy()).T + tt(cell.bias_ih.numpy()) + tt(cell.bias_hh.numpy()))
    close(h, ref, 1e-5)
    bi = nn.BiRNN(nn.GRUCell(6, 5), nn.GRUCell(6, 5))
    out, _ = bi(t(rng.randn(2, 7, 6).astype("float32")))
    assert out.shape == [2, 7, 10]
    dec = nn.TransformerDecoder(nn.TransformerDecoderLayer(8, 2, 16, dropout=0.0), 2)
    y = dec(t(rng.randn(2, 5, 8).astype("float32")), t(rng.randn(2, 7, 8).astype("float32")))
    assert y.shape == [2, 5, 8] and np.isfinite(y.numpy()).all()
    sn = nn.SpectralNorm([5, 6], dim=0, power_iters=30)
    w = rng.randn(5, 6).astype("float32")
    close(np.linalg.svd(sn(t(w)).numpy())[1][0], 1.0, 1e-2)
    bl = nn.Bilinear(6, 6, 3)
    close(bl(t(V), t(W)), TF.bilinear(tt(V), tt(W), tt(bl.weight.numpy()), tt(bl.bias.numpy()).reshape(-1)), 1e-4)
    pd = nn.ParameterDict({"a": paddle.create_parameter([2], "float32")})
    assert "a" in pd and len(list(pd.parameters())) == 1


def test_seq_losses_ctc_rnnt_hsigmoid_adaptive():
    T_, B_, C_ = 6, 2, 5
    logits = rng.randn(T_, B_, C_).astype("float32")
    labels = np.array([[1, 2, 0], [3, 3, 4]], "int32")
    il, ll = np.array([6, 5], "int64"), np.array([2, 3], "int64")
    ours = nn.CTCLoss(blank=0, reduction="none")(t(logits), t(labels), t(il), t(ll))
    ref = TF.ctc_loss(TF.log_softmax(tt(logits), -1), tt(labels).long(), tt(il), tt(ll), blank=0, reduction="none")
    close(ours.reshape([-1]), ref, 1e-4)
    acts = rng.randn(1, 4, 3, 5).astype("float32")
    l = nn.RNNTLoss(blank=0, fastemit_lambda=0.0)(t(acts), t(np.array([[1, 2]], "int32")), t(np.array([4], "int32")), t(np.array([2], "int32")))
    assert np.isfinite(float(l)) and float(l) > 0
    hs = nn.HSigmoidLoss(6, 8)
    assert hs(t(V), t(LBL[:, None])).shape == [4, 1]
    al = nn.AdaptiveLogSoftmaxWithLoss(6, 12, [4, 8])
    out, loss = al(t(V), t(np.array([0, 5, 9, 11])))
    assert out.shape == [4] and np.isfinite(float(loss))
    close(np.exp(al.log_prob(t(V)).numpy()).sum(-1), np.ones(4), 1e-4)
    fm = nn.FractionalMaxPool2D(output_size=3)(t(X2))
    assert fm.shape == [2, 4, 3, 3]
    fm3 = nn.FractionalMaxPool3D(output_size=2)(t(X3))
    assert fm3.shape == [1, 2, 2, 2, 2]
