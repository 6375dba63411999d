"""Seventh sweep: attention API family vs a dense reference, file-based vision / audio / text datasets from synthetic archives, decode,
callbacks, device / version helpers."""
import io
import os
import pickle
import tarfile
import wave

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

import paddle_b200 as paddle

rng = np.random.RandomState(61)
F = paddle.nn.functional


def t(a):
    return paddle.to_tensor(np.asarray(a))


def close(a, b, tol=1e-4):
    a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    b = b.detach().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


def _ref_attn(q, k, v, causal=False, mask=None):
    """q, k, v: [B, S, H, D] numpy."""
    qt, kt, vt = (torch.as_tensor(x).transpose(1, 2) for x in (q, k, v))
    return TF.scaled_dot_product_attention(qt, kt, vt, attn_mask=mask, is_causal=causal).transpose(1, 2)


def test_attention_api_family():
    B, S, H, D = 2, 6, 2, 8
    q, k, v = (rng.randn(B, S, H, D).astype("float32") for _ in range(3))
    ref = _ref_attn(q, k, v, causal=True)
    close(F.flash_attention(t(q), t(k), t(v), causal=True)[0], ref)
    close(F.scaled_dot_product_attention(t(q), t(k), t(v), is_causal=True), ref)
    qkv = np.stack([q, k, v], 2)                                         # [B, S, 3, H, D]
    close(F.flash_attn_qkvpacked(t(qkv), causal=True)[0], ref)
    # varlen: two sequences of length 4 and 2 packed along the token dimension
    lens = [4, 2]
    cu = np.array([0, 4, 6], "int32")
    qp = np.concatenate([q[0, :4], q[1, :2]])                             # [6, H, D]
    kp, vp = np.concatenate([k[0, :4], k[1, :2]]), np.concatenate([v[0, :4], v[1, :2]])
    out = F.flash_attn_unpadded(t(qp), t(kp), t(vp), t(cu), t(cu), 4, 4, D ** -0.5, causal=True)[0]
    close(out[:4], _ref_attn(q[:1, :4], k[:1, :4], v[:1, :4], causal=True)[0])
    close(out[4:], _ref_attn(q[1:, :2], k[1:, :2], v[1:, :2], causal=True)[0])
    qkvp = np.stack([qp, kp, vp], 1)                                      # [T, 3, H, D]
    out2 = F.flash_attn_varlen_qkvpacked(t(qkvp), t(cu), t(cu), 4, 4, D ** -0.5, causal=True, varlen_padded=False)[0]
    close(out2, out)
    # flashmask: startend_row_indices [B, H or 1, S, 1] = first masked row per key column (causal document mask)
    start = np.full((B, 1, S, 1), S, "int32")
    start[:, :, :3] = 3                                                   # keys 0..2 are invisible to queries >= 3: two documents of length 3
    fm = F.flashmask_attention(t(q), t(k), t(v), t(start), causal=True)
    blk = torch.ones(S, S, dtype=torch.bool).tril()
    blk[3:, :3] = False
    close(fm, _ref_attn(q, k, v, mask=blk))
    # sparse_attention with a full CSR pattern == dense attention ([B, H, S, D] layout)
    qh, kh, vh = (x.transpose(0, 2, 1, 3).copy() for x in (q, k, v))
    offs = np.tile(np.arange(0, S * S + 1, S, dtype="int32"), (B, H, 1))
    cols = np.tile(np.tile(np.arange(S, dtype="int32"), S), (B, H, 1))
    close(F.sparse_attention(t(qh), t(kh), t(vh), t(offs), t(cols)), _ref_attn(q, k, v).transpose(1, 2))
    # reduced scores: column sums of the attention probabilities
    lse = torch.logsumexp(torch.einsum("bhqd,bhkd->bhqk", torch.as_tensor(qh), torch.as_tensor(kh)) * D ** -0.5, -1)
    red = F.calc_reduced_attention_scores(t(q), t(k), t(lse.numpy()))          # [B, S, H, D] inputs, lse [B, H, S]
    probs = torch.softmax(torch.einsum("bhqd,bhkd->bhqk", torch.as_tensor(qh), torch.as_tensor(kh)) * D ** -0.5, -1)
    close(red.reshape([B, H, S]), probs.sum(2))
    with F.sdp_kernel(enable_flash=True, enable_math=True, enable_mem_efficient=True):
        close(F.scaled_dot_product_attention(t(q), t(k), t(v)), _ref_attn(q, k, v))


def test_pool_pad_upsample_leftovers():
    nn = paddle.nn
    x = rng.randn(2, 3, 8, 8).astype("float32")
    xt = torch.as_tensor(x)
    close(F.avg_pool2d(t(x), 2), TF.avg_pool2d(xt, 2))
    close(F.max_pool2d(t(x), 3, 2, 1), TF.max_pool2d(xt, 3, 2, 1))
    close(F.adaptive_avg_pool2d(t(x), 3), TF.adaptive_avg_pool2d(xt, 3))
    close(F.adaptive_max_pool2d(t(x), 3), TF.adaptive_max_pool2d(xt, 3))
    close(F.max_pool1d(t(x[:, :, 0]), 2), TF.max_pool1d(xt[:, :, 0], 2))
    close(nn.ZeroPad2D([1, 2, 0, 1])(t(x)), TF.pad(xt, [1, 2, 0, 1]))
    close(F.zeropad2d(t(x), [1, 1, 1, 1]), TF.pad(xt, [1, 1, 1, 1]))
    close(nn.ZeroPad1D([1, 2])(t(x[:, :, 0])), TF.pad(xt[:, :, 0], [1, 2]))
    assert nn.ZeroPad3D([1, 1, 1, 1, 1, 1])(t(rng.randn(1, 1, 2, 2, 2).astype("float32"))).shape == [1, 1, 4, 4, 4]
    close(nn.UpsamplingNearest2D(scale_factor=2)(t(x)), TF.interpolate(xt, scale_factor=2, mode="nearest"))
    close(nn.UpsamplingBilinear2D(size=[12, 12])(t(x)), TF.interpolate(xt, size=[12, 12], mode="bilinear", align_corners=True), 1e-4)
    close(F.upsample(t(x), scale_factor=2, mode="nearest"), TF.interpolate(xt, scale_factor=2, mode="nearest"))
    y = t(x.copy())
    F.tanh_(y)
    close(y, np.tanh(x), 1e-6)
    y = t(x.copy())
    F.thresholded_relu_(y, 0.5)
    close(y, np.where(x > 0.5, x, 0))
    assert F.fractional_max_pool2d(t(x), output_size=3).shape == [2, 3, 3, 3]
    assert F.fractional_max_pool3d(t(rng.randn(1, 2, 6, 6, 6).astype("float32")), output_size=2).shape == [1, 2, 2, 2, 2]
    a = F.feature_alpha_dropout(t(x), 0.5, training=False)
    close(a, x)
    lab = (rng.rand(4, 6) > 0.5).astype("float32")
    logits = rng.randn(4, 6).astype("float32")
    close(F.multi_label_soft_margin_loss(t(logits), t(lab)), TF.multilabel_soft_margin_loss(torch.as_tensor(logits), torch.as_tensor(lab)))
    hs = F.hsigmoid_loss(t(logits), t(np.array([[0], [1], [2], [3]])), 5, t(rng.randn(4, 6).astype("float32")))
    assert hs.shape == [4, 1] and np.isfinite(hs.numpy()).all()
    acts = rng.randn(1, 4, 3, 5).astype("float32")
    rl = F.rnnt_loss(t(acts), t(np.array([[1, 2]], "int32")), t(np.array([4], "int32")), t(np.array([2], "int32")), blank=0, fastemit_lambda=0.0)
    assert np.isfinite(float(rl)) and float(rl) > 0
    out, loss = F.adaptive_log_softmax_with_loss(t(rng.randn(3, 8).astype("float32")), t(np.array([0, 5, 9])), t(rng.randn(8, 6).astype("float32")),
                                                 [(t(rng.randn(8, 2).astype("float32")), t(rng.randn(2, 4).astype("float32"))), (t(rng.randn(8, 1).astype("float32")), t(rng.randn(1, 2).astype("float32")))],
                                                 [4, 8, 10])
    assert out.shape == [3] and np.isfinite(float(loss))


def test_beam_search_decode():
    nn = paddle.nn
    paddle.seed(2)
    V, Hd = 7, 8
    emb = nn.Embedding(V, Hd)
    cell = nn.GRUCell(Hd, Hd)
    proj = nn.Linear(Hd, V)
    dec = nn.BeamSearchDecoder(cell, start_token=0, end_token=1, beam_size=3, embedding_fn=emb, output_fn=proj)
    init = paddle.zeros([2, Hd])
    outs, states, lens = nn.dynamic_decode(dec, inits=init, max_step_num=5, return_length=True)
    ids = outs.numpy() if hasattr(outs, "numpy") else np.asarray(outs[0])
    assert ids.shape[0] == 2 and ids.shape[-1] == 3 and ids.max() < V and lens.shape[0] == 2


def test_vision_datasets_from_synthetic_files(tmp_path):
    V = paddle.vision.datasets
    from PIL import Image

    # CIFAR-10 python archive
    p = str(tmp_path / "cifar-10-python.tar.gz")
    with tarfile.open(p, "w:gz") as tf:
        for name, n in (("data_batch_1", 6), ("test_batch", 4)):
            blob = pickle.dumps({b"data": (rng.rand(n, 3072) * 255).astype("uint8"), b"labels": list(range(n))})
            info = tarfile.TarInfo(f"cifar-10-batches-py/{name}")
            info.size = len(blob)
            tf.addfile(info, io.BytesIO(blob))
    tr, te = V.Cifar10(p, mode="train"), V.Cifar10(p, mode="test")
    img, lab = tr[2]
    assert len(tr) == 6 and len(te) == 4 and np.asarray(img).shape[-1] in (3, 32) and int(lab) == 2
    p100 = str(tmp_path / "cifar-100-python.tar.gz")
    with tarfile.open(p100, "w:gz") as tf:
        for name, n in (("train", 5), ("test", 3)):
            blob = pickle.dumps({b"data": (rng.rand(n, 3072) * 255).astype("uint8"), b"fine_labels": list(range(n))})
            info = tarfile.TarInfo(f"cifar-100-python/{name}")
            info.size = len(blob)
            tf.addfile(info, io.BytesIO(blob))
    assert len(V.Cifar100(p100, mode="train")) == 5
    # MNIST-style idx files (FashionMNIST shares the format)
    import gzip
    import struct

    imgs = (rng.rand(5, 28, 28) * 255).astype("uint8")
    ip, lp = str(tmp_path / "img.gz"), str(tmp_path / "lab.gz")
    with gzip.open(ip, "wb") as f:
        f.write(struct.pack(">IIII", 2051, 5, 28, 28) + imgs.tobytes())
    with gzip.open(lp, "wb") as f:
        f.write(struct.pack(">II", 2049, 5) + bytes([0, 1, 2, 3, 4]))
    fm = V.FashionMNIST(ip, lp, mode="train")
    im, lb = fm[3]
    assert len(fm) == 5 and int(np.asarray(lb).reshape(-1)[0]) == 3 and np.asarray(im).shape[-2:] == (28, 28)
    # folder datasets
    for cls in ("cat", "dog"):
        os.makedirs(tmp_path / "folder" / cls)
        for i in range(2):
            Image.fromarray((rng.rand(8, 8, 3) * 255).astype("uint8")).save(tmp_path / "folder" / cls / f"{i}.png")
    df = V.DatasetFolder(str(tmp_path / "folder"))
    assert len(df) == 4 and df.classes == ["cat", "dog"] and df[3][1] == 1 and ".png" in V.IMG_EXTENSIONS
    imf = V.ImageFolder(str(tmp_path / "folder" / "cat"))
    assert len(imf) == 2 and np.asarray(imf[0][0]).shape == (8, 8, 3)
    assert np.asarray(V.default_loader(str(tmp_path / "folder" / "cat" / "0.png"))).shape == (8, 8, 3)
    paddle.vision.set_image_backend("pil")
    assert paddle.vision.get_image_backend() == "pil"
    assert np.asarray(paddle.vision.image_load(str(tmp_path / "folder" / "cat" / "0.png"))).shape == (8, 8, 3)


def test_audio_backends_and_datasets(tmp_path):
    A = paddle.audio
    sr = 8000
    sig = (np.sin(np.linspace(0, 440 * 2 * np.pi, sr // 4)) * 0.5).astype("float32")
    p = str(tmp_path / "a.wav")
    A.backends.save(p, t(sig[None]), sr)
    info = A.backends.info(p)
    assert isinstance(info, A.backends.AudioInfo) and info.sample_rate == sr and info.num_channels == 1
    wavf, sr2 = A.backends.load(p)
    assert sr2 == sr and abs(float(wavf.abs().max()) - 0.5) < 0.01
    assert A.backends.get_current_backend() in A.backends.list_available_backends()
    A.backends.set_backend(A.backends.get_current_backend())
    with wave.open(p) as w:
        assert w.getframerate() == sr

    class Tiny(A.datasets.AudioClassificationDataset):
        pass

    ds = Tiny(files=[p, p], labels=[0, 1], feat_type="raw")
    x, y = ds[1]
    assert len(ds) == 2 and int(y) == 1 and np.asarray(x).shape[-1] == len(sig)
    ds2 = Tiny(files=[p], labels=[1], feat_type="melspectrogram", sample_rate=sr, n_fft=256, n_mels=16)
    assert np.asarray(ds2[0][0]).shape[0] == 16
    assert A.datasets.ESC50 is not None and A.datasets.TESS is not None


def test_text_leftovers(tmp_path):
    T = paddle.text
    data = rng.rand(20, 14)
    p = tmp_path / "housing.data"
    p.write_text("\n".join(" ".join(f"{v:.6f}" for v in row) for row in data))
    tr, te = T.UCIHousing(str(p), mode="train"), T.UCIHousing(str(p), mode="test")
    assert len(tr) == 16 and len(te) == 4 and tr[0][0].shape == (13,) and tr[0][1].shape == (1,)
    ptb = str(tmp_path / "simple-examples.tgz")
    with tarfile.open(ptb, "w:gz") as tf:
        for name, text in (("./simple-examples/data/ptb.train.txt", "the cat sat on the mat\nthe dog sat on the log\n" * 3), ("./simple-examples/data/ptb.valid.txt", "the cat sat\n")):
            b = text.encode()
            info = tarfile.TarInfo(name)
            info.size = len(b)
            tf.addfile(info, io.BytesIO(b))
    ng = T.Imikolov(ptb, data_type="NGRAM", window_size=3, mode="train", min_word_freq=1)
    assert len(ng) > 0 and len(ng[0]) == 3
    sq = T.Imikolov(ptb, data_type="SEQ", mode="test", min_word_freq=1)
    src, trg = sq[0]
    assert len(src) == len(trg) == 4
    trans = rng.rand(5, 5).astype("float32")
    pot = rng.rand(2, 6, 5).astype("float32")
    dec = T.ViterbiDecoder(t(trans), include_bos_eos_tag=False)
    scores, path = dec(t(pot), t(np.array([6, 4])))
    s2, p2 = T.viterbi_decode(t(pot), t(trans), t(np.array([6, 4])), include_bos_eos_tag=False)
    close(scores, s2)
    assert path.numpy().tolist() == p2.numpy().tolist() and path.shape == [2, 6]
    # brute force check of the best path for the first (full-length) sequence
    import itertools

    best = max(itertools.product(range(5), repeat=6), key=lambda s: sum(pot[0, i, s[i]] for i in range(6)) + sum(trans[s[i], s[i + 1]] for i in range(5)))
    assert path.numpy()[0].tolist() == list(best)


def test_callbacks_device_version_misc(tmp_path):
    C = paddle.callbacks
    cb = C.ReduceLROnPlateau(monitor="loss", factor=0.5, patience=1, verbose=0, min_delta=0.0, cooldown=0)

    class M:
        pass

    net = paddle.nn.Linear(2, 2)
    model = paddle.Model(net)
    opt = paddle.optimizer.SGD(1.0, parameters=net.parameters())
    model.prepare(opt, paddle.nn.MSELoss())
    cb.set_model(model)
    cb.on_train_begin()
    for e, l in enumerate([1.0, 1.0, 1.0, 1.0]):
        cb.on_eval_end({"loss": [l]}) if hasattr(cb, "on_eval_end") else None
        cb.on_epoch_end(e, {"loss": [l]})
    assert opt.get_lr() < 1.0
    assert issubclass(C.ProgBarLogger, C.Callback) and C.VisualDL(str(tmp_path / "vdl")) is not None and hasattr(C, "WandbCallback")
    cl = C.CallbackList([C.Callback()]) if hasattr(C, "CallbackList") else None
    assert cl is None or len(cl.callbacks) == 1
    D = paddle.device
    assert D.get_all_device_type() and isinstance(D.get_available_device(), list) and D.get_all_custom_device_type() == [] and D.get_available_custom_device() == []
    assert D.is_compiled_with_ipu() is False and D.is_compiled_with_npu() is False and D.is_compiled_with_mlu() is False
    assert D.get_cudnn_version() is None or isinstance(D.get_cudnn_version(), int)
    with pytest.raises(RuntimeError):
        D.XPUPlace(0)
    DC = paddle.device.cuda
    assert DC.device_count() >= 0 and DC.memory_allocated() >= 0 and DC.max_memory_allocated() >= 0 and DC.memory_reserved() >= 0 and DC.max_memory_reserved() >= 0
    DC.empty_cache()
    DC.reset_max_memory_allocated()
    DC.reset_max_memory_reserved()
    v = paddle.version
    assert v.full_version == paddle.__version__ and v.b200_version and v.cuda() == v.cuda_version and isinstance(v.cudnn(), str) and v.istaged in (True, False)
    v.show()
    assert os.path.isdir(paddle.sysconfig.get_include()) and isinstance(paddle.sysconfig.get_lib(), str)
    assert issubclass(paddle.regularizer.L2Decay, paddle.regularizer.WeightDecayRegularizer)
    from paddle_b200.distributed import fleet

    assert fleet.is_worker() and fleet.server_index() == -1 and fleet.server_endpoints() == []
    assert fleet.PaddleCloudRoleMaker(is_collective=True)._is_collective and fleet.UserDefinedRoleMaker is not None and fleet.Fleet().util is not None
    hub_dir = tmp_path / "hubrepo"
    hub_dir.mkdir()
    (hub_dir / "hubconf.py").write_text("dependencies = []\ndef tiny(n=2):\n    '''doc of tiny'''\n    import paddle_b200 as paddle\n    return paddle.nn.Linear(n, n)\n")
    assert paddle.hub.list(str(hub_dir), source="local") == ["tiny"] and "doc of tiny" in paddle.hub.help(str(hub_dir), "tiny", source="local")
    assert paddle.hub.load(str(hub_dir), "tiny", source="local", n=3).weight.shape == [3, 3]
