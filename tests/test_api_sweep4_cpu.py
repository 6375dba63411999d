"""Fifth sweep: text datasets from synthetic local archives, audio functional vs closed forms, amp tensor dump / compare_accuracy,
cost_model.profile_measure, forward_grad, ctr_metric_bundle, incubate layers / graph ops, inference helpers."""
import gzip
import io
import os
import tarfile
import zipfile

import numpy as np
import pytest
import torch

import paddle_b200 as paddle

rng = np.random.RandomState(41)


def t(a):
    return paddle.to_tensor(np.asarray(a))


def close(a, b, tol=1e-5):
    a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    b = b.detach().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


def _add(tf, name, data):
    info = tarfile.TarInfo(name)
    info.size = len(data)
    tf.addfile(info, io.BytesIO(data))


def test_movielens(tmp_path):
    p = str(tmp_path / "ml-1m.zip")
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("ml-1m/movies.dat", "1::Toy Story (1995)::Animation|Comedy\n2::Heat (1995)::Action\n")
        z.writestr("ml-1m/users.dat", "1::F::1::10::48067\n2::M::56::16::70072\n")
        z.writestr("ml-1m/ratings.dat", "".join(f"{1 + i % 2}::{1 + (i // 2) % 2}::{1 + i % 5}::97830{i}\n" for i in range(40)))
    tr, te = paddle.text.Movielens(p, mode="train", test_ratio=0.25), paddle.text.Movielens(p, mode="test", test_ratio=0.25)
    assert len(tr) + len(te) == 40 and len(te) > 0
    uid, gender, age, job, mid, cats, title, rating = tr[0]
    assert uid.shape == (1,) and gender[0] in (0, 1) and age[0] in (0, 6) and len(cats) in (1, 2) and len(title) in (1, 2) and -3.0 <= rating[0] <= 5.0
    assert tr.categories_dict == {"Animation": 0, "Comedy": 1, "Action": 2} and "toy" in tr.movie_title_dict


def test_wmt14_wmt16(tmp_path):
    p = str(tmp_path / "wmt14.tgz")
    with tarfile.open(p, "w:gz") as tf:
        _add(tf, "wmt14/src.dict", b"<s>\n<e>\n<unk>\nhello\nworld\n")
        _add(tf, "wmt14/trg.dict", b"<s>\n<e>\n<unk>\nbonjour\nmonde\n")
        _add(tf, "wmt14/train/train", b"hello world\tbonjour monde\nhello there\tbonjour\n")
        _add(tf, "wmt14/test/test", b"world\tmonde\n")
    ds = paddle.text.WMT14(p, mode="train")
    src, trg, nxt = ds[1]
    assert src.tolist() == [0, 3, 2, 1] and trg.tolist() == [0, 3] and nxt.tolist() == [3, 1] and len(ds) == 2
    assert len(paddle.text.WMT14(p, mode="test")) == 1 and ds.get_dict(reverse=True)[0][3] == "hello"
    p16 = str(tmp_path / "wmt16.tar")
    with tarfile.open(p16, "w") as tf:
        _add(tf, "wmt16/train", b"a cat sat\teine katze sass\na dog\tein hund\n")
        _add(tf, "wmt16/val", b"a cat\teine katze\n")
        _add(tf, "wmt16/test", b"a bird\tein vogel\n")
    d = paddle.text.WMT16(p16, mode="val", lang="en")
    src, trg, nxt = d[0]
    assert d.src_dict["a"] == 3 and src.tolist()[0] == 0 and src.tolist()[-1] == 1 and len(trg) == len(nxt) == 3
    t16 = paddle.text.WMT16(p16, mode="test", lang="de", src_dict_size=5)
    assert len(t16.src_dict) == 5 and t16[0][0].tolist()[1] == t16.src_dict["ein"]
    assert t16.get_dict("de") is t16.src_dict and 2 in t16.get_dict("en", reverse=True)


def test_conll05(tmp_path):
    words = "The\ncat\nchased\nmice\n\nDogs\nbark\n\n"
    props = "-\t(A0*\n-\t*)\nchase\t(V*)\n-\t(A1*)\n\n-\t(A0*)\nbark\t(V*)\n\n"
    p = str(tmp_path / "conll.tar.gz")
    with tarfile.open(p, "w:gz") as tf:
        _add(tf, "conll05st-release/test.wsj/words/test.wsj.words.gz", gzip.compress(words.encode()))
        _add(tf, "conll05st-release/test.wsj/props/test.wsj.props.gz", gzip.compress(props.encode()))
    wd, vd, td = tmp_path / "w.dict", tmp_path / "v.dict", tmp_path / "t.dict"
    wd.write_text("<unk>\nThe\ncat\nchased\nmice\nbos\neos\n")
    vd.write_text("chase\nbark\n")
    td.write_text("B-A0\nI-A0\nB-A1\nI-A1\nB-V\nI-V\nO\n")
    ds = paddle.text.Conll05st(p, str(wd), str(vd), str(td))
    assert len(ds) == 2
    wid, n2, n1, c0, p1, p2, pred, mark, lab = ds[0]
    _, _, ld = ds.get_dict()
    assert wid.tolist() == [1, 2, 3, 4] and c0.tolist() == [3] * 4 and n2.tolist() == [1] * 4 and p2.tolist() == [6] * 4 and mark.tolist() == [1, 1, 1, 1]
    assert lab.tolist() == [ld["B-A0"], ld["I-A0"], ld["B-V"], ld["B-A1"]] and pred.tolist() == [0] * 4
    assert ds[1][0].tolist() == [0, 0] and ds[1][6].tolist() == [1, 1]


def test_audio_functional_closed_forms():
    AF = paddle.audio.functional
    import scipy.signal as ss

    close(AF.hz_to_mel(440.0, htk=True), 2595 * np.log10(1 + 440 / 700), 1e-4)
    close(AF.mel_to_hz(AF.hz_to_mel(t(np.array([100.0, 1000.0, 4000.0], "float32")))), [100.0, 1000.0, 4000.0], 1e-3)
    close(AF.fft_frequencies(16000, 8), np.fft.rfftfreq(8, 1 / 16000), 1e-4)
    mf = AF.mel_frequencies(5, 0.0, 8000.0).numpy()
    assert mf[0] == 0 and abs(mf[-1] - 8000) < 1 and (np.diff(mf) > 0).all()
    fb = AF.compute_fbank_matrix(16000, 64, n_mels=8).numpy()
    assert fb.shape == (8, 33) and (fb >= 0).all() and (fb.sum(1) > 0).all()
    dct = AF.create_dct(4, 8).numpy()
    close(dct.T @ dct, np.eye(4), 1e-5)
    db = AF.power_to_db(t(np.array([1.0, 10.0, 1e-12], "float32")), top_db=80.0).numpy()
    close(db, [0.0, 10.0, -70.0], 1e-4)
    for name, args in (("taylor", ()), ("taylor", (5, 40.0)), ("hamming", ()), ("blackman", ()), ("nuttall", ()), ("bohman", ())):
        ours = AF.get_window((name, *args) if args else name, 32, fftbins=False).numpy()
        ref = ss.get_window((name, *args) if args else name, 32, fftbins=False)
        close(ours, ref, 1e-6)
    feat = paddle.audio.features.LogMelSpectrogram(sr=8000, n_fft=64, n_mels=8)(t(rng.randn(1, 800).astype("float32")))
    assert feat.shape[1] == 8 and np.isfinite(feat.numpy()).all()
    assert paddle.audio.backends.list_available_backends()


def test_tensor_dump_and_compare_accuracy(tmp_path):
    D = paddle.amp.debugging
    x, w = rng.randn(4, 8).astype("float32"), rng.randn(8, 3).astype("float32")

    def run(out_dir, scale):
        D.enable_tensor_checker(D.TensorCheckerConfig(True, D.DebugMode.DUMP_ALL, output_dir=str(out_dir)))
        try:
            y = paddle.matmul(t(x), t(w * scale))
            paddle.nn.functional.relu(y)
        finally:
            D.disable_tensor_checker()

    run(tmp_path / "a", 1.0)
    run(tmp_path / "b", 1.0)
    run(tmp_path / "c", 3.0)
    assert "op=matmul" in open(tmp_path / "a" / "worker_0.log").read()
    assert D.compare_accuracy(str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "same.csv")) == []
    bad = D.compare_accuracy(str(tmp_path / "a"), str(tmp_path / "c"), str(tmp_path / "diff.csv"))
    assert {r["op"] for r in bad} >= {"matmul"} and os.path.getsize(tmp_path / "diff.csv") > 100
    D.enable_tensor_checker(D.TensorCheckerConfig(True, D.DebugMode.CHECK_ALL_AND_ABORT, output_dir=str(tmp_path / "d")))
    try:
        with pytest.raises(RuntimeError):
            paddle.log(t(np.array([-1.0], "float32")))
    finally:
        D.disable_tensor_checker()


def test_cost_model_forward_grad_ctr_metric():
    cm = paddle.cost_model.CostModel()
    r = cm.profile_measure(lambda: paddle.matmul(t(rng.randn(64, 64)), t(rng.randn(64, 64))), device="cpu", repeat=3)
    assert r["time"] > 0
    x = t(np.array([0.5, 1.5, 2.0]))
    x.stop_gradient = False
    y = paddle.sin(x) * x
    jv = paddle.incubate.autograd.forward_grad(y, x, t(np.array([1.0, 0.0, 2.0])))
    xn = x.numpy()
    close(jv, (np.cos(xn) * xn + np.sin(xn)) * np.array([1.0, 0.0, 2.0]), 1e-6)
    pred, lab = np.array([[0.2], [0.9], [0.6]], "float32"), np.array([[0], [1], [1]], "float32")
    sq, ab, prob, q, pos, total = paddle.static.ctr_metric_bundle(t(pred), t(lab))
    close(sq, [((pred - lab) ** 2).sum()])
    close(ab, [np.abs(pred - lab).sum()])
    assert float(pos) == 2 and float(total) == 3 and abs(float(prob) - 1.7) < 1e-6


def test_incubate_layers_and_graph_ops():
    IN = paddle.incubate.nn
    x = t(rng.randn(2, 5, 16).astype("float32"))
    for layer in (IN.FusedLinear(16, 8), IN.FusedFeedForward(16, 32, dropout_rate=0.0), IN.FusedMultiHeadAttention(16, 4, dropout_rate=0.0, attn_dropout_rate=0.0),
                  IN.FusedTransformerEncoderLayer(16, 4, 32, dropout_rate=0.0), IN.FusedBiasDropoutResidualLayerNorm(16, dropout_rate=0.0)):
        layer.eval()
        y = layer(x, x) if isinstance(layer, IN.FusedBiasDropoutResidualLayerNorm) else layer(x)
        assert y.shape[:2] == [2, 5] and np.isfinite(y.numpy()).all()
    da = IN.FusedDropoutAdd(0.5)
    da.eval()
    close(da(x, x), 2 * x.numpy())
    mt = IN.FusedMultiTransformer(16, 4, 32, num_layers=2)
    mt.eval()
    assert mt(x).shape == [2, 5, 16]
    q = t(rng.randn(2, 5, 4, 8).astype("float32"))
    ref = torch.nn.functional.scaled_dot_product_attention(*(torch.as_tensor(q.numpy()).transpose(1, 2),) * 3).transpose(1, 2)
    close(IN.memory_efficient_attention(q, q, q), ref, 1e-4)
    inc = paddle.incubate
    s = rng.randn(2, 2, 4, 4).astype("float32")
    mask = np.where(rng.rand(2, 1, 4, 4) > 0.5, 0.0, -1e4).astype("float32")
    close(inc.softmax_mask_fuse(t(s), t(mask)), torch.softmax(torch.as_tensor(s + mask), -1), 1e-5)
    tri = inc.softmax_mask_fuse_upper_triangle(t(s)).numpy()
    assert np.allclose(np.triu(tri[0, 0], 1), 0) and np.allclose(tri.sum(-1), 1, atol=1e-5)
    close(inc.identity_loss(t(s), "mean"), s.mean(), 1e-6)
    xg = np.array([[1., 2.], [3., 4.], [5., 6.]], "float32")
    close(inc.graph_send_recv(t(xg), t(np.array([0, 1, 2])), t(np.array([1, 1, 0])), "sum"), [[5, 6], [4, 6], [0, 0]])
    row, colptr = np.array([1, 2, 0, 2, 0, 1, 3, 0]), np.array([0, 2, 4, 7, 8])
    nb, cnt = inc.graph_sample_neighbors(t(row), t(colptr), t(np.array([0, 2])))
    assert cnt.numpy().tolist() == [2, 3]
    rs, rd, nodes = inc.graph_reindex(t(np.array([0, 2])), nb, cnt)
    assert nodes.numpy().tolist()[:2] == [0, 2] and len(rs.numpy()) == 5
    es, ed, sample_index, reindex_nodes = inc.graph_khop_sampler(t(row), t(colptr), t(np.array([0])), [2, 2])[:4]
    assert len(es.numpy()) == len(ed.numpy()) > 0
    lin = paddle.nn.Linear(16, 16)
    inc.asp.prune_model(lin)
    wv = lin.weight.numpy().reshape(-1, 4)
    assert ((wv != 0).sum(1) <= 2).all() and abs(inc.asp.calculate_density(lin.weight) - 0.5) < 1e-6
    inc.autotune.set_config({"kernel": {"enable": True}})


def test_inference_helpers():
    I = paddle.inference
    assert I.get_num_bytes_of_data_type(I.DataType.FLOAT32) == 4 and I.get_num_bytes_of_data_type(I.DataType.INT64) == 8
    assert isinstance(I.get_version(), str) and I.get_trt_compile_version() == (0, 0, 0) and I.get_trt_runtime_version() == (0, 0, 0)
    assert I.PrecisionType.Half is not None and I.PlaceType.GPU is not None
