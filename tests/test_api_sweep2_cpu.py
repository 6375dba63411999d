"""Third API sweep: vision transforms / detection ops, geometric, sparse, incubate fused functional ops (vs plain compositions)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

import paddle_b200 as paddle

rng = np.random.RandomState(21)


def t(a):
    return paddle.to_tensor(np.asarray(a))


def tt(a):
    return torch.as_tensor(np.asarray(a))


def close(a, b, tol=1e-5):
    a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    b = b.detach().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


def test_vision_transforms_functional_and_random():
    T = paddle.vision.transforms
    img = (rng.rand(12, 16, 3) * 255).astype("uint8")
    assert (T.hflip(img) == img[:, ::-1]).all() and (T.vflip(img) == img[::-1]).all()
    assert (T.crop(img, 2, 3, 5, 6) == img[2:7, 3:9]).all()
    assert T.center_crop(img, 8).shape == (8, 8, 3) and T.resize(img, (6, 8)).shape == (6, 8, 3)
    g = T.to_grayscale(img)
    assert g.shape[:2] == (12, 16) and abs(float(np.asarray(g).reshape(12, 16, -1)[..., 0].mean()) - float((img @ np.array([0.299, 0.587, 0.114])).mean())) < 1.5
    assert np.allclose(np.asarray(T.adjust_brightness(img, 1.0)), img, atol=1) and np.asarray(T.adjust_brightness(img, 0.0)).max() == 0
    assert np.allclose(np.asarray(T.adjust_contrast(img, 1.0)), img, atol=1) and np.allclose(np.asarray(T.adjust_saturation(img, 1.0)), img, atol=1)
    assert np.allclose(np.asarray(T.adjust_hue(img, 0.0)), img, atol=2)
    assert np.asarray(T.adjust_hue(img, 0.3)).shape == img.shape
    assert (np.asarray(T.rotate(img, 0)) == img).all() and np.asarray(T.rotate(img, 90, expand=True)).shape == (16, 12, 3)
    assert (np.asarray(T.affine(img, 0, (0, 0), 1.0, 0)) == img).all()
    shifted = np.asarray(T.affine(img, 0, (2, 0), 1.0, 0))
    assert (shifted[:, 2:] == img[:, :-2]).all()
    pts = [[0, 0], [15, 0], [15, 11], [0, 11]]
    assert (np.asarray(T.perspective(img, pts, pts)) == img).all()
    e = T.erase(paddle.to_tensor(img.transpose(2, 0, 1).astype("float32")), 1, 2, 3, 4, 0.0)
    assert float(e[:, 1:4, 2:6].abs().sum()) == 0
    assert T.Transpose()(img).shape == (3, 12, 16)
    np.random.seed(0)
    import random

    random.seed(0)
    for tr in (T.RandomAffine(10, translate=(0.1, 0.1), scale=(0.9, 1.1), shear=5), T.RandomPerspective(1.0), T.RandomRotation(15), T.RandomVerticalFlip(1.0),
               T.BrightnessTransform(0.3), T.ContrastTransform(0.3), T.SaturationTransform(0.3), T.HueTransform(0.2)):
        out = np.asarray(tr(img))
        assert out.shape == img.shape and isinstance(tr, T.BaseTransform)
    assert np.asarray(T.RandomResizedCrop(8)(img)).shape == (8, 8, 3)
    er = T.RandomErasing(1.0, value=0)(paddle.to_tensor(img.transpose(2, 0, 1).astype("float32") + 1))
    assert float((er == 0).sum()) > 0


def test_detection_ops():
    V = paddle.vision.ops
    x = rng.randn(1, 4, 8, 8).astype("float32")
    boxes = np.array([[0, 0, 4, 4], [2, 2, 7, 7]], "float32")
    import torchvision.ops as tvo

    tb = [tt(boxes)]
    close(V.roi_pool(t(x), t(boxes), t(np.array([2], "int32")), 2), tvo.roi_pool(tt(x), tb, 2), 1e-5)
    close(V.RoIPool(2)(t(x), t(boxes), t(np.array([2], "int32"))), tvo.roi_pool(tt(x), tb, 2), 1e-5)
    close(V.RoIAlign(2)(t(x), t(boxes), t(np.array([2], "int32"))).shape, [2, 4, 2, 2], 0)
    x8 = rng.randn(1, 8, 8, 8).astype("float32")
    close(V.psroi_pool(t(x8), t(boxes), t(np.array([2], "int32")), 2), tvo.ps_roi_pool(tt(x8), tb, 2), 1e-5)
    assert V.PSRoIPool(2)(t(x8), t(boxes), t(np.array([2], "int32"))).shape == [2, 2, 2, 2]
    # deform conv with zero offsets == plain conv
    w = rng.randn(5, 4, 3, 3).astype("float32")
    off = np.zeros((1, 18, 8, 8), "float32")
    close(V.deform_conv2d(t(x), t(off), t(w), padding=1), TF.conv2d(tt(x), tt(w), padding=1), 1e-4)
    # box coder round trip
    prior = np.array([[0, 0, 10, 10], [5, 5, 20, 25]], "float32")
    var = np.array([0.1, 0.1, 0.2, 0.2], "float32")
    target = np.array([[1, 1, 9, 11], [6, 4, 18, 22]], "float32")
    enc = V.box_coder(t(prior), t(var), t(target), "encode_center_size", box_normalized=False)
    assert enc.shape == [2, 2, 4]
    dec = V.box_coder(t(prior), t(var), enc, "decode_center_size", box_normalized=False, axis=0)
    close(np.stack([dec.numpy()[0, 0], dec.numpy()[1, 1]]), target, 1e-3)
    pb, pv = V.prior_box(t(rng.randn(1, 3, 4, 4).astype("float32")), t(rng.randn(1, 3, 32, 32).astype("float32")), min_sizes=[8.0], aspect_ratios=[1.0, 2.0], flip=True)
    assert pb.shape == [4, 4, 3, 4] and pv.shape == pb.shape
    yb, ys = V.yolo_box(t(rng.randn(1, 2 * 7, 4, 4).astype("float32")), t(np.array([[64, 64]], "int32")), [10, 13, 16, 30], 2, 0.01, 16)
    assert yb.shape == [1, 32, 4] and ys.shape == [1, 32, 2]
    yl = V.yolo_loss(t(rng.randn(1, 2 * 7, 4, 4).astype("float32")), t(np.array([[[0.5, 0.5, 0.3, 0.3]]], "float32")), t(np.array([[1]], "int32")),
                     [10, 13, 16, 30], [0, 1], 2, 0.7, 16)
    assert yl.shape == [1] and np.isfinite(yl.numpy()).all()
    bb = np.array([[[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30]]], "float32")
    sc = np.array([[[0.1, 0.1, 0.1], [0.9, 0.8, 0.7]]], "float32")
    out, num = V.matrix_nms(t(bb), t(sc), 0.05, 0.3, 10, 10)[:2]
    assert out.shape[1] == 6 and int(num.numpy().sum()) == out.shape[0] >= 2
    rois = np.array([[0, 0, 10, 10], [0, 0, 200, 200], [0, 0, 60, 60]], "float32")
    multi, restore = V.distribute_fpn_proposals(t(rois), 2, 5, 4, 224)[:2]
    assert sum(m.shape[0] for m in multi) == 3 and sorted(restore.numpy().reshape(-1).tolist()) == [0, 1, 2]
    anchors = rng.rand(4, 4, 3, 4).astype("float32") * 16
    anchors[..., 2:] += anchors[..., :2] + 4
    r, s = V.generate_proposals(t(rng.rand(1, 3, 4, 4).astype("float32")), t(rng.randn(1, 12, 4, 4).astype("float32") * 0.1), t(np.array([[64, 64]], "float32")),
                                t(anchors), t(np.ones((4, 4, 3, 4), "float32")), pre_nms_top_n=20, post_nms_top_n=5)[:2]
    assert r.shape[1] == 4 and r.shape[0] <= 5 and s.shape[0] == r.shape[0]


def test_image_file_ops(tmp_path):
    V = paddle.vision.ops
    from PIL import Image

    img = (rng.rand(8, 10, 3) * 255).astype("uint8")
    p = str(tmp_path / "a.jpg")
    Image.fromarray(img).save(p, quality=95)
    raw = V.read_file(p)
    assert raw.dtype == paddle.uint8 and raw.ndim == 1
    dec = V.decode_jpeg(raw)
    assert dec.shape == [3, 8, 10] and abs(float(dec.astype("float32").mean()) - img.mean()) < 10


def test_geometric_leftovers():
    G = paddle.geometric
    data = np.array([[1., 2.], [3., 1.], [0., 5.], [4., 4.]], "float32")
    ids = np.array([0, 0, 1, 1])
    close(G.segment_max(t(data), t(ids)), [[3, 2], [4, 5]])
    close(G.segment_min(t(data), t(ids)), [[1, 1], [0, 4]])
    close(G.segment_mean(t(data), t(ids)), [[2, 1.5], [2, 4.5]])
    x = np.array([[1., 2.], [3., 4.], [5., 6.]], "float32")
    e = np.array([[10., 10.], [20., 20.], [30., 30.]], "float32")
    src, dst = np.array([0, 1, 2]), np.array([1, 1, 0])
    close(G.send_ue_recv(t(x), t(e), t(src), t(dst), "add", "sum"), [[35, 36], [34, 36], [0, 0]])
    close(G.send_ue_recv(t(x), t(e), t(src), t(dst), "mul", "max"), [[150, 180], [60, 80], [0, 0]])
    close(G.send_uv(t(x), t(x), t(src), t(dst), "add"), x[src] + x[dst])
    # CSC graph: node i's in-neighbours are row[colptr[i]:colptr[i+1]]
    row = np.array([1, 2, 0, 2, 0, 1, 3, 0])
    colptr = np.array([0, 2, 4, 7, 8])
    nb, cnt = G.sample_neighbors(t(row), t(colptr), t(np.array([0, 2])), sample_size=-1)
    assert cnt.numpy().tolist() == [2, 3] and nb.numpy().tolist() == [1, 2, 0, 1, 3]
    nb2, cnt2 = G.sample_neighbors(t(row), t(colptr), t(np.array([2])), sample_size=2)
    assert cnt2.numpy().tolist() == [2] and set(nb2.numpy().tolist()) <= {0, 1, 3}
    nb3, cnt3 = G.weighted_sample_neighbors(t(row), t(colptr), t(np.ones(8, "float32")), t(np.array([2])), sample_size=2)
    assert cnt3.numpy().tolist() == [2]
    rs, rd, nodes = G.reindex_graph(t(np.array([10, 20])), t(np.array([30, 20, 10, 40])), t(np.array([2, 2])))
    assert nodes.numpy().tolist() == [10, 20, 30, 40] and rs.numpy().tolist() == [2, 1, 0, 3] and rd.numpy().tolist() == [0, 0, 1, 1]
    rs2, rd2, nodes2 = G.reindex_heter_graph(t(np.array([10, 20])), [t(np.array([30, 20])), t(np.array([10, 40]))], [t(np.array([1, 1])), t(np.array([1, 1]))])
    assert nodes2.numpy().tolist() == [10, 20, 30, 40] and rs2.numpy().tolist() == [2, 1, 0, 3]


def test_sparse_leftovers():
    S = paddle.sparse
    crows, cols, vals = np.array([0, 2, 3, 5]), np.array([1, 3, 2, 0, 1]), np.array([1., 2., 3., 4., 5.], "float32")
    csr = S.sparse_csr_tensor(t(crows), t(cols), t(vals), [3, 4])
    dense = np.zeros((3, 4), "float32")
    dense[[0, 0, 1, 2, 2], cols] = vals
    close(csr.to_dense(), dense)
    coo = S.sparse_coo_tensor(np.array([[0, 0, 1], [1, 1, 2]]), np.array([1., 2., 3.], "float32"), [2, 3])
    co = S.coalesce(coo)
    assert co.nnz() == 2 and float(co.to_dense()[0, 1]) == 3.0
    assert S.is_same_shape(coo, co)
    m = S.mask_as(t(np.arange(6, dtype="float32").reshape(2, 3)), co)
    close(m.to_dense(), [[0, 1, 0], [0, 0, 5]])
    a, b = rng.randn(3, 5).astype("float32"), rng.randn(5, 4).astype("float32")
    mm = S.masked_matmul(t(a), t(b), csr)
    close(mm.to_dense(), (a @ b) * (dense != 0), 1e-5)
    assert S.convert_dtype is not None
    u, s, v = S.pca_lowrank(S.sparse_coo_tensor(np.array([[0, 1, 2, 3], [0, 1, 2, 0]]), np.array([1., 2., 3., 4.], "float32"), [4, 3]), q=2)
    assert s.shape == [2]


def _ln(x, g=None, b=None, eps=1e-5):
    return TF.layer_norm(x, x.shape[-1:], g, b, eps)


def test_incubate_fused_functional_vs_composition():
    I = paddle.incubate.nn.functional
    B_, S_, H_, NH = 2, 5, 16, 4
    x = rng.randn(B_, S_, H_).astype("float32")
    w1, b1 = rng.randn(H_, 32).astype("float32") * 0.2, rng.randn(32).astype("float32") * 0.1
    w2, b2 = rng.randn(32, H_).astype("float32") * 0.2, rng.randn(H_).astype("float32") * 0.1
    g, be = rng.rand(H_).astype("float32") + 0.5, rng.randn(H_).astype("float32") * 0.1
    close(I.fused_matmul_bias(t(x), t(w1), t(b1)), tt(x) @ tt(w1) + tt(b1), 1e-4)
    close(I.fused_linear(t(x), t(w1), t(b1)), tt(x) @ tt(w1) + tt(b1), 1e-4)
    close(I.fused_linear(t(x), t(w1.T.copy()), t(b1), transpose_weight=True), tt(x) @ tt(w1) + tt(b1), 1e-4)
    close(I.fused_linear_activation(t(x), t(w1), t(b1), activation="gelu"), TF.gelu(tt(x) @ tt(w1) + tt(b1)), 1e-4)
    close(I.fused_linear_activation(t(x), t(w1), t(b1), activation="relu"), TF.relu(tt(x) @ tt(w1) + tt(b1)), 1e-4)
    close(I.fused_dropout_add(t(x), t(x * 2), p=0.3, training=False), x * 3, 1e-5)
    paddle.seed(0)
    d = I.fused_dropout_add(t(x), t(np.zeros_like(x)), p=0.5, training=True).numpy()
    assert set(np.unique(np.round(d / x, 3))) <= {0.0, 2.0}
    res = rng.randn(B_, S_, H_).astype("float32")
    close(I.fused_bias_dropout_residual_layer_norm(t(x), t(res), t(be), t(g), t(be), dropout_rate=0.0), _ln(tt(x) + tt(be) + tt(res), tt(g), tt(be)), 1e-4)
    # feed-forward, post-LN and pre-LN
    ref_post = _ln(tt(x) + (TF.relu(tt(x) @ tt(w1) + tt(b1)) @ tt(w2) + tt(b2)), tt(g), tt(be))
    close(I.fused_feedforward(t(x), t(w1), t(w2), t(b1), t(b2), ln2_scale=t(g), ln2_bias=t(be), dropout1_rate=0.0, dropout2_rate=0.0), ref_post, 1e-4)
    ref_pre = tt(x) + (TF.gelu(_ln(tt(x), tt(g), tt(be)) @ tt(w1) + tt(b1)) @ tt(w2) + tt(b2))
    close(I.fused_feedforward(t(x), t(w1), t(w2), t(b1), t(b2), ln1_scale=t(g), ln1_bias=t(be), dropout1_rate=0.0, dropout2_rate=0.0, activation="gelu",
                              pre_layer_norm=True), ref_pre, 1e-4)
    # multi-head attention (qkv_weight [3, nh, hd, H])
    hd = H_ // NH
    qkvw = rng.randn(3, NH, hd, H_).astype("float32") * 0.2
    qkvb = rng.randn(3, NH, hd).astype("float32") * 0.1
    lw, lb = rng.randn(H_, H_).astype("float32") * 0.2, rng.randn(H_).astype("float32") * 0.1
    qkv = torch.einsum("bsh,tndh->tbnsd", tt(x), tt(qkvw)) + tt(qkvb)[:, None, :, None, :]
    att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) / hd ** 0.5, -1) @ qkv[2]
    o = att.transpose(1, 2).reshape(B_, S_, H_) @ tt(lw) + tt(lb)
    close(I.fused_multi_head_attention(t(x), t(qkvw), t(lw), qkv_bias=t(qkvb), linear_bias=t(lb), ln_scale=t(g), ln_bias=t(be), dropout_rate=0.0,
                                       attn_dropout_rate=0.0), _ln(tt(x) + o, tt(g), tt(be)), 1e-4)
    # norms with residual / bias
    rms = lambda v, w_: v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-6) * w_
    out = I.fused_rms_norm(t(x), t(g), None, 1e-6, 2, bias=t(be), residual=t(res))
    close(out[0], rms(tt(x) + tt(be) + tt(res), tt(g)), 1e-4)
    close(out[1], tt(x) + tt(be) + tt(res), 1e-5)
    out = I.fused_layer_norm(t(x), t(g), t(be), 1e-5, begin_norm_axis=2, bias=t(be), residual=t(res))
    close(out[0], _ln(tt(x) + tt(be) + tt(res), tt(g), tt(be)), 1e-4)
    close(I.fused_bias_act(t(x), t(be), act_method="gelu"), TF.gelu(tt(x) + tt(be)), 1e-4)
    sw = I.fused_bias_act(t(x), None, act_method="swiglu")
    close(sw, TF.silu(tt(x)[..., :H_ // 2]) * tt(x)[..., H_ // 2:], 1e-4)
    # rotary. Reference naming: use_neox_rotary_style=True rotates adjacent pairs, False rotates front / back halves.
    q = rng.randn(B_, S_, NH, hd).astype("float32")
    pos = np.arange(S_)[:, None] / (10000.0 ** (np.arange(0, hd, 2) / hd))[None]
    sin, cos = np.sin(pos).astype("float32"), np.cos(pos).astype("float32")
    s_half, c_half = np.concatenate([sin, sin], -1)[None, :, None], np.concatenate([cos, cos], -1)[None, :, None]
    rot_half = np.concatenate([-q[..., hd // 2:], q[..., :hd // 2]], -1)
    oq, ok, ov = I.fused_rotary_position_embedding(t(q), t(q), None, sin=t(s_half), cos=t(c_half), use_neox_rotary_style=False)
    close(oq, q * c_half + rot_half * s_half, 1e-5)
    close(ok, oq, 1e-6)
    s_pair, c_pair = np.repeat(sin, 2, -1)[None, :, None], np.repeat(cos, 2, -1)[None, :, None]
    rot_pair = np.stack([-q[..., 1::2], q[..., 0::2]], -1).reshape(q.shape)
    oq2 = I.fused_rotary_position_embedding(t(q), sin=t(s_pair), cos=t(c_pair))[0]
    close(oq2, q * c_pair + rot_pair * s_pair, 1e-5)
    close(I.fused_rotary_position_embedding(t(q))[0], q * c_pair + rot_pair * s_pair, 1e-4)   # tables built internally
    # attention variants
    qh = rng.randn(B_, NH, S_, hd).astype("float32")
    ref = TF.scaled_dot_product_attention(tt(qh), tt(qh), tt(qh), is_causal=True)
    close(I.fused_dot_product_attention(t(qh.transpose(0, 2, 1, 3).copy()), t(qh.transpose(0, 2, 1, 3).copy()), t(qh.transpose(0, 2, 1, 3).copy()), is_causal=True),
          ref.transpose(1, 2), 1e-4)
    lens = np.array([5, 3], "int32")
    vo = I.variable_length_memory_efficient_attention(t(qh), t(qh), t(qh), t(lens), t(lens), causal=True)
    close(vo[0], ref[0], 1e-4)
    close(vo[1, :, :3], TF.scaled_dot_product_attention(tt(qh[1:, :, :3]), tt(qh[1:, :, :3]), tt(qh[1:, :, :3]), is_causal=True)[0], 1e-4)
    me, md = I.blha_get_max_len(t(np.array([3, 7], "int32")), t(np.array([0, 9], "int32")), 2)
    assert int(me) == 7 and int(md) == 9
    # MoE: dense reference over the top-2 experts
    E_, F_ = 4, 24
    xm = rng.randn(6, H_).astype("float32")
    gw = rng.randn(H_, E_).astype("float32")
    f1, f2 = rng.randn(E_, H_, 2 * F_).astype("float32") * 0.2, rng.randn(E_, F_, H_).astype("float32") * 0.2
    probs = torch.softmax(tt(xm) @ tt(gw), -1)
    tv, ti = probs.topk(2, -1)
    tv = tv / tv.sum(-1, keepdim=True)
    ref = torch.zeros(6, H_)
    for n in range(6):
        for k in range(2):
            e = int(ti[n, k])
            h = tt(xm[n]) @ tt(f1[e])
            ref[n] += tv[n, k] * ((TF.silu(h[:F_]) * h[F_:]) @ tt(f2[e]))
    close(I.fused_moe(t(xm), t(gw), t(f1), ffn2_weight=t(f2), moe_topk=2), ref, 1e-3)
