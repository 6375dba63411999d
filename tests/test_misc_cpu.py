"""Inference predictor, quantization, nn.quant, profiler, incubate optimizers, dlpack, audio features, regularizer, callbacks.
Parity: test/legacy_test/test_inference_api.py, test/quantization/*, test_profiler.py, test_lookahead.py, test_dlpack.py."""
import json
import os

import numpy as np
import pytest

import paddle_b200 as paddle
from paddle_b200 import nn


class TinyNet(nn.Layer):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(6, 12), nn.Linear(12, 3)

    def forward(self, x):
        return self.fc2(paddle.nn.functional.relu(self.fc1(x)))


def test_inference_predictor_roundtrip(tmp_path):
    from paddle_b200 import inference

    paddle.seed(0)
    net = TinyNet()
    net.eval()
    paddle.jit.save(net, str(tmp_path / "m"), input_spec=[paddle.static.InputSpec([None, 6], "float32", "x")])
    cfg = inference.Config(str(tmp_path / "m.pdmodel"), str(tmp_path / "m.pdiparams"))
    cfg.disable_gpu() if hasattr(cfg, "disable_gpu") else None
    pred = inference.create_predictor(cfg)
    x = np.random.rand(4, 6).astype("float32")
    h = pred.get_input_handle(pred.get_input_names()[0])
    h.copy_from_cpu(x)
    pred.run()
    out = pred.get_output_handle(pred.get_output_names()[0]).copy_to_cpu()
    np.testing.assert_allclose(out, net(paddle.to_tensor(x)).numpy(), rtol=1e-5)
    outs = pred.run([x])
    np.testing.assert_allclose(outs[0].numpy(), out, rtol=1e-6)


def test_quantization_qat_ptq_and_weight_only():
    from paddle_b200.quantization import PTQ, QAT, AbsmaxObserver, FakeQuanterWithAbsMaxObserver, QuantConfig

    paddle.seed(1)
    net = TinyNet()
    x = paddle.to_tensor(np.random.rand(8, 6).astype("float32"))
    ref = net(x).numpy()
    q = QAT(QuantConfig(activation=FakeQuanterWithAbsMaxObserver(moving_rate=0.9), weight=FakeQuanterWithAbsMaxObserver(moving_rate=0.9)))
    qnet = q.quantize(net, inplace=False)
    y = qnet(x)
    y.sum().backward()                      # straight-through estimator keeps gradients flowing
    assert np.abs(y.numpy() - ref).max() < 0.1 and any(p.grad is not None for p in qnet.parameters())
    p = PTQ(QuantConfig(activation=AbsmaxObserver(), weight=AbsmaxObserver()))
    pnet = p.quantize(net, inplace=False)
    pnet(x)
    conv = p.convert(pnet, inplace=False)
    assert np.abs(conv(x).numpy() - ref).max() < 0.1
    w = paddle.to_tensor(np.random.randn(16, 8).astype("float32"))
    qw, sc = nn.quant.weight_quantize(w, algo="weight_only_int8")
    xin = paddle.to_tensor(np.random.randn(4, 16).astype("float32"))
    out = nn.quant.weight_only_linear(xin, qw, weight_scale=sc, weight_dtype="int8")
    assert np.abs(out.numpy() - xin.numpy() @ w.numpy()).max() < 0.15
    deq = nn.quant.weight_dequantize(qw, sc, out_dtype="float32")
    assert np.abs(deq.numpy().reshape(w.shape) - w.numpy()).max() < 0.05 or np.abs(deq.numpy().T.reshape(w.shape) - w.numpy()).max() < 0.05


def test_profiler_records_and_exports(tmp_path):
    from paddle_b200 import profiler

    net = TinyNet()
    x = paddle.ones([2, 6])
    prof = profiler.Profiler(targets=[profiler.ProfilerTarget.CPU], scheduler=profiler.make_scheduler(closed=0, ready=0, record=2, repeat=1),
                             on_trace_ready=profiler.export_chrome_tracing(str(tmp_path)))
    prof.start()
    for _ in range(3):
        with profiler.RecordEvent("fwd"):
            net(x)
        prof.step()
    prof.stop()
    prof.summary()
    files = [f for f in os.listdir(tmp_path) if f.endswith(".json")]
    assert files and "traceEvents" in json.load(open(tmp_path / files[0]))


def test_incubate_optimizers():
    from paddle_b200.incubate.optimizer import GradientMergeOptimizer, LookAhead, ModelAverage

    paddle.seed(2)
    p = paddle.create_parameter([3], "float32", default_initializer=nn.initializer.Constant(1.0))
    la = LookAhead(paddle.optimizer.SGD(0.1, parameters=[p]), alpha=0.5, k=2)
    for _ in range(4):
        (p * p).sum().backward()
        la.step()
        la.clear_grad()
    assert float(p.sum()) < 3.0
    q = paddle.create_parameter([2], "float32", default_initializer=nn.initializer.Constant(1.0))
    inner = paddle.optimizer.SGD(0.1, parameters=[q])
    ma = ModelAverage(0.15, parameters=[q], min_average_window=2, max_average_window=4)
    for _ in range(4):
        (q * q).sum().backward()
        inner.step()
        ma.step()
        inner.clear_grad()
    cur = q.numpy().copy()
    with ma.apply():
        assert not np.allclose(q.numpy(), cur)     # averaged weights swapped in
    np.testing.assert_allclose(q.numpy(), cur)
    r = paddle.create_parameter([2], "float32", default_initializer=nn.initializer.Constant(1.0))
    gm = GradientMergeOptimizer(paddle.optimizer.SGD(0.1, parameters=[r]), k_steps=2, avg=True)
    for i in range(2):
        (r * (i + 1.0)).sum().backward()
        gm.step()
        gm.clear_grad()
    np.testing.assert_allclose(r.numpy(), 1.0 - 0.1 * 1.5, rtol=1e-6)


def test_dlpack_hub_regularizer_callbacks(tmp_path):
    x = paddle.to_tensor(np.arange(6, dtype="float32").reshape(2, 3))
    y = paddle.utils.dlpack.from_dlpack(paddle.utils.dlpack.to_dlpack(x))
    np.testing.assert_array_equal(y.numpy(), x.numpy())
    (tmp_path / "hubconf.py").write_text("def tiny(n=2):\n    '''doc'''\n    import paddle_b200 as p\n    return p.nn.Linear(n, n)\n")
    assert "tiny" in paddle.hub.list(str(tmp_path), source="local")
    assert paddle.hub.load(str(tmp_path), "tiny", source="local", n=3).weight.shape == [3, 3]
    p = paddle.create_parameter([2], "float32", default_initializer=nn.initializer.Constant(2.0))
    opt = paddle.optimizer.SGD(0.1, parameters=[p], weight_decay=paddle.regularizer.L1Decay(0.5))
    (p * 0).sum().backward()
    opt.step()
    np.testing.assert_allclose(p.numpy(), 2.0 - 0.1 * 0.5, rtol=1e-6)
    es = paddle.callbacks.EarlyStopping(monitor="loss", patience=1, verbose=0)
    assert hasattr(es, "on_eval_end") and paddle.callbacks.LRScheduler is not None and paddle.callbacks.ModelCheckpoint is not None


def test_audio_features():
    sig = paddle.to_tensor(np.random.randn(1, 4000).astype("float32"))
    assert paddle.audio.features.Spectrogram(n_fft=256, hop_length=128)(sig).shape[1] == 129
    mel = paddle.audio.features.MelSpectrogram(sr=8000, n_fft=256, hop_length=128, n_mels=20)(sig)
    assert mel.shape[1] == 20
    assert paddle.audio.features.MFCC(sr=8000, n_mfcc=13, n_fft=256, hop_length=128, n_mels=20)(sig).shape[1] == 13
    w = paddle.audio.functional.get_window("hann", 16)
    assert w.shape == [16] and abs(float(w[0])) < 1e-6


def test_inference_config_switches_hooks_and_memory_model(tmp_path):
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200 import inference as I

    paddle.seed(0)
    net = paddle.nn.Sequential(paddle.nn.Linear(4, 8), paddle.nn.ReLU(), paddle.nn.Linear(8, 2))
    prefix = str(tmp_path / "m")
    paddle.jit.save(net, prefix, input_spec=[paddle.static.InputSpec([None, 4], "float32", "x")])
    cfg = I.Config(prefix + ".pdmodel", prefix + ".pdiparams")
    pb = cfg.pass_builder()
    n0 = len(pb.all_passes())
    pb.append_pass("my_pass")
    cfg.delete_pass("constant_folding")
    cfg.switch_ir_debug(True)
    cfg.set_optimization_level(2)
    cfg.exp_disable_mixed_precision_ops({"softmax"})
    cfg.enable_low_precision_io(False)
    cfg.collect_shape_range_info(str(tmp_path / "shape.txt"))
    assert len(pb.all_passes()) == n0 and "my_pass" in pb.all_passes() and cfg.shape_range_info_collected() and not cfg.mkldnn_enabled()
    assert cfg.to_native_config()["opt_level"] == 2 and cfg.glog_info_disabled() is False and cfg.new_ir_enabled()
    import pytest

    with pytest.raises(RuntimeError):
        cfg.enable_xpu()
    pred = I.create_predictor(cfg)
    seen = []
    pred.register_input_hook(lambda name, t: seen.append(("in", name)))
    pred.register_output_hook(lambda name, t: seen.append(("out", name)))
    x = np.random.RandomState(0).randn(3, 4).astype("float32")
    h = pred.get_input_handle(pred.get_input_names()[0])
    h.copy_from_cpu(x)
    assert pred.zero_copy_run() and ("in", "x") in seen and ("out", "out0") in seen
    out = pred.get_output_handle(pred.get_output_names()[0])
    np.testing.assert_allclose(out.as_ndarray(), net(paddle.to_tensor(x)).numpy(), rtol=1e-5, atol=1e-6)
    assert out.tolist() == out.copy_to_cpu().tolist() and len(pred.get_serialized_program()) > 0
    # model from memory
    cfg2 = I.Config()
    pm, pp = open(prefix + ".pdmodel", "rb").read(), open(prefix + ".pdiparams", "rb").read()
    cfg2.set_model_buffer(pm, len(pm), pp, len(pp))
    assert cfg2.model_from_memory()
    outs = I.create_predictor(cfg2).run([x])
    np.testing.assert_allclose(outs[0].numpy(), net(paddle.to_tensor(x)).numpy(), rtol=1e-5, atol=1e-6)
    pool = I.PredictorPool(cfg, 2)
    assert pool.retrieve(1) is not pool.retrieve(0)


def test_detection_ops_vectorised_roi_align_and_native_nms():
    """roi_align batches the boxes (no python loop per box) and nms runs its serial scan in the native runtime: both against plain references."""
    import math

    import numpy as np
    import torch

    import paddle_b200 as paddle
    from paddle_b200.vision import ops

    rng = np.random.RandomState(0)
    feat = rng.randn(2, 5, 18, 22).astype("float32")
    k = 30
    xy = rng.rand(k, 2) * np.array([70, 60]) - 6
    wh = rng.rand(k, 2) * np.array([40, 30]) + 1
    boxes = np.concatenate([xy, xy + wh], 1).astype("float32")
    nums = [18, 12]

    def bilinear(f, y, x):
        c, h, w = f.shape
        if y < -1 or y > h or x < -1 or x > w:
            return np.zeros(c, "float32")
        y, x = min(max(y, 0), h - 1), min(max(x, 0), w - 1)
        y0, x0 = int(math.floor(y)), int(math.floor(x))
        y1, x1 = min(y0 + 1, h - 1), min(x0 + 1, w - 1)
        ly, lx = y - y0, x - x0
        return f[:, y0, x0] * (1 - ly) * (1 - lx) + f[:, y0, x1] * (1 - ly) * lx + f[:, y1, x0] * ly * (1 - lx) + f[:, y1, x1] * ly * lx

    def reference(out_hw, scale, ratio, aligned):
        oh, ow = out_hw
        res, bi = [], 0
        for img, n in enumerate(nums):
            for _ in range(n):
                x1, y1, x2, y2 = boxes[bi] * scale - (0.5 if aligned else 0.0)
                bi += 1
                rw, rh = x2 - x1, y2 - y1
                if not aligned:
                    rw, rh = max(rw, 1.0), max(rh, 1.0)
                sh = ratio if ratio > 0 else max(1, math.ceil(rh / oh))
                sw = ratio if ratio > 0 else max(1, math.ceil(rw / ow))
                o = np.zeros((feat.shape[1], oh, ow), "float32")
                for i in range(oh):
                    for j in range(ow):
                        acc = 0
                        for a in range(sh):
                            for b in range(sw):
                                acc = acc + bilinear(feat[img], y1 + i * rh / oh + (a + 0.5) * rh / oh / sh, x1 + j * rw / ow + (b + 0.5) * rw / ow / sw)
                        o[:, i, j] = acc / (sh * sw)
                res.append(o)
        return np.stack(res)

    for out_hw, scale, ratio, aligned in (((3, 3), 0.25, -1, True), ((2, 4), 0.2, 2, False)):
        got = ops.roi_align(paddle.to_tensor(feat), paddle.to_tensor(boxes), paddle.to_tensor(nums, dtype="int32"), out_hw, scale, ratio, aligned).numpy()
        np.testing.assert_allclose(got, reference(out_hw, scale, ratio, aligned), rtol=1e-4, atol=1e-5)
    assert ops.roi_align(paddle.to_tensor(feat), paddle.to_tensor(np.zeros((0, 4), "float32")), paddle.to_tensor([0, 0], dtype="int32"), 2).shape == [0, 5, 2, 2]

    # greedy NMS against the textbook loop
    n = 300
    xy = rng.rand(n, 2) * 60
    bx = np.concatenate([xy, xy + rng.rand(n, 2) * 25 + 2], 1).astype("float32")
    sc = rng.rand(n).astype("float32")

    def iou(a, b):
        iw = max(min(a[2], b[2]) - max(a[0], b[0]), 0)
        ih = max(min(a[3], b[3]) - max(a[1], b[1]), 0)
        inter = iw * ih
        return inter / max((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter, 1e-10)

    order, keep = np.argsort(-sc, kind="stable"), []
    for i in order:
        if all(iou(bx[i], bx[j]) <= 0.45 for j in keep):
            keep.append(i)
    got = ops.nms(paddle.to_tensor(bx), 0.45, paddle.to_tensor(sc)).numpy()
    assert got.tolist() == keep
    from paddle_b200._build import load

    m = load()
    if m is not None and hasattr(m, "nms_scan"):                       # the device path's scan over a suppression matrix
        sb = torch.from_numpy(bx[order])
        from paddle_b200.vision.ops import _iou

        kept = m.nms_scan((_iou(sb, sb) > 0.45).triu_(1).to(torch.uint8))
        assert order[kept.numpy()].tolist() == keep


def test_roi_pool_and_psroi_pool_batched_match_the_bin_loops():
    import math

    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200.vision import ops

    rng = np.random.RandomState(1)
    k = 24
    xy = rng.rand(k, 2) * np.array([60, 50]) - 4
    boxes = np.concatenate([xy, xy + rng.rand(k, 2) * np.array([50, 40]) + 0.5], 1).astype("float32")
    nums = [14, 10]
    img_of = [0] * 14 + [1] * 10
    x = rng.randn(2, 4, 20, 24).astype("float32")
    H, W = 20, 24
    for scale, (oh, ow) in ((0.25, (3, 3)), (1.0, (2, 4))):
        ref = np.zeros((k, 4, oh, ow), "float32")
        for b in range(k):
            x1, y1, x2, y2 = [int(round(float(v) * scale)) for v in boxes[b]]
            rh, rw = max(y2 - y1 + 1, 1), max(x2 - x1 + 1, 1)
            for i in range(oh):
                hs, he = min(max(y1 + math.floor(i * rh / oh), 0), H), min(max(y1 + math.ceil((i + 1) * rh / oh), 0), H)
                for j in range(ow):
                    ws, we = min(max(x1 + math.floor(j * rw / ow), 0), W), min(max(x1 + math.ceil((j + 1) * rw / ow), 0), W)
                    if he > hs and we > ws:
                        ref[b, :, i, j] = x[img_of[b], :, hs:he, ws:we].max((-2, -1))
        got = ops.roi_pool(paddle.to_tensor(x), paddle.to_tensor(boxes), paddle.to_tensor(nums, dtype="int32"), (oh, ow), scale).numpy()
        np.testing.assert_allclose(got, ref, rtol=0, atol=0)
    oh = ow = 3
    xp = rng.randn(2, 2 * oh * ow, H, W).astype("float32")
    ref = np.zeros((k, 2, oh, ow), "float32")
    for b in range(k):
        x1, y1, x2, y2 = [float(v) * 0.5 for v in boxes[b]]
        rh, rw = max(y2 - y1, 0.1), max(x2 - x1, 0.1)
        for i in range(oh):
            hs, he = min(max(math.floor(y1 + i * rh / oh), 0), H), min(max(math.ceil(y1 + (i + 1) * rh / oh), 0), H)
            for j in range(ow):
                ws, we = min(max(math.floor(x1 + j * rw / ow), 0), W), min(max(math.ceil(x1 + (j + 1) * rw / ow), 0), W)
                if he > hs and we > ws:
                    for c in range(2):
                        ref[b, c, i, j] = xp[img_of[b], c * oh * ow + i * ow + j, hs:he, ws:we].mean()
    got = ops.psroi_pool(paddle.to_tensor(xp), paddle.to_tensor(boxes), paddle.to_tensor(nums, dtype="int32"), 3, 0.5).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_deform_conv2d_batched_sampling_matches_a_plain_loop():
    """All sampling positions of all images / groups / taps are gathered at once; checked against a per-output-pixel loop (v2: with mask)."""
    import math

    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200.vision import ops

    rng = np.random.RandomState(0)
    N, C, H, W, Co, k, st, pd, dl, dg = 1, 4, 6, 7, 3, 3, 2, 1, 1, 2
    x = rng.randn(N, C, H, W).astype("float32")
    w = rng.randn(Co, C, k, k).astype("float32")
    Ho, Wo = (H + 2 * pd - dl * (k - 1) - 1) // st + 1, (W + 2 * pd - dl * (k - 1) - 1) // st + 1
    off = (rng.randn(N, 2 * dg * k * k, Ho, Wo) * 1.2).astype("float32")
    m = rng.rand(N, dg * k * k, Ho, Wo).astype("float32")

    def sample(f, y, xq):
        y0, x0 = math.floor(y), math.floor(xq)
        v = 0.0
        for dy in (0, 1):
            for dx in (0, 1):
                yi, xi = y0 + dy, x0 + dx
                if 0 <= yi < H and 0 <= xi < W:
                    v += f[yi, xi] * (1 - abs(y - yi)) * (1 - abs(xq - xi))
        return v

    ref = np.zeros((N, Co, Ho, Wo), "float32")
    cpg = C // dg
    for o in range(Co):
        for i in range(Ho):
            for j in range(Wo):
                acc = 0.0
                for c in range(C):
                    g = c // cpg
                    for t in range(k * k):
                        ky, kx = divmod(t, k)
                        y = i * st - pd + ky * dl + off[0, (g * k * k + t) * 2, i, j]
                        xq = j * st - pd + kx * dl + off[0, (g * k * k + t) * 2 + 1, i, j]
                        acc += w[o, c, ky, kx] * sample(x[0, c], y, xq) * m[0, g * k * k + t, i, j]
                ref[0, o, i, j] = acc
    got = ops.deform_conv2d(paddle.to_tensor(x), paddle.to_tensor(off), paddle.to_tensor(w), None, st, pd, dl, dg, 1, paddle.to_tensor(m)).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4)
