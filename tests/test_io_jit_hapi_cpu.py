"""DataLoader / datasets / samplers, paddle.save/load formats, jit.to_static + save/load, hapi Model, metrics, vision.
Parity: test/legacy_test/test_dataloader_*.py, test_paddle_save_load.py, test/dygraph_to_static, test_model.py, test_metrics.py."""
import numpy as np
import pytest

import paddle_b200 as paddle
from paddle_b200 import nn
from paddle_b200.io import BatchSampler, DataLoader, Dataset, DistributedBatchSampler, IterableDataset, RandomSampler, TensorDataset, random_split


class Sq(Dataset):
    def __init__(self, n=20):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return np.array([i, i * i], "float32"), np.array(i % 2, "int64")


def test_dataloader_basic_and_workers():
    dl = DataLoader(Sq(), batch_size=6, shuffle=False, drop_last=False)
    batches = list(dl)
    assert len(dl) == 4 and [b[0].shape[0] for b in batches] == [6, 6, 6, 2]
    assert batches[0][0].dtype == paddle.float32 and batches[0][1].dtype == paddle.int64
    np.testing.assert_array_equal(batches[1][0].numpy()[:, 0], np.arange(6, 12))
    dl = DataLoader(Sq(), batch_size=5, shuffle=True, drop_last=True, num_workers=2)
    seen = np.concatenate([b[0].numpy()[:, 0] for b in dl])
    assert sorted(seen.tolist()) == list(range(20))
    dl = DataLoader(Sq(), batch_size=4, collate_fn=lambda items: len(items))
    assert list(dl) == [4] * 5


def test_samplers_and_datasets():
    bs = BatchSampler(Sq(10), batch_size=3, drop_last=True)
    assert len(bs) == 3 and [len(b) for b in bs] == [3, 3, 3]
    assert sorted(RandomSampler(Sq(10))) == list(range(10))
    parts = [list(DistributedBatchSampler(Sq(10), batch_size=2, num_replicas=2, rank=r)) for r in range(2)]
    flat = [sorted(sum(p, [])) for p in parts]
    assert len(flat[0]) == len(flat[1]) == 5 and set(flat[0]) | set(flat[1]) == set(range(10))
    td = TensorDataset([paddle.arange(6).reshape([6, 1]), paddle.arange(6)])
    assert len(td) == 6 and int(td[2][1]) == 2
    a, b = random_split(Sq(10), [7, 3])
    assert len(a) == 7 and len(b) == 3
    w = paddle.io.WeightedRandomSampler([0.0, 1.0, 0.0], 5)
    assert list(w) == [1] * 5

    class It(IterableDataset):
        def __iter__(self):
            yield from (np.array([i], "float32") for i in range(7))

    assert [b.shape[0] for b in DataLoader(It(), batch_size=3)] == [3, 3, 1]
    assert len(paddle.io.ConcatDataset([Sq(3), Sq(4)])) == 7 and len(paddle.io.Subset(Sq(10), [1, 3])) == 2


def test_save_load_formats(tmp_path):
    obj = {"w": paddle.to_tensor(np.arange(6, dtype="float32").reshape(2, 3)), "step": 3, "nested": {"b": paddle.ones([2], dtype="bfloat16")},
           "lst": [paddle.zeros([1]), "s"]}
    p = str(tmp_path / "o.pdparams")
    paddle.save(obj, p)
    back = paddle.load(p)
    np.testing.assert_array_equal(back["w"].numpy(), obj["w"].numpy())
    assert back["step"] == 3 and back["nested"]["b"].dtype == paddle.bfloat16 and back["lst"][1] == "s"
    raw = paddle.load(p, return_numpy=True)
    assert isinstance(raw["w"], np.ndarray)
    import pickle

    with open(p, "rb") as f:
        plain = pickle.load(f)      # reference format: a plain pickle of numpy arrays readable without the framework
    assert isinstance(plain["w"], np.ndarray) and plain["nested"]["b"].dtype == np.uint16
    t = paddle.to_tensor([1.0, 2.0])
    paddle.save(t, str(tmp_path / "t.pdtensor"))
    np.testing.assert_array_equal(paddle.load(str(tmp_path / "t.pdtensor")).numpy(), [1, 2])


def test_jit_to_static_and_save_load(tmp_path):
    class Net(nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(4, 3)

        def forward(self, x):
            if x.shape[0] > 2:          # python control flow is evaluated per shape signature
                return paddle.tanh(self.fc(x))
            return self.fc(x)

    net = Net()
    snet = paddle.jit.to_static(net, input_spec=[paddle.static.InputSpec([None, 4], "float32")])
    x = paddle.to_tensor(np.random.rand(5, 4).astype("float32"))
    np.testing.assert_allclose(snet(x).numpy(), net(x).numpy(), rtol=1e-6)
    np.testing.assert_allclose(snet(x[:2]).numpy(), net.fc(x[:2]).numpy(), rtol=1e-6)

    @paddle.jit.to_static
    def f(a, b):
        return a * 2 + b

    np.testing.assert_allclose(f(paddle.ones([2]), paddle.ones([2])).numpy(), [3, 3])
    # a locally defined class cannot be re-imported by jit.load: the forward is stored as a traced Program instead
    class Plain(nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc, self.norm = nn.Linear(4, 3), nn.LayerNorm(3)

        def forward(self, x):
            return paddle.tanh(self.norm(self.fc(x)))

    net = Plain()
    net.eval()
    paddle.jit.save(net, str(tmp_path / "net"), input_spec=[paddle.static.InputSpec([None, 4], "float32")])
    loaded = paddle.jit.load(str(tmp_path / "net"))
    np.testing.assert_allclose(loaded(x).numpy(), net(x).numpy(), rtol=1e-6)
    np.testing.assert_allclose(loaded(x[:3]).numpy(), net(x[:3]).numpy(), rtol=1e-6)
    assert paddle.jit.not_to_static(lambda: 1)() == 1
    paddle.jit.enable_to_static(False)
    paddle.jit.enable_to_static(True)


def test_hapi_model_fit_evaluate_predict(tmp_path):
    paddle.seed(0)
    rng = np.random.RandomState(0)
    X = rng.rand(128, 8).astype("float32")
    Y = (X.sum(1) > 4).astype("int64")

    class DS(Dataset):
        def __len__(self):
            return 128

        def __getitem__(self, i):
            return X[i], Y[i]

    net = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 2))
    model = paddle.Model(net)
    model.prepare(paddle.optimizer.Adam(0.02, parameters=net.parameters()), nn.CrossEntropyLoss(), paddle.metric.Accuracy())
    model.fit(DS(), epochs=40, batch_size=32, verbose=0)
    res = model.evaluate(DS(), batch_size=64, verbose=0)
    assert res["acc"] > 0.8 and "loss" in res
    pred = model.predict(DS(), batch_size=64, verbose=0)
    assert np.concatenate(pred[0]).shape == (128, 2)
    model.save(str(tmp_path / "ck"))
    m2 = paddle.Model(nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 2)))
    m2.prepare(loss=nn.CrossEntropyLoss(), metrics=paddle.metric.Accuracy())
    m2.load(str(tmp_path / "ck"))
    assert abs(m2.evaluate(DS(), batch_size=64, verbose=0)["acc"] - res["acc"]) < 1e-6
    info = paddle.summary(net, (1, 8))
    assert info["total_params"] == 8 * 16 + 16 + 16 * 2 + 2
    assert paddle.flops(net, [1, 8], print_detail=False) > 0


def test_metrics():
    m = paddle.metric.Accuracy(topk=(1, 2))
    pred = paddle.to_tensor(np.array([[0.1, 0.7, 0.2], [0.5, 0.3, 0.2], [0.2, 0.3, 0.5]], "float32"))
    lab = paddle.to_tensor(np.array([[1], [1], [0]]))
    m.update(m.compute(pred, lab))
    a1, a2 = m.accumulate()
    assert a1 == pytest.approx(1 / 3) and a2 == pytest.approx(2 / 3)
    p, r = paddle.metric.Precision(), paddle.metric.Recall()
    pr, lb = np.array([0.9, 0.8, 0.2, 0.7]), np.array([1, 0, 1, 1])
    p.update(pr, lb)
    r.update(pr, lb)
    assert p.accumulate() == pytest.approx(2 / 3) and r.accumulate() == pytest.approx(2 / 3)
    auc = paddle.metric.Auc()
    auc.update(np.stack([1 - pr, pr], 1), lb.reshape(-1, 1))
    assert 0.0 <= auc.accumulate() <= 1.0
    assert float(paddle.metric.accuracy(pred, lab, k=1)) == pytest.approx(1 / 3)


def test_vision_models_transforms_ops():
    from paddle_b200.vision import models, ops, transforms

    x = paddle.to_tensor(np.random.rand(1, 3, 64, 64).astype("float32"))
    for ctor in (models.resnet18, models.mobilenet_v2, models.squeezenet1_1, models.shufflenet_v2_x0_25):
        net = ctor(num_classes=10)
        net.eval()
        assert net(x).shape == [1, 10]
    assert models.LeNet()(paddle.zeros([2, 1, 28, 28])).shape == [2, 10]
    r50 = models.resnet50()
    assert sum(p.numel() for p in r50.parameters()) == 25557032
    img = (np.random.rand(40, 50, 3) * 255).astype("uint8")
    t = transforms.Compose([transforms.Resize(32), transforms.CenterCrop(24), transforms.RandomHorizontalFlip(), transforms.ToTensor(),
                            transforms.Normalize([0.5] * 3, [0.5] * 3)])
    out = t(img)
    assert out.shape == [3, 24, 24] and float(out.min()) >= -1.0001
    assert transforms.RandomCrop(16)(img).shape[:2] == (16, 16) and transforms.Pad(2)(img).shape[:2] == (44, 54)
    assert transforms.ColorJitter(0.2, 0.2, 0.2, 0.1)(img).shape == img.shape and transforms.Grayscale()(img).shape[:2] == (40, 50)
    boxes = paddle.to_tensor(np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30]], "float32"))
    keep = ops.nms(boxes, 0.5, scores=paddle.to_tensor(np.array([0.9, 0.8, 0.7], "float32")))
    assert keep.numpy().tolist() == [0, 2]
    feat = paddle.to_tensor(np.random.rand(1, 4, 16, 16).astype("float32"))
    ra = ops.roi_align(feat, paddle.to_tensor(np.array([[0, 0, 8, 8]], "float32")), paddle.to_tensor(np.array([1], "int32")), 4)
    assert ra.shape == [1, 4, 4, 4]
    dc = ops.DeformConv2D(4, 6, 3, padding=1)
    off = paddle.zeros([1, 18, 16, 16])
    assert dc(feat, off).shape == [1, 6, 16, 16]


def test_distribution_sparse_geometric_text():
    D = paddle.distribution
    n = D.Normal(paddle.to_tensor([0.0]), paddle.to_tensor([2.0]))
    assert float(n.log_prob(paddle.to_tensor([0.0]))) == pytest.approx(-np.log(2 * np.sqrt(2 * np.pi)), rel=1e-5)
    assert float(n.entropy()) == pytest.approx(0.5 * np.log(2 * np.pi * np.e * 4), rel=1e-5)
    assert float(D.kl_divergence(n, D.Normal(paddle.to_tensor([0.0]), paddle.to_tensor([2.0])))) == pytest.approx(0.0, abs=1e-6)
    assert D.Categorical(paddle.to_tensor([0.2, 0.8])).sample([7]).shape == [7]
    assert D.Uniform(0.0, 2.0).sample([3]).shape[0] == 3 and float(D.Bernoulli(paddle.to_tensor(0.3)).mean) == pytest.approx(0.3)
    sp = paddle.sparse.sparse_coo_tensor(np.array([[0, 1], [1, 0]]), np.array([2.0, 3.0], "float32"), [2, 2])
    np.testing.assert_array_equal(sp.to_dense().numpy(), [[0, 2], [3, 0]])
    np.testing.assert_array_equal(paddle.sparse.matmul(sp, paddle.eye(2)).numpy(), [[0, 2], [3, 0]])
    csr = sp.to_sparse_csr()
    assert csr.crows().numpy().tolist() == [0, 1, 2]
    x = paddle.to_tensor(np.array([[1.0], [2.0], [4.0]], "float32"))
    src, dst = paddle.to_tensor(np.array([0, 1, 2])), paddle.to_tensor(np.array([1, 1, 0]))
    np.testing.assert_array_equal(paddle.geometric.send_u_recv(x, src, dst, "sum").numpy(), [[4], [3], [0]])
    np.testing.assert_array_equal(paddle.geometric.segment_sum(x, paddle.to_tensor(np.array([0, 0, 1]))).numpy(), [[3], [4]])
    pot = paddle.to_tensor(np.random.rand(1, 4, 3).astype("float32"))
    trans = paddle.to_tensor(np.random.rand(3, 3).astype("float32"))
    sc, path = paddle.text.viterbi_decode(pot, trans, paddle.to_tensor(np.array([4])), include_bos_eos_tag=False)
    assert path.shape == [1, 4] and sc.shape == [1]


def test_jit_save_of_local_class_is_the_eval_program(tmp_path):
    """A Layer whose class cannot be re-imported is saved as its traced inference program: traced in eval mode (dropout off, batch norm on
    running statistics) even when the layer is training, dynamic batch on reload, clear error when nothing can be saved."""
    def make():
        class Local(paddle.nn.Layer):
            def __init__(self):
                super().__init__()
                self.fc, self.drop, self.bn = paddle.nn.Linear(4, 4), paddle.nn.Dropout(0.5), paddle.nn.BatchNorm1D(4)

            def forward(self, x):
                return self.bn(self.drop(self.fc(x)))

        return Local()

    net = make()
    net.train()
    paddle.jit.save(net, str(tmp_path / "m"), input_spec=[paddle.static.InputSpec([None, 4], "float32", "x")])
    assert net.training
    loaded = paddle.jit.load(str(tmp_path / "m"))
    x = paddle.to_tensor(np.random.RandomState(0).randn(6, 4).astype("float32"))
    a, b = loaded(x).numpy(), loaded(x).numpy()
    net.eval()
    np.testing.assert_allclose(a, b)
    np.testing.assert_allclose(a, net(x).numpy(), rtol=1e-5, atol=1e-6)
    assert loaded(x[:3]).shape == [3, 4]
    with pytest.raises(RuntimeError, match="input_spec"):
        paddle.jit.save(make(), str(tmp_path / "n"))


def test_dy2static_ast_conversion():
    """to_static converts tensor-dependent `if` / logical ops (run both branches + device-side select), keeps Python control flow,
    exposes the transformed code, and gradients follow the selected branch. Parity: test/dygraph_to_static/test_ifelse.py."""
    import torch

    import paddle_b200 as paddle
    from paddle_b200.jit import dy2static as D

    class Net(paddle.nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc = paddle.nn.Linear(4, 4)

        def forward(self, x, scale):
            h = self.fc(x)
            if h.mean() > 0 and scale > 0:
                out = h * scale
                tag = paddle.ones([1])
            else:
                out = h - scale
                tag = paddle.zeros([1])
            k = 0
            while k < 2:
                out = out + 1
                k += 1
            return out, tag

    paddle.seed(0)
    net, ref = Net(), Net()
    ref.set_state_dict(net.state_dict())
    snet = paddle.jit.to_static(net)
    assert "convert_ifelse" in snet.forward.code and "convert_logical_and" in snet.forward.code
    for sign in (1.0, -1.0):
        x = paddle.to_tensor(np.full((2, 4), sign, dtype=np.float32))
        a, ta = snet(x, 3.0)
        b, tb = ref(x, 3.0)
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-6)
        np.testing.assert_allclose(ta.numpy(), tb.numpy())
    x = paddle.to_tensor(np.ones((2, 4), dtype=np.float32), stop_gradient=False)
    snet(x, 2.0)[0].sum().backward()
    xr = paddle.to_tensor(np.ones((2, 4), dtype=np.float32), stop_gradient=False)
    ref(xr, 2.0)[0].sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), xr.grad.numpy(), rtol=1e-6)
    np.testing.assert_allclose(net.fc.weight.grad.numpy(), ref.fc.weight.grad.numpy(), rtol=1e-6)

    def loop(x):
        n = paddle.zeros([1])
        while x.sum() < 50:
            x = x * 2
            n = n + 1
        return x, n

    g = D.convert_to_static(loop)
    a, b = g(paddle.ones([3])), loop(paddle.ones([3]))
    np.testing.assert_allclose(a[0].numpy(), b[0].numpy())
    assert float(a[1]) == float(b[1]) == 5.0
    pt = D.ProgramTranslator()
    assert "convert_while_loop" in pt.get_code(loop)


def test_sparse_conv_pool_rulebook_matches_dense():
    """Sparse conv3d / subm_conv3d / conv2d / max_pool3d through the gather-GEMM-scatter rulebook (no densification) == dense conv on
    the active output sites, for stride / padding / dilation / groups; gradients reach the weights."""
    import torch

    from paddle_b200.sparse.nn import functional as SF

    torch.manual_seed(0)
    def dense_ref3d(xd, w, b, stride, padding, dilation, groups):
        return torch.nn.functional.conv3d(xd.permute(0,4,1,2,3), w.permute(4,3,0,1,2), b, stride, padding, dilation, groups).permute(0,2,3,4,1)
    for (stride,padding,dilation,groups) in [(1,0,1,1),(2,1,1,1),(1,1,2,1),(1,1,1,2)]:
        xd = torch.randn(2,6,7,5,4)*(torch.rand(2,6,7,5,1)>0.7)
        x = xd.to_sparse(4)
        w = torch.randn(3,3,3,4//groups,6); b = torch.randn(6)
        out = SF.conv3d(x, w, b, stride, padding, dilation, groups).coalesce()
        ref = dense_ref3d(xd, w, None, stride, padding, dilation, groups)
        act = torch.nn.functional.conv3d((xd.abs().sum(-1,keepdim=True)>0).float().permute(0,4,1,2,3), torch.ones(1,1,3,3,3), None, stride, padding, dilation)[:,0]>0
        od = out.to_dense()
        assert torch.allclose(od[act], ref[act]+b, atol=1e-4), (stride,padding,dilation,groups)
        assert out.indices().shape[1]==int(act.sum())
        outs = SF.subm_conv3d(x, w, None, 1, 0, dilation, groups).coalesce()
        pad = tuple(dilation*(3-1)//2 for _ in range(3))
        refs = dense_ref3d(xd, w, None, 1, pad, dilation, groups)
        xi = x.coalesce().indices()
        assert torch.allclose(outs.to_dense()[xi[0],xi[1],xi[2],xi[3]], refs[xi[0],xi[1],xi[2],xi[3]], atol=1e-4)
    xd = torch.randn(1,5,5,5,3)*(torch.rand(1,5,5,5,1)>0.6)
    x = xd.to_sparse(4).requires_grad_(True)
    w = torch.randn(3,3,3,3,4, requires_grad=True)
    out = SF.conv3d(x, w, None, 1, 1, 1, 1)
    out.coalesce().values().pow(2).sum().backward()
    assert w.grad is not None and w.grad.abs().sum()>0
    xp = (torch.rand(1,4,4,4,2)+0.1)*(torch.rand(1,4,4,4,1)>0.5)
    o = SF.max_pool3d(xp.to_sparse(4), 2, 2).to_dense()
    neg = torch.where(xp==0, torch.full_like(xp,float('-inf')), xp)
    r = torch.nn.functional.max_pool3d(neg.permute(0,4,1,2,3),2,2).permute(0,2,3,4,1); r = torch.where(torch.isinf(r), torch.zeros_like(r), r)
    assert torch.allclose(o, r)
    x2 = torch.randn(2,6,6,3)*(torch.rand(2,6,6,1)>0.6)
    w2 = torch.randn(3,3,3,5)
    o2 = SF.conv2d(x2.to_sparse(3), w2, None, 1, 1).to_dense()
    r2 = torch.nn.functional.conv2d(x2.permute(0,3,1,2), w2.permute(3,2,0,1), None, 1, 1).permute(0,2,3,1)
    act2 = torch.nn.functional.conv2d((x2.abs().sum(-1,keepdim=True)>0).float().permute(0,3,1,2), torch.ones(1,1,3,3), None, 1, 1)[:,0]>0
    assert torch.allclose(o2[act2], r2[act2], atol=1e-4)


def test_jit_save_refuses_a_trace_that_bakes_constants(tmp_path):
    """A locally-defined layer is saved as a traced program; when its forward computes on raw tensors the trace is wrong and save says so."""
    import torch

    import paddle_b200 as paddle
    from paddle_b200.static import InputSpec

    class Fine(paddle.nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc = paddle.nn.Linear(4, 3)

        def forward(self, x):
            return paddle.tanh(self.fc(x))

    class Leaky(Fine):
        def forward(self, x):
            return paddle.tanh(self.fc(x)) + x.as_subclass(torch.Tensor).sum(-1, keepdim=True).cos().as_subclass(paddle.Tensor)

    class Blind(Fine):
        def forward(self, x):
            return (self.fc(x).as_subclass(torch.Tensor) * 2.0).as_subclass(paddle.Tensor)

    spec = [InputSpec([2, 4], "float32", "x")]
    paddle.jit.save(Fine(), str(tmp_path / "fine"), input_spec=spec)
    x = paddle.randn([2, 4])
    m = Fine()
    paddle.jit.save(m, str(tmp_path / "fine2"), input_spec=spec)
    loaded = paddle.jit.load(str(tmp_path / "fine2"))
    assert torch.allclose(loaded(x).as_subclass(torch.Tensor), m(x).as_subclass(torch.Tensor), atol=1e-6)
    import pytest

    with pytest.raises(RuntimeError, match="does not reproduce"):
        paddle.jit.save(Leaky(), str(tmp_path / "leaky"), input_spec=spec)
    with pytest.raises(RuntimeError, match="not values of the traced program"):
        paddle.jit.save(Blind(), str(tmp_path / "blind"), input_spec=spec)
