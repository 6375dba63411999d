"""Optimizers (closed-form single-step checks + convergence), LR schedulers, AMP, save/load of optimizer state.
Parity: test/legacy_test/test_adam_op.py, test_adamw_op.py, test_momentum_op.py, test_lr_scheduler.py, test_amp_*.py."""
import math

import numpy as np
import pytest

import paddle_b200 as paddle
from paddle_b200 import nn

W0 = np.array([1.0, -2.0, 3.0], dtype="float32")
G = np.array([0.1, -0.2, 0.3], dtype="float32")


def one_step(opt_cls, **kw):
    p = paddle.create_parameter([3], "float32", default_initializer=nn.initializer.Assign(W0))
    opt = opt_cls(parameters=[p], **kw)
    (p * paddle.to_tensor(G)).sum().backward()
    opt.step()
    opt.clear_grad()
    return p.numpy(), opt, p


def test_sgd_momentum_closed_form():
    w, _, _ = one_step(paddle.optimizer.SGD, learning_rate=0.1)
    np.testing.assert_allclose(w, W0 - 0.1 * G, rtol=1e-6)
    w, _, _ = one_step(paddle.optimizer.SGD, learning_rate=0.1, weight_decay=0.01)
    np.testing.assert_allclose(w, W0 - 0.1 * (G + 0.01 * W0), rtol=1e-6)
    w, opt, p = one_step(paddle.optimizer.Momentum, learning_rate=0.1, momentum=0.9)
    np.testing.assert_allclose(w, W0 - 0.1 * G, rtol=1e-6)
    (p * paddle.to_tensor(G)).sum().backward()
    opt.step()
    np.testing.assert_allclose(p.numpy(), W0 - 0.1 * G - 0.1 * (0.9 * G + G), rtol=1e-5)
    w, _, _ = one_step(paddle.optimizer.Momentum, learning_rate=0.1, momentum=0.9, use_nesterov=True)
    np.testing.assert_allclose(w, W0 - 0.1 * (G + 0.9 * G), rtol=1e-6)


def test_adam_family_closed_form():
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-8
    m, v = (1 - b1) * G, (1 - b2) * G * G
    upd = (m / (1 - b1)) / (np.sqrt(v / (1 - b2)) + eps)
    w, _, _ = one_step(paddle.optimizer.Adam, learning_rate=lr)
    np.testing.assert_allclose(w, W0 - lr * upd, rtol=1e-5)
    w, _, _ = one_step(paddle.optimizer.AdamW, learning_rate=lr, weight_decay=0.1)
    np.testing.assert_allclose(w, W0 * (1 - lr * 0.1) - lr * upd, rtol=1e-5)
    w, _, _ = one_step(paddle.optimizer.AdamW, learning_rate=lr, weight_decay=0.1, apply_decay_param_fun=lambda n: False)
    np.testing.assert_allclose(w, W0 - lr * upd, rtol=1e-5)
    w, _, _ = one_step(paddle.optimizer.Adamax, learning_rate=lr)
    np.testing.assert_allclose(w, W0 - lr / (1 - b1) * m / (np.abs(G) + eps), rtol=1e-5)
    w, _, _ = one_step(paddle.optimizer.Adagrad, learning_rate=lr, epsilon=1e-6)
    np.testing.assert_allclose(w, W0 - lr * G / (np.sqrt(G * G) + 1e-6), rtol=1e-5)
    w, _, _ = one_step(paddle.optimizer.RMSProp, learning_rate=lr, rho=0.95, epsilon=1e-6)
    np.testing.assert_allclose(w, W0 - lr * G / np.sqrt(0.05 * G * G + 1e-6), rtol=1e-4)
    w, _, _ = one_step(paddle.optimizer.Lamb, learning_rate=lr, lamb_weight_decay=0.0)
    r = upd
    np.testing.assert_allclose(w, W0 - lr * np.linalg.norm(W0) / np.linalg.norm(r) * r, rtol=1e-4)


def test_flat_arena_matches_per_tensor():
    def run(arena):
        paddle.seed(0)
        net = nn.Sequential(nn.Linear(6, 12), nn.Tanh(), nn.Linear(12, 2))
        opt = paddle.optimizer.AdamW(1e-2, parameters=net.parameters(), weight_decay=0.05, grad_clip=nn.ClipGradByGlobalNorm(0.5))
        if arena:
            opt.enable_flat_arena()
        x = paddle.to_tensor(np.random.RandomState(0).rand(8, 6).astype("float32"))
        for _ in range(4):
            net(x).square().mean().backward()
            opt.step()
            opt.clear_grad()
        return [p.numpy() for p in net.parameters()], opt

    a, oa = run(True)
    b, ob = run(False)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, rtol=2e-5, atol=1e-7)
    sa, sb = oa.state_dict(), ob.state_dict()
    ka, kb = sorted(k for k in sa if k.endswith("_moment1_0")), sorted(k for k in sb if k.endswith("_moment1_0"))
    assert len(ka) == len(kb) == 4   # parameter names differ between the two nets (unique-name counter), order does not
    np.testing.assert_allclose(sa[ka[0]].numpy(), sb[kb[0]].numpy(), rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("cls", ["SGD", "Momentum", "Adam", "AdamW", "Adamax", "Adagrad", "Adadelta", "RMSProp", "Lamb", "NAdam", "RAdam", "ASGD", "Rprop"])
def test_all_optimizers_reduce_quadratic(cls):
    if not hasattr(paddle.optimizer, cls):
        pytest.skip(cls)
    paddle.seed(0)
    p = paddle.create_parameter([4], "float32", default_initializer=nn.initializer.Assign(np.array([1.0, -1.0, 2.0, -2.0], "float32")))
    kw = {"learning_rate": 1.0 if cls == "Adadelta" else 0.05}
    opt = getattr(paddle.optimizer, cls)(parameters=[p], **kw)
    first = None
    for _ in range(60):
        loss = (p * p).sum()
        first = first if first is not None else float(loss)
        loss.backward()
        opt.step()
        opt.clear_grad()
    assert float((p * p).sum()) < first


def test_optimizer_state_roundtrip(tmp_path):
    paddle.seed(1)
    net = nn.Linear(3, 2)
    sched = paddle.optimizer.lr.StepDecay(0.1, step_size=2, gamma=0.5)
    opt = paddle.optimizer.Adam(sched, parameters=net.parameters())
    x = paddle.ones([4, 3])
    for _ in range(3):
        net(x).sum().backward()
        opt.step()
        opt.clear_grad()
        sched.step()
    paddle.save(opt.state_dict(), str(tmp_path / "o.pdopt"))
    paddle.save(net.state_dict(), str(tmp_path / "m.pdparams"))
    net2 = nn.Linear(3, 2)
    net2.set_state_dict(paddle.load(str(tmp_path / "m.pdparams")))
    for (_, a), (_, b) in zip(net.named_parameters(), net2.named_parameters()):
        b.name = a.name   # state is keyed by parameter name (as in the reference)
    sched2 = paddle.optimizer.lr.StepDecay(0.1, step_size=2, gamma=0.5)
    opt2 = paddle.optimizer.Adam(sched2, parameters=net2.parameters())
    opt2.set_state_dict(paddle.load(str(tmp_path / "o.pdopt")))
    assert abs(opt2.get_lr() - opt.get_lr()) < 1e-9
    for o, n in ((opt, net), (opt2, net2)):
        n(x).sum().backward()
        o.step()
    np.testing.assert_allclose(net.weight.numpy(), net2.weight.numpy(), rtol=1e-6)


def test_lr_schedulers():
    L = paddle.optimizer.lr

    def seq(s, n=6):
        out = []
        for _ in range(n):
            out.append(s())
            s.step()
        return out

    np.testing.assert_allclose(seq(L.StepDecay(1.0, 2, 0.5)), [1, 1, .5, .5, .25, .25])
    np.testing.assert_allclose(seq(L.MultiStepDecay(1.0, [1, 3], 0.1)), [1, .1, .1, .01, .01, .01], rtol=1e-6)
    np.testing.assert_allclose(seq(L.ExponentialDecay(1.0, 0.5), 3), [1, .5, .25])
    np.testing.assert_allclose(seq(L.CosineAnnealingDecay(1.0, 4), 5), [0.5 * (1 + math.cos(math.pi * t / 4)) for t in range(5)], atol=1e-6)
    np.testing.assert_allclose(seq(L.LinearWarmup(1.0, 4, 0.0, 1.0), 6), [0, .25, .5, .75, 1, 1])
    np.testing.assert_allclose(seq(L.PolynomialDecay(1.0, 4, 0.0, power=1.0), 5), [1, .75, .5, .25, 0])
    np.testing.assert_allclose(seq(L.PiecewiseDecay([2, 4], [1.0, 0.5, 0.1])), [1, 1, .5, .5, .1, .1])
    np.testing.assert_allclose(seq(L.NaturalExpDecay(1.0, 0.5), 2), [1, math.exp(-0.5)])
    np.testing.assert_allclose(seq(L.InverseTimeDecay(1.0, 0.5), 3), [1, 1 / 1.5, 1 / 2])
    np.testing.assert_allclose(seq(L.LambdaDecay(1.0, lambda e: 0.9 ** e), 3), [1, .9, .81])
    noam = seq(L.NoamDecay(64, 4, 1.0), 8)
    assert np.argmax(noam) in (3, 4)
    r = L.ReduceOnPlateau(1.0, patience=1, factor=0.5)
    for m in (1.0, 1.0, 1.0, 1.0):
        r.step(m)
    assert r() < 1.0
    oc = seq(L.OneCycleLR(1.0, 10), 10)
    assert max(oc) == pytest.approx(1.0, rel=1e-3) and oc[-1] < oc[0]
    cyc = seq(L.CyclicLR(0.1, 1.0, 2), 5)
    assert cyc[2] == pytest.approx(1.0) and cyc[4] == pytest.approx(0.1)
    sd = L.StepDecay(1.0, 2).state_dict()
    assert "last_epoch" in sd


def test_amp_autocast_and_scaler():
    net = nn.Linear(4, 4)
    x = paddle.ones([2, 4])
    with paddle.amp.auto_cast(level="O1", dtype="bfloat16"):
        y = net(x)
        z = paddle.nn.functional.softmax(y)     # black-listed op stays fp32
    assert y.dtype == paddle.bfloat16 and z.dtype == paddle.float32
    with paddle.amp.auto_cast(enable=False):
        assert net(x).dtype == paddle.float32
    with paddle.amp.auto_cast(custom_black_list={"linear", "matmul"}, dtype="bfloat16"):
        assert net(x).dtype == paddle.float32
    m2, o2 = paddle.amp.decorate(nn.Linear(4, 4), paddle.optimizer.SGD(0.1, parameters=net.parameters()), level="O2", dtype="bfloat16")
    assert m2.weight.dtype == paddle.bfloat16
    p = paddle.create_parameter([2], "float32", default_initializer=nn.initializer.Constant(1.0))
    opt = paddle.optimizer.SGD(0.1, parameters=[p])
    sc = paddle.amp.GradScaler(init_loss_scaling=8.0, incr_every_n_steps=1, decr_every_n_nan_or_inf=1)
    sc.scale((p * 2).sum()).backward()
    np.testing.assert_allclose(p.grad.numpy(), [16, 16])
    sc.step(opt)
    sc.update()
    np.testing.assert_allclose(p.numpy(), [0.8, 0.8], rtol=1e-6)
    assert sc._scale_value() == 16.0 if hasattr(sc, "_scale_value") else True
    opt.clear_grad()
    sc.scale((p * float("inf")).sum()).backward()
    sc.step(opt)          # skipped
    sc.update()
    np.testing.assert_allclose(p.numpy(), [0.8, 0.8], rtol=1e-6)
    st = sc.state_dict()
    assert st["scale"] < 16.0 or float(np.asarray(st["scale"])) < 16.0


def test_flat_arena_adamw_checkpoint_resume():
    """AdamW on flat arenas: optimizer state_dict -> new optimizer -> set_state_dict continues exactly like the uninterrupted run,
    whether the arena is enabled before or after loading."""
    import numpy as np

    import paddle_b200 as paddle

    def make(seed=0):
        paddle.seed(seed)
        net = paddle.nn.Sequential(paddle.nn.Linear(6, 12), paddle.nn.Tanh(), paddle.nn.Linear(12, 3))
        opt = paddle.optimizer.AdamW(5e-2, parameters=net.parameters(), weight_decay=0.1, grad_clip=paddle.nn.ClipGradByGlobalNorm(0.5))
        return net, opt

    rng = np.random.RandomState(0)
    data = [(paddle.to_tensor(rng.randn(5, 6).astype("float32")), paddle.to_tensor(rng.randn(5, 3).astype("float32"))) for _ in range(5)]

    def run(net, opt, batches):
        for x, y in batches:
            ((net(x) - y) ** 2).mean().backward()
            opt.step()
            opt.clear_grad()

    net_a, opt_a = make()
    opt_a.enable_flat_arena()
    run(net_a, opt_a, data)
    for arena_first in (True, False):
        net_b, opt_b = make()
        opt_b.enable_flat_arena()
        run(net_b, opt_b, data[:3])
        msd = {k: v.numpy().copy() for k, v in net_b.state_dict().items()}
        osd = opt_b.state_dict()
        net_c, opt_c = make(seed=123)                    # different init: everything must come from the checkpoint
        net_c.set_state_dict({k: paddle.to_tensor(v) for k, v in msd.items()})
        # parameter names differ between instances: remap the optimizer entries by position
        remap = {}
        for pb, pc in zip(net_b.parameters(), net_c.parameters()):
            for suf in ("_moment1_0", "_moment2_0", "_beta1_pow_acc_0", "_beta2_pow_acc_0"):
                if pb.name + suf in osd:
                    remap[pc.name + suf] = osd[pb.name + suf]
        for k, v in osd.items():
            if not any(k.startswith(p.name + "_") for p in net_b.parameters()):
                remap[k] = v
        if arena_first:
            opt_c.enable_flat_arena()
            opt_c.set_state_dict(remap)
        else:
            opt_c.set_state_dict(remap)
            opt_c.enable_flat_arena()
        run(net_c, opt_c, data[3:])
        for (k, a), (_, c) in zip(net_a.state_dict().items(), net_c.state_dict().items()):
            np.testing.assert_allclose(c.numpy(), a.numpy(), rtol=1e-5, atol=1e-6, err_msg=f"{k} arena_first={arena_first}")


def test_amp_o2_fp32_input_and_checkpoint_roundtrip(tmp_path):
    """O2-decorated bf16 model fed with fp32 data (the input is cast down at the first white-list op); .pdparams / .pdopt round trip with fp32
    master weights reproduces the run bit for bit."""
    import os

    import numpy as np

    import paddle_b200 as paddle

    def build():
        paddle.seed(0)
        net = paddle.nn.Sequential(paddle.nn.Conv1D(2, 4, 3, padding=1), paddle.nn.Flatten(), paddle.nn.Linear(16, 8), paddle.nn.LayerNorm(8), paddle.nn.Linear(8, 2))
        opt = paddle.optimizer.AdamW(1e-2, parameters=net.parameters(), multi_precision=True)
        return paddle.amp.decorate(models=net, optimizers=opt, level="O2", dtype="bfloat16")

    net, opt = build()
    assert {str(p.dtype) for p in net.parameters()} == {"torch.bfloat16", "torch.float32"}        # norm layers stay fp32
    x = paddle.to_tensor(np.random.RandomState(0).randn(3, 2, 4).astype("float32"))
    scaler = paddle.amp.GradScaler(init_loss_scaling=128)

    def loss_of(m):
        with paddle.amp.auto_cast(level="O2", dtype="bfloat16"):
            return m(x).astype("float32").square().mean()

    for _ in range(3):
        scaler.scale(loss_of(net)).backward()
        scaler.step(opt)
        scaler.update()
        opt.clear_grad()
    paddle.save(net.state_dict(), str(tmp_path / "m.pdparams"))
    paddle.save(opt.state_dict(), str(tmp_path / "m.pdopt"))
    sd, od = paddle.load(str(tmp_path / "m.pdparams")), paddle.load(str(tmp_path / "m.pdopt"))
    assert "master_weights" in od
    net2, opt2 = build()
    net2.set_state_dict(sd)
    n1, n2 = [p.name for p in net.parameters()], [p.name for p in net2.parameters()]
    remap = {}
    for k, v in od.items():
        if k == "master_weights":
            remap[k] = {n2[n1.index(n)]: t for n, t in v.items()}
            continue
        for a, b in zip(n1, n2):
            if k.startswith(a + "_"):
                k = b + k[len(a):]
                break
        remap[k] = v
    opt2.set_state_dict(remap)
    l1, l2 = loss_of(net), loss_of(net2)
    assert float(l1) == float(l2)
    l1.backward()
    opt.step()
    l2.backward()
    opt2.step()
    for a, b in zip(net.parameters(), net2.parameters()):
        np.testing.assert_array_equal(a.astype("float32").numpy(), b.astype("float32").numpy())


def test_split_master_weights_match_fp32_master():
    """bf16 arena AdamW: master weights stored as (bf16 parameter, int16 residual) follow the fp32-master trajectory bit for bit
    (except round-to-even ties, 1 fp32 ulp), and checkpoints still carry fp32 master weights that restore exactly."""
    import numpy as np
    import torch

    import paddle_b200 as paddle
    from paddle_b200.optimizer.optimizer import split_master_join, split_master_split

    x = torch.randn(4096) * 3
    x[:4] = torch.tensor([0.0, -0.0, 1e-30, -65504.0])
    w, lo = split_master_split(x)
    back = split_master_join(w, lo)
    assert (back.view(torch.int32) - x.view(torch.int32)).abs().max().item() <= 1
    assert torch.equal(w, x.to(torch.bfloat16))

    def run(split):
        paddle.set_flags({"FLAGS_b200_split_master_weights": split})
        paddle.seed(3)
        paddle.set_default_dtype("bfloat16")
        try:
            net = paddle.nn.Sequential(paddle.nn.Linear(16, 32), paddle.nn.Linear(32, 8))
            opt = paddle.optimizer.AdamW(1e-2, parameters=net.parameters(), weight_decay=0.05, multi_precision=True)
            opt.enable_flat_arena()
            g = torch.Generator().manual_seed(0)
            for _ in range(6):
                xx = torch.randn(4, 16, generator=g).to(torch.bfloat16).as_subclass(paddle.Tensor)
                (net(xx).astype("float32") ** 2).mean().backward()
                opt.step()
                opt.clear_grad()
            return net, opt
        finally:
            paddle.set_default_dtype("float32")
            paddle.set_flags({"FLAGS_b200_split_master_weights": True})

    n1, o1 = run(True)
    n0, o0 = run(False)
    slab = o1._arena.all_slabs()[0]
    assert slab.master.dtype == torch.int16
    for a, b in zip(n1.parameters(), n0.parameters()):
        assert torch.equal(a.as_subclass(torch.Tensor).float(), b.as_subclass(torch.Tensor).float())
    sd1, sd0 = o1.state_dict(), o0.state_dict()
    for m1, m0 in zip(sd1["master_weights"].values(), sd0["master_weights"].values()):
        assert m1.dtype == paddle.float32
        assert np.abs(np.asarray(m1.numpy()).view(np.int32) - np.asarray(m0.numpy()).view(np.int32)).max() <= 1
    # fp32 master weights from the checkpoint restore exactly into the split format
    masters = [np.asarray(v.numpy()) for v in sd1["master_weights"].values()]
    o1.set_state_dict(sd1)
    sd2 = o1.state_dict()
    for a, b in zip(masters, sd2["master_weights"].values()):
        assert np.abs(a.view(np.int32) - np.asarray(b.numpy()).view(np.int32)).max() <= 1
