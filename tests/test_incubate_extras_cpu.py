"""incubate.multiprocessing (tensors through shared memory), incubate.checkpoint.auto_checkpoint, incubate.layers, incubate.jit.inference."""
import os

import numpy as np

import paddle_b200 as paddle


def _child(q_in, q_out):
    import paddle_b200 as paddle  # noqa: F811

    t = q_in.get()
    assert isinstance(t, paddle.Tensor)
    t.as_subclass(__import__("torch").Tensor).add_(1.0)      # visible to the parent: same shared-memory storage
    q_out.put(float(t.sum()))


def test_multiprocessing_shares_tensor_storage():
    import paddle_b200.incubate.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q_in, q_out = ctx.Queue(), ctx.Queue()
    x = paddle.zeros([4])
    p = ctx.Process(target=_child, args=(q_in, q_out))
    p.start()
    q_in.put(x)
    assert q_out.get(timeout=120) == 4.0
    p.join(60)
    assert p.exitcode == 0 and np.allclose(x.numpy(), 1.0)      # the child's in-place update is visible here


def test_auto_checkpoint_resumes(tmp_path, monkeypatch):
    from paddle_b200.incubate import checkpoint as acp

    monkeypatch.setenv("PADDLE_EDL_FS_CHECKPOINT", str(tmp_path))
    monkeypatch.setenv("PADDLE_JOB_ID", "t1")
    net = paddle.nn.Linear(2, 2)
    opt = paddle.optimizer.SGD(0.1, parameters=net.parameters())
    acp.register(net=net, opt=opt)
    seen = []
    for epoch in acp.train_epoch_range(5, name="run"):
        seen.append(epoch)
        net.weight.set_value(paddle.full([2, 2], float(epoch)))
        if epoch == 2:
            break                                   # "crash" during epoch 2: epochs 0 and 1 were checkpointed
    assert seen == [0, 1, 2]
    net2 = paddle.nn.Linear(2, 2)
    acp.register(net=net2, opt=paddle.optimizer.SGD(0.1, parameters=net2.parameters()))
    resumed = list(acp.train_epoch_range(5, name="run"))
    assert resumed == [2, 3, 4] and np.allclose(net2.weight.numpy(), 1.0)     # restored to the state after epoch 1


def test_incubate_layers():
    L = paddle.incubate.layers
    x = paddle.to_tensor(np.arange(12, dtype="float32").reshape(4, 3))
    s = L.shuffle_batch(x, seed=3)
    assert sorted(map(tuple, s.numpy().tolist())) == sorted(map(tuple, x.numpy().tolist())) and not np.allclose(s.numpy(), x.numpy())
    a, b = x, x * 10
    np.testing.assert_allclose(L.partial_concat([a, b], start_index=1, length=2).numpy(), np.concatenate([a.numpy()[:, 1:3], b.numpy()[:, 1:3]], 1))
    np.testing.assert_allclose(L.partial_sum([a, b], start_index=0, length=2).numpy(), (a.numpy() + b.numpy())[:, :2])
    out = L.batch_fc(paddle.ones([2, 3, 4]), [2, 4, 5], None, [2, 5], None, act="relu")
    assert out.shape == [2, 3, 5]
    sched = L.pow2_decay_with_linear_warmup(2, 10, 1.0, 0.1)
    vals = []
    for _ in range(12):
        vals.append(sched())
        sched.step()
    assert abs(vals[0] - 0.5) < 1e-6 and abs(vals[1] - 1.0) < 1e-6 and vals[5] < vals[2] and abs(vals[-1] - 0.1) < 1e-6
    ids = paddle.base.create_lod_tensor(np.array([[1], [2], [3], [1]]), [[3, 1]])
    pooled = L.fused_embedding_seq_pool(ids, [5, 4])
    assert pooled.shape == [2, 4]
    y = L.fused_bn_add_act(paddle.randn([2, 4, 4, 3]), paddle.zeros([2, 4, 4, 3]))
    assert y.shape == [2, 4, 4, 3] and float(y.min()) >= 0


def test_incubate_jit_inference_decorator():
    from paddle_b200.incubate.jit import inference

    net = paddle.nn.Sequential(paddle.nn.Linear(3, 3), paddle.nn.Dropout(0.5))
    ref = net[0]
    served = inference(net)
    x = paddle.ones([2, 3])
    out = served(x)
    assert out.stop_gradient and np.allclose(out.numpy(), ref(x).numpy())     # eval mode: dropout is the identity, no graph recorded

    @inference(precision_mode="float32")
    def f(a, b):
        return a * 2 + b

    assert np.allclose(f(x, x).numpy(), 3.0) and f._inference_options["precision_mode"] == "float32"
