"""Model zoo smoke + training-signal tests on CPU (tiny configs). Parity: test/legacy_test/test_imperative_*model*.py."""
import numpy as np
import pytest

import paddle_b200 as paddle
from paddle_b200 import models


def _train(model, vocab, steps=8, lr=3e-3, seq=32):
    paddle.seed(0)
    opt = paddle.optimizer.AdamW(lr, parameters=model.parameters())
    ids = paddle.to_tensor(np.random.RandomState(0).randint(0, vocab, (4, seq + 1)))
    losses = []
    for _ in range(steps):
        loss = model(ids[:, :-1], ids[:, 1:])
        loss.backward()
        opt.step()
        opt.clear_grad()
        losses.append(float(loss))
    return losses


def test_llama_tiny_trains():
    paddle.seed(1)
    m = models.LlamaForCausalLM(models.llama_tiny())
    l = _train(m, 512)
    assert l[-1] < l[0] - 0.3 and abs(l[0] - np.log(512)) < 0.5


def test_gpt_tiny_trains_and_generates_logits():
    paddle.seed(2)
    m = models.GPTForCausalLM(models.gpt_tiny())
    l = _train(m, 512)
    assert l[-1] < l[0] - 0.3
    logits = m(paddle.to_tensor(np.zeros((1, 5), "int64")))
    assert logits.shape == [1, 5, 512]


def test_mixtral_tiny_moe_trains():
    paddle.seed(3)
    m = models.MixtralForCausalLM(models.mixtral_tiny())
    l = _train(m, 512)
    assert l[-1] < l[0] - 0.2
    moe = m.layers[0].moe
    assert moe.num_expert == 4 and moe.top_k == 2
    # every expert weight received gradient signal over a few steps (routing spreads tokens)
    loss = m(paddle.to_tensor(np.random.randint(0, 512, (4, 33)))[:, :-1], paddle.to_tensor(np.random.randint(0, 512, (4, 33)))[:, 1:])
    loss.backward()
    g = moe.experts.w1.grad
    assert (g.abs().sum([1, 2]) > 0).numpy().sum() >= 3


def test_mnist_mlp_pdparams_roundtrip(tmp_path):
    """BASELINE config 1: MNIST MLP on CPU, .pdparams save/load."""
    paddle.seed(4)
    m = models.MnistMLP(hidden=(64,))
    x = paddle.to_tensor(np.random.rand(16, 1, 28, 28).astype("float32"))
    y = paddle.to_tensor(np.random.randint(0, 10, (16,)))
    opt = paddle.optimizer.Adam(1e-2, parameters=m.parameters())
    first = None
    for _ in range(30):
        loss = paddle.nn.functional.cross_entropy(m(x), y)
        first = first if first is not None else float(loss)
        loss.backward()
        opt.step()
        opt.clear_grad()
    assert float(loss) < first * 0.5
    paddle.save(m.state_dict(), str(tmp_path / "mlp.pdparams"))
    paddle.save(opt.state_dict(), str(tmp_path / "mlp.pdopt"))
    m2 = models.MnistMLP(hidden=(64,))
    m2.set_state_dict(paddle.load(str(tmp_path / "mlp.pdparams")))
    np.testing.assert_allclose(m2(x).numpy(), m(x).numpy(), rtol=1e-6)
