"""Framework-level behaviour on a CUDA device (placement, AMP, DataLoader pinning, jit CUDA-graph replay)."""
import numpy as np
import pytest
import torch

import paddle_b200 as paddle
from paddle_b200 import nn

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu_default():
    paddle.set_device("gpu:0")
    yield
    paddle.set_device("cpu")


def test_parameters_and_buffers_follow_device():
    net = nn.Sequential(nn.Conv2D(3, 4, 3), nn.BatchNorm2D(4), nn.ReLU())
    assert all(p.is_cuda for p in net.parameters()) and all(b.is_cuda for b in net.buffers())
    x = paddle.randn([2, 3, 8, 8])
    assert x.is_cuda and net(x).is_cuda and "gpu" in str(x.place).lower()


def test_resnet_amp_step_on_gpu():
    paddle.seed(0)
    net = paddle.vision.models.resnet18(num_classes=10)
    opt = paddle.optimizer.Momentum(0.05, parameters=net.parameters(), multi_precision=True)
    x, y = paddle.randn([4, 3, 64, 64]), paddle.randint(0, 10, [4])
    losses = []
    for _ in range(4):
        with paddle.amp.auto_cast(level="O1", dtype="bfloat16"):
            loss = paddle.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        opt.step()
        opt.clear_grad()
        losses.append(float(loss))
    assert losses[-1] < losses[0]


def test_to_static_cuda_graph_replay_matches_eager():
    paddle.seed(1)
    net = nn.Sequential(nn.Linear(64, 128), nn.GELU(), nn.Linear(128, 16))
    snet = paddle.jit.to_static(net)
    x = paddle.randn([8, 64])
    ref = net(x).numpy()
    for _ in range(4):      # eager warm-ups, capture, replays
        out = snet(x)
    np.testing.assert_allclose(out.numpy(), ref, rtol=1e-4, atol=1e-5)


def test_dataloader_pinned_to_device():
    from paddle_b200.io import DataLoader, TensorDataset

    ds = TensorDataset([paddle.to_tensor(np.arange(32, dtype="float32").reshape(16, 2)).cpu(), paddle.to_tensor(np.arange(16)).cpu()])
    for xb, yb in DataLoader(ds, batch_size=4, places=paddle.CUDAPlace(0) if hasattr(paddle, "CUDAPlace") else None):
        assert xb.shape == [4, 2]
        break
