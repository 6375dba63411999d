"""Generated (paddle_b200.cinn) kernels on the device vs an fp32 PyTorch reference.

This file sorts last on purpose and its tests are non-strict xfail: the generated CUDA was cross-compiled for sm_100a and its bodies were
checked through the host target (tests/test_cinn_cpu.py), but the launch path (ctypes launcher on the current stream, vector variants,
shuffle / shared-memory reductions, register row cache) had no hardware run when it was written.  A pass shows up as XPASS."""
import pytest
import torch

import paddle_b200 as paddle
from paddle_b200 import cinn, static

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first hardware run of the generated-kernel launch path")]
F = paddle.nn.functional


def _fused(build, feeds):
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            ph = {k: static.data(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in feeds.items()}
            fetch = build(**ph)
            fetch = list(fetch) if isinstance(fetch, (list, tuple)) else [fetch]
        new, rep = cinn.compile_program(main, fetch)
        out = static.Executor().run(new, feed=feeds, fetch_list=fetch, return_numpy=False)
    finally:
        paddle.disable_static()
    torch.cuda.synchronize()
    return [o.as_subclass(torch.Tensor) for o in out], rep


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2), (torch.float16, 3e-3)])
def test_elementwise_vec4_and_scalar_variants(dtype, tol):
    torch.manual_seed(0)
    x = torch.randn(64, 12, 256, device="cuda").to(dtype)
    b = torch.randn(256, device="cuda").to(dtype)
    g = torch.randn(64, 1, 1, device="cuda").to(dtype)

    def build(x, b, g):
        return F.gelu(x * b + g) * paddle.tanh(x) - 0.5

    (out,), rep = _fused(build, dict(x=x, b=b, g=g))
    assert len(rep.groups) == 1 and rep.groups[0]["kernel"].launches == 1
    xf, bf, gf = x.float(), b.float(), g.float()
    ref = torch.nn.functional.gelu(xf * bf + gf) * torch.tanh(xf) - 0.5
    assert out.dtype == dtype and torch.allclose(out.float(), ref, rtol=tol, atol=tol)
    # an unaligned view takes the scalar variant
    x2 = torch.randn(64 * 12 * 256 + 1, device="cuda").to(dtype)[1:].view(64, 12, 256)
    (out2,), _ = _fused(build, dict(x=x2, b=b, g=g))
    ref2 = torch.nn.functional.gelu(x2.float() * bf + gf) * torch.tanh(x2.float()) - 0.5
    assert torch.allclose(out2.float(), ref2, rtol=tol, atol=tol)


@pytest.mark.parametrize("cols", [40, 256, 1000, 4096, 12000])
def test_row_kernels_every_schedule(cols):
    torch.manual_seed(1)
    x = torch.randn(37, cols, device="cuda")
    w = torch.randn(cols, device="cuda")

    def build(x, w):
        mu = x.mean(-1, keepdim=True)
        xc = x - mu
        y = xc * paddle.rsqrt((xc * xc).mean(-1, keepdim=True) + 1e-5) * w
        return F.softmax(y, -1), y.amax(-1)

    (p, m), rep = _fused(build, dict(x=x, w=w))
    assert len(rep.groups) == 1 and rep.groups[0]["kind"] == "reduce"
    xc = x - x.mean(-1, keepdim=True)
    y = xc * torch.rsqrt((xc * xc).mean(-1, keepdim=True) + 1e-5) * w
    assert torch.allclose(p, torch.softmax(y, -1), rtol=1e-4, atol=1e-6)
    assert torch.allclose(m, y.amax(-1), rtol=1e-4, atol=1e-5)


def test_to_static_backend_cinn_on_device():
    class Head(paddle.nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc = paddle.nn.Linear(256, 512)

        def forward(self, x):
            h = self.fc(x)
            return F.softmax(F.silu(h) * 1.3 - h.mean(-1, keepdim=True), -1)

    paddle.seed(0)
    paddle.set_device("gpu:0")                         # parameters and randn land on the device
    try:
        net = Head()
        net.eval()
        x = paddle.randn([64, 256])
        fast = paddle.jit.to_static(net, backend="CINN")
        with paddle.no_grad():
            ref = net(x)
            fast(x)                                    # first call: trace, compile, verify against the eager forward
            out = fast(x)                              # second call: the compiled program
        rep = fast.forward.cinn_report(x)
    finally:
        paddle.set_device("cpu")
    assert x.is_cuda
    assert rep is not None and len(rep.groups) >= 1
    assert torch.allclose(out.as_subclass(torch.Tensor), ref.as_subclass(torch.Tensor), rtol=1e-4, atol=1e-6)


def test_generated_backward_on_device():
    torch.manual_seed(2)
    x = torch.randn(64, 512, device="cuda", requires_grad=True)
    w = torch.randn(512, device="cuda", requires_grad=True)
    b = torch.randn(512, device="cuda", requires_grad=True)
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            xv, wv, bv = static.data("x", [64, 512], "float32"), static.data("w", [512], "float32"), static.data("b", [512], "float32")
            mu = xv.mean(-1, keepdim=True)
            xc = xv - mu
            out = F.gelu(xc * paddle.rsqrt((xc * xc).mean(-1, keepdim=True) + 1e-5) * wv + bv)
        _, rep = cinn.compile_program(main, [out])
    finally:
        paddle.disable_static()
    k = rep.groups[0]["kernel"]
    o = k(x, w, b).as_subclass(torch.Tensor)
    go = torch.randn_like(o)
    got = torch.autograd.grad(o, [x, w, b], go)
    xr, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    ref_o = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xr, (512,), wr, br, 1e-5))
    ref = torch.autograd.grad(ref_o, [xr, wr, br], go)
    assert torch.allclose(o, ref_o, rtol=1e-4, atol=1e-5)
    for a, r in zip(got, ref):
        assert torch.allclose(a, r, rtol=1e-3, atol=1e-4)
    (bk, plan), = k._bwd.values()
    assert plan is not None and bk.launches == 1


@pytest.mark.parametrize("shape", [(64, 128, 256), (4096, 48), (3, 2000, 40)])
def test_column_reduction_schedule(shape):
    torch.manual_seed(3)
    dy = torch.randn(*shape, device="cuda")
    x = torch.randn(*shape, device="cuda")
    lead = list(range(len(shape) - 1))

    def build(dy, x):
        t = dy * paddle.tanh(x)
        return t.sum(axis=lead), (dy * x).mean(axis=lead[-1:])

    (s, m), rep = _fused(build, dict(dy=dy, x=x))
    assert any(g["kind"] == "column" for g in rep.groups)
    assert torch.allclose(s, (dy * torch.tanh(x)).sum(lead), rtol=1e-4, atol=1e-3)
    assert torch.allclose(m, (dy * x).mean(lead[-1]), rtol=1e-4, atol=1e-5)


def test_batched_detection_ops_on_device():
    """vision.ops after their rewrite as batched tensor programs: device results equal the CPU results (no per-box host synchronisation inside)."""
    import numpy as np

    from paddle_b200.vision import ops

    rng = np.random.RandomState(0)
    feat = rng.randn(2, 8, 24, 30).astype("float32")
    k = 64
    xy = rng.rand(k, 2) * np.array([100, 80]) - 5
    boxes = np.concatenate([xy, xy + rng.rand(k, 2) * np.array([60, 50]) + 1], 1).astype("float32")
    nums = np.array([40, 24], "int32")
    sc = rng.rand(k).astype("float32")

    def run(dev):
        t = lambda a: paddle.to_tensor(a).to(dev)       # noqa: E731
        ra = ops.roi_align(t(feat), t(boxes), t(nums), 7, 0.25)
        rp = ops.roi_pool(t(feat), t(boxes), t(nums), 3, 0.25)
        keep = ops.nms(t(boxes), 0.4, t(sc))
        off = (rng.randn(2, 18, 24, 30) * 0.7).astype("float32")
        dc = ops.deform_conv2d(t(feat), t(off), t(rng.randn(4, 8, 3, 3).astype("float32")), None, 1, 1)
        return [v.cpu().numpy() for v in (ra, rp, keep, dc)]

    rng_state = rng.get_state()
    cpu = run("cpu")
    rng.set_state(rng_state)
    gpu = run("gpu:0")
    for a, b in zip(cpu, gpu):
        np.testing.assert_allclose(b, a, rtol=1e-4, atol=1e-4)
