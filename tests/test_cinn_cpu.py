"""paddle_b200.cinn: fusion groups, generated kernels (run through the host target here; the CUDA source of every kernel must also generate,
and a sample is cross-compiled with nvcc for sm_100a), numerics vs the unfused program."""
import shutil

import numpy as np
import pytest
import torch

import paddle_b200 as paddle
from paddle_b200 import cinn, static

pytestmark = pytest.mark.skipif(not cinn.is_available(), reason="native IR core not built")
F = paddle.nn.functional


def _run_both(build, feeds, rtol=1e-5, atol=1e-6, expect_kernels=None, expect_fused=None):
    """build(**placeholders) -> fetch list.  Returns the FusionResult; asserts fused == unfused."""
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            ph = {k: static.data(k, list(v.shape), str(v.dtype)) for k, v in feeds.items()}
            fetch = build(**ph)
            fetch = list(fetch) if isinstance(fetch, (list, tuple)) else [fetch]
        new, rep = cinn.compile_program(main, fetch)
        exe = static.Executor()
        a = exe.run(main, feed=feeds, fetch_list=fetch)
        b = exe.run(new, feed=feeds, fetch_list=fetch)
    finally:
        paddle.disable_static()
    for x, y in zip(a, b):
        assert x.shape == y.shape and x.dtype == y.dtype
        np.testing.assert_allclose(np.asarray(y, dtype=np.float64), np.asarray(x, dtype=np.float64), rtol=rtol, atol=atol)
    if expect_kernels is not None:
        assert len(rep.groups) == expect_kernels, (rep, rep.rejected)
    if expect_fused is not None:
        assert sum(len(g["ops"]) for g in rep.groups) == expect_fused, [g["ops"] for g in rep.groups]
    for g in rep.groups:
        assert "__global__" in g["kernel"].source("cuda")
    return rep


def test_elementwise_chain_with_broadcasts_is_one_kernel():
    rng = np.random.default_rng(0)
    x, b, c = rng.standard_normal((6, 5, 16), dtype=np.float32), rng.standard_normal((16,), dtype=np.float32), rng.standard_normal((6, 1, 1), dtype=np.float32)
    r = rng.standard_normal((5, 1), dtype=np.float32)

    def build(x, b, c, r):
        y = paddle.tanh(x * b + c) * 0.5 - r
        y = paddle.maximum(y, paddle.abs(x)) / (1.0 + paddle.exp(-y))
        return paddle.where(y > 0.25, y, 2.0 - y * y)

    rep = _run_both(build, dict(x=x, b=b, c=c, r=r), expect_kernels=1)
    g = rep.groups[0]
    assert g["kind"] == "elementwise" and g["inputs"] == 4 and g["outputs"] == 1 and len(g["ops"]) >= 10
    assert "_vec4" in g["kernel"].source("cuda")                      # 16 columns: the 4-wide variant exists


def test_softmax_and_normalisation_chains_become_row_kernels():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 7, 40), dtype=np.float32)
    w, b = rng.standard_normal((40,), dtype=np.float32), rng.standard_normal((40,), dtype=np.float32)

    def layer_norm_by_hand(x, w, b):
        mu = x.mean(-1, keepdim=True)
        xc = x - mu
        var = (xc * xc).mean(-1, keepdim=True)
        return xc * paddle.rsqrt(var + 1e-5) * w + b

    rep = _run_both(layer_norm_by_hand, dict(x=x, w=w, b=b), rtol=2e-5, atol=2e-6, expect_kernels=1)
    assert rep.groups[0]["kind"] == "reduce"
    src = rep.groups[0]["kernel"].source("cuda")
    assert "cinn_warp_reduce" in src and "c" in src                  # 40 columns: one warp per row, centred input kept in registers

    def rms_then_softmax(x, w):
        ms = (x * x).mean(-1, keepdim=True)
        y = x * paddle.rsqrt(ms + 1e-6) * w
        return F.softmax(y * 0.7, -1), F.log_softmax(y, -1)

    rep = _run_both(rms_then_softmax, dict(x=x, w=w), rtol=2e-5, atol=2e-6, expect_kernels=1)
    assert rep.groups[0]["outputs"] == 2


def test_row_results_and_nonkeepdim_reductions():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((9, 33), dtype=np.float32)

    def build(x):
        e = paddle.exp(x - x.amax(-1, keepdim=True))
        s = e.sum(-1)                                   # [9]: a per-row result that leaves the group
        lse = paddle.log(s) + x.amax(-1)
        return e / e.sum(-1, keepdim=True), lse, (x * x).sum(-1, keepdim=True)

    rep = _run_both(build, dict(x=x), rtol=2e-5, atol=2e-6)
    assert sum(g["kind"] == "reduce" for g in rep.groups) >= 1
    assert sum(len(g["ops"]) for g in rep.groups) >= 8


def test_long_rows_use_the_block_schedule():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 3000), dtype=np.float32)
    rep = _run_both(lambda x: F.softmax(x * 1.5 + 0.1, -1), dict(x=x), rtol=2e-5, atol=1e-7, expect_kernels=1)
    assert "cinn_block_reduce" in rep.groups[0]["kernel"].source("cuda")
    y = rng.standard_normal((2, 9000), dtype=np.float32)
    rep = _run_both(lambda y: F.softmax(y * 1.5 + 0.1, -1), dict(y=y), rtol=2e-5, atol=1e-7, expect_kernels=1)
    assert "_Pragma" not in rep.groups[0]["kernel"].source("cuda").split("__global__")[-1]       # too long for the register cache: plain column loops


@pytest.mark.parametrize("dtype", ["bfloat16", "float16", "float64"])
def test_other_float_types(dtype):
    rng = np.random.default_rng(4)
    td = getattr(torch, dtype)
    x = torch.from_numpy(rng.standard_normal((8, 24))).to(td)
    w = torch.from_numpy(rng.standard_normal((24,))).to(td)
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            xv, wv = static.data("x", [8, 24], dtype), static.data("w", [24], dtype)
            y = F.gelu(xv * wv) + F.silu(xv)
            out = F.softmax(y, -1) * wv
        new, rep = cinn.compile_program(main, [out])
        exe = static.Executor()
        b = exe.run(new, feed={"x": x, "w": w}, fetch_list=[out], return_numpy=False)[0].as_subclass(torch.Tensor)
    finally:
        paddle.disable_static()
    assert len(rep.groups) == 1 and b.dtype == td
    xf, wf = x.double(), w.double()
    ref = torch.softmax(torch.nn.functional.gelu(xf * wf) + torch.nn.functional.silu(xf), -1) * wf
    tol = {"bfloat16": 2e-2, "float16": 3e-3, "float64": 1e-12}[dtype]
    assert torch.allclose(b.double(), ref, rtol=tol, atol=tol)        # ONE rounding at the store: closer to fp64 than the per-op rounded eager chain


def test_casts_compares_and_integer_inputs():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4, 12), dtype=np.float32)
    k = rng.integers(-3, 4, size=(4, 12)).astype(np.int64)
    m = rng.integers(0, 2, size=(12,)).astype(bool)

    def build(x, k, m):
        y = x * k.astype("float32") + 1.0
        z = paddle.where(m, y, -y)
        return z.astype("float16"), (z >= 0.5), paddle.clip(z, -1.0, 1.0)

    rep = _run_both(build, dict(x=x, k=k, m=m), rtol=1e-3, atol=1e-3)
    assert len(rep.groups) == 1 and rep.groups[0]["outputs"] == 3


def test_groups_are_cut_at_ops_that_are_not_generated():
    rng = np.random.default_rng(6)
    x, w = rng.standard_normal((8, 16), dtype=np.float32), rng.standard_normal((16, 16), dtype=np.float32)

    def build(x, w):
        a = paddle.exp(x * 0.1) + 1.0                      # group 1
        h = paddle.matmul(a, w)                            # stays a GEMM
        g = F.relu(h) * 2.0 - a                            # group 2 reads the GEMM and group 1's value
        return g, a

    rep = _run_both(build, dict(x=x, w=w), rtol=1e-5, atol=1e-5, expect_kernels=2)
    assert all("matmul" not in g["ops"] for g in rep.groups)


def test_a_single_op_is_left_alone_and_inplace_spellings_are_skipped():
    x = np.random.default_rng(7).standard_normal((4, 4)).astype(np.float32)
    rep = _run_both(lambda x: paddle.exp(x), dict(x=x), expect_kernels=0)
    assert rep.groups == []


def test_to_static_backend_cinn():
    class Head(paddle.nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc = paddle.nn.Linear(32, 48)

        def forward(self, x):
            h = self.fc(x)
            h = F.gelu(h) * 1.702 + paddle.tanh(h)
            return F.softmax(h - h.mean(-1, keepdim=True), -1)

    paddle.seed(0)
    net = Head()
    net.eval()
    x = paddle.randn([5, 32])
    with paddle.no_grad():
        ref = net(x)
    fast = paddle.jit.to_static(net, backend="CINN")
    with paddle.no_grad():
        out = fast(x)
        out2 = fast(x)
    rep = fast.forward.cinn_report(x)
    assert rep is not None and len(rep.groups) == 1 and rep.groups[0]["kind"] == "reduce"
    assert rep.groups[0]["kernel"].launches == 2
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(out2.numpy(), ref.numpy(), rtol=2e-5, atol=1e-6)
    y = paddle.randn([3, 32])                                      # another signature: compiled separately
    with paddle.no_grad():
        np.testing.assert_allclose(fast(y).numpy(), net(y).numpy(), rtol=2e-5, atol=1e-6)


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_generated_cuda_cross_compiles_for_sm100a():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((4, 64), dtype=np.float32).astype(np.float32)
    h = torch.from_numpy(x).to(torch.bfloat16)

    rep = _run_both(lambda x: F.softmax(paddle.tanh(x) * 3.0, -1), dict(x=x), rtol=2e-5, atol=1e-7, expect_kernels=1)
    so = cinn.nvcc_check(rep.groups[0]["kernel"])
    assert so.endswith(".so")
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            xv = static.data("x", [4, 64], "bfloat16")
            out = F.silu(xv) * xv + 1.0
        new, rep = cinn.compile_program(main, [out])
    finally:
        paddle.disable_static()
    assert len(rep.groups) == 1
    assert cinn.nvcc_check(rep.groups[0]["kernel"]).endswith(".so")
    del h
    # a generated backward (reductions at three levels) builds too
    xs, ws = torch.randn(6, 300), torch.randn(300)
    k = _kernel_of(lambda x, w: F.softmax(x * w, -1) * paddle.rsqrt((x * x).mean(-1, keepdim=True) + 1e-6), dict(x=xs, w=ws))
    a = [xs.clone().requires_grad_(True), ws.clone().requires_grad_(True)]
    k(*a).as_subclass(torch.Tensor).sum().backward()
    (bk, plan), = k._bwd.values()
    assert plan is not None and cinn.nvcc_check(bk).endswith(".so")


# ---- random programs: whatever the grouping decides, the fused program must equal the unfused one -----------------------------------------
def _bshape(rng, full):
    nd = int(rng.integers(0, len(full) + 1))
    sh = list(full[len(full) - nd:])
    return [1 if rng.random() < 0.4 else d for d in sh]


_UN = [paddle.tanh, paddle.exp, paddle.abs, F.sigmoid, F.relu, paddle.sin, lambda v: v * v, lambda v: -v, lambda v: 1.0 - v, lambda v: 2.0 / (paddle.abs(v) + 1.0), F.silu, F.gelu]
_BI = [lambda a, b: a + b, lambda a, b: a - b, lambda a, b: a * b, paddle.maximum, paddle.minimum, lambda a, b: a / (paddle.abs(b) + 1.5),
       lambda a, b: paddle.where(a > b, a, b * 0.5)]
_RED = [lambda v: v.sum(-1, keepdim=True), lambda v: v.mean(-1, keepdim=True), lambda v: v.amax(-1, keepdim=True), lambda v: v.amin(-1, keepdim=True),
        lambda v: v.sum(-1), lambda v: v.mean(), lambda v: v.amax(), lambda v: v.sum(0), lambda v: v.mean(0, keepdim=True),
        lambda v: v.sum(tuple(range(v.dim() - 1))) if v.dim() > 1 else v.sum(), lambda v: v.sum(1) if v.dim() > 2 else v.sum(0)]


@pytest.mark.parametrize("seed", range(24))
def test_random_programs(seed):
    rng = np.random.default_rng(1000 + seed)
    full = [int(rng.integers(1, 7)) for _ in range(int(rng.integers(1, 5)))]
    full[-1] = int(rng.choice([1, 3, 4, 8, 33, 40, 300]))
    shapes = [full] + [_bshape(rng, full) for _ in range(int(rng.integers(0, 4)))]
    feeds = {f"i{k}": rng.standard_normal(s).astype(np.float32) for k, s in enumerate(shapes)}

    def build(**ph):
        ph = list(ph.values())
        vals, fulls = list(ph), [ph[0]]
        for _ in range(int(rng.integers(3, 14))):
            r = rng.random()
            try:
                if r < 0.4:
                    v = _UN[rng.integers(len(_UN))](vals[rng.integers(len(vals))])
                elif r < 0.85:
                    a = fulls[rng.integers(len(fulls))] if rng.random() < 0.7 else vals[rng.integers(len(vals))]
                    v = _BI[rng.integers(len(_BI))](a, vals[rng.integers(len(vals))])
                else:
                    v = _RED[rng.integers(len(_RED))](fulls[rng.integers(len(fulls))])
            except RuntimeError:                      # shapes that do not broadcast (a non-keepdim row value against the domain)
                continue
            vals.append(v)
            if list(v.shape) == full:
                fulls.append(v)
        return [vals[-1]] + [vals[int(rng.integers(len(ph), len(vals)))] for _ in range(int(rng.integers(0, 2)))]

    _run_both(build, feeds, rtol=2e-4, atol=2e-5)


# ---- backward: generated from the forward group (cinn/autodiff.py) ---------------------------------------------------------------------------
def _kernel_of(build, feeds):
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            ph = {k: static.data(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in feeds.items()}
            fetch = build(**ph)
            fetch = list(fetch) if isinstance(fetch, (list, tuple)) else [fetch]
        _, rep = cinn.compile_program(main, fetch)
    finally:
        paddle.disable_static()
    assert len(rep.groups) == 1, (rep, rep.rejected)
    return rep.groups[0]["kernel"]


def _check_grads(kernel, torch_fn, inputs, rtol=1e-4, atol=1e-5):
    a = [t.clone().requires_grad_(True) for t in inputs]
    b = [t.clone().requires_grad_(True) for t in inputs]
    oa = kernel(*a)
    oa = [o.as_subclass(torch.Tensor) for o in (oa if isinstance(oa, tuple) else (oa,))]
    ob = torch_fn(*b)
    ob = list(ob) if isinstance(ob, (tuple, list)) else [ob]
    torch.manual_seed(0)
    gos = [torch.randn_like(o) for o in ob]
    ga = torch.autograd.grad(oa, a, gos, allow_unused=True)
    gb = torch.autograd.grad(ob, b, gos, allow_unused=True)
    for x, y, t in zip(ga, gb, inputs):
        assert x is not None and x.shape == t.shape and x.dtype == t.dtype
        assert torch.allclose(x, y, rtol=rtol, atol=atol), (x - y).abs().max()
    (bk, plan), = kernel._bwd.values()
    assert plan is not None, getattr(kernel, "backward_fallback", None)          # generated, not the interpreter fall-back
    return bk, plan


def test_backward_of_a_normalisation_is_one_generated_kernel():
    torch.manual_seed(0)
    x, w, b = torch.randn(5, 6, 48), torch.randn(48), torch.randn(48)

    def build(x, w, b):
        mu = x.mean(-1, keepdim=True)
        xc = x - mu
        return xc * paddle.rsqrt((xc * xc).mean(-1, keepdim=True) + 1e-5) * w + b

    k = _kernel_of(build, dict(x=x, w=w, b=b))
    bk, plan = _check_grads(k, lambda x, w, b: torch.nn.functional.layer_norm(x, (48,), w, b, 1e-5), [x, w, b])
    assert bk.spec.has_reduce and bk.spec.max_level >= 2                # the row sums of the backward are inside the kernel
    assert set(plan.parts) == {0, 1, 2}
    assert "__global__" in bk.source("cuda")


def test_backward_softmax_ties_and_broadcast_inputs():
    torch.manual_seed(1)
    x, s = torch.randn(7, 33), torch.randn(7, 1)
    k = _kernel_of(lambda x, s: F.softmax(x * s, -1) * F.log_softmax(x, -1), dict(x=x, s=s))
    _check_grads(k, lambda x, s: torch.softmax(x * s, -1) * torch.log_softmax(x, -1), [x, s])
    # amax with tied maxima: the gradient is shared, like the eager op
    t = torch.tensor([[1.0, 3.0, 3.0, 0.0], [2.0, 2.0, 2.0, 2.0]])
    k = _kernel_of(lambda t: t.amax(-1, keepdim=True) * 2.0 + paddle.maximum(t, 2.0 - t).sum(-1, keepdim=True), dict(t=t))
    _check_grads(k, lambda t: t.amax(-1, keepdim=True) * 2.0 + torch.maximum(t, 2.0 - t).sum(-1, keepdim=True), [t])
    # an input broadcast over leading axes gets its gradient summed back to its shape
    a, c = torch.randn(3, 4, 8), torch.randn(4, 1)
    k = _kernel_of(lambda a, c: paddle.tanh(a * c) / (1.0 + paddle.exp(-a)) - c, dict(a=a, c=c))
    _check_grads(k, lambda a, c: torch.tanh(a * c) / (1.0 + torch.exp(-a)) - c, [a, c])


def test_backward_half_precision_and_missing_rule_fallback():
    torch.manual_seed(2)
    x = torch.randn(16, 64).to(torch.bfloat16)
    k = _kernel_of(lambda x: F.gelu(x) * paddle.tanh(x) + x, dict(x=x))
    xa = x.clone().requires_grad_(True)
    out = k(xa).as_subclass(torch.Tensor)
    out.float().sum().backward()
    xr = x.float().requires_grad_(True)
    (torch.nn.functional.gelu(xr) * torch.tanh(xr) + xr).sum().backward()
    assert xa.grad.dtype == torch.bfloat16 and torch.allclose(xa.grad.float(), xr.grad, rtol=2e-2, atol=2e-2)
    # fmod by a tensor has no rule: the gradient comes from differentiating the reference evaluation
    y, d = torch.rand(4, 8) * 5, torch.rand(4, 8) + 1.0
    k = _kernel_of(lambda y, d: paddle.exp(torch.fmod(y, d) * 0.1), dict(y=y, d=d))
    ya, da = y.clone().requires_grad_(True), d.clone().requires_grad_(True)
    k(ya, da).as_subclass(torch.Tensor).sum().backward()
    yr, dr = y.clone().requires_grad_(True), d.clone().requires_grad_(True)
    torch.exp(torch.fmod(yr, dr) * 0.1).sum().backward()
    assert torch.allclose(ya.grad, yr.grad, rtol=1e-4, atol=1e-6) and torch.allclose(da.grad, dr.grad, rtol=1e-4, atol=1e-5)
    assert "fmod" in k.backward_fallback


def test_to_static_cinn_trains():
    class Net(paddle.nn.Layer):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2 = paddle.nn.Linear(16, 32), paddle.nn.Linear(32, 4)
            self.g = self.create_parameter([32], default_initializer=paddle.nn.initializer.Constant(1.0))

        def forward(self, x):
            h = self.fc1(x)
            h = h * paddle.rsqrt((h * h).mean(-1, keepdim=True) + 1e-6) * self.g          # RMSNorm written out: one generated kernel
            h = F.silu(h) + 0.1 * paddle.tanh(h)
            return F.log_softmax(self.fc2(h), -1)

    def run(compiled):
        paddle.seed(3)
        net = Net()
        opt = paddle.optimizer.SGD(learning_rate=0.1, parameters=net.parameters())
        fwd = paddle.jit.to_static(net, backend="CINN") if compiled else net
        x = paddle.to_tensor(np.random.default_rng(0).standard_normal((8, 16)).astype("float32"))
        y = paddle.to_tensor(np.arange(8) % 4)
        losses = []
        for _ in range(5):
            lp = fwd(x)
            loss = -(lp * F.one_hot(y, 4)).sum(-1).mean()
            loss.backward()
            opt.step()
            opt.clear_grad()
            losses.append(float(loss))
        return losses, (fwd.forward.cinn_report(x) if compiled else None)

    ref, _ = run(False)
    got, rep = run(True)
    assert rep is not None and len(rep.groups) >= 1 and sum(g["kernel"].launches for g in rep.groups) >= 5
    assert any(getattr(g["kernel"], "_bwd", None) for g in rep.groups)              # the backward went through generated kernels
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)
    assert got[-1] < got[0]


@pytest.mark.parametrize("seed", range(10))
def test_random_program_gradients(seed):
    """Generated backward == autograd through the reference evaluation of the same group."""
    from paddle_b200.cinn import interp

    rng = np.random.default_rng(500 + seed)
    full = [int(rng.integers(1, 6)) for _ in range(int(rng.integers(1, 4)))]
    full[-1] = int(rng.choice([1, 4, 8, 33]))
    shapes = [full] + [_bshape(rng, full) for _ in range(int(rng.integers(0, 3)))]
    feeds = {f"i{k}": torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for k, s in enumerate(shapes)}

    def build(**ph):
        ph = list(ph.values())
        vals, fulls = list(ph), [ph[0]]
        for _ in range(int(rng.integers(3, 10))):
            r = rng.random()
            try:
                if r < 0.4:
                    v = _UN[rng.integers(len(_UN))](vals[rng.integers(len(vals))])
                elif r < 0.85:
                    v = _BI[rng.integers(len(_BI))](fulls[rng.integers(len(fulls))], vals[rng.integers(len(vals))])
                else:
                    v = _RED[rng.integers(len(_RED))](fulls[rng.integers(len(fulls))])
            except RuntimeError:
                continue
            vals.append(v)
            if list(v.shape) == full:
                fulls.append(v)
        return vals[-1]

    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            ph = {k: static.data(k, list(v.shape), "float32") for k, v in feeds.items()}
            out = build(**ph)
        _, rep = cinn.compile_program(main, [out])
    finally:
        paddle.disable_static()
    for g in rep.groups:
        k = g["kernel"]
        torch.manual_seed(seed)
        ins = [(torch.randn(n.shape) > 0) if n.dtype == "bool" else torch.randn(n.shape) for n in k.spec.inputs]
        a = [t.clone().requires_grad_(t.is_floating_point()) for t in ins]
        b = [t.clone().requires_grad_(t.is_floating_point()) for t in ins]
        oa = k(*a)
        oa = [o.as_subclass(torch.Tensor) for o in (oa if isinstance(oa, tuple) else (oa,))]
        ob = interp.evaluate(k.spec, b)
        gos = [torch.randn_like(o) if o.is_floating_point() else None for o in ob]
        pa = [(o, g_) for o, g_ in zip(oa, gos) if o.requires_grad]
        pb = [(o, g_) for o, g_ in zip(ob, gos) if o.requires_grad]
        if not pa:
            continue
        ga = torch.autograd.grad([o for o, _ in pa], [t for t in a if t.requires_grad], [g_ for _, g_ in pa], allow_unused=True)
        gb = torch.autograd.grad([o for o, _ in pb], [t for t in b if t.requires_grad], [g_ for _, g_ in pb], allow_unused=True)
        assert getattr(k, "backward_fallback", None) is None
        for x, y in zip(ga, gb):
            if x is None or y is None:
                assert (x is None or float(x.abs().max()) == 0.0) and (y is None or float(y.abs().max()) == 0.0)
                continue
            assert torch.allclose(x, y, rtol=2e-3, atol=2e-4, equal_nan=True)


def test_static_executor_uses_generated_kernels_under_the_flags():
    """FLAGS_enable_pir_api + FLAGS_use_cinn: Executor.run specialises a program per feed signature (first run generic, later runs compiled)."""
    paddle.enable_static()
    old = paddle.get_flags(["FLAGS_enable_pir_api", "FLAGS_use_cinn"])
    try:
        paddle.set_flags({"FLAGS_enable_pir_api": True, "FLAGS_use_cinn": True})
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [6, 20], "float32")
            out = F.softmax(paddle.tanh(x) * 2.0 - x.mean(-1, keepdim=True), -1)
        xv = np.random.default_rng(0).standard_normal((6, 20)).astype("float32")
        before = cinn.stats["launches"]
        exe = static.Executor()
        first = exe.run(main, feed={"x": xv}, fetch_list=[out])[0]         # first run with this feed signature: generic program, then specialise
        assert cinn.stats["launches"] == before
        got = exe.run(main, feed={"x": xv}, fetch_list=[out])[0]
        assert cinn.stats["launches"] == before + 1
        np.testing.assert_allclose(got, first, rtol=2e-5, atol=1e-7)
    finally:
        paddle.set_flags(old)
        paddle.disable_static()
    t = torch.from_numpy(xv)
    np.testing.assert_allclose(got, torch.softmax(torch.tanh(t) * 2.0 - t.mean(-1, keepdim=True), -1).numpy(), rtol=2e-5, atol=1e-7)


def test_inference_config_enable_cinn(tmp_path):
    """Config.enable_cinn(): the predictor's program runs its pointwise / softmax tail as one generated kernel (reference: AnalysisConfig::EnableCINN).
    The saved program declares a dynamic batch; the executor specialises per input signature (first run generic, later runs compiled)."""
    from paddle_b200 import inference

    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [-1, 16], "float32")
            w = paddle.to_tensor(np.random.RandomState(0).randn(16, 24).astype("float32"))
            h = paddle.matmul(x, w)
            y = F.softmax(paddle.tanh(h) * 1.5 + F.sigmoid(h), -1)
        static.save_inference_model(str(tmp_path / "m"), [x], [y], static.Executor(), program=main)
    finally:
        paddle.disable_static()
    cfg0 = inference.Config(str(tmp_path / "m.pdmodel"), str(tmp_path / "m.pdiparams"))
    plain = inference.create_predictor(cfg0)
    cfg = inference.Config(str(tmp_path / "m.pdmodel"), str(tmp_path / "m.pdiparams"))
    cfg.enable_cinn()
    p = inference.create_predictor(cfg)
    before = cinn.stats["launches"]
    for n_spec, batch in enumerate((4, 3), 1):
        data = np.random.RandomState(batch).randn(batch, 16).astype("float32")
        ref = plain.run([data])[0]
        first = p.run([data])[0]                              # first run with this signature: generic program, then specialise
        launches = cinn.stats["launches"]
        got = p.run([data])[0]
        assert cinn.stats["launches"] == launches + 1 and len(p.cinn_report()) == n_spec
        np.testing.assert_allclose(np.asarray(first), np.asarray(ref), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(np.asarray(got), np.asarray(ref), rtol=2e-5, atol=1e-7)
    assert cinn.stats["launches"] == before + 2
    reps = list(p.cinn_report().values())
    assert all(len(r.groups) == 1 and "softmax" in r.groups[0]["ops"] for r in reps)
    assert reps[0].groups[0]["kernel"].source("cuda") == reps[1].groups[0]["kernel"].source("cuda")      # one compiled object for both batch sizes


def test_groups_that_differ_only_in_leading_extents_share_one_object():
    """The row count is a launch argument: softmax(x * w) over [B, S, 64] compiles once for every (B, S)."""
    w = np.random.default_rng(0).standard_normal((64,)).astype(np.float32)
    paths, srcs = set(), set()
    for shape in ((2, 3, 64), (5, 64), (7, 1, 4, 64)):
        x = np.random.default_rng(1).standard_normal(shape).astype(np.float32)
        rep = _run_both(lambda x, w: F.softmax(x * w, -1), dict(x=x, w=w), rtol=2e-5, atol=1e-7, expect_kernels=1)
        k = rep.groups[0]["kernel"]
        srcs.add(k.source("cuda"))
        from paddle_b200.cinn import runtime

        paths.add(runtime.compile_source(k.source("host"), "host"))
    assert len(srcs) == 1 and len(paths) == 1
    # an input that is broadcast over SOME leading axes bakes those extents into the index math: a different kernel, still correct
    x = np.random.default_rng(2).standard_normal((2, 3, 64)).astype(np.float32)
    g = np.random.default_rng(3).standard_normal((2, 1, 1)).astype(np.float32)
    rep = _run_both(lambda x, g: F.softmax(x * g, -1), dict(x=x, g=g), rtol=2e-5, atol=1e-7, expect_kernels=1)
    assert rep.groups[0]["kernel"].source("cuda") not in srcs


def test_to_static_cinn_refuses_a_trace_that_bakes_constants():
    """A forward that leaves the recorded tensor type computes on trace-time placeholders; the first-call check catches it and the function keeps
    running as written."""
    class Leaky(paddle.nn.Layer):
        def forward(self, x):
            raw = x.as_subclass(torch.Tensor)                       # invisible to the tracer
            return (paddle.tanh(x) * 2.0 + raw.cos().as_subclass(paddle.Tensor)) * 0.5

    net = Leaky()
    net.eval()
    f = paddle.jit.to_static(net, backend="CINN")
    for seed in (0, 1):
        x = paddle.to_tensor(np.random.default_rng(seed).standard_normal((4, 8)).astype("float32"))
        with paddle.no_grad():
            got = f(x)
        t = x.as_subclass(torch.Tensor)
        assert torch.allclose(got.as_subclass(torch.Tensor), (torch.tanh(t) * 2.0 + torch.cos(t)) * 0.5, atol=1e-6)
    assert f.forward.cinn_report(x) is None and "does not reproduce" in str(f.forward._cinn_error)


def test_activation_zoo_and_recorded_norms():
    rng = np.random.default_rng(11)
    x = rng.standard_normal((6, 24)).astype(np.float32) * 3
    w, b = rng.standard_normal((24,)).astype(np.float32), rng.standard_normal((24,)).astype(np.float32)

    def acts(x):
        s = F.relu6(x) + F.elu(x, 0.7) + F.selu(x) + F.mish(x) + F.leaky_relu(x, 0.2) + F.softplus(x) + F.hardswish(x) + F.log_sigmoid(x) + F.tanhshrink(x)
        return s + F.softsign(x) + F.hardtanh(x, -2.0, 1.5) + F.hardsigmoid(x) + F.gelu(x, approximate=True)

    rep = _run_both(acts, dict(x=x), rtol=2e-5, atol=2e-5, expect_kernels=1)
    assert rep.rejected == []

    def norms(x, w, b):
        h = F.layer_norm(x, [24], w, b, 1e-5)                 # recorded as ONE op, decomposed inside the group
        h = F.gelu(h) + x
        return F.rms_norm(h, w, 1e-6) * 0.5, F.layer_norm(h, 24)

    rep = _run_both(norms, dict(x=x, w=w, b=b), rtol=3e-5, atol=3e-5, expect_kernels=1)
    assert rep.groups[0]["kind"] == "reduce" and {"layer_norm", "rms_norm", "gelu"} <= set(rep.groups[0]["ops"])
    # a lone norm is not a group: it keeps its hand-written kernel
    rep = _run_both(lambda x, w, b: F.layer_norm(x, [24], w, b), dict(x=x, w=w, b=b), expect_kernels=0)


def test_batch_norm_inference_folds_into_the_pointwise_tail():
    """conv -> BN(eval) -> relu -> + skip -> relu: everything after the convolution is one kernel; the per-channel statistics [C] are read as [C, 1, 1]."""
    paddle.seed(0)
    conv, bn = paddle.nn.Conv2D(4, 8, 3, padding=1), paddle.nn.BatchNorm2D(8)
    rs = np.random.RandomState(3)
    bn._mean.set_value(rs.randn(8).astype("float32") * 0.3)
    bn._variance.set_value(rs.rand(8).astype("float32") + 0.5)
    bn.weight.set_value(rs.randn(8).astype("float32"))
    bn.bias.set_value(rs.randn(8).astype("float32"))
    conv.eval()
    bn.eval()
    x = np.random.default_rng(1).standard_normal((2, 4, 6, 6)).astype(np.float32)
    skip = np.random.default_rng(2).standard_normal((2, 8, 6, 6)).astype(np.float32)
    rep = _run_both(lambda x, skip: F.relu(F.relu(bn(conv(x))) + skip), dict(x=x, skip=skip), rtol=1e-5, atol=1e-5, expect_kernels=1)
    g = rep.groups[0]
    assert g["ops"] == ["batch_norm", "relu", "add", "relu"] and g["kind"] == "elementwise"
    # the view survives differentiation: gradients of the statistics come back in their own shape
    k = g["kernel"]
    ins = [torch.randn(n.shape if not n.attrs.get("view") else (8,)) for n in k.spec.inputs]
    ins = [t.abs() + 0.5 if i == 2 else t for i, t in enumerate(ins)]
    a = [t.clone().requires_grad_(True) for t in ins]
    k(*a).as_subclass(torch.Tensor).sum().backward()
    assert all(t.grad is not None and t.grad.shape == t.shape for t in a)


def test_loss_tail_reduces_to_a_scalar():
    """`(-(log_softmax(z) * onehot).sum(-1)).mean()`: everything up to the per-row partials is one kernel, the scalar is finished outside; trains."""
    rng = np.random.default_rng(12)
    z = rng.standard_normal((16, 10)).astype(np.float32)
    oh = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 16)]

    def loss(z, oh):
        lp = F.log_softmax(z * 1.3, -1)
        return (-(lp * oh).sum(-1)).mean(), (lp * lp).sum(), paddle.exp(lp).amax()

    rep = _run_both(loss, dict(z=z, oh=oh), rtol=2e-5, atol=2e-6)
    assert len(rep.groups) >= 1 and any("mean" in g["ops"] or "sum" in g["ops"] for g in rep.groups)
    k = _kernel_of(lambda z, oh: ((F.log_softmax(z, -1) * oh) * (z * 0.1 + 1.0)).mean(), dict(z=torch.from_numpy(z), oh=torch.from_numpy(oh)))
    zt = torch.from_numpy(z).requires_grad_(True)
    out = k(zt, torch.from_numpy(oh)).as_subclass(torch.Tensor)
    assert out.shape == ()
    out.backward()
    zr = torch.from_numpy(z).requires_grad_(True)
    ref = ((torch.log_softmax(zr, -1) * torch.from_numpy(oh)) * (zr * 0.1 + 1.0)).mean()
    ref.backward()
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-7) and torch.allclose(zt.grad, zr.grad, rtol=1e-4, atol=1e-7)



def test_column_reductions_bias_gradient_pattern():
    """Sums over leading axes ([A, K, B] schedule): `(dy * f(x)).sum((0, 1))` and its sibling share one kernel with their elementwise producers;
    means, keepdim, a middle axis; the backward is a plain elementwise kernel that reads the gradient under the keepdim shape."""
    rng = np.random.default_rng(21)
    dy, x = rng.standard_normal((6, 5, 16)).astype(np.float32), rng.standard_normal((6, 5, 16)).astype(np.float32)
    g = rng.standard_normal((16,)).astype(np.float32)

    def grads(dy, x, g):
        xh = paddle.tanh(x) * g
        return (dy * xh).sum(axis=[0, 1]), (xh * xh).mean(axis=[0, 1])

    rep = _run_both(grads, dict(dy=dy, x=x, g=g), rtol=2e-5, atol=2e-5, expect_kernels=1)
    k = rep.groups[0]["kernel"]
    assert rep.groups[0]["kind"] == "column" and k.spec.akb == (1, 30, 16) and len(k.spec.col) == 2
    src = k.source("cuda")
    assert "cinn_k_col(" in src and "cinn_k_colfin(" in src and "atomic" not in src
    rep = _run_both(lambda dy, x: (dy * x).mean(axis=1, keepdim=True), dict(dy=dy, x=x), rtol=2e-5, atol=2e-6, expect_kernels=1)
    assert rep.groups[0]["kernel"].spec.akb == (6, 5, 16)
    rep = _run_both(lambda dy, x: ((dy * x).sum(axis=0), dy * x + 1.0), dict(dy=dy, x=x), rtol=2e-5, atol=2e-5)
    assert any(g_["kind"] == "column" for g_ in rep.groups)
    # gradient through the column reduction
    k = _kernel_of(lambda x, dy: (paddle.tanh(x) * dy).sum(axis=[0, 1]), dict(x=torch.from_numpy(x), dy=torch.from_numpy(dy)))      # operands in recording order
    bk, plan = _check_grads(k, lambda x, dy: (torch.tanh(x) * dy).sum((0, 1)), [torch.from_numpy(x), torch.from_numpy(dy)])
    assert not bk.spec.col and not bk.spec.has_reduce
    # a row reduction and a column reduction of the same chain do not share a kernel
    rep = _run_both(lambda dy, x: ((dy * x).sum(axis=0), (dy * x).sum(axis=-1)), dict(dy=dy, x=x), rtol=2e-5, atol=2e-5)
    assert all(not (g_["kernel"].spec.col and any(n.kind == "reduce" for n in g_["kernel"].spec.nodes)) for g_ in rep.groups)


def test_sibling_reductions_sharing_an_input_join_one_kernel():
    """LayerNorm-backward shape: dgamma = (dy * xhat).sum(0) and dbeta = dy.sum(0) read dy once."""
    rng = np.random.default_rng(31)
    dy, xhat = rng.standard_normal((40, 24)).astype(np.float32), rng.standard_normal((40, 24)).astype(np.float32)
    rep = _run_both(lambda dy, xhat: ((dy * xhat).sum(axis=0), dy.sum(axis=0)), dict(dy=dy, xhat=xhat), rtol=2e-5, atol=2e-5, expect_kernels=1)
    g = rep.groups[0]
    assert g["kind"] == "column" and g["inputs"] == 2 and g["outputs"] == 2 and g["ops"] == ["mul", "sum", "sum"]
    # a sideways candidate that does not fit (row reduction next to a column group) opens its own group instead of being dropped
    rep = _run_both(lambda dy, xhat: ((dy * xhat).sum(axis=0), F.softmax(paddle.tanh(dy * 2.0), -1) * 3.0), dict(dy=dy, xhat=xhat), rtol=2e-5, atol=2e-5)
    assert rep.rejected == [] and "column" in [g_["kind"] for g_ in rep.groups]
    assert all(not (g_["kernel"].spec.col and any(n.kind == "reduce" for n in g_["kernel"].spec.nodes)) for g_ in rep.groups)


def test_missing_compiler_degrades_to_the_reference_evaluation(monkeypatch):
    from paddle_b200.cinn import runtime

    x = torch.randn(4, 8)
    k = _kernel_of(lambda x: F.softmax(paddle.tanh(x) * 2.0, -1), dict(x=x))

    def boom(src, target, keep_source=True):
        raise runtime.CompileError("g++: not found")

    monkeypatch.setattr(runtime, "compile_source", boom)
    k._fn.clear()
    with pytest.warns(UserWarning, match="could not be built"):
        out = k(x).as_subclass(torch.Tensor)
    assert torch.allclose(out, torch.softmax(torch.tanh(x) * 2.0, -1), atol=1e-6)


def test_static_training_program_through_passes_and_generated_kernels():
    """FLAGS_enable_pir_api + FLAGS_use_cinn on a program with optimizer.minimize: the training node declares the loss it reads, the forward is
    rewritten (fusions, generated kernels with generated backward) and the loss curve is the unoptimised one."""
    def run(flags):
        old = paddle.get_flags(list(flags))
        paddle.set_flags(flags)
        paddle.enable_static()
        try:
            paddle.seed(0)
            main, start = static.Program(), static.Program()
            with static.program_guard(main, start):
                x, y = static.data("x", [8, 16], "float32"), static.data("y", [8, 1], "float32")
                h = static.nn.fc(x, 32)
                h = h * paddle.rsqrt((h * h).mean(-1, keepdim=True) + 1e-6)
                h = F.silu(h) + paddle.tanh(h) * 0.1
                p = static.nn.fc(h, 1)
                loss = ((p - y) * (p - y)).mean()
                paddle.optimizer.SGD(learning_rate=0.05).minimize(loss)
            exe = static.Executor()
            exe.run(start)
            rng = np.random.default_rng(0)
            xv, yv = rng.standard_normal((8, 16)).astype("float32"), rng.standard_normal((8, 1)).astype("float32")
            return [float(exe.run(main, feed={"x": xv, "y": yv}, fetch_list=[loss])[0]) for _ in range(5)]
        finally:
            paddle.disable_static()
            paddle.set_flags(old)

    ref = run({"FLAGS_enable_pir_api": False, "FLAGS_use_cinn": False})
    before = cinn.stats["launches"]
    got = run({"FLAGS_enable_pir_api": True, "FLAGS_use_cinn": True})
    assert cinn.stats["launches"] - before >= 10                 # forward and backward kernels, every step
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert got[-1] < 0.5 * got[0]


def test_build_strategy_build_cinn_pass():
    """CompiledProgram(program, BuildStrategy(build_cinn_pass=True)) and to_static(build_strategy=...) select the generated-kernel path."""
    bs = static.BuildStrategy()
    bs.build_cinn_pass = True
    paddle.enable_static()
    try:
        main = static.Program()
        with static.program_guard(main):
            x = static.data("x", [-1, 12], "float32")
            out = F.softmax(paddle.tanh(x) * 2.0, -1) * 3.0
        cp = static.CompiledProgram(main, build_strategy=bs)
        exe = static.Executor()
        xv = np.random.default_rng(0).standard_normal((5, 12)).astype("float32")
        a = exe.run(cp, feed={"x": xv}, fetch_list=[out])[0]
        before = cinn.stats["launches"]
        b = exe.run(cp, feed={"x": xv}, fetch_list=[out])[0]
        assert cinn.stats["launches"] == before + 1
        np.testing.assert_allclose(b, a, rtol=2e-5, atol=1e-7)
    finally:
        paddle.disable_static()
    f = paddle.jit.to_static(lambda t: F.softmax(paddle.tanh(t) * 2.0, -1) * 3.0, build_strategy=bs)
    t = paddle.to_tensor(xv)
    with paddle.no_grad():
        f(t)
        got = f(t)
    assert f.cinn_report(t) is not None
    np.testing.assert_allclose(got.numpy(), a, rtol=2e-5, atol=1e-7)


def _branchy(x):
    if x.sum() > 0:
        y = paddle.exp(x) * 2.0 + 1.0
    else:
        y = paddle.tanh(x) - 1.0
    return y * 0.5 + F.sigmoid(y)


def test_tensor_dependent_branch_compiles_to_one_kernel():
    """dy2static turns the `if` into run-both-and-select; under backend="CINN" both branches, the select and the tail are one generated kernel."""
    f = paddle.jit.to_static(_branchy, backend="CINN")
    for sign in (1.0, -1.0):
        t = paddle.to_tensor(np.abs(np.random.default_rng(0).standard_normal((3, 4))).astype("float32") * sign)
        with paddle.no_grad():
            f(t)
            got = f(t)
        rep = f.cinn_report(t)
        assert rep is not None and len(rep.groups) == 1 and "where" in rep.groups[0]["ops"] and len(rep.groups[0]["ops"]) >= 8
        np.testing.assert_allclose(got.numpy(), _branchy(t).numpy(), rtol=1e-6, atol=1e-6)


def _early(x, k=2):
    z = x * 1.5
    if z.sum() > 0:                # tensor condition with an early return and a branch-local temporary
        w = paddle.exp(z)
        return w + 1.0
    if k > 1:                      # python condition: ordinary semantics
        return paddle.tanh(z) - k
    return z


def test_early_returns_are_normalised_and_compile():
    """`if c: ...; return A` + `...; return B` becomes an if / else that assigns the result, so a tensor condition converts (run both, select);
    a name assigned in one branch only stays undefined unless it is used."""
    from paddle_b200.jit.dy2static import get_code

    code = get_code(_early)
    assert "convert_ifelse" in code and "_jst_ret_" in code
    f = paddle.jit.to_static(_early, backend="CINN")
    plain = paddle.jit.to_static(_early)
    for sign in (1.0, -1.0):
        t = paddle.to_tensor(np.abs(np.random.default_rng(1).standard_normal((3, 4))).astype("float32") * sign)
        with paddle.no_grad():
            f(t)
            got = f(t)
        assert f.cinn_report(t) is not None and len(f.cinn_report(t).groups) >= 1
        np.testing.assert_allclose(got.numpy(), _early(t).numpy(), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(plain(t).numpy(), _early(t).numpy(), rtol=1e-6, atol=1e-6)
