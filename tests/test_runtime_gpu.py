"""GPU checks of the native runtime pieces that are not kernels: the range tracer's device timing."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tracer_device_timing():
    import paddle_b200 as paddle
    from paddle_b200 import profiler

    x = paddle.randn([4096, 5120]).astype("bfloat16").cuda()
    w = paddle.ones([5120]).astype("bfloat16").cuda()
    from paddle_b200.kernels import norm

    norm.rms_norm(x, w, 1e-5)
    torch.cuda.synchronize()
    with profiler.Profiler(targets=[profiler.ProfilerTarget.CPU, profiler.ProfilerTarget.GPU]) as prof:
        for _ in range(3):
            norm.rms_norm(x, w, 1e-5)
        prof.step()
    stats = profiler.kernel_statistics(prof)
    assert "rms_norm_fwd" in stats, stats
    calls, host_ms, dev_ms = stats["rms_norm_fwd"]
    assert calls == 3 and 0.003 < dev_ms < 5.0, stats   # 3 × ~84 MB of traffic: tens of microseconds each


def test_symm_heap_best_fit_single_rank():
    from paddle_b200 import _build

    C = _build.load(required=True)
    heap = C.SymmHeap(64 << 20, 4096, torch.cuda.current_device())
    heap.open_self()
    a = heap.alloc(8 << 20, 1024)
    b = heap.alloc(16 << 20, 1024)
    t = heap.tensor(b, [1024, 1024], torch.float32, -1)
    t.fill_(3.0)
    assert float(t.sum()) == 3.0 * 1024 * 1024
    free_before = heap.size() - heap.cursor()
    heap.free(a)
    c = heap.alloc(4 << 20, 1024)
    assert c == a and heap.size() - heap.cursor() >= free_before
    st = dict(heap.stats())
    assert st["live_blocks"] == 2 and st["allocated"] == (16 << 20) + (4 << 20) and st["peak_allocated"] >= 24 << 20
    with pytest.raises(RuntimeError, match="out of symmetric memory"):
        heap.alloc(128 << 20, 1024)


def _mlp_and_opt(seed, lr):
    import paddle_b200 as paddle

    paddle.seed(seed)
    net = paddle.nn.Sequential(paddle.nn.Linear(64, 128), paddle.nn.GELU(), paddle.nn.Linear(128, 10))
    net.to("gpu")
    opt = paddle.optimizer.AdamW(lr, parameters=net.parameters(), weight_decay=0.01, grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0))
    opt.enable_flat_arena()
    return net, opt


def test_captured_train_step_matches_eager():
    """jit.capture_train_step: forward + backward + clip + fused AdamW + clear_grad replayed from one CUDA graph, with a changing lr."""
    import numpy as np

    import paddle_b200 as paddle

    sched_a = paddle.optimizer.lr.StepDecay(1e-2, step_size=3, gamma=0.5)
    sched_b = paddle.optimizer.lr.StepDecay(1e-2, step_size=3, gamma=0.5)
    net_a, opt_a = _mlp_and_opt(7, sched_a)
    net_b, opt_b = _mlp_and_opt(7, sched_b)
    ce = paddle.nn.CrossEntropyLoss()
    step = paddle.jit.capture_train_step(lambda x, y: ce(net_b(x), y), opt_b, warmup=2)
    rng = np.random.RandomState(0)
    losses_a, losses_b = [], []
    for i in range(10):
        x = paddle.to_tensor(rng.randn(32, 64).astype("float32")).cuda()
        y = paddle.to_tensor(rng.randint(0, 10, (32,))).cuda()
        la = ce(net_a(x), y)
        la.backward()
        opt_a.step()
        opt_a.clear_grad()
        sched_a.step()
        losses_a.append(float(la))
        losses_b.append(float(step(x, y)))
        sched_b.step()
    assert step.captured, step.failure
    assert step.replays >= 6 and opt_b._step_count == opt_a._step_count == 10
    np.testing.assert_allclose(losses_b, losses_a, rtol=2e-4, atol=2e-5)
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        np.testing.assert_allclose(pb.numpy(), pa.numpy(), rtol=2e-4, atol=2e-5)
    assert losses_a[-1] < losses_a[0]


def test_sharded_optimizer_offload_matches_device_update():
    """GroupShardedOptimizerStage2(offload=True): master weights + moments in pinned host memory, CPU update — same result as the
    device-resident update (single rank: the shard is the whole slab)."""
    import numpy as np

    import paddle_b200 as paddle
    from paddle_b200.distributed.sharding import GroupShardedOptimizerStage2

    def run(offload):
        paddle.seed(3)
        net = paddle.nn.Sequential(paddle.nn.Linear(32, 64), paddle.nn.GELU(), paddle.nn.Linear(64, 8))
        net.to("gpu")
        inner = paddle.optimizer.AdamW(1e-2, parameters=net.parameters(), weight_decay=0.01)
        opt = GroupShardedOptimizerStage2(params=net.parameters(), optim=inner, group=None, offload=offload)
        rs = np.random.RandomState(0)
        for _ in range(4):
            x = paddle.to_tensor(rs.randn(16, 32).astype("float32")).cuda()
            (net(x) ** 2).mean().backward()
            opt.step()
            opt.clear_grad()
        return [p.numpy() for p in net.parameters()]

    a, b = run(False), run(True)
    for x, y in zip(a, b):
        np.testing.assert_allclose(y, x, rtol=2e-5, atol=2e-6)


def test_to_static_captures_training_forward_backward():
    """to_static on a training Layer: after two eager warm-ups forward AND backward replay as CUDA graphs; gradients match eager."""
    import numpy as np
    import torch

    import paddle_b200 as paddle

    paddle.set_device("gpu:0")
    paddle.seed(3)

    class Net(paddle.nn.Layer):
        def __init__(self):
            super().__init__()
            self.a, self.b = paddle.nn.Linear(64, 128), paddle.nn.Linear(128, 32)

        def forward(self, x):
            h = paddle.nn.functional.gelu(self.a(x))
            if h.mean() > 0:                      # tensor-dependent branch: converted to run-both + select, capture safe
                h = h * 2
            else:
                h = h * 0.5
            return self.b(h)

    net, ref = Net(), Net()
    ref.set_state_dict(net.state_dict())
    snet = paddle.jit.to_static(net)
    xs = [paddle.to_tensor(np.random.RandomState(i).randn(16, 64).astype(np.float32)) for i in range(5)]
    for i, x in enumerate(xs):
        for m in (snet, ref):
            for p in m.parameters():
                p.clear_gradient()
        (snet(x) ** 2).mean().backward()
        (ref(x) ** 2).mean().backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            np.testing.assert_allclose(p.grad.numpy(), q.grad.numpy(), rtol=2e-4, atol=1e-6)
    graphs = [v for v in snet.forward._train_graphs.values()]
    assert graphs and graphs[0] is not None, "training call was not captured"


@pytest.mark.gpu
def test_native_auto_growth_allocator_serves_all_cuda_memory():
    """FLAGS_b200_native_allocator=1: torch's allocations go through csrc/runtime/allocator.cpp (CUDAPluggableAllocator); a small training
    loop runs, the statistics move, cross-stream frees are deferred, empty_cache returns idle chunks to cudaFree."""
    import os
    import subprocess
    import sys

    code = r"""
import torch, paddle_b200 as paddle
from paddle_b200.device import cuda as C
assert C.auto_growth_allocator_active()
x = torch.randn(1024, 1024, device="cuda")
s0 = C.allocator_stats()
assert s0["allocated"] >= x.numel() * 4 and s0["num_chunks"] >= 1 and C.memory_allocated() == s0["allocated"]
lin = paddle.nn.Linear(1024, 1024).to("cuda")
opt = paddle.optimizer.AdamW(learning_rate=1e-3, parameters=lin.parameters())
for _ in range(5):
    loss = (lin(x.as_subclass(paddle.Tensor)) ** 2).mean()
    loss.backward(); opt.step(); opt.clear_grad()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    y = torch.empty(1 << 20, device="cuda")
    y.fill_(1.0)
del y                                             # freed on the default stream, allocated on `side`
torch.cuda.synchronize()
s1 = C.allocator_stats()
assert s1["num_allocs"] > s0["num_allocs"] and s1["allocated_peak"] >= s1["allocated"] and s1["reserved"] >= s1["allocated"]
peak = C.max_memory_allocated()
big = torch.empty(96 << 20, dtype=torch.uint8, device="cuda")      # larger than a chunk: gets a chunk of its own
s_big = C.allocator_stats()
assert s_big["reserved"] >= s1["reserved"] + (96 << 20)
del big
C.empty_cache()                                                    # ... which goes back to cudaFree once it is idle
s2 = C.allocator_stats()
assert s2["reserved"] <= s_big["reserved"] - (96 << 20) and s2["num_backend_frees"] >= 1, (s_big, s2)
print("OK", s1["num_allocs"], s1["num_chunks"], s1["deferred_frees"], peak, s2["reserved"])
"""
    env = dict(os.environ, FLAGS_b200_native_allocator="1", B200_ALLOCATOR_CHUNK_MB="64")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
