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
