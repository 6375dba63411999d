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
