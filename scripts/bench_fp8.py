"""fp8 (tcgen05 kind::f8f6f4 per-tensor scaled with device-side scales, and kind::mxf8f6f4.block_scale with MX E8M0 block scales) vs bf16 GEMM on Llama-2-13B shapes + the fused quantise cost.
Writes gpurun_out/bench_fp8.json (CUDA events, 3 warm-ups, inputs >> L2 rotated between iterations)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paddle_b200  # noqa: E402,F401
from paddle_b200._build import ext  # noqa: E402

E = ext()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


rows = []
for tag, m, n, k in [("qkv fwd", 8192, 15360, 5120), ("o fwd", 8192, 5120, 5120), ("gate_up fwd", 8192, 27648, 5120), ("down fwd", 8192, 5120, 13824)]:
    x = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(k, n, device="cuda") * 0.02).to(torch.bfloat16)
    xq, _, sx = E.quantize_fp8(x, False, False)
    _, wqt, sw = E.quantize_fp8(w, False, True)
    t_bf16 = timeit(lambda: E.gemm(x, w, None, False, False, 0, None, None))
    t_fp8 = timeit(lambda: E.gemm_fp8(xq, wqt, None, 1.0, 0, torch.bfloat16, sx, sw))
    t_qx = timeit(lambda: E.quantize_fp8(x, False, True))
    t_qw = timeit(lambda: E.quantize_fp8(w, False, True))
    xm, sxm = E.quantize_mx(x)
    wm, swm = E.quantize_mx(w.t().contiguous())
    t_mx = timeit(lambda: E.gemm_fp8_mx(xm, sxm, wm, swm, None, torch.bfloat16))
    t_qmx = timeit(lambda: E.quantize_mx(x))
    ref = x.float() @ w.float()
    err_mx = ((E.gemm_fp8_mx(xm, sxm, wm, swm, None, torch.bfloat16).float() - ref).norm() / ref.norm()).item()
    out = E.gemm_fp8(xq, wqt, None, 1.0, 0, torch.bfloat16, sx, sw).float()
    err = ((out - ref).norm() / ref.norm()).item()
    fl = 2.0 * m * n * k
    rows.append({"tag": tag, "m": m, "n": n, "k": k, "bf16_ms": round(t_bf16, 4), "bf16_tflops": round(fl / t_bf16 / 1e9, 1), "fp8_ms": round(t_fp8, 4),
                 "fp8_tflops": round(fl / t_fp8 / 1e9, 1), "fp8_speedup": round(t_bf16 / t_fp8, 3), "quantize_x_ms": round(t_qx, 4), "quantize_w_ms": round(t_qw, 4),
                 "quantize_x_gbs": round((m * k * 2 + 2 * m * k) / t_qx / 1e6, 1), "rel_err_vs_fp32": round(err, 4),
                 "mx_ms": round(t_mx, 4), "mx_tflops": round(fl / t_mx / 1e9, 1), "mx_speedup_vs_bf16": round(t_bf16 / t_mx, 3), "quantize_mx_x_ms": round(t_qmx, 4),
                 "mx_rel_err_vs_fp32": round(err_mx, 4)})
    print(json.dumps(rows[-1]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_fp8.json", "w"), indent=1)
