"""Micro-benchmarks: tcgen05 GEMM vs torch.matmul (cuBLAS) on Llama-2-13B shapes, and HBM-bound kernels vs copy peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddle_b200._build import ext  # noqa: E402

E = ext()
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_bench():
    rows = []
    shapes = [  # (M, N, K, a_km, b_nk, tag)
        (4096, 15360, 5120, False, False, "qkv fwd"), (4096, 5120, 5120, False, False, "o fwd"),
        (4096, 27648, 5120, False, False, "gate_up fwd"), (4096, 5120, 13824, False, False, "down fwd"),
        (4096, 5120, 15360, False, True, "qkv dx"), (5120, 15360, 4096, True, False, "qkv dW"),
        (4096, 13824, 5120, False, True, "down dx"), (13824, 5120, 4096, True, False, "down dW"),
        (8192, 8192, 8192, False, True, "8k^3 TN"), (4096, 32000, 5120, False, False, "lm_head"),
    ]
    for (m, n, k, a_km, b_nk, tag) in shapes:
        A = torch.randn((k, m) if a_km else (m, k), device="cuda", dtype=torch.bfloat16)
        B = torch.randn((n, k) if b_nk else (k, n), device="cuda", dtype=torch.bfloat16)
        t_mine = timeit(lambda: E.gemm(A, B, None, a_km, b_nk, 0, None, None))
        Aa = A.t() if a_km else A
        Bb = B.t() if b_nk else B
        t_ref = timeit(lambda: torch.matmul(Aa, Bb))
        fl = 2.0 * m * n * k
        r = dict(tag=tag, m=m, n=n, k=k, ms=round(t_mine, 4), tflops=round(fl / t_mine / 1e9, 1), cublas_ms=round(t_ref, 4),
                 cublas_tflops=round(fl / t_ref / 1e9, 1), frac_of_measured_peak=round(fl / t_mine / 1e9 / PEAKS.get("bf16_tflops", 1689.8), 3))
        print(json.dumps(r), flush=True)
        rows.append(r)
    return rows


def bw_bench():
    rows = []
    peak = PEAKS.get("hbm_gbs", 6576.4)
    x = torch.randn(16384, 5120, device="cuda", dtype=torch.bfloat16)
    w = torch.ones(5120, device="cuda", dtype=torch.bfloat16)
    t = timeit(lambda: E.rms_norm_fwd(x, None, w, None, 1e-6))
    rows.append(dict(kernel="rms_norm_fwd", ms=round(t, 4), gbs=round(2 * x.numel() * 2 / t / 1e6, 1)))
    y, rstd, _ = E.rms_norm_fwd(x, None, w, None, 1e-6)
    t = timeit(lambda: E.rms_norm_bwd(y, x, w, rstd))
    rows.append(dict(kernel="rms_norm_bwd", ms=round(t, 4), gbs=round(3 * x.numel() * 2 / t / 1e6, 1)))
    g = torch.randn(16384, 13824, device="cuda", dtype=torch.bfloat16)
    u = torch.randn_like(g)
    t = timeit(lambda: E.swiglu_fwd(g, u))
    rows.append(dict(kernel="swiglu_fwd", ms=round(t, 4), gbs=round(3 * g.numel() * 2 / t / 1e6, 1)))
    lg = torch.randn(8192, 32000, device="cuda", dtype=torch.bfloat16)
    lab = torch.randint(0, 32000, (8192,), device="cuda")
    t = timeit(lambda: E.softmax_ce_fwd(lg, lab, -100))
    rows.append(dict(kernel="softmax_ce_fwd", ms=round(t, 4), gbs=round(lg.numel() * 2 / t / 1e6, 1)))
    n = 1 << 28
    p = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    gr = torch.randn(n, device="cuda", dtype=torch.bfloat16)
    ms_ = torch.zeros(n, device="cuda")
    m_ = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    v_ = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    t = timeit(lambda: E.adamw_step(p, gr, ms_, m_, v_, 1e-3, 0.9, 0.95, 1e-8, 0.1, 1, None, 0.0, None, None), iters=5, warmup=2)
    rows.append(dict(kernel="adamw(bf16 p/g/m/v + fp32 master)", ms=round(t, 4), gbs=round(n * (2 + 2 + 4 + 4 + 2 * 2 + 2 * 2) / t / 1e6, 1)))
    a = torch.empty(1 << 29, device="cuda", dtype=torch.bfloat16)
    b = torch.empty_like(a)
    t = timeit(lambda: b.copy_(a), iters=5, warmup=2)
    rows.append(dict(kernel="torch copy (ref)", ms=round(t, 4), gbs=round(2 * a.numel() * 2 / t / 1e6, 1)))
    for r in rows:
        r["frac_of_measured_copy_peak"] = round(r["gbs"] / peak, 3)
        print(json.dumps(r), flush=True)
    return rows


def attn_bench():
    """tcgen05 flash-attention forward vs the library flash kernel (torch SDPA) on Llama-2-13B attention shapes."""
    import torch.nn.functional as F

    rows = []
    peak = PEAKS.get("bf16_tflops_sustained", 1433.6)
    for (b, s, h, causal) in ((1, 4096, 40, True), (4, 4096, 20, True), (2, 8192, 40, True), (4, 2048, 40, False)):
        q, k, v = (torch.randn(b, s, h, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
        flops = 4.0 * b * h * s * s * 128 * (0.5 if causal else 1.0)
        t = timeit(lambda: E.attention_fwd(q, k, v, 128 ** -0.5, causal))
        qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
        t_lib = timeit(lambda: F.scaled_dot_product_attention(qt, kt, vt, is_causal=causal))
        rows.append(dict(kernel="flash_attn_fwd", b=b, s=s, h=h, causal=causal, ms=round(t, 4), tflops=round(flops / t / 1e9, 1),
                         lib_ms=round(t_lib, 4), lib_tflops=round(flops / t_lib / 1e9, 1), frac_of_measured_bf16_peak=round(flops / t / 1e9 / peak, 3)))
        print(json.dumps(rows[-1]), flush=True)
        out, lse = E.attention_fwd(q, k, v, 128 ** -0.5, causal)
        g = torch.randn_like(out)
        tb = timeit(lambda: E.attention_bwd(q, k, v, out, lse, g, 128 ** -0.5, causal))
        qr, kr, vr = (x.detach().requires_grad_(True) for x in (qt, kt, vt))
        o2 = F.scaled_dot_product_attention(qr, kr, vr, is_causal=causal)
        g2 = g.transpose(1, 2)
        tb_lib = timeit(lambda: torch.autograd.grad(o2, (qr, kr, vr), g2, retain_graph=True))
        rows.append(dict(kernel="flash_attn_bwd", b=b, s=s, h=h, causal=causal, ms=round(tb, 4), tflops=round(2.5 * flops / tb / 1e9, 1),
                         lib_ms=round(tb_lib, 4), lib_tflops=round(2.5 * flops / tb_lib / 1e9, 1), frac_of_measured_bf16_peak=round(2.5 * flops / tb / 1e9 / peak, 3)))
        print(json.dumps(rows[-1]), flush=True)
    return rows


if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0), "peaks": PEAKS}
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "gemm"):
        out["gemm"] = gemm_bench()
    if which in ("all", "bw"):
        out["bw"] = bw_bench()
    if which in ("all", "attn"):
        out["attn"] = attn_bench()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/bench_kernels_{which}.json", "w"), indent=1)
