"""Weight-only int8 / int4 linear (in-SM dequantisation) vs the bf16 tcgen05 GEMM and vs dequantise-then-GEMM, decode and prefill rows."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paddle_b200 as paddle  # noqa: E402
from paddle_b200._build import ext  # noqa: E402
from paddle_b200.nn import quant as Q  # noqa: E402

E = ext()


def timeit(fn, n=20):
    """Device time per call: `n` calls captured into one CUDA graph (no host launch overhead between them), replayed 5 times."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n)


rows = []
for k, n in [(5120, 15360), (5120, 5120), (13824, 5120), (5120, 27648)]:
    ws = [torch.randn(k, n, device="cuda") * 0.02 for _ in range(4)]       # rotate weights: 4 x >= 52 MB defeats the 126 MB L2 only partly, report as is
    qs = [Q.weight_quantize(w.as_subclass(paddle.Tensor), algo="weight_only_int8") for w in ws]
    q4 = [Q.weight_quantize(w.as_subclass(paddle.Tensor), algo="weight_only_int4") for w in ws]
    wb = [w.to(torch.bfloat16) for w in ws]
    for m in (1, 16, 64):
        x = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
        i = [0]

        def nxt():
            i[0] = (i[0] + 1) % 4
            return i[0]

        w8 = [(q[0].as_subclass(torch.Tensor), q[1].as_subclass(torch.Tensor).float()) for q in qs]
        w4 = [(q[0].as_subclass(torch.Tensor), q[1].as_subclass(torch.Tensor).float()) for q in q4]
        t8 = timeit(lambda: E.weight_only_gemm(x, *w8[nxt()], None, False))
        t4 = timeit(lambda: E.weight_only_gemm(x, *w4[nxt()], None, True))
        tb = timeit(lambda: torch.matmul(x, wb[nxt()]))
        row = {"k": k, "n": n, "m": m, "int8_ms": round(t8, 4), "int4_ms": round(t4, 4), "bf16_matmul_ms": round(tb, 4), "int8_speedup_vs_bf16": round(tb / t8, 2),
               "int4_speedup_vs_bf16": round(tb / t4, 2), "int8_weight_gbs": round(k * n / t8 / 1e6, 1), "int4_weight_gbs": round(k * n / 2 / t4 / 1e6, 1)}
        rows.append(row)
        print(json.dumps(row))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_weight_only.json", "w"), indent=1)
