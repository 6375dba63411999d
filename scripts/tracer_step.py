"""Per-kernel-entry device time of one Llama training step, measured by the framework's own range tracer
(csrc/runtime/tracer.cpp: cudaEvent pair around every `_C` entry point on the launching stream; no CUPTI, no torch.profiler).
Writes a markdown table.  usage: python scripts/tracer_step.py [layers] [out.md]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paddle_b200 as paddle  # noqa: E402
from paddle_b200 import _build  # noqa: E402
from paddle_b200.models import llama as L  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/tracer_step.md"
C = _build.load(required=True)
paddle.set_device("gpu:0")
paddle.set_default_dtype("bfloat16")
cfg = L.llama2_13b(num_hidden_layers=layers, recompute=False)
m = L.LlamaForCausalLM(cfg)
opt = paddle.optimizer.AdamW(1e-5, parameters=m.parameters(), weight_decay=0.1, multi_precision=True, moment_dtype="bfloat16",
                             grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0))
opt.enable_flat_arena()
ids = torch.randint(0, 32000, (1, 4097), device="cuda").as_subclass(paddle.Tensor)


def step():
    loss = m(ids[:, :-1], ids[:, 1:])
    loss.backward()
    opt.step()
    opt.clear_grad()


for _ in range(3):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
step()
e.record()
torch.cuda.synchronize()
plain_ms = s.elapsed_time(e)
C.tracer_collect()
C.tracer_enable(2)
s.record()
step()
e.record()
torch.cuda.synchronize()
traced_ms = s.elapsed_time(e)
evs = C.tracer_collect()
C.tracer_enable(0)
agg = {}
for name, typ, tid, depth, t0, t1, d0, dd in evs:
    c, h, g = agg.get(name, (0, 0.0, 0.0))
    agg[name] = (c + 1, h + (t1 - t0) / 1e6, g + max(dd, 0.0) / 1e3)
tot = sum(v[2] for v in agg.values())
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
with open(out, "w") as f:
    f.write(f"# Native range tracer: Llama-2-13B dims, {layers} layers, seq 4096, one training step (fwd+bwd+AdamW)\n\n")
    f.write(f"step without tracing {plain_ms:.2f} ms, with device tracing {traced_ms:.2f} ms; own-kernel entries cover {tot:.2f} ms of device time\n\n")
    f.write("| kernel entry (`_C.*`) | calls | host ms | device ms | share of traced |\n|---|---|---|---|---|\n")
    for k, (c, h, g) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        f.write(f"| {k} | {c} | {h:.3f} | {g:.3f} | {100 * g / max(tot, 1e-9):.1f} % |\n")
print(open(out).read())
