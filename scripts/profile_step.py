"""Per-kernel device-time breakdown of one Llama training step (torch.profiler, CUDA activities)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paddle_b200 as paddle  # noqa: E402
from paddle_b200.models import llama as L  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
recompute = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
paddle.set_device("gpu:0")
paddle.set_default_dtype("bfloat16")
cfg = L.llama2_13b(num_hidden_layers=layers, recompute=recompute)
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
print("step ms", s.elapsed_time(e))
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
