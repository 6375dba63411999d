"""Short single-GPU workload for ncu captures: Llama-2-13B widths, 2 decoder layers, seq 4096, 3 optimizer steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paddle_b200 as paddle  # noqa: E402
from paddle_b200.models import llama as L  # noqa: E402

paddle.set_device("gpu:0")
paddle.set_default_dtype("bfloat16")
cfg = L.llama2_13b(num_hidden_layers=int(os.environ.get("LAYERS", "2")), recompute=False)
m = L.LlamaForCausalLM(cfg)
opt = paddle.optimizer.AdamW(1e-5, parameters=m.parameters(), weight_decay=0.1, multi_precision=True, moment_dtype="bfloat16",
                             grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0))
opt.enable_flat_arena()
ids = torch.randint(0, 32000, (1, 4097), device="cuda").as_subclass(paddle.Tensor)
for _ in range(int(os.environ.get("STEPS", "3"))):
    loss = m(ids[:, :-1], ids[:, 1:])
    loss.backward()
    opt.step()
    opt.clear_grad()
torch.cuda.synchronize()
print("loss", float(loss))
