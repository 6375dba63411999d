"""torch.profiler per-kernel table of one mp2 (+sequence parallel) Llama training step (run under torchrun, rank 0 prints)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paddle_b200 as paddle  # noqa: E402
from paddle_b200.distributed import env, fleet  # noqa: E402
from paddle_b200.models import llama as L  # noqa: E402

layers = int(os.environ.get("LAYERS", "4"))
mb = int(os.environ.get("MB", "4"))
lr_ = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr_)
paddle.set_device(f"gpu:{lr_}")
s = fleet.DistributedStrategy()
s.hybrid_configs = {"dp_degree": 1, "mp_degree": 2, "pp_degree": 1}
fleet.init(is_collective=True, strategy=s)
paddle.seed(1 + env.get_rank())
paddle.set_default_dtype("bfloat16")
cfg = L.llama2_13b(num_hidden_layers=layers, tensor_parallel_degree=2, sequence_parallel=True)
model = L.LlamaForCausalLM(cfg)
opt = paddle.optimizer.AdamW(1e-5, parameters=model.parameters(), weight_decay=0.1, multi_precision=True, moment_dtype="bfloat16",
                             grad_clip=paddle.nn.ClipGradByGlobalNorm(1.0))
model = fleet.distributed_model(model)
opt = fleet.distributed_optimizer(opt)
ids = torch.randint(0, 32000, (mb, 4097), generator=torch.Generator().manual_seed(0)).cuda().as_subclass(paddle.Tensor)


def step():
    loss = model(ids[:, :-1], ids[:, 1:])
    loss.backward()
    opt.step()
    opt.clear_grad()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
if env.get_rank() == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=80))
env.destroy_process_group()
