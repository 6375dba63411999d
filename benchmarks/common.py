"""Shared timing helpers for the secondary BASELINE.json configs (device-timed, max over ranks, one JSON line)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def init_dist():
    import paddle_b200 as paddle

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    paddle.set_device(f"gpu:{local}")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        paddle.distributed.init_parallel_env()
    return paddle, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def timed(step, steps, warmup):
    import torch.distributed as dist

    for _ in range(warmup):
        step()
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        last = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), last


def report(rank, **kw):
    if rank == 0:
        print(json.dumps(kw))
