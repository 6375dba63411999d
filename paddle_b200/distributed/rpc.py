"""paddle.distributed.rpc. Parity: python/paddle/distributed/rpc/rpc.py (init_rpc, rpc_sync, rpc_async, shutdown,
get_worker_info...). Transport: torch.distributed.rpc (TensorPipe) in place of the reference's brpc."""
from __future__ import annotations

import os
from collections import namedtuple

WorkerInfo = namedtuple("WorkerInfo", ["name", "rank", "ip", "port"])
_state = {"name": None}


def init_rpc(name, rank=None, world_size=None, master_endpoint=None):
    import torch.distributed.rpc as trpc

    rank = int(os.environ.get("PADDLE_TRAINER_ID", os.environ.get("RANK", 0))) if rank is None else rank
    world_size = int(os.environ.get("PADDLE_TRAINERS_NUM", os.environ.get("WORLD_SIZE", 1))) if world_size is None else world_size
    ep = master_endpoint or os.environ.get("PADDLE_MASTER_ENDPOINT", "127.0.0.1:29600")
    opts = trpc.TensorPipeRpcBackendOptions(init_method=f"tcp://{ep}", rpc_timeout=1800)
    trpc.init_rpc(name, rank=rank, world_size=world_size, rpc_backend_options=opts)
    _state["name"] = name


def rpc_sync(to, fn, args=None, kwargs=None, timeout=-1):
    import torch.distributed.rpc as trpc

    return trpc.rpc_sync(to, fn, args=args or (), kwargs=kwargs or {}, timeout=timeout if timeout > 0 else -1)


class _Future:
    def __init__(self, f):
        self._f = f

    def wait(self):
        return self._f.wait()


def rpc_async(to, fn, args=None, kwargs=None, timeout=-1):
    import torch.distributed.rpc as trpc

    return _Future(trpc.rpc_async(to, fn, args=args or (), kwargs=kwargs or {}, timeout=timeout if timeout > 0 else -1))


def shutdown():
    import torch.distributed.rpc as trpc

    trpc.shutdown()


def get_worker_info(name):
    import torch.distributed.rpc as trpc

    w = trpc.get_worker_info(name)
    return WorkerInfo(w.name, w.id, "127.0.0.1", 0)


def get_all_worker_infos():
    import torch.distributed.rpc as trpc

    return [WorkerInfo(w.name, w.id, "127.0.0.1", 0) for w in trpc.api._get_current_rpc_agent().get_worker_infos()]


def get_current_worker_info():
    return get_worker_info(_state["name"])
