"""paddle.distributed.spawn. Parity: python/paddle/distributed/spawn.py."""
from __future__ import annotations

import multiprocessing as mp
import os
import socket


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(func, rank, nprocs, port, args, kwargs, env):
    os.environ.update(env)
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(nprocs), "LOCAL_RANK": str(rank), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port),
                       "PADDLE_TRAINER_ID": str(rank), "PADDLE_TRAINERS_NUM": str(nprocs)})
    func(*args, **kwargs)


class MultiprocessContext:
    def __init__(self, procs):
        self.processes = procs

    def join(self, timeout=None):
        for p in self.processes:
            p.join(timeout)
        bad = [p for p in self.processes if p.exitcode not in (0, None)]
        if bad:
            raise RuntimeError(f"spawned process exited with code {bad[0].exitcode}")
        return True


def spawn(func, args=(), nprocs=-1, join=True, daemon=False, **options):
    import torch

    if nprocs <= 0:
        nprocs = max(1, torch.cuda.device_count()) if torch.cuda.is_available() else 1
    ctx = mp.get_context("spawn")
    port = _free_port()
    env = {k: v for k, v in options.items() if isinstance(v, str)}
    procs = []
    for r in range(nprocs):
        p = ctx.Process(target=_worker, args=(func, r, nprocs, port, args, {}, env), daemon=daemon)
        p.start()
        procs.append(p)
    c = MultiprocessContext(procs)
    if join:
        c.join()
    return c
