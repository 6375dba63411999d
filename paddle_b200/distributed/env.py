"""Parallel environment. Parity: python/paddle/distributed/parallel.py (init_parallel_env, ParallelEnv, get_rank...).

One process per GPU; rendezvous and process groups are torch.distributed (NCCL on GPU, gloo on CPU).
"""
from __future__ import annotations

import datetime
import os

import torch
import torch.distributed as dist

_initialized = [False]


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def get_rank(group=None):
    if dist.is_available() and dist.is_initialized():
        g = getattr(group, "pg", group)
        return dist.get_rank(g) if g is not None else dist.get_rank()
    return _env_int("PADDLE_TRAINER_ID", _env_int("RANK", 0))


def get_world_size(group=None):
    if dist.is_available() and dist.is_initialized():
        g = getattr(group, "pg", group)
        return dist.get_world_size(g) if g is not None else dist.get_world_size()
    return _env_int("PADDLE_TRAINERS_NUM", _env_int("WORLD_SIZE", 1))


def local_rank():
    return _env_int("LOCAL_RANK", _env_int("PADDLE_RANK_IN_NODE", get_rank()))


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def is_available():
    return dist.is_available()


def init_parallel_env(backend=None, timeout_s=1800):
    """paddle.distributed.init_parallel_env()."""
    if is_initialized():
        return ParallelEnv()
    world = _env_int("WORLD_SIZE", _env_int("PADDLE_TRAINERS_NUM", 1))
    rank = _env_int("RANK", _env_int("PADDLE_TRAINER_ID", 0))
    use_cuda = torch.cuda.is_available()
    if backend is None:
        backend = "nccl" if use_cuda else "gloo"
    if use_cuda:
        lr = local_rank() % max(torch.cuda.device_count(), 1)
        torch.cuda.set_device(lr)
        from ..framework import place

        place.set_device(f"gpu:{lr}")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    kw = {}
    if use_cuda and backend == "nccl":
        kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=timeout_s), **kw)
    _initialized[0] = True
    return ParallelEnv()


def destroy_process_group(group=None):
    if is_initialized():
        dist.destroy_process_group(getattr(group, "pg", group))


class ParallelEnv:
    @property
    def rank(self):
        return get_rank()

    @property
    def world_size(self):
        return get_world_size()

    @property
    def local_rank(self):
        return get_rank()

    @property
    def nranks(self):
        return get_world_size()

    @property
    def device_id(self):
        return torch.cuda.current_device() if torch.cuda.is_available() else 0

    @property
    def dev_id(self):
        return self.device_id

    @property
    def current_endpoint(self):
        return f"{os.environ.get('MASTER_ADDR', '127.0.0.1')}:{os.environ.get('MASTER_PORT', '29500')}"

    @property
    def trainer_endpoints(self):
        return [self.current_endpoint]
