"""Remaining names of python/paddle/distributed/__init__.py: gloo helpers, PS-mode datasets / entry configs, enums, `io`."""
from __future__ import annotations

import os
from enum import IntEnum

import torch.distributed as dist


class ParallelMode:
    """Parity: distributed/parallel.py:ParallelMode."""
    DATA_PARALLEL = 0
    TENSOR_PARALLEL = 1
    PIPELINE_PARALLEL = 2
    SHARDING_PARALLEL = 3


class ReduceType(IntEnum):
    kRedSum = 0
    kRedMax = 1
    kRedMin = 2
    kRedProd = 3
    kRedAvg = 4
    kRedAny = 5
    kRedAll = 6


_gloo = [None]


def gloo_init_parallel_env(rank_id, rank_num, server_endpoint):
    """CPU-only rendezvous (gloo) for parameter-server style jobs. Parity: distributed/parallel.py:gloo_init_parallel_env."""
    host, port = server_endpoint.split(":")
    if not dist.is_initialized():
        store = dist.TCPStore(host, int(port), rank_num, is_master=(rank_id == 0))
        dist.init_process_group("gloo", store=store, rank=rank_id, world_size=rank_num)
    _gloo[0] = True


def gloo_barrier():
    assert _gloo[0], "call gloo_init_parallel_env first"
    dist.barrier()


def gloo_release():
    if _gloo[0] and dist.is_initialized():
        dist.destroy_process_group()
    _gloo[0] = None


class _Entry:
    def _to_attr(self):
        return self._attr


class CountFilterEntry(_Entry):
    """Sparse-table admission: a feature id enters the table after `count_filter` occurrences. Parity: distributed/entry_attr.py."""

    def __init__(self, count_filter):
        if not isinstance(count_filter, int) or count_filter < 0:
            raise ValueError("count_filter must be a non-negative integer")
        self.count_filter = count_filter
        self._attr = f"count_filter_entry:{count_filter}"


class ProbabilityEntry(_Entry):
    def __init__(self, probability):
        if not isinstance(probability, float) or not 0 < probability <= 1:
            raise ValueError("probability must be a float in (0, 1]")
        self.probability = probability
        self._attr = f"probability_entry:{probability}"


class ShowClickEntry(_Entry):
    def __init__(self, show_name, click_name):
        if not isinstance(show_name, str) or not isinstance(click_name, str):
            raise ValueError("show_name / click_name must be strings")
        self.show_name, self.click_name = show_name, click_name
        self._attr = f"show_click_entry:{show_name}:{click_name}"


class InMemoryDataset:
    """Line-oriented slot dataset held in host memory (PS-mode CTR pipelines). Parity: distributed/fleet/dataset/dataset.py.
    `init(batch_size, use_var, pipe_command, parse_fn)`: every line of the file list is parsed by `parse_fn` (default:
    whitespace-separated floats); `load_into_memory`, `local_shuffle` / `global_shuffle`, iteration yields batches."""

    def __init__(self):
        self.batch_size, self.filelist, self.parse_fn, self._rows, self.thread_num = 1, [], None, [], 1

    def init(self, batch_size=1, thread_num=1, use_var=None, pipe_command=None, input_type=0, fs_name="", fs_ugi="", download_cmd="cat", parse_fn=None, **kw):
        self.batch_size, self.thread_num, self.use_var, self.parse_fn = batch_size, thread_num, use_var or [], parse_fn

    def set_filelist(self, filelist):
        self.filelist = list(filelist)

    def _parse(self, line):
        if self.parse_fn is not None:
            return self.parse_fn(line)
        return [float(x) for x in line.split()]

    def load_into_memory(self, is_shuffle=False):
        self._rows = []
        for fn in self.filelist:
            with open(fn) as f:
                self._rows.extend(self._parse(l) for l in f if l.strip())
        if is_shuffle:
            self.local_shuffle()

    def local_shuffle(self):
        import random

        random.shuffle(self._rows)

    def global_shuffle(self, fleet=None, thread_num=12):
        """Rows are re-partitioned across trainers by hash so every trainer sees a random 1/N of the global data."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            n, me = dist.get_world_size(), dist.get_rank()
            buckets = [[r for i, r in enumerate(self._rows) if hash((me, i)) % n == d] for d in range(n)]
            gathered = [None] * n
            dist.all_gather_object(gathered, buckets)
            self._rows = [r for g in gathered for r in g[me]]
        self.local_shuffle()

    def get_memory_data_size(self, fleet=None):
        return len(self._rows)

    def get_shuffle_data_size(self, fleet=None):
        return len(self._rows)

    def release_memory(self):
        self._rows = []

    def __iter__(self):
        import numpy as np

        for i in range(0, len(self._rows), self.batch_size):
            yield np.asarray(self._rows[i:i + self.batch_size], dtype=np.float32)


class QueueDataset(InMemoryDataset):
    """Streaming variant: files are read lazily, nothing is kept in memory."""

    def load_into_memory(self, is_shuffle=False):
        raise RuntimeError("QueueDataset streams from files; load_into_memory is not available")

    def __iter__(self):
        import numpy as np

        batch = []
        for fn in self.filelist:
            with open(fn) as f:
                for l in f:
                    if not l.strip():
                        continue
                    batch.append(self._parse(l))
                    if len(batch) == self.batch_size:
                        yield np.asarray(batch, dtype=np.float32)
                        batch = []
        if batch:
            yield np.asarray(batch, dtype=np.float32)


class _IO:
    """paddle.distributed.io: persistables save / load helpers for static programs."""

    @staticmethod
    def save_persistables(executor, dirname, main_program=None, filename=None):
        from .. import static

        os.makedirs(dirname, exist_ok=True)
        static.save(main_program or static.default_main_program(), os.path.join(dirname, filename or "persistables"))

    @staticmethod
    def load_persistables(executor, dirname, main_program=None, filename=None):
        from .. import static

        static.load(main_program or static.default_main_program(), os.path.join(dirname, filename or "persistables"), executor)

    @staticmethod
    def is_persistable(var):
        return bool(getattr(var, "persistable", False) or getattr(var, "__dict__", {}).get("_pd_persistable", False))


io = _IO()
