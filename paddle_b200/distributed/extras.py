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


def _slot_type(v):
    """'float' | 'uint64' of a use_var entry (static Variable / Tensor with a dtype, (name, dtype) pair, or dtype string)."""
    d = v[1] if isinstance(v, (tuple, list)) else getattr(v, "dtype", v)
    return "float" if "float" in str(d) else "uint64"


def _slot_name(v, i):
    if isinstance(v, (tuple, list)):
        return v[0]
    return getattr(v, "name", None) or f"slot_{i}"


class InMemoryDataset:
    """Slot dataset held in host memory (PS-mode CTR pipelines). Parity: distributed/fleet/dataset/dataset.py:InMemoryDataset over
    paddle/fluid/framework/data_set.cc + data_feed.cc.

    With `use_var` (the feed variables, in slot order) files are in the multi-slot text format that
    `fleet.MultiSlotDataGenerator` emits and are parsed / stored / shuffled / batched by the native columnar feed
    (`csrc/runtime/data_feed.cpp`, multi-threaded, GIL released): a batch is `{name: (values, lod)}`, or `{name: [B, n] array}` for
    slots whose records all have the same length. With `parse_fn` (or no `use_var`) lines go through Python instead."""

    def __init__(self):
        self.batch_size, self.filelist, self.parse_fn, self._rows, self.thread_num = 1, [], None, [], 1
        self.use_var, self._feed, self.pipe_command = [], None, None

    def init(self, batch_size=1, thread_num=1, use_var=None, pipe_command=None, input_type=0, fs_name="", fs_ugi="", download_cmd="cat", parse_fn=None, **kw):
        self.batch_size, self.thread_num, self.use_var, self.parse_fn = batch_size, thread_num, list(use_var or []), parse_fn
        self.pipe_command = None if pipe_command in (None, "cat") else pipe_command

    def update_settings(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def set_filelist(self, filelist):
        self.filelist = list(filelist)

    def _native(self):
        if self.parse_fn is not None or not self.use_var:
            return None
        if self._feed is None:
            from .. import _build

            C = _build.load(required=False)
            if C is None:
                return None
            self._feed = C.MultiSlotFeed([_slot_type(v) for v in self.use_var], max(1, int(self.thread_num)))
        return self._feed

    def _piped_lines(self, fn):
        """Lines of `fn` after `pipe_command` (a shell filter, e.g. the data generator script)."""
        import subprocess

        with open(fn, "rb") as f:
            out = subprocess.run(self.pipe_command, shell=True, stdin=f, capture_output=True, check=True).stdout
        return out.decode().splitlines()

    def _parse(self, line):
        if self.parse_fn is not None:
            return self.parse_fn(line)
        return [float(x) for x in line.split()]

    def load_into_memory(self, is_shuffle=False):
        feed = self._native()
        if feed is not None:
            feed.clear()
            if self.pipe_command:
                for fn in self.filelist:
                    feed.load_lines(self._piped_lines(fn))
            else:
                feed.load(self.filelist)
        else:
            self._rows = []
            for fn in self.filelist:
                with open(fn) as f:
                    self._rows.extend(self._parse(l) for l in f if l.strip())
        if is_shuffle:
            self.local_shuffle()

    def preload_into_memory(self, thread_num=None):
        import threading

        self._preload = threading.Thread(target=self.load_into_memory, daemon=True)
        self._preload.start()

    def wait_preload_done(self):
        t = getattr(self, "_preload", None)
        if t is not None:
            t.join()

    def local_shuffle(self, seed=None):
        import random

        if self._feed is not None:
            self._feed.shuffle(random.getrandbits(63) if seed is None else seed)
        else:
            random.shuffle(self._rows)

    def global_shuffle(self, fleet=None, thread_num=12, seed=1234):
        """Every trainer ends up with a random 1/N of the global data. Native feed: all trainers loaded the same file list, so a
        shuffle with a shared seed followed by keeping positions `i % N == rank` needs no traffic at all; python rows are
        exchanged with all_gather_object."""
        world = dist.get_world_size() if dist.is_initialized() else 1
        if self._feed is not None:
            self._feed.shuffle(seed)
            if world > 1:
                self._feed.keep_partition(dist.get_rank(), world)
            return
        if world > 1:
            n, me = world, dist.get_rank()
            buckets = [[r for i, r in enumerate(self._rows) if hash((me, i)) % n == d] for d in range(n)]
            gathered = [None] * n
            dist.all_gather_object(gathered, buckets)
            self._rows = [r for g in gathered for r in g[me]]
        self.local_shuffle()

    def get_memory_data_size(self, fleet=None):
        return self._feed.size() if self._feed is not None else len(self._rows)

    def get_shuffle_data_size(self, fleet=None):
        return self.get_memory_data_size()

    def release_memory(self):
        self._rows = []
        if self._feed is not None:
            self._feed.clear()

    def _native_batches(self, feed):
        names = [_slot_name(v, i) for i, v in enumerate(self.use_var)]
        for start in range(0, feed.size(), self.batch_size):
            out = {}
            for name, (vals, lod) in zip(names, feed.batch(start, self.batch_size)):
                lens = lod[1:] - lod[:-1]
                n = int(lens[0]) if lens.numel() else 0
                if lens.numel() and bool((lens == n).all()):
                    out[name] = vals.reshape(lens.numel(), n).numpy()
                else:
                    out[name] = (vals.numpy(), lod.numpy())
            yield out

    def __iter__(self):
        import numpy as np

        if self._feed is not None:
            yield from self._native_batches(self._feed)
            return
        for i in range(0, len(self._rows), self.batch_size):
            yield np.asarray(self._rows[i:i + self.batch_size], dtype=np.float32)


class QueueDataset(InMemoryDataset):
    """Streaming variant: files are read lazily (one file at a time through the native feed), nothing else is kept in memory."""

    def load_into_memory(self, is_shuffle=False):
        raise RuntimeError("QueueDataset streams from files; load_into_memory is not available")

    def __iter__(self):
        import numpy as np

        feed = self._native()
        if feed is not None:
            for fn in self.filelist:
                feed.clear()
                feed.load_lines(self._piped_lines(fn)) if self.pipe_command else feed.load([fn])
                yield from self._native_batches(feed)
            feed.clear()
            return
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
