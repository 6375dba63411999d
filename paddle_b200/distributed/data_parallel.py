"""paddle.DataParallel. Parity: python/paddle/distributed/parallel.py:DataParallel +
paddle/fluid/distributed/collective/reducer.cc (EagerReducer: bucketing, ready-marking hooks, overlap, no_sync).

B200 design: gradients live in a flat arena, so a bucket is a [start,end) range of the gradient slab - no coalesce /
split copies.  When the last gradient of a bucket has been accumulated, its range is all-reduced on a side stream
while backward keeps running; the reduction itself is the peer-memory kernel (csrc/comm/p2p_collectives.cu) when the
symmetric heap is available, NCCL/gloo otherwise.
"""
from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist

from ..nn.layer import Layer
from ..parallel.arena import ParamArena
from . import collective as C
from . import env


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


class DataParallel(Layer):
    def __init__(self, layers, strategy=None, comm_buffer_size=25, last_comm_buffer_size=1, find_unused_parameters=False, group=None):
        super().__init__()
        self._layers = layers
        self.find_unused_parameters = find_unused_parameters
        self.group = group
        self._pg = group.pg if isinstance(group, C.Group) else group
        if not env.is_initialized():
            env.init_parallel_env()
        self._world = dist.get_world_size(self._pg) if env.is_initialized() else 1
        self._sync = True
        self._bucket_bytes = int(comm_buffer_size) << 20
        params = [p for p in layers.parameters() if not p.stop_gradient]
        self._params = params
        if self._world > 1:
            self._broadcast_initial()
        self._arena = None
        self._buckets = []
        self._pending = {}
        self._works = []
        self._stream = None
        if self._world > 1 and params:
            self._setup_reducer()

    def _broadcast_initial(self):
        src = dist.get_global_rank(self._pg, 0) if self._pg is not None else 0
        with torch.no_grad():
            for t in list(self._layers.parameters()) + list(self._layers.buffers()):
                dist.broadcast(_raw(t), src=src, group=self._pg)

    def _setup_reducer(self):
        existing = getattr(self._params[0], "__dict__", {}).get("_arena_grad") is not None
        self._arena = ParamArena(self._params, grad_allocator=self._symm_grad_allocator()) if not existing else None
        if self._arena is None:  # an optimizer already flattened them: fall back to per-parameter reduction
            self._buckets = []
            return
        self._arena.zero_grad()
        self._buckets = self._arena.buckets(self._bucket_bytes)
        # map each param to the buckets it overlaps
        self._bucket_remaining = []
        self._param_buckets = {}
        for bi, (slab, start, end) in enumerate(self._buckets):
            cnt = 0
            for p in slab.params:
                o, n = slab.offsets[p.name]
                if o < end and o + n > start:
                    self._param_buckets.setdefault(id(p), []).append(bi)
                    cnt += 1
            self._bucket_remaining.append(cnt)
        self._remaining = list(self._bucket_remaining)
        for p in self._params:
            p.register_post_accumulate_grad_hook(self._make_hook(p))
        if self._params[0].is_cuda:
            from ..device import side_stream

            self._stream = side_stream()

    def _symm_grad_allocator(self):
        """Gradient slabs are placed in the symmetric peer heap when they fit: the bucket all-reduce then runs in place over
        NVLink (two-shot peer-memory kernel) with no staging copies."""
        if not self._params[0].is_cuda:
            return None
        from ..parallel import symm

        ctx = symm.context_for(self.group if self.group is not None else self._pg)
        if ctx is None:
            return None
        need = sum((p.numel() + 127) // 128 * 128 * p.element_size() for p in self._params)
        if need > (ctx.heap.size() - ctx.heap.cursor()) // 2:
            return None

        def alloc(n, dt):
            t, _ = ctx.buffer(("dp_grad", id(self), n), (n,), dt)
            t.zero_()
            return t

        return alloc

    def _make_hook(self, p):
        def hook(_param):
            if not self._sync:
                return
            # keep .grad pointing at the arena view (autograd may have materialised a fresh tensor for the first grad)
            g = torch.Tensor.grad.__get__(p)
            view = p.__dict__["_arena_grad"]
            if g is not None and g.data_ptr() != view.data_ptr():
                view.copy_(g)
                torch.Tensor.grad.__set__(p, view)
            for bi in self._param_buckets.get(id(p), []):
                self._remaining[bi] -= 1
                if self._remaining[bi] == 0:
                    self._launch(bi)
            if not self._callback_queued:
                self._callback_queued = True
                torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

        return hook

    _callback_queued = False

    def _launch(self, bi):
        slab, start, end = self._buckets[bi]
        buf = slab.grad[start:end]
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                self._reduce(buf)
        else:
            self._reduce(buf)

    def _reduce(self, buf):
        from .fleet.hybrid import _allreduce_flat

        _allreduce_flat(buf, self.group if self.group is not None else self._pg)
        buf.mul_(1.0 / self._world)

    def _finalize(self):
        # buckets whose params produced no grad this step (unused parameters) still have to be reduced for consistency
        for bi, rem in enumerate(self._remaining):
            if rem > 0 and (self.find_unused_parameters or rem < self._bucket_remaining[bi]):
                self._launch(bi)
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        self._remaining = list(self._bucket_remaining)
        self._callback_queued = False

    def forward(self, *inputs, **kwargs):
        return self._layers(*inputs, **kwargs)

    def init_reducer(self):
        """(Re)build the gradient buckets, e.g. after parameters were added or frozen. Parity: parallel.py:DataParallel.init_reducer."""
        self._params = [p for p in self._layers.parameters() if not p.stop_gradient]
        if self._world > 1 and self._params:
            self._setup_reducer()

    @contextlib.contextmanager
    def no_sync(self):
        old = self._sync
        self._sync = False
        try:
            yield
        finally:
            self._sync = old

    def scale_loss(self, loss):
        return loss

    def apply_collective_grads(self):
        """Manual gradient reduction (used after no_sync accumulation when hooks were disabled)."""
        if self._world <= 1:
            return
        if self._arena is not None:
            for s in self._arena.all_slabs():
                self._reduce(s.grad)
        else:
            from .fleet.hybrid import _allreduce_tensors

            grads = [torch.Tensor.grad.__get__(p) for p in self._params if torch.Tensor.grad.__get__(p) is not None]
            _allreduce_tensors(grads, self.group if self.group is not None else self._pg, 1.0 / self._world)

    def parameters(self, include_sublayers=True):
        return self._layers.parameters(include_sublayers)

    def named_parameters(self, prefix="", include_sublayers=True):
        return self._layers.named_parameters(prefix, include_sublayers)

    def state_dict(self, *a, **k):
        return self._layers.state_dict(*a, **k)

    def set_state_dict(self, *a, **k):
        return self._layers.set_state_dict(*a, **k)

    set_dict = set_state_dict
    load_dict = set_state_dict
