"""Flat parameter / gradient arenas.

B200-first memory layout: every parameter of a (dtype) class is a view into one contiguous slab and every gradient a
view into a matching slab.  Consequences:
  * the optimizer is one fused kernel launch per slab (csrc/optim.cu) instead of one per tensor;
  * data-parallel / sharding buckets are just [offset, offset+len) ranges of the gradient slab -> zero-copy
    all-reduce / reduce-scatter (no coalesce + split like the reference's EagerReducer, reducer.cc);
  * slabs can be placed in the symmetric peer heap so fused P2P collectives address them directly.
Parity (role): paddle/fluid/distributed/collective/reducer.cc (grad bucketing), GradStorage/ParamStorage in
python/paddle/distributed/fleet/meta_parallel/sharding/group_sharded_storage.py.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

_ALIGN = 128  # elements; keeps every view 256B-aligned for 16B vector access and TMA


class Slab:
    def __init__(self, dtype, device, params, grad_dtype=None, allocator=None, grad_allocator=None):
        self.dtype, self.device = dtype, device
        self.grad_dtype = grad_dtype or dtype
        self.params = list(params)
        self.offsets = OrderedDict()
        off = 0
        for p in self.params:
            self.offsets[p.name] = (off, p.numel())
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        alloc = allocator or (lambda n, dt: torch.zeros(n, dtype=dt, device=device))
        self.data = alloc(self.numel, dtype)
        self.grad = (grad_allocator or alloc)(self.numel, self.grad_dtype)
        with torch.no_grad():
            for p in self.params:
                o, n = self.offsets[p.name]
                view = self.data[o:o + n].view(tuple(p.size()))
                view.copy_(p.as_subclass(torch.Tensor))
                p.data = view
                gview = self.grad[o:o + n].view(tuple(p.size()))
                old = torch.Tensor.grad.__get__(p)
                if old is not None:
                    gview.copy_(old)
                if gview.dtype == p.dtype:
                    torch.Tensor.grad.__set__(p, gview)
                p.__dict__["_arena_grad"] = gview
        self.master = None
        self.state = {}

    def param_view(self, p):
        o, n = self.offsets[p.name]
        return self.data[o:o + n]

    def grad_view(self, p):
        o, n = self.offsets[p.name]
        return self.grad[o:o + n]


class ParamArena:
    """Groups parameters by (dtype, device, decay-class) into slabs."""

    def __init__(self, params, group_fn=None, grad_dtype=None, allocator=None, grad_allocator=None):
        params = [p for p in params if not p.stop_gradient]
        groups = OrderedDict()
        for p in params:
            key = (p.dtype, p.device, group_fn(p) if group_fn else 0)
            groups.setdefault(key, []).append(p)
        self.slabs = OrderedDict()
        for key, ps in groups.items():
            self.slabs[key] = Slab(key[0], key[1], ps, grad_dtype=grad_dtype, allocator=allocator, grad_allocator=grad_allocator)

    def zero_grad(self):
        for s in self.slabs.values():
            s.grad.zero_()
            for p in s.params:
                g = p.__dict__["_arena_grad"]
                if g.dtype == p.dtype and torch.Tensor.grad.__get__(p) is not g:
                    torch.Tensor.grad.__set__(p, g)

    def all_slabs(self):
        return list(self.slabs.values())

    def numel(self):
        return sum(s.numel for s in self.slabs.values())

    def buckets(self, bucket_bytes):
        """[(slab, start, end)] contiguous gradient ranges of ~bucket_bytes, in reverse (backward) order."""
        out = []
        for s in self.slabs.values():
            per = max(_ALIGN, bucket_bytes // s.grad.element_size() // _ALIGN * _ALIGN)
            end = s.numel
            while end > 0:
                start = max(0, end - per)
                out.append((s, start, end))
                end = start
        return out
