"""Symmetric peer memory contexts (per process group) + the fused compute/collective kernels built on them.

The native heap lives in csrc/runtime/symm_heap.cpp; the kernels in csrc/comm/*.cu.  A context is created lazily per
group the first time a fused op runs on it; creation is collective (IPC handle exchange through torch.distributed).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

_contexts = {}
_disabled = [False]


def _pg(group):
    return getattr(group, "pg", group)


def available():
    if _disabled[0] or not torch.cuda.is_available() or os.environ.get("B200_DISABLE_SYMM", "0") == "1":
        return False
    from .._build import load

    m = load()
    return m is not None and hasattr(m, "SymmHeap")


def context_for(group, heap_bytes=None):
    """SymmContext for `group`, or None when fused P2P collectives are not usable. `heap_bytes` sizes the heap when this
    call creates the context (e.g. to hold a whole gradient slab); it must be the same on every rank of the group."""
    if not available():
        return None
    key = id(_pg(group)) if group is not None else 0
    ctx = _contexts.get(key)
    if ctx is None:
        try:
            ctx = SymmContext(group, heap_bytes=heap_bytes)
        except Exception as e:  # noqa: BLE001
            import warnings

            warnings.warn(f"paddle_b200: symmetric-memory setup failed ({e}); using NCCL collectives")
            _disabled[0] = True
            return None
        _contexts[key] = ctx
    return ctx


class SymmContext:
    """Symmetric heap + signal pad for one process group (all ranks on one NVSwitch domain)."""

    def __init__(self, group, heap_bytes=None, signal_bytes=1 << 20):
        from .._build import ext

        self.group = group
        pg = _pg(group)
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        heap_bytes = heap_bytes or int(os.environ.get("B200_SYMM_HEAP_MB", "2048")) << 20
        self.dev = torch.cuda.current_device()
        self.heap = ext().SymmHeap(heap_bytes, signal_bytes, self.dev)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(self.heap.ipc_handle()), group=pg)
        self.heap.open_peers([bytes(h) for h in handles], self.rank)
        dist.barrier(group=pg, device_ids=[self.dev]) if dist.get_backend(pg) == "nccl" else dist.barrier(group=pg)
        self._base_cursor = self.heap.cursor()
        self._bufs = {}
        self._epoch = 0
        self.ext = ext()

    # symmetric scratch buffers keyed by (tag, nbytes): allocated once, identical offsets on every rank
    def buffer(self, tag, shape, dtype):
        nbytes = 1
        for s in shape:
            nbytes *= int(s)
        nbytes *= torch.empty(0, dtype=dtype).element_size()
        key = (tag, nbytes)
        off = self._bufs.get(key)
        if off is None:
            off = self.heap.alloc(nbytes, 1024)
            self._bufs[key] = off
        return self.heap.tensor(off, list(shape), dtype, -1), off

    def release(self, tag=None):
        """Give symmetric scratch buffers back to the heap's best-fit allocator (all of them, or those of one tag). Collective in
        the same sense as `buffer`: every rank must release the same set so later offsets stay symmetric."""
        for key in [k for k in self._bufs if tag is None or k[0] == tag]:
            self.heap.free(self._bufs.pop(key))

    def memory_stats(self):
        return dict(self.heap.stats())

    def next_epoch(self):
        self._epoch += 1
        return self._epoch

    # ---- fused ops ------------------------------------------------------------------------------------------
    def max_message_bytes(self):
        return 1 << 62  # large messages are streamed through the staging window in chunks

    def owns(self, t):
        base = self.heap.base_ptr()
        return base <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= base + self.heap.size()

    def allreduce_(self, t):
        """In-place sum all-reduce of a contiguous CUDA tensor (two-shot over peer memory).
        Tensors that already live in the symmetric heap (flat gradient arenas placed there) are reduced in place with no
        staging copy; others stream through a staging window."""
        flat = t.reshape(-1)
        if self.owns(flat):
            self.heap.allreduce(flat.data_ptr() - self.heap.base_ptr(), flat.numel(), flat.dtype, self.next_epoch())
            return t
        window = int(os.environ.get("B200_SYMM_WINDOW_MB", "256")) << 20
        per = max(1, window // flat.element_size())
        buf, off = self.buffer("ar", (min(per, flat.numel()),), flat.dtype) if flat.numel() <= per else self.buffer("ar", (per,), flat.dtype)
        for lo in range(0, flat.numel(), per):
            hi = min(flat.numel(), lo + per)
            buf[: hi - lo].copy_(flat[lo:hi])
            self.heap.allreduce(off, hi - lo, flat.dtype, self.next_epoch())
            flat[lo:hi].copy_(buf[: hi - lo])
        return t

    def gemm_allreduce(self, x, w):
        """Y = sum_r X_r @ W_r.  GEMM writes into the symmetric buffer, then P2P two-shot all-reduce on the same stream."""
        from ..kernels import gemm as KG

        x2 = x.reshape(-1, x.shape[-1])
        m, n = x2.shape[0], w.shape[1]
        buf, off = self.buffer("gar", (m, n), x.dtype)
        KG.gemm(x2, w, out=buf)
        self.heap.allreduce(off, m * n, x.dtype, self.next_epoch())
        return buf.clone().reshape(*x.shape[:-1], n)

    def gemm_reduce_scatter(self, x, w, b_is_nk=False):
        """Y[S/p] = reduce_scatter_dim0(X @ op(W)): GEMM into symmetric memory, each rank reduces its own row slice
        by reading the peers' partial tiles over NVLink."""
        from ..kernels import gemm as KG

        x2 = x.reshape(-1, x.shape[-1])
        m = x2.shape[0]
        n = w.shape[0] if b_is_nk else w.shape[1]
        if os.environ.get("B200_RS_PUSH", "1") == "1" and m % self.world == 0 and n % 8 == 0 and KG._tc_ok(x2, w, False, b_is_nk):
            # single-pass fused path: the GEMM epilogue stores each output row into its owner's staging slot over NVLink while
            # the tensor cores work on the next tile; the tail kernel only sums `world` local slots
            rows = m // self.world
            _, off = self.buffer("grs_push", (self.world, rows, n), x.dtype)
            slot = rows * n * x.element_size()
            dst = [self.heap.peer_ptr(r) + off + self.rank * slot for r in range(self.world)]
            key = ("dummy", n, x.dtype)
            dummy = self._bufs.get(key)
            if dummy is None:
                dummy = self._bufs[key] = torch.empty((1, n), dtype=x.dtype, device=x.device)
            self.ext.gemm(x2, w, None, False, b_is_nk, 0, dummy, None, dst, rows)
            out = torch.empty((rows, n), dtype=x.dtype, device=x.device)
            self.heap.reduce_slots(off, out, rows * n, self.next_epoch())
            return out.reshape(x.shape[0] // self.world, *x.shape[1:-1], n)
        buf, off = self.buffer("grs", (m, n), x.dtype)
        KG.gemm(x2, w, b_is_nk=b_is_nk, out=buf)
        out = torch.empty((m // self.world, n), dtype=x.dtype, device=x.device)
        self.heap.reduce_scatter(off, out, m * n, self.next_epoch())
        return out.reshape(x.shape[0] // self.world, *x.shape[1:-1], n)

    def allgather_gemm(self, x, w, b_is_nk=False, return_gathered=False):
        """Y = all_gather_dim0(X) @ op(W): peers' shards are pulled over NVLink into the symmetric buffer while the GEMM
        on the local shard runs; then the remaining row blocks are multiplied."""
        from ..kernels import gemm as KG

        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        ml, k = x2.shape
        nout = w.shape[0] if b_is_nk else w.shape[1]
        if os.environ.get("B200_AG_FUSED", "1") == "1" and ml % 256 == 0 and k % 64 == 0 and nout >= 256 and KG._tc_ok(x2, w, False, b_is_nk):
            # one kernel: a copy warp per CTA pulls the peers' shards over NVLink (per-row-block flags) while the tensor cores
            # start on the local rows; the TMA producers wait on the flag of a row block before loading it.  Only this rank's
            # shard lives in symmetric memory (the peers read it there); the gathered operand is an ordinary fresh tensor, so it
            # can be handed to autograd for the weight-gradient GEMM without a copy.
            shard, soff = self.buffer("agg_shard", (ml, k), x.dtype)
            shard.copy_(x2)
            gathered = torch.empty((ml * self.world, k), dtype=x.dtype, device=x.device)
            gathered[self.rank * ml:(self.rank + 1) * ml].copy_(x2)
            src = [self.heap.peer_ptr(r) + soff for r in range(self.world)]
            pads = [self.heap.peer_ptr(r) for r in range(self.world)]
            nblk = (ml * self.world) // 128
            fkey = ("agflags", nblk)
            flags = self._bufs.get(fkey)
            if flags is None:
                flags = self._bufs[fkey] = torch.zeros(nblk + 1, dtype=torch.int32, device=x.device)
            flags.zero_()
            y = self.ext.gemm(gathered, w, None, False, b_is_nk, 0, None, None, [], 0, src, pads, flags, self.rank, ml, self.next_epoch())
            y = y.reshape(x.shape[0] * self.world, *x.shape[1:-1], nout)
            if return_gathered:
                return y, gathered.reshape(x.shape[0] * self.world, *x.shape[1:])
            return y
        full, off = self.buffer("agg", (ml * self.world, k), x.dtype)
        full[self.rank * ml:(self.rank + 1) * ml].copy_(x2)
        self.heap.allgather(off, ml * k * x.element_size(), self.next_epoch())
        y = KG.gemm(full, w, b_is_nk=b_is_nk)
        n = y.shape[-1]
        y = y.reshape(x.shape[0] * self.world, *x.shape[1:-1], n)
        if return_gathered:
            return y, full.clone().reshape(x.shape[0] * self.world, *x.shape[1:])
        return y



def _a2av(self, src, in_splits, out_rows, cap, gather=None, tag="a2av"):
    """Variable all-to-all of rows over peer memory (one push kernel; MoE dispatch when `gather` maps sorted slots to source rows).
    in_splits[r]: rows sent to rank r (slots grouped by destination); out_rows: rows this rank receives; cap: receive-buffer
    capacity in rows, IDENTICAL on every rank (symmetric allocation) - see a2av_capacity. Returns [out_rows, H] ordered by source."""
    h = src.shape[-1]
    recv, roff = self.buffer((tag, "recv"), (cap, h), src.dtype)
    meta, moff = self.buffer((tag, "meta"), (self.world,), torch.int64)
    meta.copy_(torch.as_tensor(list(in_splits), dtype=torch.int64))
    self.heap.a2av(src.contiguous(), gather, moff, roff, max(int(sum(in_splits)), 1), self.next_epoch())
    return recv[:int(out_rows)].clone()


def _a2av_capacity(self, rows_in, rows_out, x):
    """Receive-buffer capacity (rows) agreed by all ranks with one tiny MAX all-reduce; 0 = does not fit -> NCCL path everywhere."""
    h = x.shape[-1]
    t = torch.tensor([max(int(rows_in), int(rows_out), 1)], device=x.device, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_pg(self.group))
    need = int(t.item())
    cap = 1024
    while cap < 2 * need:
        cap *= 2
    limit = (self.heap.size() - self._base_cursor) // 8
    return cap if cap * h * x.element_size() <= limit else 0


def _a2av_dev(self, src, in_splits_dev, cap, gather=None, tag="a2av", rows_hint=None):
    """Variable all-to-all whose split sizes stay on the device (no host synchronisation): in_splits_dev = int64 [world] CUDA tensor of
    rows sent to every rank (slots grouped by destination).  Returns the whole symmetric receive buffer [cap, H]; the rows beyond what
    the peers sent are stale - the caller masks them with the counts it exchanged (see exchange_counts)."""
    h = src.shape[-1]
    recv, roff = self.buffer((tag, "recv"), (int(cap), h), src.dtype)
    meta, moff = self.buffer((tag, "meta"), (self.world,), torch.int64)
    meta.copy_(in_splits_dev.reshape(-1).to(torch.int64))
    hint = int(rows_hint if rows_hint is not None else (gather.numel() if gather is not None else src.shape[0]))
    self.heap.a2av(src.contiguous(), gather, moff, roff, max(hint, 1), self.next_epoch())
    return recv


def _exchange_counts(self, counts, tag="a2a_counts"):
    """counts: int64 [world, n] on the device (row r = what this rank sends to rank r). Returns [world, n]: row s = what rank s sends
    to this rank.  One tiny peer-memory all-to-all; nothing touches the host."""
    w, n = counts.shape
    npad = (n + 1) // 2 * 2                      # 16-byte chunks
    send, soff = self.buffer((tag, "send", npad), (w, npad), torch.int64)
    recv, roff = self.buffer((tag, "recv", npad), (w, npad), torch.int64)
    send.zero_()
    send[:, :n].copy_(counts)
    self.heap.alltoall(soff, roff, npad * 8, self.next_epoch())
    return recv[:, :n].clone()


SymmContext.a2av_dev = _a2av_dev
SymmContext.exchange_counts = _exchange_counts
SymmContext.a2av = _a2av
SymmContext.a2av_capacity = _a2av_capacity
