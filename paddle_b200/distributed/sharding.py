"""Group-sharded (ZeRO) training. Parity: python/paddle/distributed/sharding/group_sharded.py
(group_sharded_parallel, save_group_sharded_model) and fleet/meta_parallel/sharding/
(GroupShardedOptimizerStage2, GroupShardedStage2 :411 _get_reduce_fn / :705 task wait, GroupShardedStage3 :557-612 forward hooks with
prefetch, :857-934 sync_comm, GroupShardedScaler), DygraphShardingOptimizer.

B200 design (flat arenas, one communication stream per sharding instance):

  stage 2 ('os' / 'os_g'): parameters and gradients live in flat slabs laid out as BUCKETS (~B200_SHARD_BUCKET_MB each, padded to a
      multiple of world); rank r owns the r-th 1/world slice of every bucket.  A post-accumulate hook per parameter counts a bucket
      down during backward; the moment a bucket is complete its reduce-scatter is launched on the communication stream, so it
      overlaps the rest of the backward pass.  With the symmetric peer heap the slabs sit inside it and the reduce-scatter /
      all-gather kernels of csrc/comm/p2p_collectives.cu run IN PLACE over NVLink (no staging copies, no NCCL).  step(): one
      global-norm reduction, ONE fused AdamW launch per slab on the owned slices (split master weights, csrc/optim.cu), then the
      owners publish their slices bucket by bucket.
  stage 3 ('p_g_os'): every parameter-owning sublayer is a unit whose flat parameters are sharded 1/world; all shards of the model
      live in one flat shard arena (inside the symmetric heap when available).  A unit's full weights are pulled from the peers'
      shards right before its forward / backward (p2p_gather_pull) and dropped right after; the NEXT unit in execution order is
      prefetched on the communication stream while the current one computes.  Gradients of a unit are flattened and
      reduce-scattered on the communication stream as soon as its backward finished.  step(): global-norm clip over the shards, one
      fused AdamW launch over the whole shard arena.

CPU / gloo (tests): same code path with synchronous torch.distributed collectives.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from ..nn.layer import Layer
from ..tensor import Tensor
from . import collective as C
from . import env

_ALIGN = 128


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _pg(group):
    return group.pg if isinstance(group, C.Group) else group


def _world(group):
    return dist.get_world_size(_pg(group)) if env.is_initialized() else 1


def _rank(group):
    return dist.get_rank(_pg(group)) if env.is_initialized() else 0


def _backend(group):
    return dist.get_backend(_pg(group)) if env.is_initialized() else "none"


def _symm_ctx(group, device, heap_bytes=None):
    """Symmetric-heap context of the sharding group (None on CPU / when peer memory is off)."""
    if device.type != "cuda" or _world(group) <= 1:
        return None
    from ..framework.flags import flag

    if not flag("FLAGS_b200_p2p_collectives", True):
        return None
    from ..parallel import symm

    return symm.context_for(group, heap_bytes=heap_bytes)


def _ops(device):
    from ..optimizer.optimizer import _NativeOps, _TorchOps

    return _NativeOps() if device.type == "cuda" else _TorchOps()


class _Comm:
    """Collectives of one sharding instance, serialised on one side stream (the peer-memory kernels of a heap share its signal pad,
    and both ranks must issue them in the same order - hook order is deterministic)."""

    def __init__(self, group, device):
        self.group, self.device = group, device
        self.world, self.rank = _world(group), _rank(group)
        self.cuda = device.type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.ctx = None       # set by the owner once the heap size is known

    def on_stream(self):
        import contextlib

        if not self.cuda:
            return contextlib.nullcontext()
        self.stream.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.stream)

    def join(self):
        """The compute stream waits for everything issued on the communication stream so far."""
        if self.cuda:
            from . import comm_timer as CT

            with CT.region("sharding_comm_wait"):
                torch.cuda.current_stream().wait_stream(self.stream)

    def reduce_scatter(self, out, flat, heap_off=None):
        """out = sum over ranks of flat[rank * n : (rank + 1) * n] (unscaled sum). heap_off: byte offset of `flat` inside the heap."""
        if self.world == 1:
            out.copy_(flat[: out.numel()])
            return
        if self.ctx is not None and heap_off is not None:
            self.ctx.heap.reduce_scatter(heap_off, out, flat.numel(), self.ctx.next_epoch())
            return
        pg = _pg(self.group)
        if _backend(self.group) == "gloo":
            tmp = flat.clone()
            dist.all_reduce(tmp, group=pg)
            n = out.numel()
            out.copy_(tmp[self.rank * n:(self.rank + 1) * n])
        elif self.ctx is not None:
            from ..parallel import symm_ops

            if not symm_ops.reduce_scatter_into(self.ctx, out, flat):
                dist.reduce_scatter_tensor(out, flat, group=pg)
        else:
            dist.reduce_scatter_tensor(out, flat, group=pg)

    def all_gather_inplace(self, flat, heap_off=None):
        """flat = [world, n]: every rank has written its own row; fill in the others."""
        if self.world == 1:
            return
        n = flat.numel() // self.world
        if self.ctx is not None and heap_off is not None:
            self.ctx.heap.allgather(heap_off, n * flat.element_size(), self.ctx.next_epoch())
            return
        pg = _pg(self.group)
        mine = flat[self.rank * n:(self.rank + 1) * n].clone()
        if _backend(self.group) == "gloo":
            dist.all_gather(list(flat.chunk(self.world)), mine, group=pg)
        else:
            dist.all_gather_into_tensor(flat, mine, group=pg)

    def gather_from_shards(self, full, shard, shard_heap_off=None):
        """full = [world, n] fresh local tensor <- every rank's `shard` (n elements)."""
        if self.world == 1:
            full.copy_(shard)
            return
        if self.ctx is not None and shard_heap_off is not None:
            self.ctx.heap.gather_pull(full, shard_heap_off, shard.numel() * shard.element_size(), self.ctx.next_epoch())
            return
        pg = _pg(self.group)
        if _backend(self.group) == "gloo":
            dist.all_gather(list(full.chunk(self.world)), shard.contiguous(), group=pg)
        else:
            dist.all_gather_into_tensor(full, shard.contiguous(), group=pg)


# =====================================================================================================================================
# stage 2
# =====================================================================================================================================
class _Bucket:
    __slots__ = ("slab", "start", "numel", "params", "pending", "reduced", "shard_off")

    def __init__(self, slab, start, numel, params, shard_off):
        self.slab, self.start, self.numel, self.params, self.shard_off = slab, start, numel, params, shard_off
        self.pending, self.reduced = len(params), False


class _ShardedArena:
    """Flat parameter / gradient slabs laid out as buckets; rank r owns slice r of every bucket ("shard space" = those slices
    concatenated in bucket order)."""

    def __init__(self, params, group, bucket_bytes):
        self.group = group
        self.world, self.rank = _world(group), _rank(group)
        self.params = [p for p in params if not p.stop_gradient]
        by = {}
        for p in self.params:
            by.setdefault((p.dtype, p.device), []).append(p)
        self.slabs = []
        need = 0
        plans = []
        for (dt, dev), ps in by.items():
            es = torch.empty(0, dtype=dt).element_size()
            per_bucket = max(self.world * _ALIGN, bucket_bytes // es)
            buckets, cur, cur_n = [], [], 0
            for p in reversed(ps):                       # late layers finish their backward first -> bucket 0 fills first
                cur.append(p)
                cur_n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
                if cur_n >= per_bucket:
                    buckets.append(cur)
                    cur, cur_n = [], 0
            if cur:
                buckets.append(cur)
            plans.append((dt, dev, es, buckets))
            for b in buckets:
                n = sum((p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN for p in b)
                need += 2 * ((n + self.world * _ALIGN - 1) // (self.world * _ALIGN) * (self.world * _ALIGN)) * es
        dev0 = self.params[0].device if self.params else torch.device("cpu")
        self.ctx = _symm_ctx(group, dev0, heap_bytes=need + (320 << 20))
        if self.ctx is not None and need > self.ctx.heap.size() - self.ctx.heap.cursor() - (256 << 20):
            self.ctx = None                                # an existing (smaller) heap of this group: fall back to ordinary slabs
        for dt, dev, es, buckets in plans:
            offs, off, blist = {}, 0, []
            layout = []
            for b in buckets:
                start = off
                for p in b:
                    offs[p.name] = (off, p.numel())
                    off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
                k = off - start
                k = (k + self.world * _ALIGN - 1) // (self.world * _ALIGN) * (self.world * _ALIGN)
                off = start + k
                layout.append((start, k, b))
            total = off
            if self.ctx is not None:
                data, doff = self.ctx.buffer(("s2_data", len(self.slabs)), (total,), dt)
                grad, goff = self.ctx.buffer(("s2_grad", len(self.slabs)), (total,), dt)
                data.zero_()
                grad.zero_()
            else:
                data, grad = torch.zeros(total, dtype=dt, device=dev), torch.zeros(total, dtype=dt, device=dev)
                doff = goff = None
            slab = dict(dtype=dt, device=dev, params=[p for b in buckets for p in b], offsets=offs, data=data, grad=grad, doff=doff, goff=goff,
                        es=es, numel=total, shard=total // self.world, state={}, master=None, index=len(self.slabs))
            soff = 0
            for start, k, b in layout:
                blist.append(_Bucket(slab, start, k, b, soff))
                soff += k // self.world
            slab["buckets"] = blist
            with torch.no_grad():
                for p in slab["params"]:
                    o, n = offs[p.name]
                    v = data[o:o + n].view(tuple(p.size()))
                    v.copy_(_raw(p))
                    p.data = v
                    gv = grad[o:o + n].view(tuple(p.size()))
                    torch.Tensor.grad.__set__(p, gv)
                    p.__dict__["_arena_grad"] = gv
            self.slabs.append(slab)

    def owned(self, slab, which, bucket):
        """This rank's slice of a bucket inside the full slab `which` ('data' / 'grad')."""
        per = bucket.numel // self.world
        lo = bucket.start + self.rank * per
        return slab[which][lo:lo + per]

    def zero_grad(self):
        for s in self.slabs:
            s["grad"].zero_()
            for p in s["params"]:
                if torch.Tensor.grad.__get__(p) is not p.__dict__["_arena_grad"]:
                    torch.Tensor.grad.__set__(p, p.__dict__["_arena_grad"])


class GroupShardedOptimizerStage2:
    """Optimizer-state (+ gradient) sharding over bucketed flat arenas. Parity: group_sharded_optimizer_stage2.py."""

    def __init__(self, params, optim, group=None, offload=False, device="gpu", pretrain_sync_models=True, dp_group=None, **kw):
        self._optim = optim
        self.group = group
        self.world, self.rank = _world(group), _rank(group)
        self._params = list(params)
        self.offload = offload
        name = type(optim).__name__
        if name not in ("Adam", "AdamW", "SGD", "Momentum"):
            raise NotImplementedError(f"GroupShardedOptimizerStage2 supports Adam / AdamW / SGD / Momentum inner optimizers, got {name}")
        if name == "Adam" and float(getattr(optim, "_weight_decay", 0.0) or 0.0) != 0.0 and not getattr(optim, "_decoupled", False):
            raise NotImplementedError("GroupShardedOptimizerStage2: Adam with coupled (L2) weight decay is not supported; use AdamW")
        if pretrain_sync_models and self.world > 1:
            src = dist.get_global_rank(_pg(group), 0) if _pg(group) is not None else 0
            with torch.no_grad():
                for p in self._params:
                    dist.broadcast(_raw(p), src=src, group=_pg(group))
        bucket_bytes = int(float(os.environ.get("B200_SHARD_BUCKET_MB", "64")) * (1 << 20))
        self.arena = _ShardedArena(self._params, group, bucket_bytes)
        dev = self.arena.slabs[0]["device"] if self.arena.slabs else torch.device("cpu")
        self.comm = _Comm(group, dev)
        self.comm.ctx = self.arena.ctx
        self._overlap = False          # GroupShardedStage2 turns the backward hooks on
        self._step = 0
        self._aux = {}
        for slab in self.arena.slabs:
            slab["g_shard"] = torch.zeros(slab["shard"], dtype=slab["dtype"], device=slab["device"])
            slab["g_shard_valid"] = False

    @property
    def _parameter_list(self):
        return self._params

    # ---- gradient reduction ----------------------------------------------------------------------------------------------------
    def _reduce_bucket(self, b):
        """Reduce-scatter one complete bucket on the communication stream (overlaps the rest of the backward pass)."""
        slab = b.slab
        per = b.numel // self.world
        flat = slab["grad"][b.start:b.start + b.numel]
        dst = slab["g_shard"][b.shard_off:b.shard_off + per]
        hoff = slab["goff"] + b.start * slab["es"] if slab["goff"] is not None else None
        with self.comm.on_stream():
            if slab["g_shard_valid"] or b.reduced:       # a second backward before step(): accumulate the reduced gradients
                tmp = torch.empty_like(dst)
                self.comm.reduce_scatter(tmp, flat, hoff)
                dst.add_(tmp)
                if self.comm.cuda:
                    tmp.record_stream(self.comm.stream)
            else:
                self.comm.reduce_scatter(dst, flat, hoff)
            flat.zero_()                                  # consumed: the next backward accumulates from zero
        b.reduced = True
        b.pending = len(b.params)
        self._finished = False

    def _param_ready(self, b):
        b.pending -= 1
        if b.pending == 0:
            self._reduce_bucket(b)

    def _finish_reduction(self):
        if getattr(self, "_finished", False):            # e.g. GroupShardedScaler.unscale_ already completed this step's reduction
            self.comm.join()
            return
        for slab in self.arena.slabs:
            for b in slab["buckets"]:
                if not b.reduced:                         # no hooks (plain optimizer use), unused parameters, world 1
                    self._reduce_bucket(b)
                b.reduced = False
                b.pending = len(b.params)
            slab["g_shard_valid"] = True
        self._finished = True
        self.comm.join()

    # ---- update --------------------------------------------------------------------------------------------------------------------
    def _init_state(self, slab):
        st = slab["state"]
        if "p_shard" in st:
            return
        dev, n = slab["device"], slab["shard"]
        st["p_shard"] = torch.cat([self.arena.owned(slab, "data", b) for b in slab["buckets"]]) if slab["buckets"] else torch.zeros(0, dtype=slab["dtype"], device=dev)
        o = self._optim
        if type(o).__name__ in ("Adam", "AdamW"):
            mdt = getattr(o, "_moment_dtype", None) or torch.float32
            host = self.offload and dev.type == "cuda"
            st["m"] = torch.zeros(n, dtype=torch.float32 if host else mdt, device="cpu" if host else dev)
            st["v"] = torch.zeros_like(st["m"])
            if host:
                st["m"], st["v"] = st["m"].pin_memory(), st["v"].pin_memory()
                st["master"] = st["p_shard"].float().cpu().pin_memory()
                st["g_host"] = torch.zeros(n, dtype=torch.float32).pin_memory()
            elif slab["dtype"] != torch.float32:
                from ..framework.flags import flag

                if slab["dtype"] == torch.bfloat16 and flag("FLAGS_b200_split_master_weights", True):
                    st["master"] = torch.zeros(n, dtype=torch.int16, device=dev)
                else:
                    st["master"] = st["p_shard"].float()
            else:
                st["master"] = None
        elif slab["dtype"] != torch.float32:
            st["master"] = st["p_shard"].float()
        else:
            st["master"] = None

    def _update_slab(self, slab, sq, max_norm, inv_scale):
        o = self._optim
        st = slab["state"]
        p_sh, g_sh = st["p_shard"], slab["g_shard"]
        lr = o.get_lr()
        name = type(o).__name__
        if name in ("Adam", "AdamW"):
            b1, b2 = o._betas()
            wd = float(o._weight_decay or 0.0) if getattr(o, "_decoupled", False) else 0.0
            if "g_host" in st:     # offload: fp32 master + moments in pinned host memory, update on the CPU, bf16 shard travels
                st["g_host"].copy_(g_sh.float(), non_blocking=True)
                torch.cuda.current_stream().synchronize()
                gs = float(inv_scale.item()) if inv_scale is not None else 1.0
                if sq is not None and max_norm > 0:
                    norm = float(sq.sqrt().item()) * gs
                    if norm > max_norm:
                        gs *= max_norm / (norm + 1e-6)
                gf = st["g_host"] * gs
                pf = st["master"]
                pf.mul_(1 - lr * wd)
                st["m"].mul_(b1).add_(gf, alpha=1 - b1)
                st["v"].mul_(b2).addcmul_(gf, gf, value=1 - b2)
                denom = (st["v"] / (1 - b2 ** self._step)).sqrt_().add_(float(o._epsilon))
                pf.addcdiv_(st["m"], denom, value=-lr / (1 - b1 ** self._step))
                p_sh.copy_(pf, non_blocking=True)
                return
            _ops(slab["device"]).adamw_step(p_sh, g_sh, st["master"], st["m"], st["v"], lr, b1, b2, float(o._epsilon), wd, self._step,
                                            sq, max_norm, None, inv_scale)
            return
        # SGD / Momentum (plain tensor math on the shard)
        gs = float(inv_scale.item()) if inv_scale is not None else 1.0
        if sq is not None and max_norm > 0:
            norm = float(sq.sqrt().item()) * gs
            if norm > max_norm:
                gs *= max_norm / (norm + 1e-6)
        mom = getattr(o, "_momentum", 0.0)
        wdv = o._weight_decay
        wd = float(getattr(wdv, "coeff", wdv) or 0.0)
        pf = st["master"] if st["master"] is not None else p_sh
        gf = g_sh.float() * gs + wd * pf.float()
        if mom:
            if "vel" not in st:
                st["vel"] = torch.zeros_like(gf)
            st["vel"].mul_(mom).add_(gf)
            gf = st["vel"]
        pf.add_(gf.to(pf.dtype), alpha=-lr)
        if pf is not p_sh:
            p_sh.copy_(pf)

    def _publish(self, slab):
        """Owners write their updated slices into the full parameter slab and the peers fetch them, bucket by bucket."""
        st = slab["state"]
        for b in slab["buckets"]:
            per = b.numel // self.world
            self.arena.owned(slab, "data", b).copy_(st["p_shard"][b.shard_off:b.shard_off + per])
        if self.world > 1:
            for b in slab["buckets"]:
                hoff = slab["doff"] + b.start * slab["es"] if slab["doff"] is not None else None
                self.comm.all_gather_inplace(slab["data"][b.start:b.start + b.numel], hoff)

    @torch.no_grad()
    def step(self):
        self._step += 1
        self._finish_reduction()
        clip = self._optim._grad_clip
        dev = self.arena.slabs[0]["device"] if self.arena.slabs else torch.device("cpu")
        inv = self._aux.get("inv_world")
        if inv is None:
            inv = self._aux["inv_world"] = torch.full((1,), 1.0 / self.world, dtype=torch.float32, device=dev)
        sq, max_norm = None, 0.0
        if clip is not None and hasattr(clip, "clip_norm"):
            sq = torch.zeros(1, dtype=torch.float32, device=dev)
            E = _ops(dev)
            for slab in self.arena.slabs:
                E.grad_sq_norm(slab["g_shard"], sq, None)
            if self.world > 1:
                dist.all_reduce(sq, group=_pg(self.group))
            max_norm = float(clip.clip_norm)
        elif clip is not None:
            raise NotImplementedError("GroupShardedOptimizerStage2 supports ClipGradByGlobalNorm only")
        for slab in self.arena.slabs:
            self._init_state(slab)
            self._update_slab(slab, sq, max_norm, inv)
            self._publish(slab)
            slab["g_shard_valid"] = False
        self._finished = False

    def clear_grad(self, set_to_zero=True):
        self._finished = False
        self.comm.join()
        self.arena.zero_grad()
        for slab in self.arena.slabs:
            slab["g_shard_valid"] = False
            for b in slab["buckets"]:
                b.reduced, b.pending = False, len(b.params)

    clear_gradients = clear_grad

    def get_lr(self):
        return self._optim.get_lr()

    def set_lr(self, v):
        self._optim.set_lr(v)

    # ---- checkpoints: full-length, per-parameter, independent of the sharding degree --------------------------------------------------
    def _full(self, slab, t):
        """Shard-space tensor of every rank -> full slab layout."""
        if self.world == 1:
            parts = [t]
        else:
            parts = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(parts, t.contiguous(), group=_pg(self.group))
        full = torch.zeros(slab["numel"], dtype=t.dtype, device=t.device)
        for b in slab["buckets"]:
            per = b.numel // self.world
            for r in range(self.world):
                full[b.start + r * per:b.start + (r + 1) * per].copy_(parts[r][b.shard_off:b.shard_off + per])
        return full

    def _own(self, slab, full):
        return torch.cat([full[b.start + self.rank * (b.numel // self.world):b.start + (self.rank + 1) * (b.numel // self.world)] for b in slab["buckets"]])

    def state_dict(self):
        from ..optimizer.optimizer import split_master_join

        sd = {"@step@": self._step}
        for slab in self.arena.slabs:
            st = slab["state"]
            if "p_shard" not in st:
                continue
            fulls = {}
            for k, key in (("m", "moment1"), ("v", "moment2"), ("vel", "velocity")):
                if k in st:
                    fulls[key] = self._full(slab, st[k].to(slab["device"]))
            master = st.get("master")
            if master is not None:
                mf = split_master_join(st["p_shard"], master) if master.dtype == torch.int16 else master.to(slab["device"])
                fulls["master"] = self._full(slab, mf)
            for p in slab["params"]:
                o, n = slab["offsets"][p.name]
                for key, full in fulls.items():
                    v = full[o:o + n].view(tuple(p.size())).clone().as_subclass(Tensor)
                    if key == "master":
                        sd.setdefault("master_weights", {})[p.name] = v
                    else:
                        sd[f"{p.name}_{key}_0"] = v
        return sd

    def set_state_dict(self, sd):
        from ..optimizer.optimizer import split_master_split

        self._step = int(sd.get("@step@", 0))
        masters = sd.get("master_weights", {}) or {}
        with torch.no_grad():
            for slab in self.arena.slabs:
                if not any(f"{p.name}_moment1_0" in sd or f"{p.name}_velocity_0" in sd or p.name in masters for p in slab["params"]):
                    continue
                self._init_state(slab)
                st = slab["state"]
                for k, key in (("m", "moment1"), ("v", "moment2"), ("vel", "velocity")):
                    if not any(f"{p.name}_{key}_0" in sd for p in slab["params"]):
                        continue
                    full = torch.zeros(slab["numel"], dtype=torch.float32, device=slab["device"])
                    for p in slab["params"]:
                        src = sd.get(f"{p.name}_{key}_0")
                        if src is not None:
                            o, n = slab["offsets"][p.name]
                            full[o:o + n].copy_(torch.as_tensor(_raw(src)).reshape(-1).float())
                    own = self._own(slab, full)
                    if k not in st:
                        st[k] = torch.zeros_like(own)
                    st[k].copy_(own.to(st[k].dtype))
                if masters and st.get("master") is not None:
                    full = slab["data"].float()
                    for p in slab["params"]:
                        src = masters.get(p.name)
                        if src is not None:
                            o, n = slab["offsets"][p.name]
                            full[o:o + n].copy_(torch.as_tensor(_raw(src)).reshape(-1).float())
                    own = self._own(slab, full)
                    if st["master"].dtype == torch.int16:
                        w, lo = split_master_split(own)
                        st["p_shard"].copy_(w)
                        st["master"].copy_(lo)
                    else:
                        st["master"].copy_(own)
                        st["p_shard"].copy_(own.to(st["p_shard"].dtype))
                    self._publish(slab)

    def __getattr__(self, name):
        return getattr(self._optim, name)


DygraphShardingOptimizer = GroupShardedOptimizerStage2


class GroupShardedStage2(Layer):
    """Model wrapper: gradient buckets are reduce-scattered DURING backward (per-parameter hooks -> bucket countdown -> collective on the
    communication stream). Parity: group_sharded_stage2.py:411 (_get_reduce_fn), :705 (task wait before the update)."""

    def __init__(self, layer, sharding_optimizer, group=None, sync_buffers=False, buffer_max_size=2 ** 23, auto_refresh_trainable=True,
                 device="gpu", dp_group=None):
        super().__init__()
        self._layer = layer
        self._opt = sharding_optimizer if not isinstance(sharding_optimizer, (list, tuple)) else sharding_optimizer[0]
        self.group = group
        self.world = _world(group)
        self._opt._overlap = True
        for slab in self._opt.arena.slabs:
            for b in slab["buckets"]:
                for p in b.params:
                    p.register_post_accumulate_grad_hook(lambda param, _b=b: self._opt._param_ready(_b))
        if sync_buffers and self.world > 1:
            src = dist.get_global_rank(_pg(group), 0) if _pg(group) is not None else 0
            for b in layer.buffers():
                dist.broadcast(_raw(b), src=src, group=_pg(group))

    def forward(self, *a, **k):
        self._opt.comm.join()      # reductions (and the zeroing of consumed buckets) of an earlier backward finish before new gradients land
        return self._layer(*a, **k)

    def parameters(self, include_sublayers=True):
        return self._layer.parameters(include_sublayers)

    def state_dict(self, *a, **k):
        return self._layer.state_dict(*a, **k)

    def set_state_dict(self, *a, **k):
        return self._layer.set_state_dict(*a, **k)

    def to_static_state_dict(self, *a, **k):
        return self._layer.state_dict(*a, **k)


# =====================================================================================================================================
# stage 3
# =====================================================================================================================================
class GroupShardedStage3(Layer):
    """Parameter + gradient + optimizer-state sharding with prefetch. Parity: group_sharded_stage3.py."""

    def __init__(self, layer, optimizer, group=None, sync_buffers=False, device="gpu", segment_size=2 ** 20, pretrain_sync_models=True,
                 offload=False, sync_comm=False, dp_group=None, exclude_layer=None):
        super().__init__()
        self._layer = layer
        self._optim = optimizer
        self.group = group
        self.world, self.rank = _world(group), _rank(group)
        self._sync_comm = bool(sync_comm)
        name = type(optimizer).__name__
        if name not in ("Adam", "AdamW", "SGD", "Momentum"):
            raise NotImplementedError(f"GroupShardedStage3 supports Adam / AdamW / SGD / Momentum optimizers, got {name}")
        if name == "Adam" and float(getattr(optimizer, "_weight_decay", 0.0) or 0.0) != 0.0 and not getattr(optimizer, "_decoupled", False):
            raise NotImplementedError("GroupShardedStage3: Adam with coupled (L2) weight decay is not supported; use AdamW")
        clip = optimizer._grad_clip
        if clip is not None and not hasattr(clip, "clip_norm"):
            raise NotImplementedError("GroupShardedStage3 supports ClipGradByGlobalNorm only")
        if pretrain_sync_models and self.world > 1:
            src = dist.get_global_rank(_pg(group), 0) if _pg(group) is not None else 0
            with torch.no_grad():
                for p in layer.parameters():
                    dist.broadcast(_raw(p), src=src, group=_pg(group))
        self._step = 0
        self._units = []
        self._order = []           # execution order of the units, learnt in the first forward
        self._order_frozen = False
        self._build_units()
        self._patch_optimizer()

    # ---- layout --------------------------------------------------------------------------------------------------------------------
    def _collect_units(self):
        """Sharding units: the largest sub-trees of the model whose parameters fit B200_S3_UNIT_MB (default 1 GB: one transformer block
        of a 6.7B model), so that a block is gathered with ONE collective before its forward / backward instead of one per Linear or
        LayerNorm.  0 = one unit per parameter-owning sublayer (the reference's granularity)."""
        max_bytes = int(float(os.environ.get("B200_S3_UNIT_MB", "1024")) * (1 << 20))
        seen = set()

        def own(layer):
            return [p for p in layer._parameters.values() if p is not None and not p.stop_gradient]

        def under(layer):
            out = []
            for sub in layer.sublayers(include_self=True):
                out.extend(own(sub))
            return out

        units = []

        def add(root, ps):
            ps = [p for p in ps if id(p) not in seen]
            keyed = {}
            for p in ps:
                keyed.setdefault((p.dtype, p.device), []).append(p)
            for group in keyed.values():          # a unit holds one dtype / device (mixed-precision models: one unit per class)
                for p in group:
                    seen.add(id(p))
                units.append((root, group))

        def walk(layer):
            kids = [c for c in layer.children() if under(c)]
            ps = under(layer)
            total = sum(p.numel() * p.element_size() for p in ps)
            if max_bytes > 0 and (total <= max_bytes or not kids):
                add(layer, ps)
                return
            if max_bytes <= 0 and not kids:
                add(layer, own(layer))
                return
            if own(layer):
                add(layer, own(layer))
            for c in kids:
                walk(c)

        walk(self._layer)
        return units

    def _build_units(self):
        w = self.world
        units = self._collect_units()
        # shard arena per (dtype, device): [unit shards back to back]
        self._arenas = {}
        sizes = {}
        for sub, ps in units:
            key = (ps[0].dtype, ps[0].device)
            assert all((p.dtype, p.device) == key for p in ps), "a sharding unit must hold parameters of one dtype / device"
            flat_n = sum((p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN for p in ps)
            flat_n = (flat_n + w * _ALIGN - 1) // (w * _ALIGN) * (w * _ALIGN)
            sizes.setdefault(key, []).append(flat_n // w)
        dev0 = units[0][1][0].device if units else torch.device("cpu")
        need = sum(sum(v) * torch.empty(0, dtype=k[0]).element_size() for k, v in sizes.items())
        ctx = _symm_ctx(self.group, dev0, heap_bytes=need + (640 << 20))
        if ctx is not None and need > ctx.heap.size() - ctx.heap.cursor() - (512 << 20):
            ctx = None
        self.comm = _Comm(self.group, dev0)
        self.comm.ctx = ctx
        for key, per_unit in sizes.items():
            n = sum(per_unit)
            dt, dev = key
            if ctx is not None:
                p_sh, off = ctx.buffer(("s3_shards", str(dt)), (n,), dt)
                p_sh.zero_()
            else:
                p_sh, off = torch.zeros(n, dtype=dt, device=dev), None
            self._arenas[key] = dict(dtype=dt, device=dev, p_shard=p_sh, heap_off=off, g_shard=torch.zeros(n, dtype=dt, device=dev), state={},
                                     es=p_sh.element_size(), cursor=0, has_grad=False)
        for sub, ps in units:
            key = (ps[0].dtype, ps[0].device)
            ar = self._arenas[key]
            offs, off = [], 0
            for p in ps:
                offs.append((off, p.numel(), tuple(p.size())))
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            flat_n = (off + w * _ALIGN - 1) // (w * _ALIGN) * (w * _ALIGN)
            per = flat_n // w
            lo = ar["cursor"]
            ar["cursor"] += per
            full = torch.zeros(flat_n, dtype=ps[0].dtype, device=ps[0].device)
            for p, (o, n, _) in zip(ps, offs):
                full[o:o + n].copy_(_raw(p).detach().reshape(-1))
            ar["p_shard"][lo:lo + per].copy_(full[self.rank * per:(self.rank + 1) * per])
            unit = dict(layer=sub, params=ps, offs=offs, flat=flat_n, per=per, arena=ar, lo=lo, full=None, event=None, gathered=False,
                        pending=0, index=len(self._units))
            for p in ps:
                p.__dict__["_s3_unit"] = unit
                p.data = torch.empty(0, dtype=p.dtype, device=p.device)   # released
            self._units.append(unit)
            sub.register_forward_pre_hook(lambda l, inp, u=unit: self._pre_forward(u))
            sub.register_forward_post_hook(lambda l, inp, out, u=unit: self._after_forward(u, out))
            for p in ps:
                p.register_post_accumulate_grad_hook(lambda param, u=unit: self._grad_ready(u, param))

    # ---- gather / release --------------------------------------------------------------------------------------------------------
    def _issue_gather(self, unit):
        """Start the all-gather of a unit's weights on the communication stream (no-op when already in flight / resident)."""
        if unit["full"] is not None:
            return
        ar = unit["arena"]
        shard = ar["p_shard"][unit["lo"]:unit["lo"] + unit["per"]]
        hoff = ar["heap_off"] + unit["lo"] * ar["es"] if ar["heap_off"] is not None else None
        with self.comm.on_stream():
            full = torch.empty(unit["flat"], dtype=ar["dtype"], device=ar["device"])
            self.comm.gather_from_shards(full, shard, hoff)
            if self.comm.cuda:
                ev = torch.cuda.Event()
                ev.record(self.comm.stream)
                unit["event"] = ev
        unit["full"] = full

    def _gather(self, unit):
        """Make the unit's full weights usable on the compute stream."""
        if unit["gathered"]:
            return
        self._issue_gather(unit)
        if unit["event"] is not None:
            from . import comm_timer as CT

            with CT.region("sharding_param_wait"):
                torch.cuda.current_stream().wait_event(unit["event"])
            unit["full"].record_stream(torch.cuda.current_stream())
        for p, (o, n, shape) in zip(unit["params"], unit["offs"]):
            p.data = unit["full"][o:o + n].view(shape)
        unit["gathered"] = True

    def _release(self, unit):
        for p in unit["params"]:
            p.data = torch.empty(0, dtype=p.dtype, device=p.device)
        unit["full"], unit["event"], unit["gathered"] = None, None, False

    def _neighbour(self, unit, step):
        if not self._order_frozen or self._sync_comm:
            return None
        try:
            i = self._order.index(unit["index"]) + step
        except ValueError:
            return None
        return self._units[self._order[i]] if 0 <= i < len(self._order) else None

    def _pre_forward(self, unit):
        if not self._order_frozen and unit["index"] not in self._order:
            self._order.append(unit["index"])
        self._gather(unit)
        nxt = self._neighbour(unit, +1)
        if nxt is not None:
            self._issue_gather(nxt)          # prefetch: the next unit's weights travel while this unit computes

    def _after_forward(self, unit, out):
        if not torch.is_grad_enabled():
            self._release(unit)
            return None

        def pre_backward(_g, u=unit):
            self._gather(u)
            prv = self._neighbour(u, -1)
            if prv is not None:
                self._issue_gather(prv)      # backward runs the units in reverse order
            return None

        t = out[0] if isinstance(out, (tuple, list)) else out
        if isinstance(t, torch.Tensor) and t.requires_grad:
            t.register_hook(pre_backward)
        unit["pending"] = len(unit["params"])
        # autograd saved the parameter VARIABLES; swapping their .data frees the gathered storage now, the pre-backward hook swaps the
        # re-gathered weights back in before this unit's grad functions run
        self._release(unit)
        return None

    def _grad_ready(self, unit, param):
        unit["pending"] -= 1
        if unit["pending"] > 0:
            return
        ar = unit["arena"]
        flat = torch.zeros(unit["flat"], dtype=ar["dtype"], device=ar["device"])
        for p, (o, n, _) in zip(unit["params"], unit["offs"]):
            g = torch.Tensor.grad.__get__(p)
            if g is not None:
                flat[o:o + n].copy_(g.reshape(-1))
                torch.Tensor.grad.__set__(p, None)
        dst = ar["g_shard"][unit["lo"]:unit["lo"] + unit["per"]]
        with self.comm.on_stream():
            tmp = torch.empty(unit["per"], dtype=ar["dtype"], device=ar["device"])
            self.comm.reduce_scatter(tmp, flat, None)
            dst.add_(tmp)
            if self.comm.cuda:
                flat.record_stream(self.comm.stream)
        ar["has_grad"] = True
        self._release(unit)

    # ---- optimizer -------------------------------------------------------------------------------------------------------------------
    def _init_state(self, ar):
        st = ar["state"]
        if "ready" in st:
            return
        o = self._optim
        n, dev = ar["p_shard"].numel(), ar["device"]
        if type(o).__name__ in ("Adam", "AdamW"):
            mdt = getattr(o, "_moment_dtype", None) or torch.float32
            st["m"], st["v"] = torch.zeros(n, dtype=mdt, device=dev), torch.zeros(n, dtype=mdt, device=dev)
        if ar["dtype"] != torch.float32:
            from ..framework.flags import flag

            if ar["dtype"] == torch.bfloat16 and flag("FLAGS_b200_split_master_weights", True) and type(o).__name__ in ("Adam", "AdamW"):
                st["master"] = torch.zeros(n, dtype=torch.int16, device=dev)
            else:
                st["master"] = ar["p_shard"].float()
        else:
            st["master"] = None
        st["ready"] = True

    def grad_shards(self):
        """Gradient shards of this rank (for GroupShardedScaler)."""
        return [ar["g_shard"] for ar in self._arenas.values() if ar["has_grad"]]

    def _patch_optimizer(self):
        outer = self
        optim = self._optim
        optim.__dict__["_s3"] = self

        @torch.no_grad()
        def step():
            outer._step += 1
            outer._order_frozen = True
            outer.comm.join()
            arenas = [ar for ar in outer._arenas.values() if ar["has_grad"]]
            if not arenas:
                return
            dev = arenas[0]["device"]
            inv = outer.__dict__.get("_inv_world")
            if inv is None:
                inv = outer.__dict__["_inv_world"] = torch.full((1,), 1.0 / outer.world, dtype=torch.float32, device=dev)
            clip = optim._grad_clip
            sq, max_norm = None, 0.0
            E = _ops(dev)
            if clip is not None:
                sq = torch.zeros(1, dtype=torch.float32, device=dev)
                for ar in arenas:
                    E.grad_sq_norm(ar["g_shard"], sq, None)
                if outer.world > 1:
                    dist.all_reduce(sq, group=_pg(outer.group))
                max_norm = float(clip.clip_norm)
            lr = optim.get_lr()
            name = type(optim).__name__
            for ar in arenas:
                outer._init_state(ar)
                st = ar["state"]
                if name in ("Adam", "AdamW"):
                    b1, b2 = optim._betas()
                    wd = float(optim._weight_decay or 0.0) if getattr(optim, "_decoupled", False) else 0.0
                    E.adamw_step(ar["p_shard"], ar["g_shard"], st["master"], st["m"], st["v"], lr, b1, b2, float(optim._epsilon), wd, outer._step,
                                 sq, max_norm, None, inv)
                else:
                    gs = 1.0 / outer.world
                    if sq is not None and max_norm > 0:
                        norm = float(sq.sqrt().item()) * gs
                        if norm > max_norm:
                            gs *= max_norm / (norm + 1e-6)
                    wdv = optim._weight_decay
                    wd = float(getattr(wdv, "coeff", wdv) or 0.0)
                    pf = st["master"] if st["master"] is not None else ar["p_shard"]
                    gf = ar["g_shard"].float() * gs + wd * pf.float()
                    mom = getattr(optim, "_momentum", 0.0)
                    if mom:
                        if "vel" not in st:
                            st["vel"] = torch.zeros_like(gf)
                        st["vel"].mul_(mom).add_(gf)
                        gf = st["vel"]
                    pf.add_(gf.to(pf.dtype), alpha=-lr)
                    if pf is not ar["p_shard"]:
                        ar["p_shard"].copy_(pf)
                ar["g_shard"].zero_()
                ar["has_grad"] = False

        def clear_grad(set_to_zero=True):
            outer.comm.join()
            for ar in outer._arenas.values():
                ar["g_shard"].zero_()
                ar["has_grad"] = False
            for unit in outer._units:
                for p in unit["params"]:
                    torch.Tensor.grad.__set__(p, None)

        optim.step = step
        optim.clear_grad = clear_grad
        optim.clear_gradients = clear_grad

    def forward(self, *a, **k):
        return self._layer(*a, **k)

    def get_all_parameters(self, convert2cpu=False):
        """Materialise full parameters on every rank (for saving / evaluation)."""
        for unit in self._units:
            self._gather(unit)
        return self._layer.parameters()

    def state_dict(self, *a, **k):
        self.get_all_parameters()
        sd = {kk: vv.clone() for kk, vv in self._layer.state_dict(*a, **k).items()}
        for unit in self._units:
            self._release(unit)
        return sd

    def parameters(self, include_sublayers=True):
        return self._layer.parameters(include_sublayers)


class GroupShardedScaler:
    """GradScaler whose unscale / found_inf work on the SHARDED gradients and are agreed across the sharding group.
    Parity: group_sharded_utils.py:GroupShardedScaler."""

    def __new__(cls, scaler, group=None):
        orig = scaler.unscale_

        def unscale_(optimizer):
            s3 = getattr(optimizer, "__dict__", {}).get("_s3")
            s2 = optimizer if isinstance(optimizer, GroupShardedOptimizerStage2) else None
            if s3 is None and s2 is None:
                orig(getattr(optimizer, "_optim", optimizer))
            else:
                # the full-size .grad tensors are gone (stage 3) or about to be reduced (stage 2): unscale the reduced shards
                if s2 is not None:
                    s2._finish_reduction()
                    shards = [slab["g_shard"] for slab in s2.arena.slabs]
                    for slab in s2.arena.slabs:
                        slab["g_shard_valid"] = True
                else:
                    s3.comm.join()
                    shards = s3.grad_shards()
                if shards:
                    scaler._lazy(shards[0].device)
                    scaler._found_inf.zero_()
                    inv = 1.0 / scaler._scale
                    for g in shards:
                        g.mul_(inv.to(g.dtype))
                        if not bool(torch.isfinite(g.float().sum())):
                            scaler._found_inf.fill_(1.0)
                from ..amp.grad_scaler import OptimizerState

                scaler._opt_states[id(optimizer)] = OptimizerState.UNSCALED
            if scaler._found_inf is not None and _world(group) > 1:
                dist.all_reduce(scaler._found_inf, op=dist.ReduceOp.MAX, group=_pg(group))

        scaler.unscale_ = unscale_
        return scaler


def group_sharded_parallel(model, optimizer, level, scaler=None, group=None, offload=False, sync_buffers=False, buffer_max_size=2 ** 23,
                           segment_size=2 ** 20, sync_comm=False, dp_group=None, exclude_layer=None):
    """Parity: python/paddle/distributed/sharding/group_sharded.py:group_sharded_parallel."""
    if level not in ("os", "os_g", "p_g_os"):
        raise ValueError("level must be one of 'os', 'os_g', 'p_g_os'")
    if not env.is_initialized():
        env.init_parallel_env()
    if level in ("os", "os_g"):
        opt = GroupShardedOptimizerStage2(params=optimizer._parameter_list, optim=optimizer, group=group, offload=offload)
        model = GroupShardedStage2(model, opt, group=group, sync_buffers=sync_buffers, buffer_max_size=buffer_max_size)
        optimizer = opt
    else:
        model = GroupShardedStage3(model, optimizer=optimizer, group=group, sync_buffers=sync_buffers, segment_size=segment_size,
                                   offload=offload, sync_comm=sync_comm)
    if scaler is not None:
        scaler = GroupShardedScaler(scaler, group)
    return model, optimizer, scaler


def save_group_sharded_model(model, output, optimizer=None):
    from ..framework.io import save

    os.makedirs(output, exist_ok=True)
    sd = model.state_dict()
    if env.get_rank() == 0:
        save(sd, os.path.join(output, "model.pdmodel"))
    if optimizer is not None:
        osd = optimizer.state_dict()          # stage 2: full-length and rank-independent (gathered); stage 3 / plain: per rank
        if isinstance(optimizer, GroupShardedOptimizerStage2):
            if env.get_rank() == 0:
                save(osd, os.path.join(output, "model.pdopt"))
        else:
            save(osd, os.path.join(output, f"model.pdopt.rank{env.get_rank()}"))
