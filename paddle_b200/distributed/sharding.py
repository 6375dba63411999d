"""Group-sharded (ZeRO) training. Parity: python/paddle/distributed/sharding/group_sharded.py
(group_sharded_parallel, save_group_sharded_model) and fleet/meta_parallel/sharding/
(GroupShardedOptimizerStage2, GroupShardedStage2, GroupShardedStage3, GroupShardedScaler), DygraphShardingOptimizer.

B200 design (flat arenas):
  stage 1 ('os')   : optimizer state sharded - every rank updates 1/N of the flat parameter slab, then all-gathers it.
  stage 2 ('os_g') : + gradients reduce-scattered straight out of the flat gradient slab (bucket ranges, overlapped
                     with backward on a side stream); each rank only keeps/uses its shard of the reduced gradients.
  stage 3 ('p_g_os'): + parameters sharded: a layer's full weights are all-gathered right before its forward /
                     backward and dropped right after; gradients are reduce-scattered as soon as the layer's backward
                     finishes.
Collectives use the peer-memory kernels when the symmetric heap is available (reduce-scatter / all-gather over
NVSwitch), NCCL/gloo otherwise.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from ..nn.layer import Layer
from ..tensor import Tensor
from . import collective as C
from . import env


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _pg(group):
    return group.pg if isinstance(group, C.Group) else group


def _world(group):
    return dist.get_world_size(_pg(group)) if env.is_initialized() else 1


def _rank(group):
    return dist.get_rank(_pg(group)) if env.is_initialized() else 0


def _reduce_scatter(out, flat, group):
    pg = _pg(group)
    if dist.get_backend(pg) == "gloo":
        tmp = flat.clone()
        dist.all_reduce(tmp, group=pg)
        n = out.numel()
        out.copy_(tmp[_rank(group) * n:(_rank(group) + 1) * n])
    else:
        sc = _symm(group, flat)
        if sc is not None:
            from ..parallel import symm_ops

            if symm_ops.reduce_scatter_into(sc, out, flat):   # peer-memory kernel over NVLink (no NCCL)
                return
        dist.reduce_scatter_tensor(out, flat, group=pg)


def _all_gather(flat_out, shard, group):
    pg = _pg(group)
    if dist.get_backend(pg) == "gloo":
        outs = list(flat_out.chunk(_world(group)))
        dist.all_gather(outs, shard.contiguous(), group=pg)
    else:
        sc = _symm(group, flat_out)
        if sc is not None:
            from ..parallel import symm_ops

            if symm_ops.all_gather_into(sc, flat_out, shard.contiguous()):
                return
        dist.all_gather_into_tensor(flat_out, shard.contiguous(), group=pg)


def _symm(group, t):
    if not t.is_cuda or t.dtype not in (torch.bfloat16, torch.float16, torch.float32) or not t.is_contiguous():
        return None
    from ..framework.flags import flag

    if not flag("FLAGS_b200_p2p_collectives", True):
        return None
    from ..parallel import symm

    return symm.context_for(group)


class _ShardedArena:
    """Flat fp-param / grad slabs padded to a multiple of world; rank r owns [r*shard, (r+1)*shard)."""

    def __init__(self, params, group):
        self.group = group
        self.world, self.rank = _world(group), _rank(group)
        self.params = [p for p in params if not p.stop_gradient]
        by = {}
        for p in self.params:
            by.setdefault((p.dtype, p.device), []).append(p)
        self.slabs = []
        for (dt, dev), ps in by.items():
            offs, off = {}, 0
            for p in ps:
                offs[p.name] = (off, p.numel())
                off += (p.numel() + 127) // 128 * 128
            total = (off + self.world * 128 - 1) // (self.world * 128) * (self.world * 128)
            data = torch.zeros(total, dtype=dt, device=dev)
            grad = torch.zeros(total, dtype=dt, device=dev)
            with torch.no_grad():
                for p in ps:
                    o, n = offs[p.name]
                    v = data[o:o + n].view(tuple(p.size()))
                    v.copy_(_raw(p))
                    p.data = v
                    gv = grad[o:o + n].view(tuple(p.size()))
                    torch.Tensor.grad.__set__(p, gv)
                    p.__dict__["_arena_grad"] = gv
            self.slabs.append(dict(dtype=dt, params=ps, offsets=offs, data=data, grad=grad, shard=total // self.world, state={}, master=None))

    def shard_view(self, slab, which):
        s = slab["shard"]
        return slab[which][self.rank * s:(self.rank + 1) * s]

    def zero_grad(self):
        for s in self.slabs:
            s["grad"].zero_()
            for p in s["params"]:
                if torch.Tensor.grad.__get__(p) is not p.__dict__["_arena_grad"]:
                    torch.Tensor.grad.__set__(p, p.__dict__["_arena_grad"])


class GroupShardedOptimizerStage2:
    """Optimizer-state (+ gradient) sharding over flat arenas. Parity: group_sharded_optimizer_stage2.py."""

    def __init__(self, params, optim, group=None, offload=False, device="gpu", pretrain_sync_models=True, dp_group=None, **kw):
        self._optim = optim
        self.group = group
        self.world, self.rank = _world(group), _rank(group)
        self._params = list(params)
        self.offload = offload
        if pretrain_sync_models and self.world > 1:
            src = dist.get_global_rank(_pg(group), 0) if _pg(group) is not None else 0
            with torch.no_grad():
                for p in self._params:
                    dist.broadcast(_raw(p), src=src, group=_pg(group))
        self.arena = _ShardedArena(self._params, group)
        self._reduce_grads_in_step = True   # stage 2 wrapper flips this off and reduces during backward
        self._step = 0

    @property
    def _parameter_list(self):
        return self._params

    def _shard_update(self, slab):
        """AdamW/SGD-style update of this rank's shard using the inner optimizer's hyper-parameters."""
        o = self._optim
        p_sh, g_sh = self.arena.shard_view(slab, "data"), slab["grad_shard"]
        lr = o.get_lr()
        name = type(o).__name__
        if name in ("Adam", "AdamW"):
            b1, b2 = o._betas()
            st = slab["state"]
            wd = float(o._weight_decay or 0.0) if o._decoupled else 0.0
            if self.offload and p_sh.is_cuda:
                # offload: fp32 master shard + moments live in pinned host memory and the update runs on the CPU; only the bf16 / fp16
                # shard travels (gradient down, updated parameters up).  Parity: GroupShardedOptimizerStage2(offload=True).
                if "m" not in st:
                    st["m"] = torch.zeros(p_sh.shape, dtype=torch.float32).pin_memory()
                    st["v"] = torch.zeros(p_sh.shape, dtype=torch.float32).pin_memory()
                    slab["master"] = p_sh.float().cpu().pin_memory()
                    st["g_host"] = torch.zeros(p_sh.shape, dtype=torch.float32).pin_memory()
                st["g_host"].copy_(g_sh.float(), non_blocking=True)
                torch.cuda.current_stream().synchronize()
                self._adam_math(slab["master"], st["g_host"], st, lr, wd, b1, b2, float(o._epsilon))
                p_sh.copy_(slab["master"], non_blocking=True)
                return
            if "m" not in st:
                st["m"] = torch.zeros_like(p_sh, dtype=torch.float32)
                st["v"] = torch.zeros_like(p_sh, dtype=torch.float32)
                if p_sh.dtype != torch.float32:
                    slab["master"] = p_sh.float()
            if p_sh.is_cuda:
                from .._build import ext

                ext().adamw_step(p_sh, g_sh.contiguous(), slab["master"], st["m"], st["v"], lr, b1, b2, float(o._epsilon), wd, self._step, None, 0.0, None, None)
            else:
                pf = slab["master"] if slab["master"] is not None else p_sh
                self._adam_math(pf, g_sh.float(), st, lr, wd, b1, b2, float(o._epsilon))
                if pf is not p_sh:
                    p_sh.copy_(pf)
        else:  # SGD / Momentum
            mom = getattr(o, "_momentum", 0.0)
            st = slab["state"]
            gf = g_sh.float()
            wd = float(o._weight_decay or 0.0) if not hasattr(o._weight_decay, "coeff") else float(o._weight_decay.coeff)
            pf = p_sh.float()
            gf = gf + wd * pf
            if mom:
                if "vel" not in st:
                    st["vel"] = torch.zeros_like(gf)
                st["vel"].mul_(mom).add_(gf)
                gf = st["vel"]
            p_sh.copy_(pf - lr * gf)

    def _adam_math(self, pf, gf, st, lr, wd, b1, b2, eps):
        pf.mul_(1 - lr * wd)
        st["m"].mul_(b1).add_(gf, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(gf, gf, value=1 - b2)
        denom = (st["v"] / (1 - b2 ** self._step)).sqrt_().add_(eps)
        pf.addcdiv_(st["m"], denom, value=-lr / (1 - b1 ** self._step))

    @torch.no_grad()
    def step(self):
        self._step += 1
        clip = self._optim._grad_clip
        for slab in self.arena.slabs:
            if self.world > 1:
                if self._reduce_grads_in_step:
                    sh = torch.empty(slab["shard"], dtype=slab["grad"].dtype, device=slab["grad"].device)
                    _reduce_scatter(sh, slab["grad"], self.group)
                    sh.mul_(1.0 / self.world)
                    slab["grad_shard"] = sh
                else:
                    slab["grad_shard"] = self.arena.shard_view(slab, "grad")
            else:
                slab["grad_shard"] = slab["grad"]
        if clip is not None and hasattr(clip, "clip_norm"):
            sq = torch.zeros(1, dtype=torch.float32, device=self.arena.slabs[0]["grad"].device)
            for slab in self.arena.slabs:
                sq += slab["grad_shard"].float().pow(2).sum()
            if self.world > 1:
                dist.all_reduce(sq, group=_pg(self.group))
            coef = float(clip.clip_norm) / torch.clamp(sq.sqrt(), min=float(clip.clip_norm))
            for slab in self.arena.slabs:
                slab["grad_shard"].mul_(coef.to(slab["grad_shard"].dtype))
        for slab in self.arena.slabs:
            self._shard_update(slab)
            if self.world > 1:
                _all_gather(slab["data"], self.arena.shard_view(slab, "data").clone(), self.group)

    def clear_grad(self, set_to_zero=True):
        self.arena.zero_grad()

    clear_gradients = clear_grad

    def get_lr(self):
        return self._optim.get_lr()

    def set_lr(self, v):
        self._optim.set_lr(v)

    def state_dict(self):
        sd = {"@step@": self._step}
        for i, slab in enumerate(self.arena.slabs):
            for k, v in slab["state"].items():
                sd[f"slab{i}_{k}_rank{self.rank}"] = v.as_subclass(Tensor)
        return sd

    def set_state_dict(self, sd):
        self._step = int(sd.get("@step@", 0))
        for i, slab in enumerate(self.arena.slabs):
            for k in ("m", "v", "vel"):
                key = f"slab{i}_{k}_rank{self.rank}"
                if key in sd:
                    slab["state"][k] = _raw(sd[key]).to(slab["data"].device).clone()

    def __getattr__(self, name):
        return getattr(self._optim, name)


DygraphShardingOptimizer = GroupShardedOptimizerStage2


class GroupShardedStage2(Layer):
    """Model wrapper: reduce-scatter gradient buckets during backward. Parity: group_sharded_stage2.py."""

    def __init__(self, layer, sharding_optimizer, group=None, sync_buffers=False, buffer_max_size=2 ** 23, auto_refresh_trainable=True,
                 device="gpu", dp_group=None):
        super().__init__()
        self._layer = layer
        self._opt = sharding_optimizer if not isinstance(sharding_optimizer, (list, tuple)) else sharding_optimizer[0]
        self.group = group
        self.world = _world(group)
        self._opt._reduce_grads_in_step = True  # reduce-scatter happens at step() on the flat slab (single large collective)
        if sync_buffers and self.world > 1:
            src = dist.get_global_rank(_pg(group), 0) if _pg(group) is not None else 0
            for b in layer.buffers():
                dist.broadcast(_raw(b), src=src, group=_pg(group))

    def forward(self, *a, **k):
        return self._layer(*a, **k)

    def parameters(self, include_sublayers=True):
        return self._layer.parameters(include_sublayers)

    def state_dict(self, *a, **k):
        return self._layer.state_dict(*a, **k)

    def set_state_dict(self, *a, **k):
        return self._layer.set_state_dict(*a, **k)

    def to_static_state_dict(self, *a, **k):
        return self._layer.state_dict(*a, **k)


class GroupShardedStage3(Layer):
    """Parameter + gradient + optimizer-state sharding. Parity: group_sharded_stage3.py."""

    def __init__(self, layer, optimizer, group=None, sync_buffers=False, device="gpu", segment_size=2 ** 20, pretrain_sync_models=True,
                 offload=False, sync_comm=False, dp_group=None, exclude_layer=None):
        super().__init__()
        self._layer = layer
        self._optim = optimizer
        self.group = group
        self.world, self.rank = _world(group), _rank(group)
        self._units = []
        if pretrain_sync_models and self.world > 1:
            src = dist.get_global_rank(_pg(group), 0) if _pg(group) is not None else 0
            with torch.no_grad():
                for p in layer.parameters():
                    dist.broadcast(_raw(p), src=src, group=_pg(group))
        self._step = 0
        self._build_units()
        self._patch_optimizer()

    # every sublayer that directly owns parameters is a gather/release unit
    def _build_units(self):
        for sub in self._layer.sublayers(include_self=True):
            ps = [p for p in sub._parameters.values() if p is not None and not p.stop_gradient]
            if not ps:
                continue
            unit = {"layer": sub, "params": ps, "shards": [], "full_shape": [], "gathered": False}
            for p in ps:
                full = _raw(p).detach().reshape(-1)
                n = full.numel()
                per = (n + self.world - 1) // self.world
                padded = torch.zeros(per * self.world, dtype=full.dtype, device=full.device)
                padded[:n] = full
                shard = padded[self.rank * per:(self.rank + 1) * per].clone()
                unit["shards"].append(shard)
                unit["full_shape"].append((tuple(p.size()), n, per))
                p.__dict__["_s3_shard"] = shard
                p.__dict__["_s3_grad_shard"] = None
                p.data = torch.empty(0, dtype=p.dtype, device=p.device)   # released
            self._units.append(unit)
            sub.register_forward_pre_hook(lambda l, inp, u=unit: self._gather(u))
            sub.register_forward_post_hook(lambda l, inp, out, u=unit: self._after_forward(u, out))
            for p in ps:
                p.register_post_accumulate_grad_hook(lambda param, u=unit: self._grad_ready(u, param))

    def _gather(self, unit):
        if unit["gathered"]:
            return
        for p, shard, (shape, n, per) in zip(unit["params"], unit["shards"], unit["full_shape"]):
            if self.world > 1:
                full = torch.empty(per * self.world, dtype=shard.dtype, device=shard.device)
                _all_gather(full, shard, self.group)
            else:
                full = shard
            p.data = full[:n].view(shape)
        unit["gathered"] = True

    def _release(self, unit):
        for p in unit["params"]:
            p.data = torch.empty(0, dtype=p.dtype, device=p.device)
        unit["gathered"] = False

    def _after_forward(self, unit, out):
        if not torch.is_grad_enabled():
            self._release(unit)
            return None
        # keep params until this unit's backward ran: re-gather lazily when the backward reaches the unit's output
        def pre_backward(_g, u=unit):
            self._gather(u)
            return None

        t = out[0] if isinstance(out, (tuple, list)) else out
        if isinstance(t, torch.Tensor) and t.requires_grad:
            t.register_hook(pre_backward)
        unit["pending"] = len(unit["params"])
        # autograd saved the parameter *variables*; swapping their .data releases the gathered storage now and the
        # pre-backward hook above swaps the re-gathered weights back in before this unit's grad function runs
        self._release(unit)
        return None

    def _grad_ready(self, unit, param):
        g = torch.Tensor.grad.__get__(param)
        idx = [id(p) for p in unit["params"]].index(id(param))
        shape, n, per = unit["full_shape"][idx]
        flat = torch.zeros(per * self.world, dtype=g.dtype, device=g.device)
        flat[:n] = g.reshape(-1)
        if self.world > 1:
            sh = torch.empty(per, dtype=g.dtype, device=g.device)
            _reduce_scatter(sh, flat, self.group)
            sh.mul_(1.0 / self.world)
        else:
            sh = flat
        prev = param.__dict__.get("_s3_grad_shard")
        param.__dict__["_s3_grad_shard"] = sh if prev is None else prev + sh
        torch.Tensor.grad.__set__(param, None)
        unit["pending"] = unit.get("pending", 1) - 1
        if unit["pending"] <= 0:
            self._release(unit)

    def _patch_optimizer(self):
        outer = self
        optim = self._optim

        def step():
            outer._step += 1
            lr = optim.get_lr()
            name = type(optim).__name__
            for unit in outer._units:
                for p, shard in zip(unit["params"], unit["shards"]):
                    g = p.__dict__.get("_s3_grad_shard")
                    if g is None:
                        continue
                    st = p.__dict__.setdefault("_s3_state", {})
                    gf = g.float()
                    if "master" not in st:
                        st["master"] = shard.float().clone()
                    pf = st["master"]
                    if name in ("Adam", "AdamW"):
                        b1, b2 = optim._betas()
                        if "m" not in st:
                            st["m"], st["v"] = torch.zeros_like(pf), torch.zeros_like(pf)
                        wd = float(optim._weight_decay or 0.0) if optim._decoupled else 0.0
                        pf.mul_(1 - lr * wd)
                        st["m"].mul_(b1).add_(gf, alpha=1 - b1)
                        st["v"].mul_(b2).addcmul_(gf, gf, value=1 - b2)
                        denom = (st["v"] / (1 - b2 ** outer._step)).sqrt_().add_(optim._epsilon)
                        pf.addcdiv_(st["m"], denom, value=-lr / (1 - b1 ** outer._step))
                    else:
                        mom = getattr(optim, "_momentum", 0.0)
                        if mom:
                            if "vel" not in st:
                                st["vel"] = torch.zeros_like(pf)
                            st["vel"].mul_(mom).add_(gf)
                            gf = st["vel"]
                        pf.add_(gf, alpha=-lr)
                    shard.copy_(pf)
                    p.__dict__["_s3_grad_shard"] = None

        def clear_grad(set_to_zero=True):
            for unit in outer._units:
                for p in unit["params"]:
                    p.__dict__["_s3_grad_shard"] = None
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
    """GradScaler whose found_inf is agreed across the sharding group. Parity: group_sharded_utils.py:GroupShardedScaler."""

    def __new__(cls, scaler, group=None):
        orig = scaler.unscale_

        def unscale_(optimizer):
            orig(getattr(optimizer, "_optim", optimizer))
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
    if isinstance(model, GroupShardedStage3):
        sd = model.state_dict()
    else:
        sd = model.state_dict()
    if env.get_rank() == 0:
        save(sd, os.path.join(output, "model.pdmodel"))
    if optimizer is not None:
        save(optimizer.state_dict(), os.path.join(output, f"model.pdopt.rank{env.get_rank()}"))
