"""Hybrid-parallel model wrappers and optimizer.

Parity: python/paddle/distributed/fleet/model.py (distributed_model), meta_parallel/tensor_parallel.py,
sharding_parallel.py, segment_parallel.py, meta_optimizers/dygraph_optimizer/hybrid_parallel_optimizer.py
(HybridParallelOptimizer, HybridParallelClipGrad), utils/hybrid_parallel_util.py (broadcast_*_parameters,
fused_allreduce_gradients).

B200 design: parameters/gradients of every rank live in flat arenas grouped by (decay, distributed, sequence-parallel)
class, so (a) the dp / sharding gradient reduction is a handful of large in-place collectives over the grad slab
(peer-memory kernels when the symmetric heap is up), (b) the global-norm clip needs one reduction kernel per slab and
(c) the AdamW update is one fused kernel per slab reading the clip coefficient from device memory.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...nn.clip import ClipGradByGlobalNorm
from ...nn.layer import Layer
from ...tensor import Tensor
from .. import collective as C
from .. import env
from . import mp_layers as mpu
from .pipeline import PipelineLayer, PipelineParallel


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _pg(group):
    return group.pg if isinstance(group, C.Group) else group


def _n(group):
    return group.nranks if isinstance(group, C.Group) else (dist.get_world_size(group) if group is not None else 1)


def _broadcast_params(model, group, src_rank, only_not_distributed=False):
    if group is None or _n(group) <= 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            if only_not_distributed and getattr(t, "is_distributed", False):
                continue
            dist.broadcast(_raw(t), src=src_rank, group=_pg(group))


def broadcast_mp_parameters(model, hcg):
    _broadcast_params(model, hcg.get_model_parallel_group(), hcg.get_model_parallel_group_src_rank(), only_not_distributed=True)


def broadcast_dp_parameters(model, hcg):
    _broadcast_params(model, hcg.get_data_parallel_group(), hcg.get_data_parallel_group_src_rank())


def broadcast_sharding_parameters(model, hcg):
    _broadcast_params(model, hcg.get_sharding_parallel_group(), hcg.get_sharding_parallel_group_src_rank())


def broadcast_sep_parameters(model, hcg):
    g = hcg.get_sep_parallel_group()
    if g is not None:
        _broadcast_params(model, g, g.ranks[0])


def fused_allreduce_gradients(parameter_list, hcg, scale=None):
    """Sum-then-average gradients over the data-parallel (x sep) group. Parity: hybrid_parallel_util.fused_allreduce_gradients."""
    group = hcg.get_dp_sep_parallel_group() if hcg is not None and hcg.get_sep_parallel_world_size() > 1 else (hcg.get_data_parallel_group() if hcg is not None else None)
    n = _n(group) if group is not None else env.get_world_size()
    if n <= 1:
        return
    grads = [torch.Tensor.grad.__get__(p) for p in parameter_list if torch.Tensor.grad.__get__(p) is not None]
    _allreduce_tensors(grads, group, 1.0 / n if scale is None else scale)


def _allreduce_tensors(tensors, group, scale):
    """Coalesced all-reduce of a list of tensors (bucketed flat buffers)."""
    if not tensors:
        return
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dt, ts in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in ts])
        _allreduce_flat(flat, group)
        if scale != 1.0:
            flat.mul_(scale)
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n


def _allreduce_flat(flat, group):
    """In-place all-reduce of one contiguous buffer: peer-memory kernel on NVSwitch when available, else NCCL/gloo."""
    from .. import comm_timer as CT

    with CT.region("grad_all_reduce"):
        _allreduce_flat_impl(flat, group)


def _allreduce_flat_impl(flat, group):
    if flat.is_cuda and flat.dtype in (torch.bfloat16, torch.float16, torch.float32):
        from ...framework.flags import flag

        if flag("FLAGS_b200_nvls", False):                   # opt-in: the switch reduces (parallel/nvls.py)
            from ...parallel import nvls

            nc = nvls.context_for(group)
            if nc is not None and nc.supports(flat):
                nc.all_reduce_(flat)
                return
        if flag("FLAGS_b200_p2p_collectives", True):
            from ...parallel import symm

            sc = symm.context_for(group)
            if sc is not None and flat.numel() * flat.element_size() <= sc.max_message_bytes():
                sc.allreduce_(flat)
                return
    dist.all_reduce(flat, group=_pg(group))


class MetaParallelBase(Layer):
    def __init__(self, layers, hcg, strategy):
        super().__init__()
        self._layers, self._hcg, self._strategy = layers, hcg, strategy
        self._prepare_for_model()

    def _prepare_for_model(self):
        pass

    def forward(self, *inputs, **kwargs):
        return self._layers(*inputs, **kwargs)

    def parameters(self, include_sublayers=True):
        return self._layers.parameters(include_sublayers)

    def named_parameters(self, prefix="", include_sublayers=True):
        return self._layers.named_parameters(prefix, include_sublayers)

    def state_dict(self, *a, **k):
        return self._layers.state_dict(*a, **k)

    def set_state_dict(self, *a, **k):
        return self._layers.set_state_dict(*a, **k)

    def train(self):
        self._layers.train()
        return super().train()

    def eval(self):
        self._layers.eval()
        return super().eval()


class TensorParallel(MetaParallelBase):
    def _prepare_for_model(self):
        broadcast_mp_parameters(self._layers, self._hcg)
        if self._hcg.get_sharding_parallel_world_size() > 1:
            broadcast_sharding_parameters(self._layers, self._hcg)
        broadcast_dp_parameters(self._layers, self._hcg)
        mpu.register_sequence_parallel_allreduce_hooks(self._layers)

    def forward(self, *inputs, **kwargs):
        # all ranks of a tensor-parallel group must consume the same batch: broadcast it from the group's first rank
        # (parity: fleet/utils/hybrid_parallel_util.py:broadcast_input_data called by TensorParallel._pre_forward)
        inputs, kwargs = broadcast_input_data(self._hcg, *inputs, **kwargs)
        return self._layers(*inputs, **kwargs)


def broadcast_input_data(hcg, *inputs, **kwargs):
    group = hcg.get_model_parallel_group()
    if _n(group) <= 1:
        return inputs, kwargs
    src = hcg.get_model_parallel_group_src_rank()

    def bc(t):
        if isinstance(t, torch.Tensor):
            raw = _raw(t)
            if not raw.is_contiguous():
                raw = raw.contiguous()
                dist.broadcast(raw, src=src, group=_pg(group))
                return raw.as_subclass(type(t)) if type(t) is not torch.Tensor else raw
            dist.broadcast(raw, src=src, group=_pg(group))
        return t

    return tuple(bc(t) for t in inputs), {k: bc(v) for k, v in kwargs.items()}


class ShardingParallel(MetaParallelBase):
    def _prepare_for_model(self):
        broadcast_sharding_parameters(self._layers, self._hcg)
        broadcast_dp_parameters(self._layers, self._hcg)


class SegmentParallel(MetaParallelBase):
    def _prepare_for_model(self):
        broadcast_sep_parameters(self._layers, self._hcg)
        broadcast_dp_parameters(self._layers, self._hcg)


def distributed_model(model, hcg, strategy):
    if hcg is None or env.get_world_size() == 1:
        return model
    if hcg.get_pipe_parallel_world_size() > 1:
        if not isinstance(model, PipelineLayer):
            raise TypeError("pp_degree > 1 requires the model to be a PipelineLayer")
        broadcast_mp_parameters(model, hcg)
        if hcg.get_sharding_parallel_world_size() > 1:
            broadcast_sharding_parameters(model, hcg)
        broadcast_dp_parameters(model, hcg)
        mpu.register_sequence_parallel_allreduce_hooks(model)
        from .pipeline import PipelineParallelWithInterleave, PipelineParallelWithInterleaveFthenB, PipelineParallelZeroBubble

        mode = str(strategy.pipeline_configs.get("schedule_mode", "1F1B") or "1F1B").upper().replace("-", "")
        if getattr(model, "_num_virtual", 1) > 1:
            cls = PipelineParallelWithInterleaveFthenB if mode == "FTHENB" else PipelineParallelWithInterleave
            return cls(model, hcg, strategy)
        if mode.startswith("ZB"):
            return PipelineParallelZeroBubble(model, hcg, strategy)
        return PipelineParallel(model, hcg, strategy)
    if hcg.get_model_parallel_world_size() > 1:
        return TensorParallel(model, hcg, strategy)
    if hcg.get_sharding_parallel_world_size() > 1:
        return ShardingParallel(model, hcg, strategy)
    if hcg.get_sep_parallel_world_size() > 1:
        return SegmentParallel(model, hcg, strategy)
    from ..data_parallel import DataParallel

    return DataParallel(model, group=hcg.get_data_parallel_group())


class HybridParallelClipGrad:
    """Global-norm clip across the hybrid topology. Parity: hybrid_parallel_optimizer.py:HybridParallelClipGrad."""

    def __init__(self, clip, hcg):
        self._clip, self._hcg = clip, hcg
        self.clip_norm = clip.clip_norm

    def global_norm_sq(self, params_grads):
        dev = None
        dist_sq = nondist_sq = None
        for p, g in params_grads:
            if g is None or not getattr(p, "need_clip", True):
                continue
            gr = _raw(g)
            dev = gr.device
            s = gr.float().pow(2).sum()
            if getattr(p, "is_distributed", False):
                dist_sq = s if dist_sq is None else dist_sq + s
            else:
                nondist_sq = s if nondist_sq is None else nondist_sq + s
        z = torch.zeros((), dtype=torch.float32, device=dev or "cpu")
        return _reduce_norm_sq(dist_sq if dist_sq is not None else z, nondist_sq if nondist_sq is not None else z, self._hcg)

    def __call__(self, params_grads):
        total = self.global_norm_sq(params_grads)
        gn = torch.sqrt(total)
        coef = self.clip_norm / torch.clamp(gn, min=self.clip_norm)
        out = []
        for p, g in params_grads:
            if g is None or not getattr(p, "need_clip", True):
                out.append((p, g))
            else:
                out.append((p, g * coef.to(g.dtype)))
        return out


def _reduce_norm_sq(dist_sq, nondist_sq, hcg, grads_sharded=False):
    """total = sum over the model: mp-sharded params summed over mp; replicated params counted once; then the pp sum.  The sum over
    the sharding group applies only when every rank holds a distinct shard of the gradients (`grads_sharded`); on this non-arena
    path the gradients were averaged over the sharding group and are replicated, so they are counted once."""
    mp = hcg.get_model_parallel_world_size()
    total = dist_sq + nondist_sq / mp
    total = total.reshape(1).clone()
    if mp > 1:
        dist.all_reduce(total, group=_pg(hcg.get_model_parallel_group()))
    if hcg.get_pipe_parallel_world_size() > 1:
        dist.all_reduce(total, group=_pg(hcg.get_pipe_parallel_group()))
    if grads_sharded and hcg.get_sharding_parallel_world_size() > 1:
        dist.all_reduce(total, group=_pg(hcg.get_sharding_parallel_group()))
    return total.reshape([])


class HybridParallelOptimizer:
    """Wraps an optimizer for hybrid parallel training. Parity: hybrid_parallel_optimizer.py:HybridParallelOptimizer."""

    def __init__(self, optimizer, hcg, strategy):
        self._inner_opt, self._hcg, self._strategy = optimizer, hcg, strategy
        self._dp_enable = hcg is not None and hcg.get_data_parallel_world_size() > 1
        self._sharding_enable = hcg is not None and hcg.get_sharding_parallel_world_size() > 1
        self._need_hybrid_clip = hcg is not None and (hcg.get_model_parallel_world_size() > 1 or hcg.get_pipe_parallel_world_size() > 1 or self._sharding_enable)
        clip = optimizer._grad_clip
        self._params = optimizer._parameter_list
        self._use_arena = bool(self._params) and hasattr(optimizer, "_arena_step") and optimizer._arena_ok_static()
        if self._use_arena:
            from ...parallel.arena import ParamArena

            decay_fn = getattr(optimizer, "_apply_decay_param_fun", None)

            def group_fn(p):
                d = 1 if (decay_fn is None or decay_fn(p.name)) else 0
                return (d, 1 if getattr(p, "is_distributed", False) else 0, 1 if mpu.is_sequence_parallel_parameter(p) else 0)

            arena = ParamArena(self._params, group_fn=group_fn, grad_allocator=self._dp_grad_allocator())
            mp = hcg.get_model_parallel_world_size() if hcg is not None else 1
            for key, slab in arena.slabs.items():
                d, is_dist, is_sp = key[2]
                slab.decay, slab.is_distributed, slab.sequence_parallel = bool(d), bool(is_dist), bool(is_sp)
                if is_sp:
                    for p in slab.params:
                        p.__dict__["_sp_reduce_in_optimizer"] = True
            optimizer.enable_flat_arena(arena)
            if self._sharding_enable:
                # sharding stage 1 (DygraphShardingOptimizer): each rank of the sharding group keeps moments / master weights for, and
                # updates, 1/N of every slab, then the owners publish their ranges; the clip norm is summed over the group
                g = hcg.get_sharding_parallel_group()
                optimizer._aux["shard"] = (hcg.get_sharding_parallel_rank(), hcg.get_sharding_parallel_world_size(), g)
            if isinstance(clip, ClipGradByGlobalNorm) and self._need_hybrid_clip:
                optimizer._aux["norm_allreduce"] = self._arena_norm_allreduce
                optimizer._arena_norm_split = True
        elif isinstance(clip, ClipGradByGlobalNorm) and self._need_hybrid_clip:
            optimizer._grad_clip = HybridParallelClipGrad(clip, hcg)

    def _dp_grad_allocator(self):
        """Data-parallel replicas all-reduce the whole gradient slab every step: place it in a symmetric heap sized for it, so
        the peer-memory all-reduce runs in place (no staging copies, no NCCL)."""
        hcg = self._hcg
        if hcg is None or not (self._dp_enable or self._sharding_enable) or not self._params or not self._params[0].is_cuda:
            return None
        group = hcg.get_dp_sharding_parallel_group() if (self._dp_enable and self._sharding_enable) else \
            (hcg.get_data_parallel_group() if self._dp_enable else hcg.get_sharding_parallel_group())
        from ...parallel import symm

        need = sum((p.numel() + 127) // 128 * 128 * p.element_size() for p in self._params if not p.stop_gradient)
        ctx = symm.context_for(group, heap_bytes=need + (768 << 20))
        if ctx is None or need > ctx.heap.size() - ctx.heap.cursor() - (256 << 20):
            return None

        def alloc(n, dt):
            t, _ = ctx.buffer(("hybrid_dp_grad", n, str(dt)), (n,), dt)
            t.zero_()
            return t

        return alloc

    # arena path: sq holds sum over ALL local slabs; replicated slabs must be counted once across mp -> recompute split
    def _arena_norm_allreduce(self, sq):
        hcg = self._hcg
        mp = hcg.get_model_parallel_world_size()
        if mp > 1:
            # subtract (1 - 1/mp) of the replicated part
            rep = torch.zeros(1, dtype=torch.float32, device=sq.device)
            for s in self._inner_opt._arena.all_slabs():
                if not s.is_distributed:
                    lo, hi = self._inner_opt._shard_bounds(s.numel)     # the range this rank contributed to `sq`
                    if hi <= lo:
                        continue
                    if s.grad.is_cuda:
                        from ..._build import ext

                        ext().grad_sq_norm(s.grad[lo:hi], rep, None)
                    else:
                        rep.add_(s.grad[lo:hi].float().pow(2).sum())
            sq.sub_(rep * (1.0 - 1.0 / mp))
            dist.all_reduce(sq, group=_pg(hcg.get_model_parallel_group()))
        if hcg.get_pipe_parallel_world_size() > 1:
            dist.all_reduce(sq, group=_pg(hcg.get_pipe_parallel_group()))
        if hcg.get_sharding_parallel_world_size() > 1:
            dist.all_reduce(sq, group=_pg(hcg.get_sharding_parallel_group()))

    def _sync_grads(self):
        hcg = self._hcg
        if hcg is None:
            return
        arena = self._inner_opt._arena
        # sequence-parallel replicated params (norm weights): grads are partial sums over the mp group
        mp_group = hcg.get_model_parallel_group()
        if hcg.get_model_parallel_world_size() > 1 and arena is not None:
            for s in arena.all_slabs():
                if getattr(s, "sequence_parallel", False):
                    _allreduce_flat(s.grad, mp_group)
        # data-parallel (and sharding-as-dp) average
        group = None
        if self._dp_enable and self._sharding_enable:
            group = hcg.get_dp_sharding_parallel_group()
        elif self._dp_enable:
            group = hcg.get_data_parallel_group()
        elif self._sharding_enable:
            group = hcg.get_sharding_parallel_group()
        if hcg.get_sep_parallel_world_size() > 1:
            # sep (segment / context parallel) ranks hold replicated parameters and see different sequence segments: their gradients
            # are averaged like data-parallel ones (reference: fused_allreduce_gradients over the dp x sep group)
            if self._sharding_enable:
                sep_group = hcg.get_sep_parallel_group()
                ns = _n(sep_group)
                if arena is not None:
                    for s in arena.all_slabs():
                        _allreduce_flat(s.grad, sep_group)
                        s.grad.mul_(1.0 / ns)
                else:
                    grads = [torch.Tensor.grad.__get__(p) for p in self._params if torch.Tensor.grad.__get__(p) is not None]
                    _allreduce_tensors(grads, sep_group, 1.0 / ns)
            else:
                group = hcg.get_dp_sep_parallel_group()
        if group is None or _n(group) <= 1:
            return
        n = _n(group)
        if arena is not None:
            for s in arena.all_slabs():
                _allreduce_flat(s.grad, group)
                s.grad.mul_(1.0 / n)
        else:
            grads = [torch.Tensor.grad.__get__(p) for p in self._params if torch.Tensor.grad.__get__(p) is not None]
            _allreduce_tensors(grads, group, 1.0 / n)

    @torch.no_grad()
    def step(self):
        if not self.__dict__.get("_grads_synced_externally", False):
            self._sync_grads()
        self.__dict__["_grads_synced_externally"] = False
        self._inner_opt.step()

    def clear_grad(self, set_to_zero=True):
        self.__dict__["_grads_synced_externally"] = False
        self._inner_opt.clear_grad(set_to_zero)

    clear_gradients = clear_grad

    def minimize(self, loss, startup_program=None, parameters=None, no_grad_set=None):
        loss.backward()
        self.step()
        return None, None

    def __getattr__(self, name):
        return getattr(self._inner_opt, name)


def distributed_scaler(scaler, hcg):
    """found_inf must be agreed on by every rank of the model replica. Parity: fleet/scaler.py:distributed_scaler."""
    orig_unscale = scaler.unscale_

    def unscale_(optimizer):
        # The dp / sep / sharding gradient reduction lives in HybridParallelOptimizer.step(), which GradScaler.step() skips on an
        # overflow: run it BEFORE the found_inf decision, so (a) every replica enters the collective on every step and (b) an
        # inf / nan produced by one replica reaches all of them through the reduced gradients.  The check group (mp / pp / sharding)
        # then agrees on the flag; data-parallel replicas already agree because they now hold identical gradients.
        if isinstance(optimizer, HybridParallelOptimizer) and not optimizer.__dict__.get("_grads_synced_externally", False):
            optimizer._sync_grads()
            optimizer.__dict__["_grads_synced_externally"] = True
        orig_unscale(getattr(optimizer, "_inner_opt", optimizer))
        if scaler._found_inf is not None and hcg is not None and env.get_world_size() > 1:
            dist.all_reduce(scaler._found_inf, op=dist.ReduceOp.MAX, group=_pg(hcg.get_check_parallel_group()))
            if hcg.get_data_parallel_world_size() > 1:     # belt and braces: low-precision sums can turn one replica's nan into a finite value
                dist.all_reduce(scaler._found_inf, op=dist.ReduceOp.MAX, group=_pg(hcg.get_data_parallel_group()))

    scaler.unscale_ = unscale_
    return scaler
