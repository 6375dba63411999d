"""Pipeline parallelism. Parity: python/paddle/distributed/fleet/meta_parallel/parallel_layers/pp_layers.py
(LayerDesc, SharedLayerDesc, PipelineLayer, segment methods) and pipeline_parallel.py (PipelineParallel 1F1B /
FThenB / interleaved-VPP entry points, train_batch / eval_batch), pp_utils/p2p_communication.py.

One engine runs every schedule: `pp_schedule.build` turns (schedule_mode, stages, chunks, micro-batches) into per-stage op lists
(F / B / W), the engine executes its own list and moves activations through a transport:

* CUDA + symmetric peer heap: a **mailbox** per hop (csrc/comm/p2p_collectives.cu:signal_flag_kernel / wait_flag_kernel): the producer
  copies the tensor into a slot of the consumer's heap with the copy engine on a side stream (no SM work, the compute stream never
  waits on a send), then publishes the slot with a release store; the consumer's compute stream runs a one-thread wait kernel before
  the first kernel that reads the slot and then uses the slot in place.  The wait kernel accounts the nanoseconds it spun, which is
  the exposed pipeline wait per step reported by bench.py.
* otherwise (NCCL without peer memory, gloo in the CPU tests): one communicator per directed hop and direction, asynchronous isend,
  irecv at the point of use - the role of the reference's overlap_p2p_comm (pp_utils/p2p_communication.py:837-864).

Zero-bubble mode (schedule_mode="ZBH1") splits the backward: B computes input gradients only and parks the weight-gradient GEMMs
(kernels/wgrad.py); W ops run them where the schedule has gaps.
"""
from __future__ import annotations

import contextlib
import math
import os
import re

import torch
import torch.distributed as dist

from ...nn.layer import Layer
from ...tensor import Tensor
from . import pp_schedule
from . import topology as topo
from .recompute import recompute as _recompute


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


class LayerDesc:
    def __init__(self, layer_func, *inputs, **kwargs):
        self.layer_func, self.inputs, self.kwargs = layer_func, inputs, kwargs
        if not (isinstance(layer_func, type) and issubclass(layer_func, Layer)):
            raise TypeError("The input(layer_func) should be a derived class of Layer.")

    def build_layer(self):
        return self.layer_func(*self.inputs, **self.kwargs)

    def __repr__(self):
        return f"LayerDesc({self.layer_func.__name__})"


class SharedLayerDesc(LayerDesc):
    def __init__(self, key, layer_func, forward_func=None, shared_weight_attr="weight", *inputs, **kwargs):
        super().__init__(layer_func, *inputs, **kwargs)
        self.layer_name, self.forward_func, self.shared_weight_attr = key, forward_func, shared_weight_attr


class SegmentLayers:
    def __init__(self, layers_desc, num_parts, method="uniform", num_virtual_pipeline_stage=None):
        self.descs, self.num_parts, self.method = layers_desc, num_parts, method
        self.total_parts = num_parts * (num_virtual_pipeline_stage or 1)

    def do_segment(self):
        n, parts = len(self.descs), self.total_parts
        if isinstance(self.method, (list, tuple)):
            return list(self.method)
        if self.method == "uniform":
            return self._uniform(n, parts)
        if self.method.startswith("layer:"):
            name = self.method.split(":", 1)[1]
            flags = [1 if re.search(name, (d.layer_func.__name__ if isinstance(d, LayerDesc) else type(d).__name__)) else 0 for d in self.descs]
            total = sum(flags)
            assert total >= parts, f"only {total} '{name}' layers for {parts} pipeline parts"
            per = self._uniform(total, parts)   # per[k] = number of matched layers that precede part k
            idxs = [i for i, f in enumerate(flags) if f]
            return [0] + [idxs[per[k]] for k in range(1, parts)] + [n]
        raise ValueError(f"unknown seg_method {self.method}")

    @staticmethod
    def _uniform(n, parts):
        base, extra = divmod(n, parts)
        out = [0]
        for i in range(parts):
            out.append(out[-1] + base + (1 if i >= parts - extra else 0))
        return out


class PipelineLayer(Layer):
    def __init__(self, layers, num_stages=None, topology=None, loss_fn=None, seg_method="uniform", recompute_interval=0,
                 recompute_ctx=None, num_virtual_pipeline_stages=None, use_cudagraph=False):
        super().__init__()
        self._use_cudagraph = use_cudagraph       # accepted for parity; whole-step capture lives in jit.capture_train_step
        hcg = topo.get_hybrid_communicate_group()
        self._hcg = hcg
        self._loss_fn = loss_fn
        self._num_stages = num_stages or (hcg.get_pipe_parallel_world_size() if hcg is not None else 1)
        self._stage_id = hcg.get_stage_id() if hcg is not None else 0
        self._recompute_interval = recompute_interval
        self._layers_desc = list(layers)
        self._num_virtual = num_virtual_pipeline_stages or 1
        seg = SegmentLayers(self._layers_desc, self._num_stages, seg_method, self._num_virtual)
        self.segment_parts = seg.do_segment()
        self.shared_layers = {}
        self.shared_weight_attrs = {}
        self._chunks = []
        for v in range(self._num_virtual):
            part = v * self._num_stages + self._stage_id
            lo, hi = self.segment_parts[part], self.segment_parts[part + 1]
            self._chunks.append(self._build(lo, hi, v))
        self.run_function = self._chunks[0]
        self._start, self._end = self.segment_parts[self._stage_id], self.segment_parts[self._stage_id + 1]
        self._shared_groups = {}
        self._init_shared_weight_comm()

    def _init_shared_weight_comm(self):
        """For every SharedLayerDesc key: a communicator over exactly the stages that own a copy, and the first owner's initial value
        broadcast to the others (stages build their layers in different order, so equal seeds do not give equal copies).
        Parity: pp_layers.py:_construct_shared_comm / _synchronize_shared_weights."""
        hcg = self._hcg
        keys = {}
        for idx, d in enumerate(self._layers_desc):
            if isinstance(d, SharedLayerDesc):
                part = next(p for p in range(len(self.segment_parts) - 1) if self.segment_parts[p] <= idx < self.segment_parts[p + 1])
                keys.setdefault(d.layer_name, set()).add(part % self._num_stages)
        if hcg is None or hcg.get_pipe_parallel_world_size() <= 1 or not keys or not dist.is_initialized():
            return
        from ..collective import new_group

        my_pp = list(hcg.get_pipe_parallel_group().ranks)
        for key in sorted(keys):                                     # every process creates every group, in the same order
            stages = sorted(keys[key])
            for pp_ranks in hcg.topology().get_comm_list("pipe"):
                ranks = [pp_ranks[s] for s in stages]
                g = new_group(ranks) if len(ranks) > 1 else None
                if list(pp_ranks) == my_pp and self._stage_id in stages and g is not None:
                    self._shared_groups[key] = (g, ranks[0])
        for key, (g, src) in self._shared_groups.items():
            w = getattr(self.shared_layers[key], self.shared_weight_attrs[key])
            with torch.no_grad():
                dist.broadcast(_raw(w), src=src, group=g.pg)

    def _build(self, lo, hi, chunk):
        fns = []
        for idx in range(lo, hi):
            d = self._layers_desc[idx]
            if isinstance(d, SharedLayerDesc):
                if d.layer_name not in self.shared_layers:
                    layer = d.build_layer()
                    self.shared_layers[d.layer_name] = layer
                    self.shared_weight_attrs[d.layer_name] = d.shared_weight_attr
                    self.add_sublayer(f"shared_{d.layer_name}", layer)
                layer = self.shared_layers[d.layer_name]
                if d.forward_func is None:
                    fns.append(layer)
                else:
                    fns.append(lambda x, _l=layer, _f=d.forward_func: _f(_l, x))
            elif isinstance(d, LayerDesc):
                layer = d.build_layer()
                self.add_sublayer(f"{chunk}_{idx}" if self._num_virtual > 1 else str(idx), layer)
                fns.append(layer)
            elif isinstance(d, Layer):
                self.add_sublayer(str(idx), d)
                fns.append(d)
            elif callable(d):
                fns.append(d)
            else:
                raise TypeError(f"unsupported pipeline element {type(d)}")
        return fns

    def get_stage_from_index(self, layer_idx):
        for s in range(self._num_stages):
            if self.segment_parts[s] <= layer_idx < self.segment_parts[s + 1]:
                return s
        raise IndexError(layer_idx)

    def forward_function(self, start, end, chunk=0):
        fns = self._chunks[chunk][start:end]

        def run(x):
            for f in fns:
                x = f(*x) if isinstance(x, tuple) else f(x)
            return x

        return run

    def forward(self, input, chunk_id=None):
        fns = self._chunks[chunk_id or 0]
        if self._recompute_interval == 0:
            x = input
            for f in fns:
                x = f(*x) if isinstance(x, tuple) else f(x)
            return x
        x = input
        for lo in range(0, len(fns), self._recompute_interval):
            hi = min(len(fns), lo + self._recompute_interval)
            fn = self.forward_function(lo, hi, chunk_id or 0)
            x = _recompute(fn, x) if (self.training and torch.is_grad_enabled()) else fn(x)
        return x

    def allreduce_shared_weight_gradients(self):
        """Tied weights living on several stages (e.g. embedding / lm head): sum their grads over the owning stages."""
        if self._hcg is None or not self._shared_groups:
            return
        for key in sorted(self._shared_groups):
            group, _ = self._shared_groups[key]
            w = getattr(self.shared_layers[key], self.shared_weight_attrs[key])
            g = torch.Tensor.grad.__get__(w)
            if g is None:
                g = torch.zeros_like(_raw(w))
                torch.Tensor.grad.__set__(w, g)
            dist.all_reduce(g, group=group.pg)


# ---------------------------------------------------------------------------------------------------- p2p
_DT = [torch.float32, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.float64, torch.bool, torch.uint8]


class _GroupTransport:
    """torch.distributed transport (NCCL or gloo): one communicator per directed hop and kind, so every channel is a FIFO with a
    single sender and a single receiver - asynchronous sends can never block a receive of the opposite direction."""

    name = "torch.distributed"

    def __init__(self, hcg, ring):
        from ..collective import new_group

        self.hcg = hcg
        pipe = hcg.get_pipe_parallel_group()
        self.pipe_pg = pipe.pg
        self.my_ranks = list(pipe.ranks)
        self.stage = hcg.get_stage_id()
        S = len(self.my_ranks)
        self.S = S
        self.chan = {}
        backend = dist.get_backend(self.pipe_pg)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        hops = [(a, a + 1) for a in range(S - 1)] + ([(S - 1, 0)] if ring and S > 1 else [])
        if S == 2 and ring:
            hops = [(0, 1), (1, 0)]
        # every process creates every communicator in the same order (new_group is collective over the world)
        for pp_ranks in hcg.topology().get_comm_list("pipe"):
            for a, b in hops:
                for kind in ("f", "b"):
                    ranks = sorted({pp_ranks[a], pp_ranks[b]})
                    g = new_group(ranks)
                    if list(pp_ranks) == self.my_ranks and self.stage in (a, b):
                        self.chan[(a, b, kind)] = g.pg
        self.pending = []

    def _pg(self, src, dst, kind):
        # forward traffic of hop (a -> b) uses channel (a, b, 'f'); backward traffic (b -> a) uses (a, b, 'b')
        return self.chan[(src, dst, "f")] if kind == "f" else self.chan[(dst, src, "b")]

    def send(self, kind, key, t, dst_stage, step):
        pg = self._pg(self.stage, dst_stage, kind)
        t = t.contiguous()
        w = dist.isend(t, self.my_ranks[dst_stage], group=pg)
        self.pending.append((w, t))

    def recv(self, kind, key, shape, dtype, src_stage, step):
        pg = self._pg(src_stage, self.stage, kind)
        from .. import comm_timer as CT

        buf = torch.empty(shape, dtype=dtype, device=self.dev)
        with CT.region("pp_recv_wait"):
            w = dist.irecv(buf, self.my_ranks[src_stage], group=pg)
            w.wait()
        return buf

    def end_step(self):
        for w, _ in self.pending:
            w.wait()
        self.pending.clear()

    def exposed_wait(self, reset=True):
        return None


class _MailboxTransport:
    """Peer-memory mailbox transport (see the module docstring)."""

    name = "mailbox"

    def __init__(self, hcg, ctx, slot_bytes, n_keys):
        self.hcg, self.ctx = hcg, ctx
        self.stage = hcg.get_stage_id()
        self.slot_bytes = (int(slot_bytes) + 1023) // 1024 * 1024
        self.n_keys = n_keys                        # slots per (kind, parity)
        total = 2 * 2 * n_keys
        self.data, self.data_off = ctx.buffer("pp_mailbox", (total * self.slot_bytes,), torch.uint8)
        self.flags, self.flag_off = ctx.buffer("pp_mailbox_flags", (total,), torch.int32)
        self.side = torch.cuda.Stream()
        self.dev = self.data.device
        self._peer_views = {}

    @staticmethod
    def heap_bytes(slot_bytes, n_keys):
        slot = (int(slot_bytes) + 1023) // 1024 * 1024
        return 2 * 2 * n_keys * slot + (8 << 20)

    def _index(self, kind, key, step):
        return ((0 if kind == "f" else 1) * 2 + (step & 1)) * self.n_keys + key

    def send(self, kind, key, t, dst_stage, step):
        idx = self._index(kind, key, step)
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        assert nbytes <= self.slot_bytes, "pipeline message larger than the mailbox slot"
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        dst = self._peer_views.get((dst_stage, idx))
        if dst is None:
            dst = self._peer_views[(dst_stage, idx)] = self.ctx.heap.tensor(self.data_off + idx * self.slot_bytes, [self.slot_bytes], torch.uint8, dst_stage)
        with torch.cuda.stream(self.side):
            dst[:nbytes].copy_(t.view(-1).view(torch.uint8), non_blocking=True)     # copy engine over NVLink into the consumer's slot
            self.ctx.heap.signal_flag(dst_stage, self.flag_off + idx * 4, step)
        t.record_stream(self.side)

    def recv(self, kind, key, shape, dtype, src_stage, step):
        idx = self._index(kind, key, step)
        self.ctx.heap.wait_flag(self.flag_off + idx * 4, step, float(os.environ.get("B200_PP_WAIT_TIMEOUT_S", "600")))
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        lo = idx * self.slot_bytes
        return self.data[lo:lo + nbytes].view(dtype).view(tuple(shape))     # used in place: the slot is not rewritten before step + 2

    def end_step(self):
        torch.cuda.current_stream().wait_stream(self.side)

    def exposed_wait(self, reset=True):
        ns, n = self.ctx.heap.wait_stats(reset)
        return {"wait_ms": ns / 1e6, "waits": int(n)}


class PipelineParallel(Layer):
    """Pipeline engine (1F1B / FThenB / interleaved VPP / zero-bubble H1). Parity: fleet/meta_parallel/pipeline_parallel.py."""

    _default_mode = "1F1B"

    def __init__(self, layers, hcg, strategy):
        super().__init__()
        if not isinstance(layers, PipelineLayer):
            raise TypeError("The Layer should be a derived class of PipelineLayer.")
        self._layers = layers
        self._hcg, self._strategy = hcg, strategy
        pc = strategy.pipeline_configs
        self.accumulate_steps = int(pc.get("accumulate_steps", 1))
        self.micro_batch_size = int(pc.get("micro_batch_size", 1))
        self.num_stages = hcg.get_pipe_parallel_world_size()
        self.stage_id = hcg.get_stage_id()
        self.is_first = self.stage_id == 0
        self.is_last = self.stage_id == self.num_stages - 1
        self.V = int(getattr(layers, "_num_virtual", 1) or 1)
        self.G = self.V * self.num_stages
        mode = pc.get("schedule_mode", None) or self._default_mode
        if self.V > 1 and str(mode).upper() == "1F1B":
            mode = self._default_mode if self._default_mode != "1F1B" else "VPP"
        self.schedule_mode = mode
        self.zb_pending = pc.get("zero_bubble_max_pending", None)
        self._fallback = _GroupTransport(hcg, ring=self.V > 1) if dist.is_initialized() and self.num_stages > 1 else None
        self._transport = None
        self._meta = None
        self._step = 0
        self._ops_cache = {}
        self.total_loss = None
        self._user_hooks = {}

    def parameters(self, include_sublayers=True):
        return self._layers.parameters(include_sublayers)

    def named_parameters(self, prefix="", include_sublayers=True):
        return self._layers.named_parameters(prefix, include_sublayers)

    def state_dict(self, *a, **k):
        return self._layers.state_dict(*a, **k)

    def set_state_dict(self, *a, **k):
        return self._layers.set_state_dict(*a, **k)

    def forward(self, *a, **k):
        return self._layers(*a, **k)

    # ---- helpers -------------------------------------------------------------------------------------------------
    def _micro(self, data, i):
        if data is None:
            return None
        lo, hi = i * self.micro_batch_size, (i + 1) * self.micro_batch_size
        if isinstance(data, (tuple, list)):
            return type(data)(self._micro(d, i) for d in data)
        return data[lo:hi]

    def _ops(self, M):
        key = (self.schedule_mode, M)
        ops = self._ops_cache.get(key)
        if ops is None:
            kw = {}
            if str(self.schedule_mode).upper().startswith("ZB") and self.zb_pending is not None:
                kw["max_pending_w"] = int(self.zb_pending)
            all_ops = pp_schedule.build(self.schedule_mode, self.num_stages, M, self.V, **kw)
            pp_schedule.simulate(all_ops, self.num_stages, self.V, merged_w=not any(o[0] == "W" for o in all_ops[0]))   # raises on a deadlock
            ops = self._ops_cache[key] = all_ops[self.stage_id]
        return ops

    def _learn_meta(self, inputs):
        """Shape / dtype of the tensor that travels between stages, learnt once: global stage 0 runs its first chunk on one micro-batch
        without autograd and broadcasts the result's meta over the pipe group; the mailbox is sized from it."""
        if self._meta is not None:
            return
        info = [None]
        if self.stage_id == 0:
            with torch.no_grad():
                o = self._layers(self._micro(inputs, 0), chunk_id=0)
            info = [(tuple(o.shape), _DT.index(_raw(o).dtype))]
            del o
        pg = self._hcg.get_pipe_parallel_group().pg
        dist.broadcast_object_list(info, src=self._hcg.get_rank_from_stage(0), group=pg)
        self._meta = (tuple(info[0][0]), _DT[info[0][1]])
        self._transport = self._fallback
        from ...framework.flags import flag
        from ...parallel import symm

        dev_ok = torch.cuda.is_available() and dist.get_backend(pg) == "nccl"
        if dev_ok and flag("FLAGS_b200_pp_mailbox", True) and symm.available():
            shape, dt = self._meta
            nbytes = torch.empty(0, dtype=dt).element_size()
            for d in shape:
                nbytes *= int(d)
            n_keys = self.V * self.accumulate_steps
            ctx = symm.context_for(self._hcg.get_pipe_parallel_group(), heap_bytes=_MailboxTransport.heap_bytes(nbytes, n_keys))
            if ctx is not None and ctx.heap.size() >= _MailboxTransport.heap_bytes(nbytes, n_keys):
                self._transport = _MailboxTransport(self._hcg, ctx, nbytes, n_keys)

    def transport_name(self):
        return self._transport.name if self._transport is not None else None

    def exposed_wait(self, reset=True):
        """Exposed pipeline wait accounted on the device by the mailbox wait kernels since the last reset (None for NCCL / gloo)."""
        return self._transport.exposed_wait(reset) if self._transport is not None else None

    # ---- the engine ----------------------------------------------------------------------------------------------
    def forward_backward_pipeline(self, data, scaler=None):
        from ...kernels import wgrad as WG

        inputs, labels = data if isinstance(data, (tuple, list)) and len(data) == 2 else (data, None)
        M, S, V, G, s = self.accumulate_steps, self.num_stages, self.V, self.G, self.stage_id
        self._learn_meta(inputs)
        ops = self._ops(M)
        zb = any(o[0] == "W" for o in ops)
        self._step += 1
        step, tr, meta = self._step, self._transport, self._meta
        scale = 1.0 / M
        self.total_loss = None
        acts, wq = {}, {}
        nxt, prv = (s + 1) % S, (s - 1) % S
        plan = WG.planning() if zb else contextlib.nullcontext()
        with plan:
            for kind, v, m in ops:
                g = v * S + s
                key = v * M + m
                if kind == "F":
                    if g == 0:
                        leaf, x = None, self._micro(inputs, m)
                    else:
                        # the message of global stage g-1 was sent under the key of ITS chunk (v for s > 0, v-1 across the ring)
                        src_key = (v if s > 0 else v - 1) * M + m
                        leaf = tr.recv("f", src_key, meta[0], meta[1], prv, step).detach().requires_grad_(True)
                        x = _w(leaf)
                    out = self._layers(x, chunk_id=v)
                    if g == G - 1:
                        assert self._layers._loss_fn is not None, "loss_fn is required on the last stage"
                        loss = self._layers._loss_fn(out, self._micro(labels, m)) * scale
                        self.total_loss = loss.detach() if self.total_loss is None else self.total_loss + loss.detach()
                        acts[key] = (leaf, loss)
                    else:
                        acts[key] = (leaf, out)
                        tr.send("f", key, _raw(out).detach(), nxt, step)
                elif kind == "B":
                    leaf, out = acts.pop(key)
                    q = wq.setdefault(key, []) if zb else None
                    with (WG.deferring(q) if zb else contextlib.nullcontext()):
                        if g == G - 1:
                            (scaler.scale(out) if scaler is not None else out).backward()
                        else:
                            og = tr.recv("b", key, tuple(_raw(out).shape), _raw(out).dtype, nxt, step)
                            torch.autograd.backward(_raw(out), grad_tensors=og)
                    if g != 0:
                        dst_key = (v if s > 0 else v - 1) * M + m     # keyed like the forward message it answers
                        tr.send("b", dst_key, torch.Tensor.grad.__get__(leaf), prv, step)
                    del leaf, out
                else:   # W
                    WG.flush(wq.pop(key, []))
        for q in wq.values():
            WG.flush(q)
        tr.end_step()
        self._layers.allreduce_shared_weight_gradients()
        return self._broadcast_loss()

    def _broadcast_loss(self):
        pg = self._hcg.get_pipe_parallel_group().pg
        dev = self._transport.dev if self._transport is not None else torch.device("cpu")
        last = self.stage_id == self.num_stages - 1
        loss = self.total_loss.float().reshape(1).to(dev) if last else torch.zeros(1, dtype=torch.float32, device=dev)
        src = self._hcg.get_rank_from_stage(self.num_stages - 1)
        dist.broadcast(loss, src=src, group=pg)
        return _w(loss.reshape([]))

    # ---- small parity helpers (pipeline_parallel.py) ---------------------------------------------------------------------------
    def is_pipeline_first_stage(self, ignore_virtual=False):
        return self.is_first

    def is_pipeline_last_stage(self, ignore_virtual=False):
        return self.is_last

    def set_virtual_pipeline_rank(self, rank):
        self._virtual_pp_rank = rank

    def register_hook(self, location, hook):
        """User callbacks around the schedule: location in {"forward_begin", "forward_end", "backward_begin", "backward_end"}."""
        self.__dict__.setdefault("_user_hooks", {}).setdefault(str(location), []).append(hook)

    def timer_printer(self):
        pass

    def get_static_scheduler(self):
        """The op list of this stage as text, e.g. 'f0;f1;b0;f2;b1;w0;...' (chunk suffix '@v' with virtual stages)."""
        return ";".join(f"{k.lower()}{m}" + (f"@{v}" if self.V > 1 else "") for k, v, m in self._ops(self.accumulate_steps))

    def train_batch(self, data, optimizer, lr_scheduler=None, scaler=None, loss_fn_idx=0, return_micro_batch_loss=False):
        self._layers.train()
        loss = self.forward_backward_pipeline(data, scaler)
        if scaler is not None:
            scaler.step(optimizer)
            scaler.update()
        else:
            optimizer.step()
        optimizer.clear_grad()
        if lr_scheduler is not None:
            lr_scheduler.step()
        return loss

    @torch.no_grad()
    def eval_batch(self, data, compute_loss=False, loss_fn_idx=0):
        self._layers.eval()
        inputs, labels = data if isinstance(data, (tuple, list)) and len(data) == 2 else (data, None)
        self._learn_meta(inputs)
        M, S, V, G, s = self.accumulate_steps, self.num_stages, self.V, self.G, self.stage_id
        self._step += 1
        step, tr, meta = self._step, self._transport, self._meta
        nxt, prv = (s + 1) % S, (s - 1) % S
        outs = []
        self.total_loss = None
        for kind, v, m in pp_schedule.build("FThenB", S, M, V)[s]:
            if kind != "F":
                continue
            g, key = v * S + s, v * M + m
            if g == 0:
                x = self._micro(inputs, m)
            else:
                x = _w(tr.recv("f", (v if s > 0 else v - 1) * M + m, meta[0], meta[1], prv, step))
            out = self._layers(x, chunk_id=v)
            if g == G - 1:
                if compute_loss:
                    l = self._layers._loss_fn(out, self._micro(labels, m)) / M
                    self.total_loss = l if self.total_loss is None else self.total_loss + l
                outs.append(out)
            else:
                tr.send("f", key, _raw(out), nxt, step)
        tr.end_step()
        if compute_loss:
            return self._broadcast_loss()
        return outs


class PipelineParallelWithInterleave(PipelineParallel):
    """Virtual-pipeline (interleaved 1F1B) engine. Parity: pipeline_parallel.py:PipelineParallelWithInterleave.
    Every rank owns V model chunks (global stage g = v * pp + rank); micro-batches advance in groups of pp through the chunks, which
    cuts the ramp (bubble) to 1/V of the plain schedule at the price of V times as many hops (see pp_schedule._interleaved)."""

    _default_mode = "VPP"

    def __init__(self, layers, hcg, strategy):
        super().__init__(layers, hcg, strategy)
        self.pp = self.num_stages


class PipelineParallelWithInterleaveFthenB(PipelineParallelWithInterleave):
    """All forwards of all chunks, then all backwards. Parity: pipeline_parallel.py:2256 PipelineParallelWithInterleaveFthenB."""

    _default_mode = "FThenB"


class PipelineParallelZeroBubble(PipelineParallel):
    """ZB-H1: input-gradient (B) and weight-gradient (W) passes scheduled separately. Parity (role):
    passes/pipeline_scheduler_pass/pipeline_zero_bubble.py:62."""

    _default_mode = "ZBH1"
