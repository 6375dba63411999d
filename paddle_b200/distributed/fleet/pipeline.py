"""Pipeline parallelism. Parity: python/paddle/distributed/fleet/meta_parallel/parallel_layers/pp_layers.py
(LayerDesc, SharedLayerDesc, PipelineLayer, segment methods) and pipeline_parallel.py (PipelineParallel 1F1B /
FThenB / interleaved-VPP entry points, train_batch / eval_batch), pp_utils/p2p_communication.py.

Stage-to-stage activations travel over NCCL p2p (batched isend/irecv on a dedicated stream so they overlap compute);
shapes/dtypes are exchanged once and cached.
"""
from __future__ import annotations

import math
import re

import torch
import torch.distributed as dist

from ...nn.layer import Layer
from ...tensor import Tensor
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
class _P2P:
    """Stage <-> stage tensor transport with cached meta."""

    def __init__(self, hcg):
        self.hcg = hcg
        self.group = hcg.get_pipe_parallel_group().pg
        self.next_rank, self.prev_rank = hcg.next_rank, hcg.prev_rank
        self.fwd_meta = None   # (shape, dtype) of activations received from prev
        self.bwd_meta = None   # (shape, dtype) of grads received from next
        self.dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and dist.get_backend(self.group) == "nccl" else torch.device("cpu")

    _DT = [torch.float32, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.float64, torch.bool, torch.uint8]

    def _send_meta(self, t, dst):
        meta = torch.zeros(10, dtype=torch.int64, device=self.dev)
        meta[0] = t.dim()
        meta[1] = self._DT.index(t.dtype)
        for i, s in enumerate(t.shape):
            meta[2 + i] = s
        dist.send(meta, dst, group=self.group)

    def _recv_meta(self, src):
        meta = torch.zeros(10, dtype=torch.int64, device=self.dev)
        dist.recv(meta, src, group=self.group)
        m = meta.tolist()
        return tuple(m[2:2 + m[0]]), self._DT[m[1]]

    def send_forward(self, t, first_time):
        if first_time:
            self._send_meta(t, self.next_rank)
        dist.send(t.contiguous(), self.next_rank, group=self.group)

    def recv_forward(self):
        if self.fwd_meta is None:
            self.fwd_meta = self._recv_meta(self.prev_rank)
        shape, dt = self.fwd_meta
        t = torch.empty(shape, dtype=dt, device=self.dev)
        dist.recv(t, self.prev_rank, group=self.group)
        return t

    def send_backward(self, g):
        dist.send(g.contiguous(), self.prev_rank, group=self.group)

    def recv_backward(self, like):
        g = torch.empty_like(like)
        dist.recv(g, self.next_rank, group=self.group)
        return g

    def send_forward_recv_backward(self, t, first_time):
        if first_time:
            self._send_meta(t, self.next_rank)
        g = torch.empty_like(t)
        ops = [dist.P2POp(dist.isend, t.contiguous(), self.next_rank, self.group), dist.P2POp(dist.irecv, g, self.next_rank, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return g

    def send_backward_recv_forward(self, g):
        shape, dt = self.fwd_meta
        t = torch.empty(shape, dtype=dt, device=self.dev)
        ops = [dist.P2POp(dist.isend, g.contiguous(), self.prev_rank, self.group), dist.P2POp(dist.irecv, t, self.prev_rank, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return t


class PipelineParallel(Layer):
    """1F1B pipeline engine. Parity: fleet/meta_parallel/pipeline_parallel.py:PipelineParallel."""

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
        self._p2p = _P2P(hcg)
        self._sent_meta = False
        self.total_loss = None
        self.schedule_mode = pc.get("schedule_mode", "1F1B")

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

    # ---- micro-batch helpers ---------------------------------------------------------------------------------
    def _micro(self, data, i):
        if data is None:
            return None
        lo, hi = i * self.micro_batch_size, (i + 1) * self.micro_batch_size
        if isinstance(data, (tuple, list)):
            return type(data)(self._micro(d, i) for d in data)
        return data[lo:hi]

    def _forward_step(self, inp, labels_mb, scale):
        if not self.is_first:
            inp = _w(inp.requires_grad_(True))
        out = self._layers(inp)
        if self.is_last:
            assert self._layers._loss_fn is not None, "loss_fn is required on the last stage"
            loss = self._layers._loss_fn(out, labels_mb)
            loss = loss * scale
            self.total_loss = loss.detach() if self.total_loss is None else self.total_loss + loss.detach()
            return loss
        return out

    def _backward_step(self, inp, out, out_grad, scaler=None):
        if self.is_last:
            (scaler.scale(out) if scaler is not None else out).backward()
        else:
            torch.autograd.backward(_raw(out), grad_tensors=_raw(out_grad))
        if self.is_first:
            return None
        return torch.Tensor.grad.__get__(inp)

    # ---- schedule ------------------------------------------------------------------------------------------------
    def forward_backward_pipeline(self, data, scaler=None):
        inputs, labels = data if isinstance(data, (tuple, list)) and len(data) == 2 else (data, None)
        M = self.accumulate_steps
        scale = 1.0 / M
        self.total_loss = None
        p2p = self._p2p
        warm = min(self.num_stages - self.stage_id - 1, M)
        steady = M - warm
        in_q, out_q = [], []
        fwd_i = 0

        def first_input(i):
            return self._micro(inputs, i) if self.is_first else None

        def fwd(inp_recv):
            nonlocal fwd_i
            i = fwd_i
            fwd_i += 1
            x = first_input(i) if self.is_first else inp_recv
            out = self._forward_step(x, self._micro(labels, i) if self.is_last else None, scale)
            return x, out

        for _ in range(warm):
            r = None if self.is_first else p2p.recv_forward()
            x, out = fwd(r)
            if not self.is_last:
                p2p.send_forward(_raw(out).detach(), not self._sent_meta)
                self._sent_meta = True
            in_q.append(x)
            out_q.append(out)
        r = None
        if steady > 0 and not self.is_first:
            r = p2p.recv_forward()
        for k in range(steady):
            last_iter = k == steady - 1
            x, out = fwd(r)
            if self.is_last:
                og = None
            else:
                og = p2p.send_forward_recv_backward(_raw(out).detach(), not self._sent_meta)
                self._sent_meta = True
            in_q.append(x)
            out_q.append(out)
            x0, o0 = in_q.pop(0), out_q.pop(0)
            ig = self._backward_step(x0, o0, og, scaler)
            if not self.is_first:
                if last_iter:
                    p2p.send_backward(_raw(ig))
                    r = None
                else:
                    r = p2p.send_backward_recv_forward(_raw(ig))
        for _ in range(warm):
            x0, o0 = in_q.pop(0), out_q.pop(0)
            og = None if self.is_last else p2p.recv_backward(_raw(o0).detach())
            ig = self._backward_step(x0, o0, og, scaler)
            if not self.is_first:
                p2p.send_backward(_raw(ig))
        self._layers.allreduce_shared_weight_gradients()
        return self._broadcast_loss()

    def _broadcast_loss(self):
        pg = self._hcg.get_pipe_parallel_group().pg
        dev = self._p2p.dev
        loss = self.total_loss.float().reshape(1).to(dev) if self.is_last else torch.zeros(1, dtype=torch.float32, device=dev)
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
        M, S, r = self.accumulate_steps, self.num_stages, self.stage_id
        warm = min(S - r - 1, M)
        return ";".join([f"f{i}" for i in range(warm)] + [f"f{warm + i};b{i}" for i in range(M - warm)] + [f"b{M - warm + i}" for i in range(warm)])

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
        p2p = self._p2p
        outs = []
        self.total_loss = None
        for i in range(self.accumulate_steps):
            x = self._micro(inputs, i) if self.is_first else _w(p2p.recv_forward())
            out = self._layers(x)
            if self.is_last:
                if compute_loss:
                    l = self._layers._loss_fn(out, self._micro(labels, i)) / self.accumulate_steps
                    self.total_loss = l if self.total_loss is None else self.total_loss + l
                outs.append(out)
            else:
                p2p.send_forward(_raw(out), not self._sent_meta)
                self._sent_meta = True
        if compute_loss:
            return self._broadcast_loss()
        return outs


class PipelineParallelWithInterleave(PipelineParallel):
    """Virtual-pipeline (interleaved) engine. Parity: pipeline_parallel.py:PipelineParallelWithInterleave.

    Every rank owns V model chunks (global stage g = v * pp + rank).  The schedule is the lock-step diagonal: at forward tick
    t global stage g runs micro-batch t - g, so a rank works on up to V chunks per tick and the pipeline fills after pp - 1
    ticks of ONE CHUNK each — the ramp (bubble) is 1/V of the plain schedule.  Each tick ends with exactly one batched
    isend/irecv per rank (sends to the next rank, receives what it needs for the next tick), which keeps the ring
    (last rank -> first rank for the next chunk) deadlock-free.  Backward mirrors it.  Activations of all micro-batches stay
    alive between the two phases (F-then-B)."""

    def __init__(self, layers, hcg, strategy):
        super().__init__(layers, hcg, strategy)
        self.V = layers._num_virtual
        self.pp = self.num_stages
        self.G = self.V * self.pp

    def _exchange(self, sends, recv_shapes, dst, src):
        """One batched p2p round: `sends` to rank dst, receive len(recv_shapes) tensors from rank src (same order on both sides)."""
        p2p = self._p2p
        bufs = [torch.empty(shape, dtype=dt, device=p2p.dev) for shape, dt in recv_shapes]
        ops = [dist.P2POp(dist.isend, t.contiguous(), dst, p2p.group) for t in sends] + [dist.P2POp(dist.irecv, b, src, p2p.group) for b in bufs]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return bufs

    def forward_backward_pipeline(self, data, scaler=None):
        inputs, labels = data if isinstance(data, (tuple, list)) and len(data) == 2 else (data, None)
        M, V, pp, G, r = self.accumulate_steps, self.V, self.pp, self.G, self.stage_id
        p2p = self._p2p
        nxt, prv = p2p.next_rank, p2p.prev_rank
        scale = 1.0 / M
        self.total_loss = None
        acts = {}

        def work(t, backward=False):
            out = []
            for v in (range(V - 1, -1, -1) if backward else range(V)):
                g = r + v * pp
                m = t - ((G - 1 - g) if backward else g)
                if 0 <= m < M:
                    out.append((v, m))
            return out

        # ---- activation shape between stages: learnt once (rank 0 runs its first chunk on one micro-batch without autograd) ----
        if getattr(self, "_meta", None) is None:
            info = [None]
            if r == 0:
                with torch.no_grad():
                    o = self._layers(self._micro(inputs, 0), chunk_id=0)
                info = [(tuple(o.shape), str(o.dtype).split(".")[-1])]
            dist.broadcast_object_list(info, src=self._hcg.get_rank_from_stage(0), group=p2p.group)
            self._meta = (tuple(info[0][0]), getattr(torch, info[0][1]))
        meta = self._meta
        # ---- forward ticks ---------------------------------------------------------------------------------------------
        recv_next = {}
        for t in range(G + M - 1):
            sends = []
            for v, m in work(t):
                g = r + v * pp
                if g == 0:
                    leaf = None
                    x = self._micro(inputs, m)
                else:
                    leaf = recv_next.pop((v, m)).requires_grad_(True)    # the received buffer is the autograd leaf of this chunk
                    x = _w(leaf)
                out = self._layers(x, chunk_id=v)
                if g == G - 1:
                    loss = self._layers._loss_fn(out, self._micro(labels, m)) * scale
                    self.total_loss = loss.detach() if self.total_loss is None else self.total_loss + loss.detach()
                    acts[(v, m)] = (leaf, loss)
                else:
                    acts[(v, m)] = (leaf, out)
                    sends.append(_raw(out).detach())
            # what arrives for tick t+1: my (v, m) at t+1 with g != 0 (sent by the previous rank at this tick, same order)
            want = [(v, m) for v, m in work(t + 1) if r + v * pp != 0] if t + 1 < G + M - 1 else []
            got = self._exchange(sends, [meta] * len(want), nxt, prv)
            for key, buf in zip(want, got):
                recv_next[key] = buf
        # ---- backward ticks ------------------------------------------------------------------------------------------
        grad_next = {}
        for t in range(G + M - 1):
            sends = []
            for v, m in work(t, backward=True):
                g = r + v * pp
                x, out = acts.pop((v, m))
                if g == G - 1:
                    (scaler.scale(out) if scaler is not None else out).backward()
                else:
                    torch.autograd.backward(_raw(out), grad_tensors=grad_next.pop((v, m)))
                if g != 0:
                    sends.append(x.grad)
            want = [(v, m) for v, m in work(t + 1, backward=True) if r + v * pp != G - 1] if t + 1 < G + M - 1 else []
            got = self._exchange(sends, [meta] * len(want), prv, nxt)
            for key, buf in zip(want, got):
                grad_next[key] = buf
        self._layers.allreduce_shared_weight_gradients()
        return self._broadcast_loss()

    def _broadcast_loss(self):
        # the loss lives on the last rank (global stage G - 1 = chunk V - 1 of the last rank)
        self.is_last = self.stage_id == self.num_stages - 1
        return super()._broadcast_loss()
