"""Whole-training-step CUDA-graph capture.

The reference turns a dygraph train step into one static program (dy2static: forward program + backward via the `run_program`
op, python/paddle/jit/dy2static/partial_program.py) so that per-op Python / dispatch cost disappears.  On B200 the same goal is
reached without a tracing compiler: the step — forward, backward, gradient clipping, fused AdamW, gradient zeroing — is recorded
once into a CUDA graph and replayed.  What makes this possible here:

  * parameters, gradients and optimizer state live in flat arenas (parallel/arena.py): addresses never change, `clear_grad` is a
    memset of the gradient slab;
  * the fused AdamW reads lr and the bias corrections from a device tensor (`AdamWArgs::dyn`), and the clip norm / found-inf /
    loss-scale through device pointers, so no launch argument depends on the step number;
  * every hand-written kernel launches on the current stream and does not synchronise.

Usage:

    step = paddle.jit.capture_train_step(lambda x, y: loss_fn(model(x), y), optimizer)   # fn returns the loss
    for x, y in loader:
        loss = step(x, y)            # first `warmup` calls run eagerly, then one capture, then replays

`fn` must not read tensors on the host (`.item()`, `.numpy()`, data-dependent Python branches) and input shapes must stay fixed;
if the capture fails the step silently keeps running eagerly (`step.captured` tells which)."""
from __future__ import annotations

import torch

from ..tensor import Tensor


def _tensors(args, kwargs):
    return [a for a in list(args) + [kwargs[k] for k in sorted(kwargs)] if isinstance(a, torch.Tensor)]


def _sig(args, kwargs):
    return tuple((tuple(t.shape), t.dtype, t.device) for t in _tensors(args, kwargs)) + tuple(
        (k, v) for k, v in sorted(kwargs.items()) if not isinstance(v, torch.Tensor) and isinstance(v, (int, float, bool, str, type(None))))


def _clone_tree(o):
    if isinstance(o, torch.Tensor):
        return o.clone()
    if isinstance(o, (list, tuple)):
        return type(o)(_clone_tree(i) for i in o)
    if isinstance(o, dict):
        return {k: _clone_tree(v) for k, v in o.items()}
    return o


def _detach_tree(o):
    if isinstance(o, torch.Tensor):
        return o.detach()
    if isinstance(o, (list, tuple)):
        return type(o)(_detach_tree(i) for i in o)
    if isinstance(o, dict):
        return {k: _detach_tree(v) for k, v in o.items()}
    return o


class CapturedTrainStep:
    def __init__(self, fn, optimizer, warmup=3, backward=True):
        self._fn, self._opt, self._warmup, self._backward = fn, optimizer, max(int(warmup), 1), backward
        self._calls = 0
        self._graph = None
        self._static_in, self._static_out, self._key = None, None, None
        self._failed = None
        self.replays = 0

    # -- one eager step ---------------------------------------------------------------------------------------------------
    def _eager_step(self, args, kwargs):
        out = self._fn(*args, **kwargs)
        if self._backward:
            loss = out[0] if isinstance(out, (tuple, list)) else out
            loss.backward()
            self._opt.step()
            self._opt.clear_grad()
            # hand back values, not graph handles: a loss that keeps last iteration's autograd graph alive also keeps its
            # AccumulateGrad nodes (bound to the stream they were created on), which breaks the stream capture later
            out = _detach_tree(out)
        return out

    @property
    def captured(self):
        return self._graph is not None

    @property
    def failure(self):
        return self._failed

    def _capturable(self, args, kwargs):
        ts = _tensors(args, kwargs)
        if not ts or not all(t.is_cuda for t in ts):
            return False
        opt = self._opt
        return getattr(opt, "_arena", None) is not None and hasattr(opt, "_refresh_dyn_hparams") and opt._arena_ok()

    def __call__(self, *args, **kwargs):
        self._calls += 1
        if self._failed is not None or not self._capturable(args, kwargs):
            return self._eager_step(args, kwargs)
        if self._graph is not None:
            if _sig(args, kwargs) != self._key:      # new shapes: the recorded graph does not apply
                self._opt._aux.pop("dyn_hparams", None)
                out = self._eager_step(args, kwargs)
                self._opt._refresh_dyn_hparams()
                return out
            return self._replay(args, kwargs)
        if self._calls <= self._warmup:
            return self._eager_step(args, kwargs)
        return self._capture(args, kwargs)

    def _replay(self, args, kwargs):
        opt = self._opt
        opt._refresh_dyn_hparams()
        for s, a in zip(self._static_in, _tensors(args, kwargs)):
            torch.Tensor.copy_(s, a, non_blocking=True)
        self._graph.replay()
        opt._step_count += 1
        self.replays += 1
        return _clone_tree(self._static_out)

    def _capture(self, args, kwargs):
        opt = self._opt
        result = None
        try:
            static_args = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
            static_kwargs = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in kwargs.items()}
            opt._refresh_dyn_hparams()
            # one more step on a side stream with the device-side hyper-parameters, so every lazily created buffer exists and the
            # autograd engine's stream bookkeeping is settled before recording
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                out = self._eager_step(static_args, static_kwargs)
            torch.cuda.current_stream().wait_stream(side)
            result = _clone_tree(out)
            opt._refresh_dyn_hparams()
            count = opt._step_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self._eager_step(static_args, static_kwargs)
            opt._step_count = count          # recording launches nothing; the replay below is the real next step
            self._graph, self._static_out, self._key = g, static_out, _sig(args, kwargs)
            self._static_in = _tensors(static_args, static_kwargs)
            return result
        except Exception as e:  # noqa: BLE001  (host sync / unsupported op inside fn: stay eager)
            import traceback

            first = e
            while first.__context__ is not None:      # the error raised inside the recording, not the one from ending it
                first = first.__context__
            self._failed = f"{type(first).__name__}: {first}" if first is e else f"{type(first).__name__}: {first}  [then {type(e).__name__}]"
            self.failure_traceback = "".join(traceback.format_exception(type(first), first, first.__traceback__))
            torch.cuda.synchronize()
            self._reset_generator_capture_state()
            opt._aux.pop("dyn_hparams", None)
            return result if result is not None else self._eager_step(args, kwargs)

    @staticmethod
    def _reset_generator_capture_state():
        """A recording that dies half-way never reaches the generator's capture epilogue, and every later eager RNG call then fails
        with "Offset increment outside graph capture".  Recording one empty graph to completion runs prologue + epilogue again."""
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                pass
            del g
        except Exception:  # noqa: BLE001
            pass


def capture_train_step(fn, optimizer, warmup=3):
    """Wrap `fn(*inputs) -> loss` + backward + `optimizer.step()` + `optimizer.clear_grad()` into a replayable CUDA graph."""
    return CapturedTrainStep(fn, optimizer, warmup)
