"""Autograd. Parity: python/paddle/autograd/ (backward_mode.py, py_layer.py, autograd.py),
paddle/fluid/eager/backward.cc.  The engine is torch.autograd; this module gives it the paddle surface.
"""
from __future__ import annotations

import contextlib

import torch

from ..tensor import Tensor

__all__ = ["backward", "grad", "PyLayer", "PyLayerContext", "no_grad", "enable_grad", "set_grad_enabled",
           "is_grad_enabled", "jacobian", "hessian", "saved_tensors_hooks"]


def _wrap(x):
    if isinstance(x, torch.Tensor) and not isinstance(x, Tensor):
        return x.as_subclass(Tensor)
    if isinstance(x, (tuple, list)):
        return type(x)(_wrap(i) for i in x)
    return x


class no_grad(torch.no_grad):
    """paddle.no_grad: context manager and decorator."""


class enable_grad(torch.enable_grad):
    pass


class set_grad_enabled(torch.set_grad_enabled):
    pass


def is_grad_enabled():
    return torch.is_grad_enabled()


def backward(tensors, grad_tensors=None, retain_graph=False):
    if isinstance(tensors, torch.Tensor):
        tensors = [tensors]
    torch.autograd.backward(list(tensors), grad_tensors=grad_tensors, retain_graph=retain_graph)


def grad(outputs, inputs, grad_outputs=None, retain_graph=None, create_graph=False, only_inputs=True,
         allow_unused=False, no_grad_vars=None):
    """paddle.grad. Parity: python/paddle/base/dygraph/base.py:grad."""
    single = isinstance(inputs, torch.Tensor)
    outs = [outputs] if isinstance(outputs, torch.Tensor) else list(outputs)
    ins = [inputs] if single else list(inputs)
    if grad_outputs is not None and isinstance(grad_outputs, torch.Tensor):
        grad_outputs = [grad_outputs]
    if retain_graph is None:
        retain_graph = create_graph
    res = torch.autograd.grad(outs, ins, grad_outputs=grad_outputs, retain_graph=retain_graph,
                              create_graph=create_graph, allow_unused=allow_unused)
    res = [_wrap(r) for r in res]
    return res


class PyLayerContext:
    """ctx object handed to PyLayer.forward/backward (thin view over torch's FunctionCtx)."""

    def __init__(self, ctx):
        object.__setattr__(self, "_ctx", ctx)

    def save_for_backward(self, *tensors):
        self._ctx.save_for_backward(*tensors)

    def saved_tensor(self):
        return tuple(_wrap(t) for t in self._ctx.saved_tensors)

    @property
    def saved_tensors(self):
        return self.saved_tensor()

    def mark_not_inplace(self, *a):
        pass

    def mark_non_differentiable(self, *tensors):
        self._ctx.mark_non_differentiable(*tensors)

    def set_materialize_grads(self, value):
        self._ctx.set_materialize_grads(value)

    @property
    def needs_input_grad(self):
        return self._ctx.needs_input_grad

    def __getattr__(self, k):
        return getattr(self._ctx, k)

    def __setattr__(self, k, v):
        setattr(self._ctx, k, v)


class _PyLayerMeta(type):
    def __init__(cls, name, bases, attrs):
        super().__init__(name, bases, attrs)
        if name == "PyLayer" and not bases:
            return
        user_fwd, user_bwd = cls.forward, cls.backward

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *args, **kwargs):
                return user_fwd(PyLayerContext(ctx), *args, **kwargs)

            @staticmethod
            def backward(ctx, *grads):
                out = user_bwd(PyLayerContext(ctx), *[_wrap(g) for g in grads])
                return out

        _Fn.__name__ = name
        cls._fn = _Fn


class PyLayer(metaclass=_PyLayerMeta):
    """Custom differentiable op. Parity: python/paddle/autograd/py_layer.py:PyLayer."""

    @staticmethod
    def forward(ctx, *args, **kwargs):
        raise NotImplementedError

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError

    @classmethod
    def apply(cls, *args, **kwargs):
        if kwargs:
            # torch Functions take positional args only; bind kwargs by closure
            keys = list(kwargs)
            vals = [kwargs[k] for k in keys]
            n = len(args)
            user_fwd, user_bwd = cls.forward, cls.backward

            class _FnKw(torch.autograd.Function):
                @staticmethod
                def forward(ctx, *a):
                    return user_fwd(PyLayerContext(ctx), *a[:n], **dict(zip(keys, a[n:])))

                @staticmethod
                def backward(ctx, *grads):
                    out = user_bwd(PyLayerContext(ctx), *[_wrap(g) for g in grads])
                    out = out if isinstance(out, tuple) else (out,)
                    return out + (None,) * (n + len(keys) - len(out))

            return _wrap(_FnKw.apply(*args, *vals))
        return _wrap(cls._fn.apply(*args))


def jacobian(ys, xs, batch_axis=None):
    """Dense Jacobian d ys / d xs. Parity: python/paddle/autograd/autograd.py:jacobian (eagerly materialised)."""
    def _one(y, x):
        yf = y.reshape(-1) if batch_axis is None else y.reshape(y.shape[0], -1)
        rows = []
        if batch_axis is None:
            for i in range(yf.numel()):
                (g,) = torch.autograd.grad(yf[i], x, retain_graph=True, create_graph=True, allow_unused=True)
                rows.append(torch.zeros_like(x).reshape(-1) if g is None else g.reshape(-1))
            return _wrap(torch.stack(rows, 0))
        for i in range(yf.size(1)):
            (g,) = torch.autograd.grad(yf[:, i].sum(), x, retain_graph=True, create_graph=True, allow_unused=True)
            rows.append(torch.zeros_like(x).reshape(x.shape[0], -1) if g is None else g.reshape(x.shape[0], -1))
        return _wrap(torch.stack(rows, 1))

    if isinstance(ys, (list, tuple)):
        return tuple(jacobian(y, xs, batch_axis) for y in ys)
    if isinstance(xs, (list, tuple)):
        return tuple(_one(ys, x) for x in xs)
    return _one(ys, xs)


def hessian(ys, xs, batch_axis=None):
    """Hessian of a scalar (or per-sample scalar) w.r.t. xs. Parity: python/paddle/autograd/autograd.py:hessian."""
    if isinstance(xs, (list, tuple)):
        gs = torch.autograd.grad(ys.sum() if batch_axis is not None else ys, list(xs), create_graph=True)
        return tuple(tuple(jacobian(g, x, batch_axis) for x in xs) for g in gs)
    (g,) = torch.autograd.grad(ys.sum() if batch_axis is not None else ys, xs, create_graph=True)
    return jacobian(g, xs, batch_axis)


class saved_tensors_hooks(torch.autograd.graph.saved_tensors_hooks):
    """Parity: python/paddle/autograd/saved_tensors_hooks.py."""
