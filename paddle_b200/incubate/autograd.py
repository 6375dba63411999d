"""paddle.incubate.autograd: functional jvp / vjp / Jacobian / Hessian and the prim switches.
Parity: python/paddle/incubate/autograd/{functional.py,primapi.py}."""
from __future__ import annotations

import torch

from ..tensor import Tensor

_prim = [False]


def enable_prim():
    _prim[0] = True


def disable_prim():
    _prim[0] = False


def prim_enabled():
    return _prim[0]


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def _seq(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def vjp(func, xs, v=None):
    """Returns (func(xs), v^T J)."""
    single = not isinstance(xs, (list, tuple))
    ins = [_raw(x).detach().requires_grad_(True) for x in _seq(xs)]
    ys = func(*[_w(i) for i in ins])
    ys_l = [_raw(y) for y in _seq(ys)]
    vs = [torch.ones_like(y) for y in ys_l] if v is None else [_raw(t) for t in _seq(v)]
    grads = torch.autograd.grad(ys_l, ins, vs, allow_unused=True)
    grads = [_w(g if g is not None else torch.zeros_like(i)) for g, i in zip(grads, ins)]
    return ys, (grads[0] if single else grads)


def jvp(func, xs, v=None):
    """Returns (func(xs), J v) by the double-vjp trick."""
    single = not isinstance(xs, (list, tuple))
    ins = [_raw(x).detach().requires_grad_(True) for x in _seq(xs)]
    ys = func(*[_w(i) for i in ins])
    ys_l = [_raw(y) for y in _seq(ys)]
    vs = [torch.ones_like(i) for i in ins] if v is None else [_raw(t) for t in _seq(v)]
    us = [torch.zeros_like(y, requires_grad=True) for y in ys_l]
    g = torch.autograd.grad(ys_l, ins, us, create_graph=True, allow_unused=True)
    g = [gi if gi is not None else torch.zeros_like(i) for gi, i in zip(g, ins)]
    out = torch.autograd.grad(g, us, vs, allow_unused=True)
    out = [_w(o if o is not None else torch.zeros_like(y)) for o, y in zip(out, ys_l)]
    return ys, (out[0] if (single and not isinstance(ys, (list, tuple))) else out)


class Jacobian:
    """Lazy Jacobian J[i, j] = d ys_flat[i] / d xs_flat[j] (is_batched: leading dim is a batch)."""

    def __init__(self, func, xs, is_batched=False):
        self._func, self._xs, self._batched = func, xs, is_batched
        self._mat = None

    def _compute(self):
        if self._mat is None:
            ins = [_raw(x).detach() for x in _seq(self._xs)]

            def f(*a):
                out = self._func(*[_w(t) for t in a])
                outs = [_raw(o) for o in _seq(out)]
                if self._batched:
                    return torch.cat([o.reshape(o.shape[0], -1) for o in outs], 1)
                return torch.cat([o.reshape(-1) for o in outs])

            j = torch.autograd.functional.jacobian(f, tuple(ins), vectorize=False)
            if self._batched:   # [B, M, B, N_k] -> take the batch diagonal
                b = ins[0].shape[0]
                parts = [jk.reshape(b, jk.shape[1], b, -1)[torch.arange(b), :, torch.arange(b)] for jk in j]
                self._mat = torch.cat(parts, -1)
            else:
                self._mat = torch.cat([jk.reshape(jk.shape[0], -1) for jk in j], -1)
        return self._mat

    @property
    def shape(self):
        return list(self._compute().shape)

    def __getitem__(self, idx):
        return _w(self._compute()[idx])


class Hessian(Jacobian):
    def __init__(self, func, xs, is_batched=False):
        def grad_fn(*a):
            ins = [_raw(t) if t.requires_grad else _raw(t).requires_grad_(True) for t in a]
            y = _raw(func(*[_w(t) for t in ins]))
            g = torch.autograd.grad(y.sum(), ins, create_graph=True)
            return [_w(t) for t in g] if len(g) > 1 else _w(g[0])

        super().__init__(grad_fn, xs, is_batched)


def forward_grad(outputs, inputs, grad_inputs=None):
    """Forward-mode derivative of already-computed `outputs` w.r.t. `inputs` in direction `grad_inputs` (default ones), by the
    double-backward identity  J v = d/du [ (J^T u) . v ].  Parity: python/paddle/incubate/autograd/primapi.py:forward_grad."""
    outs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
    ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
    vs = [torch.ones_like(_raw(i)) for i in ins] if grad_inputs is None else [_raw(v) for v in (grad_inputs if isinstance(grad_inputs, (list, tuple)) else [grad_inputs])]
    res = []
    for o in outs:
        u = torch.zeros_like(_raw(o), requires_grad=True)
        with torch._C.DisableTorchFunctionSubclass():   # the tensors themselves (an as_subclass alias is not the graph leaf)
            gs = torch.autograd.grad(o, ins, u, create_graph=True, allow_unused=True)
            terms = [(g * v).sum() for g, v in zip(gs, vs) if g is not None]
            jv = torch.autograd.grad(torch.stack(terms).sum(), u, allow_unused=True)[0] if terms else None
        res.append(_w(torch.zeros_like(_raw(o)) if jv is None else _raw(jv)))
    return res if isinstance(outputs, (list, tuple)) else res[0]


def grad(outputs, inputs, grad_outputs=None):
    from ..autograd import grad as _g

    return _g(outputs, inputs, grad_outputs)
