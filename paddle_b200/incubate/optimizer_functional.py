"""paddle.incubate.optimizer.functional: quasi-Newton minimisers on a callable objective.
Parity: python/paddle/incubate/optimizer/functional/{bfgs,lbfgs,line_search}.py (returns the same result tuples)."""
from __future__ import annotations

import torch

from ..tensor import Tensor


def _val_grad(f, x):
    x = x.detach().clone().requires_grad_(True)
    y = f(x.as_subclass(Tensor))
    y = y.as_subclass(torch.Tensor) if isinstance(y, torch.Tensor) else torch.as_tensor(y, dtype=x.dtype)
    (g,) = torch.autograd.grad(y.reshape(()), x)
    return y.detach().reshape(()), g.detach()


def _strong_wolfe(f, x, d, f0, g0, alpha0, max_iters, c1=1e-4, c2=0.9):
    """Bracketing + zoom line search for the strong Wolfe conditions. Returns (alpha, f_new, g_new, evaluations)."""
    d0 = float(g0 @ d)
    evals = 0

    def phi(a):
        nonlocal evals
        evals += 1
        v, g = _val_grad(f, x + a * d)
        return float(v), g, float(g @ d)

    def zoom(lo, hi, flo):
        for _ in range(max_iters):
            a = 0.5 * (lo + hi)
            v, g, dv = phi(a)
            if v > float(f0) + c1 * a * d0 or v >= flo:
                hi = a
            else:
                if abs(dv) <= -c2 * d0:
                    return a, v, g
                if dv * (hi - lo) >= 0:
                    hi = lo
                lo, flo = a, v
        v, g, _ = phi(lo)
        return lo, v, g

    a_prev, f_prev, a = 0.0, float(f0), float(alpha0)
    for i in range(max_iters):
        v, g, dv = phi(a)
        if v > float(f0) + c1 * a * d0 or (i > 0 and v >= f_prev):
            a, v, g = zoom(a_prev, a, f_prev)
            return a, v, g, evals
        if abs(dv) <= -c2 * d0:
            return a, v, g, evals
        if dv >= 0:
            a, v, g = zoom(a, a_prev, v)
            return a, v, g, evals
        a_prev, f_prev, a = a, v, a * 2.0
    return a_prev, f_prev, _val_grad(f, x + a_prev * d)[1], evals


def minimize_bfgs(objective_func, initial_position, max_iters=50, tolerance_grad=1e-7, tolerance_change=1e-9, initial_inverse_hessian_estimate=None,
                  line_search_fn="strong_wolfe", max_line_search_iters=50, initial_step_length=1.0, dtype="float32", name=None):
    """Returns (is_converge, num_func_calls, position, objective_value, objective_gradient, inverse_hessian_estimate)."""
    if line_search_fn != "strong_wolfe":
        raise NotImplementedError(f"line_search_fn {line_search_fn!r}: only 'strong_wolfe' is provided")
    x = initial_position.as_subclass(torch.Tensor).detach().clone().to(getattr(torch, dtype))
    n = x.numel()
    H = torch.eye(n, dtype=x.dtype) if initial_inverse_hessian_estimate is None else initial_inverse_hessian_estimate.as_subclass(torch.Tensor).detach().clone().to(x.dtype)
    v, g = _val_grad(objective_func, x)
    calls, converged = 1, False
    for _ in range(max_iters):
        if float(g.abs().max()) < tolerance_grad:
            converged = True
            break
        d = -(H @ g)
        a, v_new, g_new, ev = _strong_wolfe(objective_func, x, d, v, g, initial_step_length, max_line_search_iters)
        calls += ev
        s = a * d
        y = g_new - g
        x = x + s
        if float(s.abs().max()) < tolerance_change or abs(float(v_new) - float(v)) < tolerance_change * 1e-3:
            v, g = torch.as_tensor(v_new, dtype=x.dtype), g_new
            converged = True
            break
        sy = float(s @ y)
        if sy > 1e-12:
            rho = 1.0 / sy
            I = torch.eye(n, dtype=x.dtype)
            H = (I - rho * torch.outer(s, y)) @ H @ (I - rho * torch.outer(y, s)) + rho * torch.outer(s, s)
        v, g = torch.as_tensor(v_new, dtype=x.dtype), g_new
    w = lambda t: t.as_subclass(Tensor)  # noqa: E731
    return w(torch.tensor(converged)), w(torch.tensor(calls)), w(x), w(v.reshape(())), w(g), w(H)


def minimize_lbfgs(objective_func, initial_position, history_size=100, max_iters=50, tolerance_grad=1e-8, tolerance_change=1e-8,
                   initial_inverse_hessian_estimate=None, line_search_fn="strong_wolfe", max_line_search_iters=50, initial_step_length=1.0, dtype="float32", name=None):
    """Returns (is_converge, num_func_calls, position, objective_value, objective_gradient)."""
    if line_search_fn != "strong_wolfe":
        raise NotImplementedError(f"line_search_fn {line_search_fn!r}: only 'strong_wolfe' is provided")
    x = initial_position.as_subclass(torch.Tensor).detach().clone().to(getattr(torch, dtype))
    v, g = _val_grad(objective_func, x)
    calls, converged = 1, False
    S, Y = [], []
    for _ in range(max_iters):
        if float(g.abs().max()) < tolerance_grad:
            converged = True
            break
        q = g.clone()
        alphas = []
        for s, y in zip(reversed(S), reversed(Y)):          # two-loop recursion
            rho = 1.0 / float(y @ s)
            al = rho * float(s @ q)
            alphas.append((al, rho, s, y))
            q -= al * y
        gamma = float(S[-1] @ Y[-1]) / float(Y[-1] @ Y[-1]) if S else 1.0
        r = gamma * q
        for al, rho, s, y in reversed(alphas):
            r += s * (al - rho * float(y @ r))
        d = -r
        a, v_new, g_new, ev = _strong_wolfe(objective_func, x, d, v, g, initial_step_length, max_line_search_iters)
        calls += ev
        s, y = a * d, g_new - g
        x = x + s
        done = float(s.abs().max()) < tolerance_change
        if float(s @ y) > 1e-12:
            S.append(s)
            Y.append(y)
            if len(S) > history_size:
                S.pop(0)
                Y.pop(0)
        v, g = torch.as_tensor(v_new, dtype=x.dtype), g_new
        if done:
            converged = True
            break
    w = lambda t: t.as_subclass(Tensor)  # noqa: E731
    return w(torch.tensor(converged)), w(torch.tensor(calls)), w(x), w(v.reshape(())), w(g)
