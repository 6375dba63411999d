"""nn.utils. Parity: python/paddle/nn/utils/*.py."""
from __future__ import annotations

import torch

from ..tensor import Parameter, Tensor


def parameters_to_vector(parameters, name=None):
    return torch.cat([p.reshape([-1]) for p in parameters], 0)


def vector_to_parameters(vec, parameters, name=None):
    off = 0
    with torch.no_grad():
        for p in parameters:
            n = p.numel()
            torch.Tensor.copy_(p, vec[off:off + n].reshape(p.size()))
            off += n


def clip_grad_norm_(parameters, max_norm, norm_type=2.0, error_if_nonfinite=False):
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad.as_subclass(torch.Tensor) for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros([]).as_subclass(Tensor)
    if norm_type == float("inf"):
        total = torch.stack([g.abs().max() for g in grads]).max()
    else:
        total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.float(), norm_type) for g in grads]), norm_type)
    if error_if_nonfinite and not torch.isfinite(total):
        raise RuntimeError("non-finite gradient norm")
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    with torch.no_grad():
        for g in grads:
            g.mul_(coef.to(g.dtype))
    return total.as_subclass(Tensor)


def clip_grad_value_(parameters, clip_value):
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    with torch.no_grad():
        for p in parameters:
            if p.grad is not None:
                p.grad.as_subclass(torch.Tensor).clamp_(-clip_value, clip_value)


def _norm_except(w, dim):
    if dim is None or dim == -1:
        return torch.linalg.vector_norm(w)
    dims = [d for d in range(w.dim()) if d != dim]
    return torch.linalg.vector_norm(w, dim=dims, keepdim=True)


def weight_norm(layer, name="weight", dim=0):
    """w = g * v / ||v||. Parity: nn/utils/weight_norm_hook.py."""
    w = getattr(layer, name)
    wr = w.detach().as_subclass(torch.Tensor)
    g = Parameter(_norm_except(wr, dim).clone(), name=None)
    v = Parameter(wr.clone(), name=None)
    del layer._parameters[name]
    layer.add_parameter(name + "_g", g)
    layer.add_parameter(name + "_v", v)

    def hook(l, inputs):
        vv, gg = getattr(l, name + "_v"), getattr(l, name + "_g")
        object.__setattr__(l, name, vv * (gg / _norm_except(vv, dim)))

    layer._weight_norm_hook = layer.register_forward_pre_hook(hook)
    layer._weight_norm_cfg = (name, dim)
    hook(layer, None)
    return layer


def remove_weight_norm(layer, name="weight"):
    vv, gg = getattr(layer, name + "_v"), getattr(layer, name + "_g")
    _, dim = layer._weight_norm_cfg
    w = (vv * (gg / _norm_except(vv, dim))).detach()
    layer._weight_norm_hook.remove()
    del layer._parameters[name + "_g"], layer._parameters[name + "_v"]
    layer.__dict__.pop(name, None)
    layer.add_parameter(name, Parameter(w))
    return layer


def spectral_norm(layer, name="weight", n_power_iterations=1, eps=1e-12, dim=None):
    from .conv_norm_pool import SpectralNorm

    w = getattr(layer, name)
    if dim is None:
        dim = 1 if layer.__class__.__name__.endswith("Transpose") or layer.__class__.__name__ == "Linear" else 0
    sn = SpectralNorm(list(w.shape), dim=dim, power_iters=n_power_iterations, epsilon=eps)
    orig = Parameter(w.detach().clone())
    del layer._parameters[name]
    layer.add_parameter(name + "_orig", orig)
    layer.add_sublayer("_spectral_norm", sn)

    def hook(l, inputs):
        object.__setattr__(l, name, sn(getattr(l, name + "_orig")))

    layer._spectral_hook = layer.register_forward_pre_hook(hook)
    hook(layer, None)
    return layer
