"""RMSNorm / LayerNorm. Parity: paddle.incubate.nn.functional.fused_rms_norm / fused_layer_norm, F.layer_norm."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..framework.recording import recordable
from . import ext, raw, use_fused, wrap

_FUSED_DTYPES = (torch.float32, torch.float16, torch.bfloat16)


def _fusable(x, width):
    n = 16 // x.element_size()
    return use_fused(x) and x.dtype in _FUSED_DTYPES and width % n == 0 and width <= 256 * 8 * n


class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps, residual):
        xc = x.contiguous()
        y, rstd, res_out = ext().rms_norm_fwd(xc, residual.contiguous() if residual is not None else None, w, None, eps)
        h = res_out if residual is not None else xc
        ctx.save_for_backward(h, w, rstd)
        ctx.has_res = residual is not None
        if residual is not None:
            ctx.mark_non_differentiable()
            return y, res_out
        return y

    @staticmethod
    def backward(ctx, dy, dres=None):
        h, w, rstd = ctx.saved_tensors
        dx, dw = ext().rms_norm_bwd(dy.contiguous(), h, w, rstd)
        if ctx.has_res:
            if dres is not None:
                dx = dx + dres
            return dx, dw, None, dx
        return dx, dw, None, None


def rms_norm_ref(x, w, eps, bias=None):
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    if w is not None:
        y = y * w.float()
    if bias is not None:
        y = y + bias.float()
    return y.to(x.dtype)


@recordable
def rms_norm(x, weight=None, eps=1e-6, bias=None, residual=None):
    """y = rmsnorm(x [+ residual]) * weight (+ bias).  With ``residual`` returns (y, x + residual)."""
    x, weight, bias, residual = raw(x), raw(weight), raw(bias), raw(residual)
    if _fusable(x, x.shape[-1]) and (weight is None or weight.dtype == x.dtype):
        out = _RMSNorm.apply(x, weight, float(eps), residual)
        if residual is not None:
            y, h = out
            if bias is not None:
                y = y + bias
            return wrap(y), wrap(h)
        if bias is not None:
            out = out + bias
        return wrap(out)
    if residual is not None:
        h = x + residual
        return wrap(rms_norm_ref(h, weight, eps, bias)), wrap(h)
    return wrap(rms_norm_ref(x, weight, eps, bias))


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        xc = x.contiguous()
        y, mean, rstd = ext().layer_norm_fwd(xc, w, b, eps)
        ctx.save_for_backward(xc, w, mean, rstd)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dx, dw, db = ext().layer_norm_bwd(dy.contiguous(), x, w, mean, rstd, ctx.has_b)
        return dx, (dw if w is not None else None), (db if ctx.has_b else None), None


@recordable
def layer_norm(x, normalized_shape, weight=None, bias=None, eps=1e-5):
    x, weight, bias = raw(x), raw(weight), raw(bias)
    width = 1
    for s in normalized_shape:
        width *= int(s)
    if _fusable(x, width) and len(normalized_shape) == 1 and (weight is None or weight.dtype == x.dtype) and (bias is None or bias.dtype == x.dtype):
        return wrap(_LayerNorm.apply(x, weight, bias, float(eps)))
    return wrap(F.layer_norm(x, tuple(int(s) for s in normalized_shape), weight, bias, eps))
