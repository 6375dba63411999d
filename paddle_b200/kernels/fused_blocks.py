"""Memory-lean fused blocks: RMSNorm -> Linear and SwiGLU -> Linear as single autograd nodes that do NOT keep the normalised /
activated tensor alive for the weight-gradient GEMM; the backward re-creates it with one cheap HBM-bound kernel (the norm
statistics / the gate-up projection are saved anyway).  Per Llama-2-13B layer this drops ~200 MB of saved activations at
seq 4096, which is what lets the single-GPU run keep every layer's activations instead of recomputing whole layers.

Parity (role): the reference's fused_rms_norm + fused_linear / swiglu + fused_linear ops composed under recompute
(python/paddle/incubate/nn/functional/*), here with the recompute folded into the op's own backward."""
from __future__ import annotations

import torch

from . import ext, raw, use_fused, wrap
from . import wgrad as WG


def _2d(t):
    return t.reshape(-1, t.shape[-1])


class _NormLinear(torch.autograd.Function):
    """y = rmsnorm(x [+ residual]) * g @ W ; returns y (and h = x + residual when a residual is given)."""

    @staticmethod
    def forward(ctx, x, residual, g, w, eps, sink=None):
        ctx.sink = sink
        xc = x.contiguous()
        n, rstd, res_out = ext().rms_norm_fwd(xc, residual.contiguous() if residual is not None else None, g, None, eps)
        h = res_out if residual is not None else xc
        out = torch.empty((*x.shape[:-1], w.shape[1]), dtype=x.dtype, device=x.device)
        ext().gemm(_2d(n), w, None, False, False, 0, out.view(-1, w.shape[1]), None)
        ctx.save_for_backward(h, g, rstd, w)     # `n` is dropped here
        ctx.eps, ctx.has_res = eps, residual is not None
        if residual is not None:
            return out, res_out
        return out

    @staticmethod
    def backward(ctx, dy, dres=None):
        h, g, rstd, w = ctx.saved_tensors
        dy2 = _2d(dy.contiguous())
        n, _, _ = ext().rms_norm_fwd(h, None, g, None, ctx.eps)                 # recompute the normalised activations (HBM-bound)
        dw = WG.emit(ctx.sink, _2d(n), dy2)                                    # dW = n^T dy (accumulated in the arena when it can be)
        del n
        dn = ext().gemm(dy2, w, None, False, True, 0, None, None).reshape(h.shape)   # dn = dy W^T
        dx, dg = ext().rms_norm_bwd(dn, h, g, rstd)
        if ctx.has_res:
            if dres is not None:
                dx = dx + dres
            return dx, dx, dg, dw, None, None
        return dx, None, dg, dw, None, None


class _SwigluLinear(torch.autograd.Function):
    """y = swiglu(gu) @ W with gu = [gate | up] packed along the last dim."""

    @staticmethod
    def forward(ctx, gu, w, sink=None):
        ctx.sink = sink
        guc = gu.contiguous()
        act = ext().swiglu_fwd(guc, None)
        out = torch.empty((*gu.shape[:-1], w.shape[1]), dtype=gu.dtype, device=gu.device)
        ext().gemm(_2d(act), w, None, False, False, 0, out.view(-1, w.shape[1]), None)
        ctx.save_for_backward(guc, w)            # `act` is dropped here
        return out

    @staticmethod
    def backward(ctx, dy):
        gu, w = ctx.saved_tensors
        dy2 = _2d(dy.contiguous())
        act = ext().swiglu_fwd(gu, None)
        dw = WG.emit(ctx.sink, _2d(act), dy2)
        del act
        dact = ext().gemm(dy2, w, None, False, True, 0, None, None).reshape(*gu.shape[:-1], w.shape[0])
        dgu, _ = ext().swiglu_bwd(dact, gu, None)
        return dgu, dw, None


def _ok(x, w):
    from ..framework.flags import flag

    if flag("FLAGS_b200_gemm_backend", "tcgen05") != "tcgen05":
        return False
    return use_fused(x) and x.dtype in (torch.bfloat16, torch.float16) and w.dtype == x.dtype and w.dim() == 2 and w.is_contiguous() \
        and x.shape[-1] % 8 == 0 and w.shape[1] % 8 == 0 and (x.numel() // x.shape[-1]) % 8 == 0 and bool(ext().gemm_supported(_2d(x), w, False, False))


def norm_linear(x, norm_weight, weight, eps, residual=None):
    """(rmsnorm(x [+ residual]) * norm_weight) @ weight. With `residual` returns (y, x + residual)."""
    sink = WG.sink_for(weight)
    x, norm_weight, weight, residual = raw(x), raw(norm_weight), raw(weight), raw(residual)
    if _ok(x, weight) and norm_weight.dtype == x.dtype:
        out = _NormLinear.apply(x, residual, norm_weight, weight, float(eps), sink)
        return (wrap(out[0]), wrap(out[1])) if residual is not None else wrap(out)
    from . import gemm as KG
    from . import norm as KN

    if residual is not None:
        n, h = KN.rms_norm(x, norm_weight, eps, residual=residual)
        return KG.linear(n, weight), h
    return KG.linear(KN.rms_norm(x, norm_weight, eps), weight)


def swiglu_linear(gu, weight):
    """swiglu(gu) @ weight with gu = [gate | up]."""
    sink = WG.sink_for(weight)
    gu, weight = raw(gu), raw(weight)
    half = gu.shape[-1] // 2
    from ..framework.flags import flag

    if use_fused(gu) and flag("FLAGS_b200_gemm_backend", "tcgen05") == "tcgen05" and gu.dtype in (torch.bfloat16, torch.float16) and weight.dtype == gu.dtype and weight.dim() == 2 and weight.is_contiguous() \
            and half % 8 == 0 and weight.shape[1] % 8 == 0 and (gu.numel() // gu.shape[-1]) % 8 == 0 and weight.shape[0] == half:
        return wrap(_SwigluLinear.apply(gu, weight, sink))
    from . import activation as KA
    from . import gemm as KG

    return KG.linear(KA.swiglu(gu), weight)
