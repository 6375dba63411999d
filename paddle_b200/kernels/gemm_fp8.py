"""fp8 GEMM (per-tensor scaled e4m3 / e5m2 -> half) and an fp8 Linear for O2-fp8 training.
Parity: paddle.linalg.fp8_fp8_half_gemm_fused (python/paddle/tensor/linalg.py) -> phi fp8_gemm fusion kernels (cuBLASLt).

CUDA path: csrc/gemm_fp8_sm100.cu — tcgen05 `kind::f8f6f4` MMAs, fp32 accumulation in TMEM, dequantisation scale + bias +
activation fused in the epilogue.  Both operands must be K-major (x [M,K], y [N,K], i.e. transpose_x=False, transpose_y=True);
other layouts are brought into that form with one transposed copy.  CPU / unsupported shapes: fp32 reference."""
from __future__ import annotations

import torch

from ..framework.recording import recordable
from . import ext, raw, use_fused, wrap

E4M3_MAX, E5M2_MAX = 448.0, 57344.0
_FP8 = (torch.float8_e4m3fn, torch.float8_e5m2)


def _ref(x, y, transpose_x, transpose_y, bias, scale, out_dtype, act):
    xf, yf = x.to(torch.float32), y.to(torch.float32)
    if transpose_x:
        xf = xf.transpose(-1, -2)
    if transpose_y:
        yf = yf.transpose(-1, -2)
    out = torch.matmul(xf, yf) * scale
    if bias is not None:
        out = out + bias.float()
    if act == "gelu":
        out = torch.nn.functional.gelu(out)
    elif act == "relu":
        out = torch.relu(out)
    return out.to(out_dtype or torch.float16)


@recordable
def fp8_gemm(x, y, transpose_x=False, transpose_y=False, bias=None, scale=1.0, out_dtype=torch.float16, act="identity"):
    x, y, bias = raw(x), raw(y), raw(bias)
    out_dtype = out_dtype or torch.float16
    if use_fused(x) and x.dtype in _FP8 and y.dtype in _FP8 and x.dim() == 2 and y.dim() == 2 and out_dtype in (torch.float16, torch.bfloat16, torch.float32):
        a = x.t().contiguous() if transpose_x else x.contiguous()            # [M, K]
        b = y.contiguous() if transpose_y else y.t().contiguous()            # [N, K]
        if a.shape[1] % 16 == 0 and a.shape[1] == b.shape[1] and b.shape[0] % 8 == 0:
            bb = bias.to(out_dtype).contiguous() if bias is not None else None
            return wrap(ext().gemm_fp8(a, b, bb, float(scale), {"identity": 0, None: 0, "gelu": 1, "relu": 2}[act], out_dtype))
    return wrap(_ref(x, y, transpose_x, transpose_y, bias, scale, out_dtype, act))


def quantize_fp8(t, dtype=torch.float8_e4m3fn, amax=None):
    """Per-tensor scaling: returns (t_fp8, inv_scale) with t ~= t_fp8 * inv_scale."""
    t = raw(t)
    fmax = E4M3_MAX if dtype == torch.float8_e4m3fn else E5M2_MAX
    amax = t.detach().abs().amax().float().clamp_min(1e-12) if amax is None else amax
    scale = fmax / amax
    q = (t.float() * scale).clamp(-fmax, fmax).to(dtype)
    return q, (1.0 / scale)


class _Fp8Linear(torch.autograd.Function):
    """y = x @ W (W: [in, out]) with e4m3 activations / weights in the forward and e5m2 output gradients in the backward
    (the fp8 recipe of the reference's O2-fp8 AMP): three tcgen05 fp8 GEMMs, all TN."""

    @staticmethod
    def forward(ctx, x, w, bias):
        x2 = x.reshape(-1, x.shape[-1])
        xq, sx = quantize_fp8(x2)
        wq, sw = quantize_fp8(w.t())                     # [out, in]: K-major B operand
        y = raw(fp8_gemm(xq, wq, False, True, bias, float(sx * sw), x.dtype))
        ctx.save_for_backward(xq, wq)
        ctx.scales = (sx, sw)
        ctx.has_bias = bias is not None
        ctx.xshape = x.shape
        return y.reshape(*x.shape[:-1], w.shape[1])

    @staticmethod
    def backward(ctx, dy):
        xq, wq = ctx.saved_tensors
        sx, sw = ctx.scales
        dy2 = dy.reshape(-1, dy.shape[-1])
        gq, sg = quantize_fp8(dy2, torch.float8_e5m2)
        # dx[M,in] = dy[M,out] @ W^T : B operand [N=in, K=out] = W (row-major [in,out]) -> wq^T copy
        dx = raw(fp8_gemm(gq, wq.t().contiguous(), False, True, None, float(sg * sw), dy.dtype)).reshape(ctx.xshape)
        # dW[in,out] = x^T[in,M] @ dy[M,out] : A = x^T [in, M] (K = tokens), B = dy^T [out, M]
        dw = raw(fp8_gemm(xq.t().contiguous(), gq.t().contiguous(), False, True, None, float(sx * sg), dy.dtype))
        db = dy2.sum(0) if ctx.has_bias else None
        return dx, dw, db


def fp8_linear(x, weight, bias=None):
    x, weight, bias = raw(x), raw(weight), raw(bias)
    if x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 16 == 0 and weight.shape[1] % 16 == 0 \
            and (x.numel() // x.shape[-1]) % 16 == 0:
        return wrap(_Fp8Linear.apply(x, weight, bias))
    y = torch.matmul(x, weight)
    return wrap(y if bias is None else y + bias)
