"""Weight-only / llm.int8 quantised linear. Parity: python/paddle/nn/quant/quantized_linear.py
(weight_quantize, weight_dequantize, weight_only_linear, llm_int8_linear)."""
from __future__ import annotations

import torch

from ..tensor import Tensor


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def weight_quantize(x, algo="weight_only_int8", arch=None, group_size=-1):
    """x: [in, out] -> (int8 weight [out, in] (int4 packed two per byte), scale [out] or [groups, out])."""
    w = x.as_subclass(torch.Tensor).float()
    bits = 4 if "int4" in algo else 8
    qmax = 2 ** (bits - 1) - 1
    if group_size and group_size > 0:
        k, n = w.shape
        wg = w.reshape(k // group_size, group_size, n)
        scale = wg.abs().amax(1).clamp(min=1e-8) / qmax
        q = torch.round(wg / scale.unsqueeze(1)).clamp(-qmax - 1, qmax).reshape(k, n)
    else:
        scale = w.abs().amax(0).clamp(min=1e-8) / qmax
        q = torch.round(w / scale).clamp(-qmax - 1, qmax)
    q = q.t().contiguous().to(torch.int8)
    if bits == 4:
        lo, hi = q[:, 0::2] & 0xF, q[:, 1::2] & 0xF
        q = (lo | (hi << 4)).to(torch.int8)
    return _w(q), _w(scale.to(x.dtype if x.dtype != torch.float32 else torch.float32))


def weight_dequantize(x, scale, algo="weight_only_int8", out_dtype="float16", group_size=-1):
    from ..framework.dtype import convert_dtype

    q = x.as_subclass(torch.Tensor)
    if "int4" in algo:
        lo = (q << 4).to(torch.int8) >> 4
        hi = q >> 4
        q = torch.stack([lo, hi], -1).reshape(q.shape[0], -1)
    w = q.float().t()
    s = scale.as_subclass(torch.Tensor).float()
    if group_size and group_size > 0:
        w = (w.reshape(-1, group_size, w.shape[1]) * s.unsqueeze(1)).reshape(w.shape)
    else:
        w = w * s
    return _w(w.to(convert_dtype(out_dtype)))


def weight_only_linear(x, weight, bias=None, weight_scale=None, weight_dtype="int8", arch=None, group_size=-1):
    """y = x @ dequant(weight)^T + bias.  CUDA bf16 / fp16 activations with per-channel scales run csrc/gemm_wo_sm100.cu: the int8 / int4
    weights are streamed by TMA and expanded to 16-bit INSIDE the SM (no dequantised copy of the weight is ever written to HBM), which is
    what makes weight-only decoding faster than the 16-bit GEMM (it is bound by weight bytes).  Grouped scales / CPU: dequantise + linear."""
    xr = x.as_subclass(torch.Tensor)
    wr = weight.as_subclass(torch.Tensor)
    if xr.is_cuda and xr.dtype in (torch.bfloat16, torch.float16) and weight_scale is not None and (group_size is None or group_size <= 0) \
            and weight_dtype in ("int8", "int4") and wr.dtype == torch.int8 and xr.shape[-1] % 64 == 0 and wr.shape[0] % 8 == 0:
        from ..framework.flags import flag

        if flag("FLAGS_use_fused_kernels", True):
            from .._build import ext

            x2 = xr.reshape(-1, xr.shape[-1]).contiguous()
            if x2.shape[0] > 64:
                # prefill: compute bound, and every token tile would dequantise the weight tile again - expand the weight once instead
                # (one pass over the int weights) and run the bf16 tcgen05 GEMM; the fused kernel is the decode / small-batch path
                w = weight_dequantize(weight, weight_scale, "weight_only_" + weight_dtype, x.dtype, group_size)
                from . import functional as F

                return F.linear(x, w, bias)
            sc = weight_scale.as_subclass(torch.Tensor).float().contiguous()
            b = None if bias is None else bias.as_subclass(torch.Tensor).to(xr.dtype).contiguous()
            y = ext().weight_only_gemm(x2, wr.contiguous(), sc, b, weight_dtype == "int4")
            return _w(y.reshape(*xr.shape[:-1], wr.shape[0]))
    w = weight_dequantize(weight, weight_scale, "weight_only_" + weight_dtype, x.dtype, group_size)
    from . import functional as F

    return F.linear(x, w, bias)


def llm_int8_linear(x, weight, bias=None, weight_scale=None, threshold=6.0):
    xr = x.as_subclass(torch.Tensor)
    w = weight_dequantize(weight, weight_scale, "weight_only_int8", x.dtype).as_subclass(torch.Tensor)
    outlier = (xr.abs() > threshold).any(dim=tuple(range(xr.dim() - 1)))
    x_in = xr.clone()
    x_in[..., outlier] = 0
    sx = x_in.abs().amax(-1, keepdim=True).clamp(min=1e-8) / 127.0
    xq = torch.round(x_in / sx).clamp(-128, 127)
    out = (xq.float() @ torch.round(w.float() / weight_scale.as_subclass(torch.Tensor).float()).float()) * sx.float() * weight_scale.as_subclass(torch.Tensor).float()
    out = out + xr[..., outlier].float() @ w[outlier].float()
    if bias is not None:
        out = out + bias.as_subclass(torch.Tensor).float()
    return _w(out.to(x.dtype))


def apply_per_channel_scale(x, scales):
    return x * scales


class Stub:
    """Placeholder marking where an observer / quanter is inserted by QAT / PTQ. Parity: nn/quant/stub.py:Stub."""

    def __init__(self, observer=None):
        self._observer = observer

    def __call__(self, x):
        return x

    forward = __call__


# static programs record these as single ops (their bodies compute on raw tensors; framework/recording.py)
from ..framework.recording import make_recordable as _make_recordable  # noqa: E402

_make_recordable(globals(), ['weight_quantize', 'weight_dequantize', 'llm_int8_linear', 'apply_per_channel_scale'])
