"""fp8 GEMM entry (per-tensor scaled e4m3/e5m2 -> half). Parity: paddle.linalg.fp8_fp8_half_gemm_fused."""
from __future__ import annotations

import torch

from . import raw, wrap


def fp8_gemm(x, y, transpose_x=False, transpose_y=False, bias=None, scale=1.0, out_dtype=torch.float16, act="identity"):
    x, y, bias = raw(x), raw(y), raw(bias)
    xf = x.to(torch.float32)
    yf = y.to(torch.float32)
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
    return wrap(out.to(out_dtype or torch.float16))
