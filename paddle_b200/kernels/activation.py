"""SwiGLU. Parity: python/paddle/incubate/nn/functional/swiglu.py."""
from __future__ import annotations

import torch

from ..framework.recording import recordable
import torch.nn.functional as F

from . import ext, raw, use_fused, wrap


class _SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up):
        g = gate.contiguous()
        u = up.contiguous() if up is not None else None
        ctx.save_for_backward(g, u)
        return ext().swiglu_fwd(g, u)

    @staticmethod
    def backward(ctx, dout):
        g, u = ctx.saved_tensors
        dg, du = ext().swiglu_bwd(dout.contiguous(), g, u)
        return dg, (du if u is not None else None)


@recordable
def swiglu(x, y=None):
    x, y = raw(x), raw(y)
    cols = x.shape[-1] if y is not None else x.shape[-1] // 2
    if use_fused(x) and x.dtype in (torch.float32, torch.float16, torch.bfloat16) and cols % (16 // x.element_size()) == 0:
        return wrap(_SwiGLU.apply(x, y))
    if y is None:
        x, y = x.chunk(2, -1)
    return wrap(F.silu(x) * y)
