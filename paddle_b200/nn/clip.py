"""Gradient clipping. Parity: python/paddle/nn/clip.py."""
from __future__ import annotations

import torch

from ..tensor import Tensor


class ClipGradBase:
    def __call__(self, params_grads):
        return self._dygraph_clip(params_grads)


class ClipGradByValue(ClipGradBase):
    def __init__(self, max, min=None):  # noqa: A002
        self.max = float(max)
        self.min = -self.max if min is None else float(min)

    def _dygraph_clip(self, params_grads):
        out = []
        for p, g in params_grads:
            if g is None or not getattr(p, "need_clip", True):
                out.append((p, g))
            else:
                out.append((p, torch.clamp(g, self.min, self.max)))
        return out


class ClipGradByNorm(ClipGradBase):
    def __init__(self, clip_norm):
        self.clip_norm = float(clip_norm)

    def _dygraph_clip(self, params_grads):
        out = []
        for p, g in params_grads:
            if g is None or not getattr(p, "need_clip", True):
                out.append((p, g))
                continue
            n = torch.linalg.vector_norm(g.float())
            out.append((p, (g.float() * (self.clip_norm / torch.clamp(n, min=self.clip_norm))).to(g.dtype)))
        return out


class ClipGradByGlobalNorm(ClipGradBase):
    def __init__(self, clip_norm, group_name="default_group", auto_skip_clip=False):
        self.clip_norm = float(clip_norm)
        self.group_name, self.auto_skip_clip = group_name, auto_skip_clip

    def global_norm_sq(self, params_grads):
        sq = None
        for p, g in params_grads:
            if g is None or not getattr(p, "need_clip", True):
                continue
            s = g.as_subclass(torch.Tensor).float().pow(2).sum()
            sq = s if sq is None else sq + s
        return sq

    def _dygraph_clip(self, params_grads):
        sq = self.global_norm_sq(params_grads)
        if sq is None:
            return params_grads
        gn = torch.sqrt(sq)
        coef = self.clip_norm / torch.clamp(gn, min=self.clip_norm)
        out = []
        for p, g in params_grads:
            if g is None or not getattr(p, "need_clip", True):
                out.append((p, g))
            else:
                out.append((p, (g * coef.to(g.dtype))))
        return out


GradientClipByValue = ClipGradByValue
GradientClipByNorm = ClipGradByNorm
GradientClipByGlobalNorm = ClipGradByGlobalNorm
