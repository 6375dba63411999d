"""Gradient clipping. Parity: python/paddle/nn/clip.py."""
from __future__ import annotations

import torch

from ..tensor import Tensor


def _is_rows(g):
    return type(g).__name__ == "SelectedRows"


def _sq(g):
    """sum of squares of a gradient; SelectedRows: of its MERGED rows (python/paddle/nn/clip.py merge_selected_rows + get_tensor_from_selected_rows)."""
    return g.squared_l2_norm() if _is_rows(g) else g.as_subclass(torch.Tensor).float().pow(2).sum()


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
                if _is_rows(g):
                    m = g.merge()                      # clip the summed rows, as the dense gradient would be
                    out.append((p, type(g)(m.rows, torch.clamp(m.value, self.min, self.max), m.height)))
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
            if _is_rows(g):
                g = g.merge()
                out.append((p, g.scale(float(self.clip_norm / max(float(torch.sqrt(_sq(g))), self.clip_norm)))))
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
            s = _sq(g)
            sq = s if sq is None else sq + s.to(sq.device)
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
                out.append((p, g.scale(coef.to(g.value.dtype).to(g.value.device)) if _is_rows(g) else (g * coef.to(g.dtype))))
        return out


GradientClipByValue = ClipGradByValue
GradientClipByNorm = ClipGradByNorm
GradientClipByGlobalNorm = ClipGradByGlobalNorm
