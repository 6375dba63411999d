"""GradScaler (dynamic loss scaling). Parity: python/paddle/amp/grad_scaler.py
(check_finite_and_unscale + update_loss_scaling kernels in paddle/phi/kernels/gpu/amp_kernel.cu).

Device-side design: scale, found_inf and the good/bad step counters live on the device; ``step`` never reads them back
(the fused optimizer kernels take ``found_inf`` as a device pointer and skip the update themselves).
"""
from __future__ import annotations

from enum import Enum

import torch

from ..tensor import Tensor


class OptimizerState(Enum):
    INIT = 0
    UNSCALED = 1
    STEPPED = 2


class AmpScaler:
    def __init__(self, enable=True, init_loss_scaling=2.0 ** 16, incr_ratio=2.0, decr_ratio=0.5, incr_every_n_steps=2000,
                 decr_every_n_nan_or_inf=1, use_dynamic_loss_scaling=True):
        self._enable = enable
        self._init_loss_scaling = float(init_loss_scaling)
        self._incr_ratio, self._decr_ratio = incr_ratio, decr_ratio
        self._incr_every_n_steps, self._decr_every_n_nan_or_inf = incr_every_n_steps, decr_every_n_nan_or_inf
        self._use_dynamic = use_dynamic_loss_scaling
        self._scale = None
        self._found_inf = None
        self._good = 0
        self._bad = 0
        self._opt_states = {}

    def _lazy(self, device):
        if self._scale is None or self._scale.device != device:
            self._scale = torch.full((1,), self._init_loss_scaling, dtype=torch.float32, device=device)
            self._found_inf = torch.zeros(1, dtype=torch.float32, device=device)

    def scale(self, var):
        if not self._enable:
            return var
        self._lazy(var.device)
        return var * self._scale.to(var.dtype)

    def _grads(self, optimizer):
        return [torch.Tensor.grad.__get__(p) for p in optimizer._parameter_list if torch.Tensor.grad.__get__(p) is not None]

    def unscale_(self, optimizer):
        if not self._enable:
            return
        st = self._opt_states.get(id(optimizer), OptimizerState.INIT)
        if st == OptimizerState.UNSCALED:
            raise RuntimeError("unscale_() has already been called on this optimizer since the last update().")
        grads = self._grads(optimizer)
        if grads:
            self._lazy(grads[0].device)
            self._found_inf.zero_()
            inv = 1.0 / self._scale
            if grads[0].is_cuda:
                from .._build import ext

                sq = torch.zeros(1, dtype=torch.float32, device=grads[0].device)
                for g in grads:
                    gc = g if g.is_contiguous() else g.contiguous()
                    ext().grad_sq_norm(gc, sq, self._found_inf)
                    ext().scale_inplace(g, inv, 1.0) if g.is_contiguous() else g.mul_(inv.to(g.dtype))
            else:
                for g in grads:
                    g.mul_(inv.to(g.dtype))
                    if not torch.isfinite(g).all():
                        self._found_inf.fill_(1.0)
        self._opt_states[id(optimizer)] = OptimizerState.UNSCALED

    def step(self, optimizer):
        if not self._enable:
            return optimizer.step()
        if self._opt_states.get(id(optimizer), OptimizerState.INIT) != OptimizerState.UNSCALED:
            self.unscale_(optimizer)
        # skipping needs one host read in the generic path; fused arena optimizers read found_inf on-device instead
        if self._found_inf is None or float(self._found_inf.item()) == 0.0:
            optimizer.step()
        self._opt_states[id(optimizer)] = OptimizerState.STEPPED

    def update(self):
        if not self._enable or not self._use_dynamic or self._scale is None:
            self._opt_states.clear()
            return
        if float(self._found_inf.item()) != 0.0:
            self._good, self._bad = 0, self._bad + 1
            if self._bad >= self._decr_every_n_nan_or_inf:
                self._scale.mul_(self._decr_ratio).clamp_(min=1.0)
                self._bad = 0
        else:
            self._bad, self._good = 0, self._good + 1
            if self._good >= self._incr_every_n_steps:
                self._scale.mul_(self._incr_ratio)
                self._good = 0
        self._found_inf.zero_()
        self._opt_states.clear()

    def minimize(self, optimizer, *args, **kwargs):
        self.step(optimizer)
        self.update()

    def is_enable(self):
        return self._enable

    def is_use_dynamic_loss_scaling(self):
        return self._use_dynamic

    def get_init_loss_scaling(self):
        return self._init_loss_scaling

    def set_init_loss_scaling(self, new_init_loss_scaling):
        self._init_loss_scaling = float(new_init_loss_scaling)
        if self._scale is not None:
            self._scale.fill_(float(new_init_loss_scaling))

    def get_incr_ratio(self):
        return self._incr_ratio

    def set_incr_ratio(self, new_incr_ratio):
        assert new_incr_ratio > 1.0, "The new_incr_ratio must be > 1.0."
        self._incr_ratio = new_incr_ratio

    def get_decr_ratio(self):
        return self._decr_ratio

    def set_decr_ratio(self, new_decr_ratio):
        assert new_decr_ratio < 1.0, "The new_decr_ratio must be < 1.0."
        self._decr_ratio = new_decr_ratio

    def get_incr_every_n_steps(self):
        return self._incr_every_n_steps

    def set_incr_every_n_steps(self, new_incr_every_n_steps):
        self._incr_every_n_steps = new_incr_every_n_steps

    def get_decr_every_n_nan_or_inf(self):
        return self._decr_every_n_nan_or_inf

    def set_decr_every_n_nan_or_inf(self, new_decr_every_n_nan_or_inf):
        self._decr_every_n_nan_or_inf = new_decr_every_n_nan_or_inf

    def get_loss_scaling(self):
        return None if self._scale is None else self._scale.as_subclass(Tensor)

    def state_dict(self):
        if not self._enable:
            return {}
        return {"scale": (self._scale.cpu().numpy() if self._scale is not None else self._init_loss_scaling),
                "incr_ratio": self._incr_ratio, "decr_ratio": self._decr_ratio, "incr_every_n_steps": self._incr_every_n_steps,
                "decr_every_n_nan_or_inf": self._decr_every_n_nan_or_inf, "incr_count": self._good, "decr_count": self._bad,
                "use_dynamic_loss_scaling": self._use_dynamic}

    def load_state_dict(self, state_dict):
        sd = state_dict
        if not sd:
            return
        import numpy as np

        self._init_loss_scaling = float(np.asarray(sd["scale"]).reshape(-1)[0])
        if self._scale is not None:
            self._scale.fill_(self._init_loss_scaling)
        self._incr_ratio, self._decr_ratio = sd["incr_ratio"], sd["decr_ratio"]
        self._incr_every_n_steps, self._decr_every_n_nan_or_inf = sd["incr_every_n_steps"], sd["decr_every_n_nan_or_inf"]
        self._good, self._bad = sd.get("incr_count", 0), sd.get("decr_count", 0)
        self._use_dynamic = sd.get("use_dynamic_loss_scaling", True)


class GradScaler(AmpScaler):
    def __init__(self, enable=True, init_loss_scaling=2.0 ** 16, incr_ratio=2.0, decr_ratio=0.5, incr_every_n_steps=2000,
                 decr_every_n_nan_or_inf=1, use_dynamic_loss_scaling=True):
        super().__init__(enable, init_loss_scaling, incr_ratio, decr_ratio, incr_every_n_steps, decr_every_n_nan_or_inf, use_dynamic_loss_scaling)
