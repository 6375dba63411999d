"""paddle.incubate.optimizer. Parity: python/paddle/incubate/optimizer/{lookahead,modelaverage,distributed_fused_lamb,gradient_merge}.py."""
from __future__ import annotations

import torch

from ..optimizer.optimizer import Lamb, Optimizer


class LookAhead:
    def __init__(self, inner_optimizer, alpha=0.5, k=5, name=None):
        self.inner_optimizer, self.alpha, self.k = inner_optimizer, alpha, k
        self._step, self._slow = 0, {}

    @property
    def _parameter_list(self):
        return self.inner_optimizer._parameter_list

    @torch.no_grad()
    def step(self):
        self.inner_optimizer.step()
        self._step += 1
        if self._step % self.k == 0:
            for p in self._parameter_list:
                pr = p.as_subclass(torch.Tensor)
                s = self._slow.get(p.name)
                if s is None:
                    s = self._slow[p.name] = pr.detach().clone()
                s.add_(pr - s, alpha=self.alpha)
                pr.copy_(s)

    def clear_grad(self, set_to_zero=True):
        self.inner_optimizer.clear_grad(set_to_zero)

    def minimize(self, loss, **k):
        loss.backward()
        self.step()

    def state_dict(self):
        sd = self.inner_optimizer.state_dict()
        sd["@lookahead_step@"] = self._step
        return sd

    def set_state_dict(self, sd):
        self._step = int(sd.pop("@lookahead_step@", 0)) if isinstance(sd, dict) else 0
        self.inner_optimizer.set_state_dict(sd)

    def __getattr__(self, n):
        return getattr(self.inner_optimizer, n)


class ModelAverage:
    """Running average of parameters for evaluation. Parity: incubate/optimizer/modelaverage.py."""

    def __init__(self, average_window_rate, parameters=None, min_average_window=10000, max_average_window=10000, name=None):
        self._params = list(parameters)
        self._rate, self._min, self._max = average_window_rate, min_average_window, max_average_window
        self._sum = {p.name: torch.zeros_like(p.as_subclass(torch.Tensor), dtype=torch.float32) for p in self._params}
        self._n, self._backup = 0, {}

    @torch.no_grad()
    def step(self):
        self._n += 1
        for p in self._params:
            self._sum[p.name].add_(p.as_subclass(torch.Tensor).float())
        if self._n > self._max:
            for p in self._params:
                self._sum[p.name].mul_(0.5)
            self._n = self._n // 2

    def minimize(self, loss, **k):
        self.step()

    import contextlib as _c

    @_c.contextmanager
    def apply(self, executor=None, need_restore=True):
        with torch.no_grad():
            for p in self._params:
                pr = p.as_subclass(torch.Tensor)
                self._backup[p.name] = pr.clone()
                if self._n > 0:
                    pr.copy_((self._sum[p.name] / self._n).to(pr.dtype))
        try:
            yield
        finally:
            if need_restore:
                self.restore()

    @torch.no_grad()
    def restore(self, executor=None):
        for p in self._params:
            if p.name in self._backup:
                p.as_subclass(torch.Tensor).copy_(self._backup[p.name])


class DistributedFusedLamb(Lamb):
    """LAMB whose moments are sharded over the data-parallel group and updated by the fused two-stage kernel
    (csrc/optim.cu:lamb_stage1/2). Parity: incubate/optimizer/distributed_fused_lamb.py."""

    def __init__(self, learning_rate=0.001, lamb_weight_decay=0.01, beta1=0.9, beta2=0.999, epsilon=1e-6, parameters=None, grad_clip=None,
                 exclude_from_weight_decay_fn=None, clip_after_allreduce=True, is_grad_scaled_by_nranks=True, alignment=128, use_master_param_norm=True,
                 gradient_accumulation_steps=1, use_master_acc_grad=True, nproc_per_node=None, use_hierarchical_allreduce=False, name=None):
        super().__init__(learning_rate, lamb_weight_decay, beta1, beta2, epsilon, parameters, grad_clip, exclude_from_weight_decay_fn, multi_precision=True)
        self._acc_steps, self._acc = gradient_accumulation_steps, 0

    def step(self):
        import torch.distributed as dist

        self._acc += 1
        if self._acc % self._acc_steps:
            return
        if dist.is_initialized() and dist.get_world_size() > 1:
            n = dist.get_world_size()
            for p in self._parameter_list:
                g = torch.Tensor.grad.__get__(p)
                if g is not None:
                    dist.all_reduce(g)
                    g.mul_(1.0 / n)
        super().step()


class GradientMergeOptimizer:
    """Accumulate k steps of gradients before one update. Parity: fleet gradient_merge meta optimizer."""

    def __init__(self, inner_optimizer, k_steps=1, avg=True):
        self.inner_optimizer, self.k_steps, self.avg, self._i = inner_optimizer, k_steps, avg, 0

    def step(self):
        self._i += 1
        if self._i % self.k_steps == 0:
            if self.avg and self.k_steps > 1:
                with torch.no_grad():
                    for p in self.inner_optimizer._parameter_list:
                        g = torch.Tensor.grad.__get__(p)
                        if g is not None:
                            g.mul_(1.0 / self.k_steps)
            self.inner_optimizer.step()
            self.inner_optimizer.clear_grad()

    def clear_grad(self, set_to_zero=True):
        pass  # gradients are intentionally kept across the merged steps

    def __getattr__(self, n):
        return getattr(self.inner_optimizer, n)


class RecomputeOptimizer:
    def __init__(self, optimizer):
        self._opt = optimizer

    def _set_checkpoints(self, checkpoints):
        self._checkpoints = checkpoints

    def __getattr__(self, n):
        return getattr(self._opt, n)


from ..optimizer import LBFGS  # noqa: F401,E402  (paddle.incubate.optimizer.LBFGS is the same class)
