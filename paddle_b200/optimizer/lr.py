"""Learning-rate schedulers. Parity: python/paddle/optimizer/lr.py (18 schedulers)."""
from __future__ import annotations

import math

__all__ = ["LRScheduler", "NoamDecay", "PiecewiseDecay", "NaturalExpDecay", "InverseTimeDecay", "PolynomialDecay", "LinearWarmup",
           "ExponentialDecay", "MultiStepDecay", "StepDecay", "LambdaDecay", "ReduceOnPlateau", "CosineAnnealingDecay",
           "MultiplicativeDecay", "OneCycleLR", "CyclicLR", "LinearLR", "CosineAnnealingWarmRestarts"]


class LRScheduler:
    def __init__(self, learning_rate=0.1, last_epoch=-1, verbose=False):
        self.base_lr = float(learning_rate)
        self.last_lr = float(learning_rate)
        self.last_epoch = last_epoch
        self.verbose = verbose
        self.step()

    def __call__(self):
        return self.last_lr

    def step(self, epoch=None):
        if epoch is None:
            self.last_epoch += 1
        else:
            self.last_epoch = epoch
        self.last_lr = self.get_lr()
        if self.verbose:
            print(f"Epoch {self.last_epoch}: {type(self).__name__} set learning rate to {self.last_lr}.")

    def get_lr(self):
        raise NotImplementedError

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if isinstance(v, (int, float, bool, str, list, tuple)) or v is None}

    def set_state_dict(self, state_dict):
        for k, v in state_dict.items():
            if k in self.__dict__:
                self.__dict__[k] = v

    set_dict = set_state_dict
    state_keys = state_dict


class NoamDecay(LRScheduler):
    def __init__(self, d_model, warmup_steps, learning_rate=1.0, last_epoch=-1, verbose=False):
        self.d_model, self.warmup_steps = d_model, warmup_steps
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        a = 1.0 if self.last_epoch == 0 else self.last_epoch ** -0.5
        b = self.warmup_steps ** -1.5 * self.last_epoch
        return self.base_lr * (self.d_model ** -0.5) * min(a, b)


class PiecewiseDecay(LRScheduler):
    def __init__(self, boundaries, values, last_epoch=-1, verbose=False):
        self.boundaries, self.values = list(boundaries), list(values)
        super().__init__(values[0], last_epoch, verbose)

    def get_lr(self):
        for i, b in enumerate(self.boundaries):
            if self.last_epoch < b:
                return self.values[i]
        return self.values[len(self.values) - 1]


class NaturalExpDecay(LRScheduler):
    def __init__(self, learning_rate, gamma, last_epoch=-1, verbose=False):
        self.gamma = gamma
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        return self.base_lr * math.exp(-self.gamma * self.last_epoch)


class InverseTimeDecay(LRScheduler):
    def __init__(self, learning_rate, gamma, last_epoch=-1, verbose=False):
        self.gamma = gamma
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        return self.base_lr / (1 + self.gamma * self.last_epoch)


class PolynomialDecay(LRScheduler):
    def __init__(self, learning_rate, decay_steps, end_lr=0.0001, power=1.0, cycle=False, last_epoch=-1, verbose=False):
        self.decay_steps, self.end_lr, self.power, self.cycle = decay_steps, end_lr, power, cycle
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        t, steps = self.last_epoch, self.decay_steps
        if self.cycle:
            div = math.ceil(t / float(steps)) if t > 0 else 1
            steps = steps * div
        else:
            t = min(t, steps)
        return (self.base_lr - self.end_lr) * ((1 - float(t) / float(steps)) ** self.power) + self.end_lr


class LinearWarmup(LRScheduler):
    def __init__(self, learning_rate, warmup_steps, start_lr, end_lr, last_epoch=-1, verbose=False):
        self.learning_rate, self.warmup_steps, self.start_lr, self.end_lr = learning_rate, warmup_steps, start_lr, end_lr
        super().__init__(start_lr, last_epoch, verbose)

    def get_lr(self):
        if self.last_epoch < self.warmup_steps:
            return (self.end_lr - self.start_lr) * float(self.last_epoch) / float(self.warmup_steps) + self.start_lr
        if isinstance(self.learning_rate, LRScheduler):
            self.learning_rate.step(self.last_epoch - self.warmup_steps)
            return self.learning_rate()
        return self.learning_rate

    def state_dict(self):
        d = super().state_dict()
        if isinstance(self.learning_rate, LRScheduler):
            d["LinearWarmup_LR"] = self.learning_rate.state_dict()
        return d

    def set_state_dict(self, state_dict):
        inner = state_dict.pop("LinearWarmup_LR", None) if isinstance(state_dict, dict) else None
        super().set_state_dict(state_dict)
        if inner is not None and isinstance(self.learning_rate, LRScheduler):
            self.learning_rate.set_state_dict(inner)


class ExponentialDecay(LRScheduler):
    def __init__(self, learning_rate, gamma, last_epoch=-1, verbose=False):
        self.gamma = gamma
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        return self.base_lr * (self.gamma ** self.last_epoch)


class MultiStepDecay(LRScheduler):
    def __init__(self, learning_rate, milestones, gamma=0.1, last_epoch=-1, verbose=False):
        self.milestones, self.gamma = list(milestones), gamma
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        n = sum(1 for m in self.milestones if self.last_epoch >= m)
        return self.base_lr * (self.gamma ** n)


class StepDecay(LRScheduler):
    def __init__(self, learning_rate, step_size, gamma=0.1, last_epoch=-1, verbose=False):
        self.step_size, self.gamma = step_size, gamma
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        return self.base_lr * (self.gamma ** (self.last_epoch // self.step_size))


class LambdaDecay(LRScheduler):
    def __init__(self, learning_rate, lr_lambda, last_epoch=-1, verbose=False):
        self.lr_lambda = lr_lambda
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        return self.base_lr * self.lr_lambda(self.last_epoch)


class MultiplicativeDecay(LRScheduler):
    def __init__(self, learning_rate, lr_lambda, last_epoch=-1, verbose=False):
        self.lr_lambda = lr_lambda
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        lr = self.base_lr
        for e in range(1, self.last_epoch + 1):
            lr *= self.lr_lambda(e)
        return lr


class CosineAnnealingDecay(LRScheduler):
    def __init__(self, learning_rate, T_max, eta_min=0, last_epoch=-1, verbose=False):
        self.T_max, self.eta_min = T_max, float(eta_min)
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * self.last_epoch / self.T_max)) / 2


class CosineAnnealingWarmRestarts(LRScheduler):
    def __init__(self, learning_rate, T_0, T_mult=1, eta_min=0, last_epoch=-1, verbose=False):
        self.T_0, self.T_mult, self.eta_min = T_0, T_mult, float(eta_min)
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        e = self.last_epoch
        if self.T_mult == 1:
            t_cur, t_i = e % self.T_0, self.T_0
        else:
            n = int(math.log(e / self.T_0 * (self.T_mult - 1) + 1, self.T_mult)) if e > 0 else 0
            t_cur = e - self.T_0 * (self.T_mult ** n - 1) / (self.T_mult - 1)
            t_i = self.T_0 * self.T_mult ** n
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2


class LinearLR(LRScheduler):
    def __init__(self, learning_rate, total_steps, start_factor=1.0 / 3, end_factor=1.0, last_epoch=-1, verbose=False):
        self.total_steps, self.start_factor, self.end_factor = total_steps, start_factor, end_factor
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        t = min(self.last_epoch, self.total_steps)
        return self.base_lr * (self.start_factor + (self.end_factor - self.start_factor) * t / self.total_steps)


class OneCycleLR(LRScheduler):
    def __init__(self, max_learning_rate, total_steps, divide_factor=25.0, end_learning_rate=0.0001, phase_pct=0.3,
                 anneal_strategy="cos", three_phase=False, last_epoch=-1, verbose=False):
        self.max_lr, self.total_steps = max_learning_rate, total_steps
        self.initial_lr = max_learning_rate / divide_factor
        self.min_lr = end_learning_rate
        self.anneal = anneal_strategy
        if three_phase:
            self._ends = [phase_pct * total_steps - 1, 2 * phase_pct * total_steps - 2, total_steps - 1]
            self._vals = [(self.initial_lr, self.max_lr), (self.max_lr, self.initial_lr), (self.initial_lr, self.min_lr)]
        else:
            self._ends = [phase_pct * total_steps - 1, total_steps - 1]
            self._vals = [(self.initial_lr, self.max_lr), (self.max_lr, self.min_lr)]
        super().__init__(self.initial_lr, last_epoch, verbose)

    def _interp(self, a, b, pct):
        if self.anneal == "cos":
            return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1)
        return (b - a) * pct + a

    def get_lr(self):
        step = self.last_epoch
        start = 0.0
        for i, end in enumerate(self._ends):
            if step <= end or i == len(self._ends) - 1:
                pct = (step - start) / max(end - start, 1e-12)
                a, b = self._vals[i]
                return self._interp(a, b, min(max(pct, 0.0), 1.0))
            start = end
        return self.min_lr


class CyclicLR(LRScheduler):
    def __init__(self, base_learning_rate, max_learning_rate, step_size_up, step_size_down=None, mode="triangular", exp_gamma=1.0,
                 scale_fn=None, scale_mode="cycle", last_epoch=-1, verbose=False):
        self.max_lr = max_learning_rate
        self.up = float(step_size_up)
        self.down = float(step_size_down) if step_size_down is not None else self.up
        self.total = self.up + self.down
        self.ratio = self.up / self.total
        self.mode, self.gamma = mode, exp_gamma
        if scale_fn is None:
            if mode == "triangular":
                self.scale_fn, self.scale_mode = (lambda x: 1.0), "cycle"
            elif mode == "triangular2":
                self.scale_fn, self.scale_mode = (lambda x: 1 / (2.0 ** (x - 1))), "cycle"
            else:
                self.scale_fn, self.scale_mode = (lambda x: self.gamma ** x), "iterations"
        else:
            self.scale_fn, self.scale_mode = scale_fn, scale_mode
        super().__init__(base_learning_rate, last_epoch, verbose)

    def get_lr(self):
        it = self.last_epoch
        cycle = math.floor(1 + it / self.total)
        x = 1.0 + it / self.total - cycle
        pct = x / self.ratio if x <= self.ratio else (x - 1) / (self.ratio - 1)
        amp = (self.max_lr - self.base_lr) * pct
        return self.base_lr + amp * self.scale_fn(cycle if self.scale_mode == "cycle" else it)

    def state_dict(self):
        d = super().state_dict()
        d.pop("scale_fn", None)
        return d


class ReduceOnPlateau(LRScheduler):
    def __init__(self, learning_rate, mode="min", factor=0.1, patience=10, threshold=1e-4, threshold_mode="rel", cooldown=0,
                 min_lr=0, epsilon=1e-8, verbose=False):
        self.mode, self.factor, self.patience = mode, factor, patience
        self.threshold, self.threshold_mode, self.cooldown = threshold, threshold_mode, cooldown
        self.min_lr, self.epsilon = min_lr, epsilon
        self.cooldown_counter, self.best, self.num_bad_epochs = 0, None, 0
        self.base_lr = float(learning_rate)
        self.last_lr = float(learning_rate)
        self.last_epoch = 0
        self.verbose = verbose

    def get_lr(self):
        return self.last_lr

    def _better(self, cur, best):
        if self.mode == "min":
            return cur < (best - best * self.threshold if self.threshold_mode == "rel" else best - self.threshold)
        return cur > (best + best * self.threshold if self.threshold_mode == "rel" else best + self.threshold)

    def step(self, metrics=None, epoch=None):
        if metrics is None:
            return
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        cur = float(metrics.item() if hasattr(metrics, "item") else metrics)
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
        if self.best is None or self._better(cur, self.best):
            self.best, self.num_bad_epochs = cur, 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            self.cooldown_counter, self.num_bad_epochs = self.cooldown, 0
            new_lr = max(self.last_lr * self.factor, self.min_lr)
            if self.last_lr - new_lr > self.epsilon:
                self.last_lr = new_lr
