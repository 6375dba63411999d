"""paddle.callbacks. Parity: python/paddle/hapi/callbacks.py."""
from __future__ import annotations

import numbers
import os
import time

import numpy as np


class Callback:
    def __init__(self):
        self.model = None
        self.params = {}

    def set_params(self, params):
        self.params = params

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None): pass
    def on_train_end(self, logs=None): pass
    def on_eval_begin(self, logs=None): pass
    def on_eval_end(self, logs=None): pass
    def on_predict_begin(self, logs=None): pass
    def on_predict_end(self, logs=None): pass
    def on_epoch_begin(self, epoch, logs=None): pass
    def on_epoch_end(self, epoch, logs=None): pass
    def on_train_batch_begin(self, step, logs=None): pass
    def on_train_batch_end(self, step, logs=None): pass
    def on_eval_batch_begin(self, step, logs=None): pass
    def on_eval_batch_end(self, step, logs=None): pass
    def on_predict_batch_begin(self, step, logs=None): pass
    def on_predict_batch_end(self, step, logs=None): pass


class CallbackList:
    def __init__(self, callbacks=None):
        self.callbacks = list(callbacks or [])

    def set_params(self, params):
        for c in self.callbacks:
            c.set_params(params)

    def set_model(self, model):
        for c in self.callbacks:
            c.set_model(model)

    def _call(self, name, *args):
        for c in self.callbacks:
            getattr(c, name)(*args)

    def __getattr__(self, name):
        if name.startswith("on_"):
            return lambda *a: self._call(name, *a)
        raise AttributeError(name)


class ProgBarLogger(Callback):
    def __init__(self, log_freq=1, verbose=2):
        super().__init__()
        self.log_freq, self.verbose = log_freq, verbose

    def on_train_begin(self, logs=None):
        self.epochs = self.params.get("epochs")

    def on_epoch_begin(self, epoch, logs=None):
        self.epoch, self.t0 = epoch, time.time()
        if self.verbose:
            print(f"Epoch {epoch + 1}/{self.epochs}")

    def _fmt(self, logs):
        out = []
        for k, v in (logs or {}).items():
            if isinstance(v, (list, tuple, np.ndarray)):
                v = np.asarray(v).reshape(-1)
                out.append(f"{k}: " + " ".join(f"{x:.4f}" for x in v))
            elif isinstance(v, numbers.Number):
                out.append(f"{k}: {v:.4f}")
        return " - ".join(out)

    def on_train_batch_end(self, step, logs=None):
        if self.verbose == 2 and (step + 1) % self.log_freq == 0:
            print(f"step {step + 1}/{self.params.get('steps', '?')} - {self._fmt(logs)}")

    def on_epoch_end(self, epoch, logs=None):
        if self.verbose:
            print(f"Epoch {epoch + 1} done in {time.time() - self.t0:.1f}s - {self._fmt(logs)}")

    def on_eval_end(self, logs=None):
        if self.verbose:
            print(f"Eval - {self._fmt(logs)}")


class ModelCheckpoint(Callback):
    def __init__(self, save_freq=1, save_dir=None):
        super().__init__()
        self.save_freq, self.save_dir = save_freq, save_dir

    def on_epoch_end(self, epoch, logs=None):
        if self.save_dir and (epoch + 1) % self.save_freq == 0:
            self.model.save(os.path.join(self.save_dir, str(epoch)))

    def on_train_end(self, logs=None):
        if self.save_dir:
            self.model.save(os.path.join(self.save_dir, "final"))


class LRScheduler(Callback):
    def __init__(self, by_step=True, by_epoch=False):
        super().__init__()
        if by_step and by_epoch:
            raise ValueError("by_step option is mutually exclusive with by_epoch")
        self.by_step, self.by_epoch = by_step, by_epoch

    def _sched(self):
        from .optimizer.lr import LRScheduler as S

        lr = getattr(self.model._optimizer, "_learning_rate", None)
        return lr if isinstance(lr, S) else None

    def on_epoch_end(self, epoch, logs=None):
        if self.by_epoch and self._sched() is not None:
            self._sched().step()

    def on_train_batch_end(self, step, logs=None):
        if self.by_step and self._sched() is not None:
            self._sched().step()


class EarlyStopping(Callback):
    def __init__(self, monitor="loss", mode="auto", patience=0, verbose=1, min_delta=0, baseline=None, save_best_model=True):
        super().__init__()
        self.monitor, self.patience, self.verbose, self.min_delta = monitor, patience, verbose, abs(min_delta)
        self.baseline, self.save_best_model = baseline, save_best_model
        if mode == "auto":
            mode = "max" if "acc" in monitor else "min"
        self.op = np.greater if mode == "max" else np.less
        self.min_delta *= 1 if mode == "max" else -1
        self.wait, self.best, self.stopped_epoch = 0, None, 0

    def on_train_begin(self, logs=None):
        self.wait = 0
        self.best = self.baseline if self.baseline is not None else (-np.inf if self.op == np.greater else np.inf)

    def on_eval_end(self, logs=None):
        if logs is None or self.monitor not in logs:
            return
        cur = logs[self.monitor]
        cur = float(np.asarray(cur).reshape(-1)[0])
        if self.op(cur - self.min_delta, self.best):
            self.best, self.wait = cur, 0
            if self.save_best_model and getattr(self.model, "_save_dir", None):
                self.model.save(os.path.join(self.model._save_dir, "best_model"))
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.model.stop_training = True
                if self.verbose:
                    print(f"Epoch {self.stopped_epoch + 1}: Early stopping.")


class ReduceLROnPlateau(Callback):
    def __init__(self, monitor="loss", factor=0.1, patience=10, verbose=1, mode="auto", min_delta=1e-4, cooldown=0, min_lr=0):
        super().__init__()
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.min_delta, self.cooldown, self.min_lr = min_delta, cooldown, min_lr
        if mode == "auto":
            mode = "max" if "acc" in monitor else "min"
        self.mode = mode
        self.best = -np.inf if mode == "max" else np.inf
        self.wait = self.cooldown_counter = 0

    def on_eval_end(self, logs=None):
        if logs is None or self.monitor not in logs:
            return
        cur = float(np.asarray(logs[self.monitor]).reshape(-1)[0])
        better = cur > self.best + self.min_delta if self.mode == "max" else cur < self.best - self.min_delta
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if better:
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                opt = self.model._optimizer
                old = opt.get_lr()
                new = max(old * self.factor, self.min_lr)
                if old - new > 1e-12:
                    opt.set_lr(new)
                    if self.verbose:
                        print(f"ReduceLROnPlateau reducing learning rate to {new}.")
                self.cooldown_counter, self.wait = self.cooldown, 0


class VisualDL(Callback):
    """Scalar logger (writes a JSONL next to where VisualDL would; the visualdl package is not available offline)."""

    def __init__(self, log_dir):
        super().__init__()
        self.log_dir = log_dir
        self._step = 0

    def on_train_batch_end(self, step, logs=None):
        import json

        os.makedirs(self.log_dir, exist_ok=True)
        self._step += 1
        with open(os.path.join(self.log_dir, "scalars.jsonl"), "a") as f:
            f.write(json.dumps({"step": self._step, **{k: float(np.asarray(v).reshape(-1)[0]) for k, v in (logs or {}).items() if isinstance(v, (numbers.Number, list, np.ndarray))}}) + "\n")


class WandbCallback(Callback):
    def __init__(self, *a, **k):
        super().__init__()
        raise RuntimeError("wandb is not available in this offline environment")
