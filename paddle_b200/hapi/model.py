"""hapi.Model. Parity: python/paddle/hapi/model.py:Model (dygraph adapter)."""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import callbacks as cbks
from ..framework.io import load as _load
from ..framework.io import save as _save
from ..io import DataLoader, Dataset
from ..metric import Metric
from ..tensor import Tensor, to_tensor


def _to_list(x):
    if x is None:
        return []
    return list(x) if isinstance(x, (list, tuple)) else [x]


class Model:
    def __init__(self, network, inputs=None, labels=None):
        self.network = network
        self._inputs, self._labels = _to_list(inputs), _to_list(labels)
        self._optimizer = self._loss = None
        self._metrics = []
        self._scaler = None
        self._amp_level = "O0"
        self.stop_training = False
        self._save_dir = None
        self._nranks, self._rank, self._ddp = 1, 0, None

    def _init_distributed(self):
        """Launched with several trainers (paddle.distributed.launch / spawn): train through DataParallel, shard the data with a
        DistributedBatchSampler, gather evaluation outputs. Parity: hapi/model.py (DynamicGraphAdapter with ParallelEnv().nranks > 1)."""
        import os as _os

        from ..distributed import env as _env

        world = int(_os.environ.get("PADDLE_TRAINERS_NUM", _os.environ.get("WORLD_SIZE", "1")))
        if world <= 1 and not (_env.is_initialized() and _env.get_world_size() > 1):
            return
        if not _env.is_initialized():
            _env.init_parallel_env()
        self._nranks, self._rank = _env.get_world_size(), _env.get_rank()
        if self._nranks > 1 and self._ddp is None:
            from ..distributed.data_parallel import DataParallel

            self._ddp = DataParallel(self.network)

    def prepare(self, optimizer=None, loss=None, metrics=None, amp_configs=None):
        self._init_distributed()
        self._optimizer, self._loss = optimizer, loss
        self._metrics = _to_list(metrics)
        for m in self._metrics:
            if not isinstance(m, Metric):
                raise TypeError(f"{type(m).__name__} is not a sub class of Metric")
        if amp_configs:
            from .. import amp

            cfg = {"level": amp_configs} if isinstance(amp_configs, str) else dict(amp_configs)
            self._amp_level = cfg.pop("level", "O1")
            self._amp_dtype = cfg.pop("dtype", "float16")
            if self._amp_level != "O0":
                self._scaler = amp.GradScaler(**{k: v for k, v in cfg.items() if k in ("init_loss_scaling", "incr_ratio", "decr_ratio", "incr_every_n_steps", "decr_every_n_nan_or_inf", "use_dynamic_loss_scaling")})

    def _tensors(self, data):
        return [d if isinstance(d, torch.Tensor) else to_tensor(np.asarray(d)) for d in _to_list(data)]

    def _split(self, batch):
        batch = _to_list(batch)
        n_in = len(self._inputs) or (len(batch) - (len(self._labels) or (1 if self._loss is not None and len(batch) > 1 else 0)))
        return batch[:n_in], batch[n_in:]

    def _forward(self, inputs):
        net = self._ddp if (self._ddp is not None and self.network.training) else self.network
        if self._amp_level != "O0":
            from .. import amp

            with amp.auto_cast(level=self._amp_level, dtype=self._amp_dtype):
                return net(*inputs)
        return net(*inputs)

    def _gather(self, tensors):
        """Evaluation under several trainers: metrics see every rank's outputs (equal batch shapes per rank)."""
        if self._nranks <= 1:
            return tensors
        import torch.distributed as dist

        out = []
        for t in tensors:
            raw = t.as_subclass(torch.Tensor).contiguous()
            parts = [torch.empty_like(raw) for _ in range(self._nranks)]
            dist.all_gather(parts, raw)
            out.append(torch.cat(parts, 0).as_subclass(type(t)))
        return out

    def train_batch(self, inputs, labels=None, update=True):
        self.network.train()
        inputs, labels = self._tensors(inputs), self._tensors(labels)
        outs = _to_list(self._forward(inputs))
        losses = _to_list(self._loss(*(outs + labels))) if self._loss is not None else outs
        total = losses[0]
        for l in losses[1:]:
            total = total + l
        if self._scaler is not None:
            self._scaler.scale(total).backward()
            if update:
                self._scaler.step(self._optimizer)
                self._scaler.update()
                self._optimizer.clear_grad()
        else:
            total.backward()
            if update:
                self._optimizer.step()
                self._optimizer.clear_grad()
        metrics = []
        for m in self._metrics:
            res = m.update(*_to_list(m.compute(*(outs + labels))))
            metrics.append(res)
        lv = [float(l.item()) for l in losses]
        return (lv, metrics) if metrics else lv

    @torch.no_grad()
    def eval_batch(self, inputs, labels=None):
        self.network.eval()
        inputs, labels = self._tensors(inputs), self._tensors(labels)
        outs = _to_list(self._forward(inputs))
        losses = _to_list(self._loss(*(outs + labels))) if self._loss is not None else []
        metrics = []
        if self._metrics and self._nranks > 1:
            outs, labels = self._gather(outs), self._gather(labels)
        for m in self._metrics:
            metrics.append(m.update(*_to_list(m.compute(*(outs + labels)))))
        lv = [float(l.item()) for l in losses]
        return (lv, metrics) if metrics else lv

    @torch.no_grad()
    def predict_batch(self, inputs):
        self.network.eval()
        return [o.numpy() for o in _to_list(self._forward(self._tensors(inputs)))]

    def _loader(self, data, batch_size, shuffle, num_workers, drop_last=False):
        if data is None or isinstance(data, DataLoader):
            return data
        if isinstance(data, Dataset):
            if self._nranks > 1:
                from ..io import DistributedBatchSampler

                bs = DistributedBatchSampler(data, batch_size=batch_size, num_replicas=self._nranks, rank=self._rank, shuffle=shuffle, drop_last=drop_last)
                return DataLoader(data, batch_sampler=bs, num_workers=num_workers)
            return DataLoader(data, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, drop_last=drop_last)
        return data

    def fit(self, train_data=None, eval_data=None, batch_size=1, epochs=1, eval_freq=1, log_freq=10, save_dir=None, save_freq=1, verbose=2,
            drop_last=False, shuffle=True, num_workers=0, callbacks=None, accumulate_grad_batches=1, num_iters=None):
        train_loader = self._loader(train_data, batch_size, shuffle, num_workers, drop_last)
        eval_loader = self._loader(eval_data, batch_size, False, num_workers)
        self._save_dir = save_dir
        cb = [cbks.ProgBarLogger(log_freq, verbose)] + ([cbks.ModelCheckpoint(save_freq, save_dir)] if save_dir else []) + _to_list(callbacks)
        if not any(isinstance(c, cbks.LRScheduler) for c in cb):
            cb.append(cbks.LRScheduler())
        cl = cbks.CallbackList(cb)
        cl.set_model(self)
        try:
            steps = len(train_loader)
        except Exception:
            steps = None
        cl.set_params({"epochs": epochs, "steps": steps, "verbose": verbose, "metrics": self._metric_names()})
        cl.on_train_begin({})
        self.stop_training = False
        it = 0
        for epoch in range(epochs):
            cl.on_epoch_begin(epoch, {})
            for m in self._metrics:
                m.reset()
            logs = {}
            for step, batch in enumerate(train_loader):
                cl.on_train_batch_begin(step, {})
                ins, labs = self._split(batch)
                update = (step + 1) % accumulate_grad_batches == 0
                res = self.train_batch(ins, labs, update)
                logs = self._logs(res)
                logs["step"] = step
                cl.on_train_batch_end(step, logs)
                it += 1
                if num_iters is not None and it >= num_iters:
                    self.stop_training = True
                    break
            cl.on_epoch_end(epoch, logs)
            if eval_loader is not None and (epoch + 1) % eval_freq == 0:
                self.evaluate(eval_loader, verbose=0, callbacks=cb, _internal=True)
            if self.stop_training:
                break
        cl.on_train_end({})

    def _metric_names(self):
        names = ["loss"]
        for m in self._metrics:
            names += _to_list(m.name())
        return names

    def _logs(self, res):
        logs = {}
        if isinstance(res, tuple):
            losses, metrics = res
        else:
            losses, metrics = res, []
        if losses:
            logs["loss"] = losses[0] if len(losses) == 1 else losses
        for m, r in zip(self._metrics, metrics):
            for n, v in zip(_to_list(m.name()), _to_list(r)):
                logs[n] = v
        return logs

    def evaluate(self, eval_data, batch_size=1, log_freq=10, verbose=2, num_workers=0, callbacks=None, num_iters=None, _internal=False):
        loader = self._loader(eval_data, batch_size, False, num_workers)
        cl = cbks.CallbackList(_to_list(callbacks) if _internal else [cbks.ProgBarLogger(log_freq, verbose)] + _to_list(callbacks))
        cl.set_model(self)
        cl.on_eval_begin({})
        for m in self._metrics:
            m.reset()
        losses = []
        for step, batch in enumerate(loader):
            ins, labs = self._split(batch)
            res = self.eval_batch(ins, labs)
            lv = res[0] if isinstance(res, tuple) else res
            if lv:
                losses.append(lv[0])
            if num_iters is not None and step + 1 >= num_iters:
                break
        logs = {}
        if losses:
            logs["loss"] = [float(np.mean(losses))]
        for m in self._metrics:
            for n, v in zip(_to_list(m.name()), _to_list(m.accumulate())):
                logs[n] = v
        cl.on_eval_end(logs)
        return logs

    def predict(self, test_data, batch_size=1, num_workers=0, stack_outputs=False, verbose=1, callbacks=None):
        loader = self._loader(test_data, batch_size, False, num_workers)
        outs = []
        for batch in loader:
            ins, _ = self._split(batch) if (self._inputs or self._loss) else (_to_list(batch), [])
            outs.append(self.predict_batch(ins))
        res = [list(o) for o in zip(*outs)]
        if stack_outputs:
            res = [np.vstack(o) for o in res]
        return res

    def save(self, path, training=True):
        if self._rank != 0:       # one writer per job
            return
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        if training:
            _save(self.network.state_dict(), path + ".pdparams")
            if self._optimizer is not None:
                _save(self._optimizer.state_dict(), path + ".pdopt")
        else:
            from .. import jit

            jit.save(self.network, path, input_spec=self._inputs or None)

    def load(self, path, skip_mismatch=False, reset_optimizer=False):
        p = path if path.endswith(".pdparams") else path + ".pdparams"
        sd = _load(p)
        if skip_mismatch:
            own = self.network.state_dict()
            sd = {k: v for k, v in sd.items() if k in own and list(v.shape) == list(own[k].shape)}
        self.network.set_state_dict(sd)
        op = p[: -len(".pdparams")] + ".pdopt"
        if not reset_optimizer and self._optimizer is not None and os.path.exists(op):
            self._optimizer.set_state_dict(_load(op))

    def parameters(self, *a, **k):
        return self.network.parameters(*a, **k)

    def summary(self, input_size=None, dtype=None):
        from .summary import summary

        return summary(self.network, input_size or [tuple(i.shape) for i in self._inputs], dtype)
