"""auto_parallel high-level API: Engine (fit / evaluate / predict / save / load), LocalLayer, to_distributed, enable_auto_dp.
Parity: python/paddle/distributed/auto_parallel/static/engine.py:Engine, auto_parallel/local_layer.py:LocalLayer,
auto_parallel/high_level_api.py:to_distributed, auto_parallel/api.py:enable_auto_dp.

The reference's Engine builds, partitions and reshards a static program; here the same user contract runs the semi-auto eager
path: inputs are sharded along the batch over a 1-D mesh of all ranks, parameters stay replicated unless the model annotated them
with shard_tensor / shard_layer, gradients of replicated parameters are averaged over the mesh, and one step is
`DistModel.__call__` (forward, loss, backward, update)."""
from __future__ import annotations

import os

import numpy as np
import torch

from ...nn.layer import Layer
from ...tensor import Tensor
from .. import env
from . import api as _api
from .placement import Replicate, Shard
from .process_mesh import ProcessMesh

_auto_dp = [False]


def enable_auto_dp():
    """Batch-shard the inputs of `shard_dataloader` / Engine over all ranks without an explicit mesh annotation."""
    _auto_dp[0] = True


def in_auto_dp_mode():
    return _auto_dp[0]


def _world_mesh():
    return ProcessMesh(list(range(env.get_world_size())), dim_names=["dp"])


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else t


class LocalLayer(Layer):
    """A layer whose forward is written for *local* tensors: DistTensor inputs are handed over as their local shards and the outputs
    are re-wrapped with the given (mesh, placements). Used for ops that have no sharding rule (custom losses, masked reductions)."""

    def __init__(self, out_dist_attrs=None, grad_dist_attrs=None):
        super().__init__()
        self.out_dist_attrs, self.grad_dist_attrs = out_dist_attrs or [], grad_dist_attrs or []

    def __call__(self, *inputs, **kwargs):
        ins = [_api.dtensor_to_local(i) if _api._is_dt(i) else i for i in inputs]
        kw = {k: (_api.dtensor_to_local(v) if _api._is_dt(v) else v) for k, v in kwargs.items()}
        outs = super().__call__(*ins, **kw)
        single = not isinstance(outs, (list, tuple))
        outs = [outs] if single else list(outs)
        for i, (mesh, placements) in enumerate(self.out_dist_attrs[:len(outs)]):
            outs[i] = _api.dtensor_from_local(outs[i], mesh, placements)
        return outs[0] if single else outs


def to_distributed(model, optimizer, dataloader, device_num=None, node_num=1, config=None):
    """One call parallelisation: data parallel over all devices; `config` may carry an `mp_degree` + `parallelize` plan.
    Returns (model, optimizer, dataloader) ready for the usual eager loop."""
    world = env.get_world_size()
    if world <= 1:
        return model, optimizer, dataloader
    mesh = _world_mesh()
    plan = (config or {}).get("parallelize_plan") if isinstance(config, dict) else None
    if plan:
        from .intermediate import parallelize

        model, optimizer = parallelize(model, optimizer, mesh=mesh, config=plan)
    dataloader = _api.shard_dataloader(dataloader, mesh, shard_dims="dp")
    optimizer = optimizer if isinstance(optimizer, _api._ShardOptimizer) else _api.shard_optimizer(optimizer)
    _sync_replicated_grads(model, mesh)
    return model, optimizer, dataloader


def _sync_replicated_grads(model, mesh):
    """Replicated (plain) parameters of a model fed with batch-sharded data: average their gradients over the mesh."""
    import torch.distributed as dist

    if env.get_world_size() <= 1 or getattr(model, "_auto_dp_hooked", False):
        return
    n = env.get_world_size()

    def hook(p):
        g = torch.Tensor.grad.__get__(p)
        if g is not None and not _api._is_dt(p):
            dist.all_reduce(g)
            g.div_(n)

    for p in model.parameters():
        if not p.stop_gradient and not _api._is_dt(p):
            p.register_post_accumulate_grad_hook(hook)
    model.__dict__["_auto_dp_hooked"] = True


class Engine:
    def __init__(self, model=None, loss=None, optimizer=None, metrics=None, cluster=None, strategy=None):
        self._model, self._loss, self._optimizer = model, loss, optimizer
        self._metrics = list(metrics) if isinstance(metrics, (list, tuple)) else ([metrics] if metrics is not None else [])
        self._strategy = strategy or _api.Strategy()
        self._mesh = _world_mesh()
        self._dist = None
        self.history = {"loss": []}

    # -- plumbing -----------------------------------------------------------------------------------------------------------------
    def _dist_model(self):
        if self._dist is None:
            self._dist = _api.to_static(self._model, None, self._loss, self._optimizer, self._strategy)
            _sync_replicated_grads(self._model, self._mesh)
        return self._dist

    def _loader(self, data, batch_size, shuffle=False, collate_fn=None, num_workers=0):
        from ...io import DataLoader, Dataset, DistributedBatchSampler

        if data is None or isinstance(data, DataLoader) or not isinstance(data, Dataset):
            return data
        world, rank = env.get_world_size(), env.get_rank()
        if world > 1:   # global batch -> every rank reads its own slice
            bs = DistributedBatchSampler(data, batch_size=max(1, batch_size // world), num_replicas=world, rank=rank, shuffle=shuffle)
            return DataLoader(data, batch_sampler=bs, collate_fn=collate_fn, num_workers=num_workers)
        return DataLoader(data, batch_size=batch_size, shuffle=shuffle, collate_fn=collate_fn, num_workers=num_workers)

    @staticmethod
    def _split(batch, n_labels=1):
        batch = list(batch) if isinstance(batch, (list, tuple)) else [batch]
        return batch[:-n_labels] if len(batch) > n_labels else batch, batch[-n_labels:] if len(batch) > n_labels else []

    def prepare(self, inputs_spec=None, labels_spec=None, inputs=None, labels=None, main_program=None, startup_program=None, mode=None, init_parameters=True):
        self._dist_model()
        return self

    # -- loops --------------------------------------------------------------------------------------------------------------------
    def fit(self, train_data, train_sample_split=None, batch_size=1, epochs=1, steps_per_epoch=None, log_freq=10, save_dir=None, save_freq=1,
            valid_data=None, valid_sample_split=None, valid_freq=1, valid_steps=None, collate_fn=None, callbacks=None, verbose=0, num_workers=0):
        dm = self._dist_model()
        loader = self._loader(train_data, batch_size, shuffle=False, collate_fn=collate_fn, num_workers=num_workers)
        logs = {"loss": []}
        for epoch in range(epochs):
            dm.train()
            for step, batch in enumerate(loader):
                if steps_per_epoch is not None and step >= steps_per_epoch:
                    break
                xs, ys = self._split(batch)
                loss = dm(*xs, *ys)
                logs["loss"].append(float(_raw(loss)))
                if verbose and step % log_freq == 0 and env.get_rank() == 0:
                    print(f"[auto.Engine] epoch {epoch} step {step} loss {logs['loss'][-1]:.5f}")
            if valid_data is not None and (epoch + 1) % valid_freq == 0:
                logs.setdefault("eval", []).append(self.evaluate(valid_data, batch_size=batch_size, steps=valid_steps, collate_fn=collate_fn))
            if save_dir and (epoch + 1) % save_freq == 0:
                self.save(os.path.join(save_dir, f"epoch{epoch}"))
        self.history = logs
        return logs

    def evaluate(self, valid_data, valid_sample_split=None, batch_size=1, steps=None, log_freq=10, collate_fn=None, callbacks=None, verbose=0, num_workers=0):
        dm = self._dist_model()
        dm.eval()
        for m in self._metrics:
            m.reset()
        losses = []
        for step, batch in enumerate(self._loader(valid_data, batch_size, collate_fn=collate_fn, num_workers=num_workers)):
            if steps is not None and step >= steps:
                break
            xs, ys = self._split(batch)
            with torch.no_grad():
                out = self._model(*xs)
                if self._loss is not None and ys:
                    losses.append(float(_raw(self._loss(out, *ys))))
            for m in self._metrics:
                m.update(*[np.asarray(_raw(t).cpu()) for t in (m.compute(out, *ys) if hasattr(m, "compute") else (out, *ys))])
        res = {"loss": float(np.mean(losses)) if losses else None}
        for m in self._metrics:
            names = m.name() if isinstance(m.name(), (list, tuple)) else [m.name()]
            vals = m.accumulate()
            vals = vals if isinstance(vals, (list, tuple)) else [vals]
            res.update(dict(zip(names, vals)))
        return res

    def predict(self, test_data, test_sample_split=None, batch_size=1, steps=None, collate_fn=None, callbacks=None, verbose=0, num_workers=0):
        dm = self._dist_model()
        dm.predict()
        outs = []
        for step, batch in enumerate(self._loader(test_data, batch_size, collate_fn=collate_fn, num_workers=num_workers)):
            if steps is not None and step >= steps:
                break
            xs = list(batch) if isinstance(batch, (list, tuple)) else [batch]
            n_in = self._model.forward.__code__.co_argcount - 1 if hasattr(self._model.forward, "__code__") else len(xs)
            outs.append(dm(*xs[:max(1, n_in)]))
        return outs

    def dataloader(self, dataset, batch_size=1, shuffle=False, collate_fn=None, num_workers=0, **kw):
        return self._loader(dataset, batch_size, shuffle, collate_fn, num_workers)

    def run(self, data=None, feed=None, fetch_list=None, mode=None):
        xs, ys = self._split(data)
        return self._dist_model()(*xs, *ys)

    def cost(self, inputs_spec=None, labels_spec=None, mode=None):
        """(estimated step time ms, parameter + state bytes per rank) from the analytic cost model."""
        from ...cost_model import CostModel

        n = sum(int(p.numel()) for p in self._model.parameters())
        cm = CostModel()
        return cm.mem_ms(n * 16), n * 16

    # -- checkpoints ---------------------------------------------------------------------------------------------------------------
    def save(self, path, training=True):
        from ...framework.io import save

        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        if env.get_rank() == 0:
            save(self._model.state_dict(), path + ".pdparams")
            if training and self._optimizer is not None:
                save(self._optimizer.state_dict(), path + ".pdopt")

    def load(self, path, strict=True, load_optimizer=True):
        from ...framework.io import load

        self._model.set_state_dict(load(path + ".pdparams"))
        if load_optimizer and self._optimizer is not None and os.path.exists(path + ".pdopt"):
            self._optimizer.set_state_dict(load(path + ".pdopt"))

    @property
    def main_program(self):
        return None

    @property
    def startup_program(self):
        return None
