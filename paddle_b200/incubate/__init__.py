"""paddle.incubate. Parity: python/paddle/incubate/__init__.py."""
from __future__ import annotations

import torch

from ..tensor import Tensor
from . import nn  # noqa: F401
from . import moe, optimizer  # noqa: F401
from .optimizer import LookAhead, ModelAverage  # noqa: F401


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def softmax_mask_fuse(x, mask, name=None):
    return _w(torch.softmax(_raw(x) + _raw(mask), -1))


def softmax_mask_fuse_upper_triangle(x):
    xr = _raw(x)
    s = xr.shape[-1]
    m = torch.ones(xr.shape[-2], s, dtype=torch.bool, device=xr.device).tril()
    return _w(torch.softmax(xr.masked_fill(~m, float("-inf")), -1))


def identity_loss(x, reduction="none"):
    r = {0: "sum", 1: "mean", 2: "none"}.get(reduction, reduction)
    return x.sum() if r == "sum" else (x.mean() if r == "mean" else x)


from ..geometric import segment_max, segment_mean, segment_min, segment_sum  # noqa: E402,F401
from ..geometric import reindex_graph as graph_reindex  # noqa: E402,F401
from ..geometric import sample_neighbors as graph_sample_neighbors  # noqa: E402,F401
from ..geometric import send_u_recv as graph_send_recv  # noqa: E402,F401


def graph_khop_sampler(row, colptr, input_nodes, sample_sizes, sorted_eids=None, return_eids=False, name=None):
    from ..geometric import reindex_graph, sample_neighbors

    nodes = _raw(input_nodes)
    all_src, all_dst = [], []
    frontier = nodes
    for k in sample_sizes:
        nb, cnt = sample_neighbors(row, colptr, frontier, k)
        all_src.append(_raw(nb))
        all_dst.append(torch.repeat_interleave(frontier, _raw(cnt).long()))
        frontier = torch.unique(_raw(nb))
    src, dst = torch.cat(all_src), torch.cat(all_dst)
    uniq, inv = torch.unique(torch.cat([nodes, src, dst]), return_inverse=True)
    n0 = nodes.numel()
    return _w(inv[n0:n0 + src.numel()]), _w(inv[n0 + src.numel():]), _w(uniq), _w(inv[:n0])


class _ASP:
    """2:4 structured sparsity helpers. Parity: python/paddle/incubate/asp/."""

    _excluded = set()
    _masks = {}

    @staticmethod
    def calculate_density(x):
        t = _raw(x) if isinstance(x, torch.Tensor) else torch.as_tensor(x)
        return float((t != 0).float().mean())

    @classmethod
    def set_excluded_layers(cls, param_names, main_program=None):
        cls._excluded.update(param_names)

    @classmethod
    def reset_excluded_layers(cls, main_program=None):
        cls._excluded.clear()

    @classmethod
    def prune_model(cls, model, n=2, m=4, mask_algo="mask_1d", with_mask=True):
        from ..nn import Conv2D, Linear

        for name, layer in model.named_sublayers(include_self=True):
            extra = tuple(t for t, _ in cls._supported if isinstance(t, type))
            names = {t for t, _ in cls._supported if isinstance(t, str)}
            if (isinstance(layer, (Linear, Conv2D) + extra) or type(layer).__name__ in names) and getattr(layer, "weight", None) is not None \
                    and layer.weight.name not in cls._excluded:
                w = _raw(layer.weight)
                flat = w.detach().reshape(-1, m) if w.numel() % m == 0 else None
                if flat is None:
                    continue
                idx = flat.abs().topk(n, -1).indices
                mask = torch.zeros_like(flat).scatter_(1, idx, 1.0).reshape(w.shape)
                with torch.no_grad():
                    w.mul_(mask)
                cls._masks[layer.weight.name] = mask
        return cls._masks

    _supported = []

    @classmethod
    def add_supported_layer(cls, layer, pruning_func=None):
        """Register an extra layer type (class or name) whose `weight` takes part in prune_model."""
        cls._supported.append((layer, pruning_func))

    @classmethod
    def decorate(cls, optimizer):
        step = optimizer.step

        def masked_step():
            step()
            with torch.no_grad():
                for p in optimizer._parameter_list:
                    m = cls._masks.get(p.name)
                    if m is not None:
                        _raw(p).mul_(m.to(p.device))

        optimizer.step = masked_step
        return optimizer


asp = _ASP


class _Autotune:
    """paddle.incubate.autotune. Parity: python/paddle/incubate/autotune.py:set_config (kernel / layout / dataloader tuning switches).
    Kernel choice here is static (hand-written sm_100a kernels, shapes dispatched in C++), so `kernel` and `layout` are recorded
    only; `dataloader.enable` turns on the DataLoader's worker-count probe."""

    config = {"kernel": {"enable": False, "tuning_range": [1, 10]}, "layout": {"enable": False}, "dataloader": {"enable": False}}

    @classmethod
    def set_config(cls, config=None):
        import json

        if config is None:
            for v in cls.config.values():
                v["enable"] = True
            return
        if isinstance(config, str):
            with open(config) as f:
                config = json.load(f)
        for k, v in config.items():
            if k not in cls.config:
                raise ValueError(f"autotune.set_config: unknown section {k!r}")
            cls.config[k].update(v)


autotune = _Autotune


def __getattr__(name):
    import importlib

    if name in ("autograd", "distributed", "multiprocessing", "checkpoint", "tensor", "layers", "operators", "jit", "framework", "passes", "xpu"):
        try:
            return importlib.import_module("." + name, __name__)
        except ModuleNotFoundError as e:
            raise AttributeError(name) from e
    if name == "inference":   # paddle.incubate.inference: the predictor API lives in paddle_b200.inference
        return importlib.import_module("paddle_b200.inference")
    raise AttributeError(name)


# static programs record these as single ops (their bodies compute on raw tensors; framework/recording.py)
from ..framework.recording import make_recordable as _make_recordable  # noqa: E402

_make_recordable(globals(), ["softmax_mask_fuse", "softmax_mask_fuse_upper_triangle", "identity_loss"])
