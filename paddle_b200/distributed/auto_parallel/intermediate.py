"""`parallelize` — the intermediate semi-auto API. Parity: python/paddle/distributed/auto_parallel/intermediate/
{parallelize.py,tensor_parallel.py,pipeline_parallel.py,sharded_data_parallel.py}.

`parallelize(model, optimizer, mesh, config)`:
  * `mp_config.parallelize_plan`: {sublayer-name pattern -> ColWiseParallel / RowWiseParallel / SequenceParallel*} — matching
    Linear / Embedding sublayers are replaced in place by the tensor-parallel fleet layers (built from slices of the existing
    weights), which run the tcgen05 GEMM and, on a symmetric heap, the fused GEMM + peer-memory collective paths.
  * `dp_config.sharding_level` in {0,1,2,3}: gradients are averaged over the "dp" mesh dim; levels >= 1 partition the
    optimizer state (and the gradients / parameters for 2 / 3) through distributed.sharding.
  * `pp_config.split_spec`: {sublayer-name -> SplitPoint.END/BEGINNING}: records stage boundaries; the model is cut into
    per-stage segments executed by the 1F1B engine of fleet.pipeline."""
from __future__ import annotations

import fnmatch
import re
from enum import Enum

import torch

from ...nn.layer import Layer
from .process_mesh import get_mesh


class SplitPoint(Enum):
    BEGINNING = 0
    END = 1


class _Plan:
    pass


class ColWiseParallel(_Plan):
    def __init__(self, gather_output=False):
        self.gather_output = gather_output


class RowWiseParallel(_Plan):
    def __init__(self, is_input_parallel=True):
        self.is_input_parallel = is_input_parallel


class SequenceParallelBegin(_Plan):
    """Output of the marked layer is scattered along the sequence dim."""

    def __init__(self, need_transpose=True):
        self.need_transpose = need_transpose


class SequenceParallelEnd(_Plan):
    """Input of the marked layer is gathered along the sequence dim."""

    def __init__(self, need_transpose=True):
        self.need_transpose = need_transpose


class SequenceParallelEnable(_Plan):
    pass


class SequenceParallelDisable(_Plan):
    def __init__(self, need_transpose=True):
        self.need_transpose = need_transpose


class PrepareLayerInput(_Plan):
    def __init__(self, fn=None):
        self.fn = fn


class PrepareLayerOutput(_Plan):
    def __init__(self, fn=None):
        self.fn = fn


def _match(pattern, name):
    return fnmatch.fnmatchcase(name, pattern) or re.fullmatch(pattern, name) is not None


def _mp_group(mesh):
    return mesh.group_along("mp") if "mp" in mesh.dim_names else None


def _replace(parent, attr, new):
    parent._sub_layers[attr] = new


def _apply_mp_plan(model, mesh, plan):
    from ...nn import Embedding, Linear
    from ..fleet import mp_layers as M

    if "mp" not in mesh.dim_names or mesh.get_dim_size("mp") == 1:
        return
    group = _mp_group(mesh)
    n, r = group.nranks, group.rank
    named = dict(model.named_sublayers(include_self=False))
    parents = {}
    for pname, parent in [("", model)] + list(named.items()):
        for attr, child in parent._sub_layers.items():
            parents[(pname + "." + attr) if pname else attr] = (parent, attr)
    for pattern, spec in plan.items():
        specs = spec if isinstance(spec, (list, tuple)) else [spec]
        wname = None
        if pattern.endswith(".weight") or pattern.endswith(".bias"):
            pattern, wname = pattern.rsplit(".", 1)
        for lname, layer in list(named.items()):
            if not _match(pattern, lname):
                continue
            parent, attr = parents[lname]
            for s in specs:
                if isinstance(s, ColWiseParallel) and isinstance(layer, Linear):
                    new = M.ColumnParallelLinear(layer.weight.shape[0], layer.weight.shape[1], has_bias=layer.bias is not None,
                                                 gather_output=s.gather_output, mp_group=group)
                    with torch.no_grad():
                        torch.Tensor.copy_(new.weight, layer.weight.chunk(n, 1)[r])
                        if layer.bias is not None:
                            torch.Tensor.copy_(new.bias, layer.bias.chunk(n, 0)[r])
                    _replace(parent, attr, new)
                elif isinstance(s, RowWiseParallel) and isinstance(layer, Linear):
                    new = M.RowParallelLinear(layer.weight.shape[0], layer.weight.shape[1], has_bias=layer.bias is not None,
                                              input_is_parallel=s.is_input_parallel, mp_group=group)
                    with torch.no_grad():
                        torch.Tensor.copy_(new.weight, layer.weight.chunk(n, 0)[r])
                        if layer.bias is not None:
                            torch.Tensor.copy_(new.bias, layer.bias)
                    _replace(parent, attr, new)
                elif isinstance(s, (ColWiseParallel, RowWiseParallel)) and isinstance(layer, Embedding):
                    new = M.VocabParallelEmbedding(layer.weight.shape[0], layer.weight.shape[1], mp_group=group)
                    with torch.no_grad():
                        torch.Tensor.copy_(new.weight, layer.weight.chunk(n, 0)[r])
                    _replace(parent, attr, new)
                elif isinstance(s, SequenceParallelBegin):
                    layer.register_forward_post_hook(lambda l, i, o: M.ScatterOp.apply(o))
                elif isinstance(s, (SequenceParallelEnd, SequenceParallelDisable)):
                    layer.register_forward_pre_hook(lambda l, i: tuple(M.AllGatherOp.apply(x) if isinstance(x, torch.Tensor) else x for x in i))
                elif isinstance(s, PrepareLayerInput) and s.fn is not None:
                    layer.register_forward_pre_hook(s.fn(mesh))
                elif isinstance(s, PrepareLayerOutput) and s.fn is not None:
                    layer.register_forward_post_hook(s.fn(mesh))


class _DPModel(Layer):
    """Averages gradients over the dp mesh dim after backward (bucketed, overlapping with the rest of backward)."""

    def __init__(self, inner, group):
        super().__init__()
        from ..data_parallel import DataParallel

        self._layers = DataParallel(inner, group=group) if group is not None and group.nranks > 1 else inner

    def forward(self, *a, **k):
        return self._layers(*a, **k)


def parallelize(model, optimizer=None, mesh=None, config=None):
    mesh = mesh or get_mesh()
    config = config or {}
    assert mesh is not None, "parallelize needs a ProcessMesh (argument or dist.auto_parallel.set_mesh)"
    mp = config.get("mp_config") or {}
    if mp.get("parallelize_plan"):
        _apply_mp_plan(model, mesh, mp["parallelize_plan"])
    pp = config.get("pp_config") or {}
    if pp.get("split_spec"):
        model._pp_split_spec = pp["split_spec"]   # consumed by fleet.pipeline when the mesh has a "pp" dim
    dp = config.get("dp_config") or {}
    level = int(dp.get("sharding_level", 0) or 0)
    if "dp" in mesh.dim_names and mesh.get_dim_size("dp") > 1:
        g = mesh.group_along("dp")
        if level >= 1 and optimizer is not None:
            from ..sharding import group_sharded_parallel

            lvl = {1: "os", 2: "os_g", 3: "p_g_os"}[level]
            model, optimizer, _ = group_sharded_parallel(model, optimizer, lvl, group=g)
        else:
            model = _DPModel(model, g)
    if optimizer is not None and mp.get("parallelize_plan") and len(getattr(optimizer, "_param_groups", ())) == 1:
        # sublayers were replaced: re-point the optimizer at the live parameters
        optimizer._param_groups[0]["params"] = list(model.parameters())
    return (model, optimizer) if optimizer is not None else model
