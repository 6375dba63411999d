"""Distributed (sharded) checkpoint. Parity: python/paddle/distributed/checkpoint/{save_state_dict,load_state_dict,metadata}.py.

Each rank writes its local shards (`{rank}_0.distcp`, same pickle-of-numpy format as paddle.save) and rank 0 writes
`0.metadata` describing, for every tensor key, the global shape and the (rank, offsets, local shape) of each shard.
Loading re-shards: a rank assembles every requested local tensor from whatever shards overlap it, so dp/mp/pp layouts
may differ between save and load."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from ..framework.io import load as _load
from ..framework.io import save as _save
from ..tensor import Tensor
from . import env


def _shard_info(t):
    """(global_shape, offsets) of a local tensor. Tensors may carry `_dist_shard = (global_shape, offsets)`; mp layers with
    `is_distributed` + `split_axis` are described by fleet mp metadata; anything else is treated as replicated."""
    d = getattr(t, "__dict__", {})
    if "_dist_shard" in d:
        return tuple(d["_dist_shard"][0]), tuple(d["_dist_shard"][1])
    return _ls(t), tuple(0 for _ in _ls(t))


def save_state_dict(state_dict, path, process_group=None, coordinator_rank=0, unique_id=None, async_save=False):
    os.makedirs(path, exist_ok=True)
    rank, world = env.get_rank(), env.get_world_size()
    local, meta = {}, {}
    for k, v in state_dict.items():
        if isinstance(v, dict):   # nested (e.g. optimizer master_weights)
            for kk, vv in v.items():
                local[f"{k}.{kk}"] = vv
        else:
            local[k] = v
    tensors = {k: v for k, v in local.items() if isinstance(v, torch.Tensor)}
    for k, t in tensors.items():
        gshape, offs = _shard_info(t)
        replicated = gshape == _ls(t)
        if replicated and rank != coordinator_rank and world > 1 and not getattr(t, "is_distributed", False):
            continue   # one copy of replicated tensors is enough
        meta[k] = {"global_shape": list(gshape), "offsets": list(offs), "local_shape": list(_ls(t)), "rank": rank, "dtype": str(t.dtype)}
    keep = {k: tensors[k] for k in meta}
    _save({k: (v if isinstance(v, Tensor) else v.as_subclass(Tensor)) for k, v in keep.items()}, os.path.join(path, f"{rank}_0.distcp"))
    others = {k: v for k, v in local.items() if not isinstance(v, torch.Tensor)}
    all_meta = [None] * world
    if world > 1:
        import torch.distributed as dist

        dist.all_gather_object(all_meta, meta)
    else:
        all_meta = [meta]
    if rank == coordinator_rank:
        merged = {}
        for m in all_meta:
            for k, e in m.items():
                merged.setdefault(k, []).append(e)
        with open(os.path.join(path, "0.metadata"), "w") as f:
            json.dump({"state_dict_metadata": merged, "non_tensor": {k: v for k, v in others.items() if isinstance(v, (int, float, str, bool, list))}}, f)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()


def load_state_dict(state_dict, path, process_group=None, coordinator_rank=0, unique_id=None, offload=False):
    with open(os.path.join(path, "0.metadata")) as f:
        md = json.load(f)
    meta = md["state_dict_metadata"]
    cache = {}

    def shard_file(r):
        if r not in cache:
            cache[r] = _load(os.path.join(path, f"{r}_0.distcp"), return_numpy=True)
        return cache[r]

    flat_targets = {}
    for k, v in state_dict.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat_targets[f"{k}.{kk}"] = vv
        else:
            flat_targets[k] = v
    for k, t in flat_targets.items():
        if not isinstance(t, torch.Tensor) or k not in meta:
            continue
        gshape, offs = _shard_info(t)
        lo = np.array(offs)
        hi = lo + np.array(_ls(t))
        out = np.zeros(_ls(t), dtype=np.float64)
        filled = False
        entries = meta[k]
        # several ranks saved a block with identical placement (e.g. a tensor-parallel parameter written without shard metadata by an
        # older checkpoint, or replicated copies): take this rank's own copy when there is one, never silently "last writer wins"
        same = {}
        for e in entries:
            same.setdefault((tuple(e["offsets"]), tuple(e["local_shape"])), []).append(e)
        dedup = []
        for group in same.values():
            if len(group) > 1:
                mine = [e for e in group if e["rank"] == env.get_rank()]
                if not mine and getattr(t, "is_distributed", False):
                    raise ValueError(f"distributed checkpoint: {len(group)} ranks saved '{k}' with identical offsets but the tensor is "
                                     "model-parallel; the checkpoint carries no shard metadata for it (re-save with this version)")
                dedup.append(mine[0] if mine else group[0])
            else:
                dedup.append(group[0])
        for e in dedup:
            so, ss = np.array(e["offsets"]), np.array(e["local_shape"])
            a, b = np.maximum(lo, so), np.minimum(hi, so + ss)
            if (a >= b).any() and out.ndim > 0:
                continue
            src = shard_file(e["rank"])[k]
            if src.dtype == np.uint16:
                src = torch.from_numpy(src.view(np.int16).copy()).view(torch.bfloat16).float().numpy()
            sl_src = tuple(slice(int(x - y), int(z - y)) for x, z, y in zip(a, b, so))
            sl_dst = tuple(slice(int(x - y), int(z - y)) for x, z, y in zip(a, b, lo))
            out[sl_dst] = src[sl_src]
            filled = True
        if filled:
            with torch.no_grad(), torch._C.DisableTorchFunctionSubclass():   # raw local copy (no DistTensor propagation)
                torch.Tensor.copy_(t,torch.from_numpy(out).to(t.dtype).reshape(_ls(t)))
    for k, v in md.get("non_tensor", {}).items():
        if k in state_dict and not isinstance(state_dict[k], torch.Tensor):
            state_dict[k] = v
    return state_dict


def _ls(t):
    """Local (stored) shape — DistTensor.shape is the global one."""
    return tuple(int(s) for s in torch.Tensor.size(t))
