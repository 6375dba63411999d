"""paddle.save / paddle.load — ``.pdparams`` / ``.pdopt`` compatible.

Parity: python/paddle/framework/io.py (save:773, load:1020, _build_saved_state_dict:163, _pickle_save:413).
Format: pickle (protocol 2..4) of a dict whose tensor values are numpy arrays (bf16 stored as uint16), plus the
``StructuredToParameterName@@`` name table for state dicts; bare tensors pickle as ``(name, ndarray)`` tuples.
Arrays above 2**30-1 elements... are sliced into ``key@@.N`` pieces recorded under ``UnpackBigParamInfor@@``.
"""
from __future__ import annotations

import copyreg
import io as _io
import os
import pickle
import threading
from collections import OrderedDict

import numpy as np
import torch

from ..tensor import Parameter, Tensor, _np_to_torch

_MAX_NUMEL = 2 ** 30 - 1  # reference: big params are split so a single pickle opcode stays < 4 GB (protocol 2/3)
_async_tasks = []


def _to_numpy(t: torch.Tensor):
    if isinstance(t, Tensor):
        return t.numpy()
    return t.as_subclass(Tensor).numpy()


def _is_state_dict(obj):
    if not isinstance(obj, dict):
        return False
    for v in obj.values():
        if isinstance(v, dict):
            for vv in v.values():
                if not isinstance(vv, (torch.Tensor, np.ndarray, int, float, str, bool, type(None), list, tuple, dict)):
                    return False
        elif not isinstance(v, (torch.Tensor, np.ndarray, int, float, str, bool, type(None), list, tuple)):
            return False
    return any(isinstance(v, torch.Tensor) for v in obj.values()) or any(
        isinstance(v, dict) and any(isinstance(vv, torch.Tensor) for vv in v.values()) for v in obj.values())


def _build_saved_state_dict(state_dict):
    save_dict, name_table = {}, {}
    for k, v in state_dict.items():
        if isinstance(v, torch.Tensor):
            save_dict[k] = _to_numpy(v)
            name_table[k] = v.name if isinstance(v, Tensor) else k
        elif isinstance(v, dict):
            save_dict[k] = {kk: (_to_numpy(vv) if isinstance(vv, torch.Tensor) else vv) for kk, vv in v.items()}
        else:
            save_dict[k] = v
    save_dict["StructuredToParameterName@@"] = name_table
    return save_dict


def _unpack_big(save_dict, protocol):
    if protocol >= 4:
        return save_dict
    info = {}
    out = {}
    for k, v in save_dict.items():
        if isinstance(v, np.ndarray) and v.size > _MAX_NUMEL:
            flat = v.reshape(-1)
            parts = []
            for i in range(0, flat.size, _MAX_NUMEL):
                name = f"{k}@@.{len(parts)}"
                out[name] = flat[i:i + _MAX_NUMEL]
                parts.append(name)
            info[k] = {"OriginShape": v.shape, "slices": parts}
        else:
            out[k] = v
    if info:
        out["UnpackBigParamInfor@@"] = info
    return out


def _pack_big(load_dict):
    info = load_dict.pop("UnpackBigParamInfor@@", None) if isinstance(load_dict, dict) else None
    if info:
        for k, meta in info.items():
            parts = [load_dict.pop(n) for n in meta["slices"]]
            load_dict[k] = np.concatenate(parts).reshape(meta["OriginShape"])
    return load_dict


def _reduce_tensor(t):
    return (tuple, ((t.name if isinstance(t, Tensor) else "tensor", _to_numpy(t)),))


def _pickle_dump(obj, f, protocol):
    p = pickle.Pickler(f, protocol)
    p.dispatch_table = copyreg.dispatch_table.copy()
    p.dispatch_table[Tensor] = _reduce_tensor
    p.dispatch_table[Parameter] = _reduce_tensor
    p.dispatch_table[torch.Tensor] = _reduce_tensor
    p.dump(obj)


def save(obj, path, protocol=4, **configs):
    """paddle.save(obj, path): state dicts, nested containers of tensors, or arbitrary picklable objects."""
    if not isinstance(protocol, int) or protocol < 2 or protocol > 4:
        raise ValueError(f"Expected 1<'protocol'<5, but received protocol={protocol}")
    from ..nn.layer import Layer

    if isinstance(obj, Layer):
        raise ValueError("paddle do not support saving `paddle.nn.Layer` object; save layer.state_dict() instead")
    if _is_state_dict(obj):
        saved = _unpack_big(_build_saved_state_dict(obj), protocol)
    else:
        saved = obj
    if isinstance(path, (_io.BytesIO,)) or hasattr(path, "write"):
        _pickle_dump(saved, path, protocol)
        return
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        _pickle_dump(saved, f, protocol)
    os.replace(tmp, path)  # atomic publish: a crash never leaves a half-written checkpoint


def async_save(obj, path, protocol=4, sync_other_task=False, **configs):
    """Stage tensors to host (pinned when possible) then write on a background thread. Parity: io.py:async_save."""
    if sync_other_task:
        clear_async_save_task_queue()

    def to_cpu(o):
        if isinstance(o, torch.Tensor):
            t = o.detach().to("cpu", non_blocking=False)
            out = t.as_subclass(Tensor)
            if isinstance(o, Tensor):
                out.name = o.name
            return out
        if isinstance(o, dict):
            return type(o)((k, to_cpu(v)) for k, v in o.items())
        if isinstance(o, (list, tuple)):
            return type(o)(to_cpu(v) for v in o)
        return o

    staged = to_cpu(obj)
    t = threading.Thread(target=save, args=(staged, path, protocol), daemon=False)
    t.start()
    _async_tasks.append(t)
    return t


def clear_async_save_task_queue():
    while _async_tasks:
        _async_tasks.pop().join()


def _convert(obj, return_numpy):
    if isinstance(obj, np.ndarray):
        return obj if return_numpy else _np_to_torch(obj).as_subclass(Tensor)
    if isinstance(obj, tuple) and len(obj) == 2 and isinstance(obj[0], str) and isinstance(obj[1], np.ndarray):
        if return_numpy:
            return obj[1]
        t = _np_to_torch(obj[1]).as_subclass(Tensor)
        t.name = obj[0]
        return t
    if isinstance(obj, dict):
        return type(obj)((k, _convert(v, return_numpy)) for k, v in obj.items()) if type(obj) in (dict, OrderedDict) else obj
    if isinstance(obj, list):
        return [_convert(v, return_numpy) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_convert(v, return_numpy) for v in obj)
    return obj


def load(path, **configs):
    """paddle.load(path, return_numpy=False, keep_name_table=False)."""
    return_numpy = configs.get("return_numpy", False)
    keep_name_table = configs.get("keep_name_table", False)
    if hasattr(path, "read"):
        obj = pickle.load(path, encoding="latin1")
    else:
        if not os.path.exists(path):
            raise ValueError(f"The path `{path}` does not exist")
        with open(path, "rb") as f:
            obj = pickle.load(f, encoding="latin1")
    if isinstance(obj, dict):
        obj = _pack_big(obj)
        names = obj.get("StructuredToParameterName@@") if isinstance(obj.get("StructuredToParameterName@@", None), dict) else None
        if names is not None and not keep_name_table:
            del obj["StructuredToParameterName@@"]
        out = _convert(obj, return_numpy)
        if names and not return_numpy:
            for k, n in names.items():
                if k in out and isinstance(out[k], Tensor):
                    out[k].name = n
        return out
    return _convert(obj, return_numpy)
