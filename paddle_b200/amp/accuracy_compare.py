"""Tensor dump + run-to-run comparison for precision debugging. Parity: python/paddle/amp/accuracy_compare.py and the
DUMP_ALL / CHECK_ALL modes of TensorCheckerConfig (paddle/fluid/eager/nan_inf_utils.cc writes one log line per op output).

`TensorDumpMode` is a TorchFunctionMode: while active every floating-point op output is summarised (max / min / mean / #nan /
#inf) into `<output_dir>/worker_<rank>.log`.  `compare_accuracy(dir_a, dir_b, out)` pairs the two logs op by op (same op name,
same occurrence index) and writes a table with the fp32-vs-low-precision deltas, flagging non-finite and diverging entries."""
from __future__ import annotations

import csv
import os
import re

import torch

_LINE = re.compile(r"\[op=(?P<op>[^\]]+)\] \[tensor=(?P<tensor>[^\]]+)\] dtype=(?P<dtype>\S+) numel=(?P<numel>\d+) max=(?P<max>\S+) min=(?P<min>\S+) "
                   r"mean=(?P<mean>\S+) nan=(?P<nan>\d+) inf=(?P<inf>\d+)")


class TensorDumpMode(torch.overrides.TorchFunctionMode):
    def __init__(self, output_dir, checked_op_list=None, skipped_op_list=None, abort_on_nonfinite=False):
        super().__init__()
        from ..distributed import env

        os.makedirs(output_dir, exist_ok=True)
        self.path = os.path.join(output_dir, f"worker_{env.get_rank()}.log")
        self._f = open(self.path, "a")
        self.checked = set(checked_op_list) if checked_op_list else None
        self.skipped = set(skipped_op_list or ())
        self.abort = abort_on_nonfinite
        self._busy = False

    def __torch_function__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        if self._busy:
            return out
        name = getattr(func, "__name__", str(func))
        if name.startswith("_") or name in self.skipped or (self.checked is not None and name not in self.checked):
            return out
        self._busy = True
        try:
            outs = out if isinstance(out, (tuple, list)) else (out,)
            for i, o in enumerate(outs):
                if isinstance(o, torch.Tensor) and o.is_floating_point() and o.numel() > 0 and o.layout == torch.strided:
                    self._dump(name, f"out{i}", o)
        finally:
            self._busy = False
        return out

    def _dump(self, op, tensor_name, o):
        with torch.no_grad():
            f = o.detach().as_subclass(torch.Tensor).float()
            n_nan, n_inf = int(torch.isnan(f).sum()), int(torch.isinf(f).sum())
            fin = f[torch.isfinite(f)]
            mx, mn, mean = (float(fin.max()), float(fin.min()), float(fin.mean())) if fin.numel() else (0.0, 0.0, 0.0)
        self._f.write(f"[op={op}] [tensor={tensor_name}] dtype={str(o.dtype).replace('torch.', '')} numel={o.numel()} max={mx:.9g} min={mn:.9g} "
                      f"mean={mean:.9g} nan={n_nan} inf={n_inf}\n")
        if self.abort and (n_nan or n_inf):
            self._f.flush()
            raise RuntimeError(f"[tensor checker] op={op}: {n_nan} nan, {n_inf} inf in {tensor_name}")

    def close(self):
        self._f.flush()
        self._f.close()


def _read(path):
    files = sorted(os.path.join(path, f) for f in os.listdir(path) if f.startswith("worker_")) if os.path.isdir(path) else [path]
    rows, seen = [], {}
    for fn in files:
        for line in open(fn):
            m = _LINE.search(line)
            if not m:
                continue
            d = m.groupdict()
            key = (os.path.basename(fn), d["op"], d["tensor"])
            k = seen.get(key, 0)
            seen[key] = k + 1
            rows.append(((key[0], d["op"], d["tensor"], k), d))
    return rows


def compare_accuracy(dump_path, another_dump_path, output_filename, loss_scale=1, dump_all_tensors=False):
    """Pair the op logs of two runs (typically fp32 vs fp16/bf16) and write a CSV table; returns the rows that look suspicious."""
    a, b = dict(_read(dump_path)), dict(_read(another_dump_path))
    suspicious, table = [], []
    for key, da in a.items():
        db = b.get(key)
        if db is None:
            continue
        row = {"worker": key[0], "op": key[1], "tensor": key[2], "index": key[3], "dtype_a": da["dtype"], "dtype_b": db["dtype"]}
        bad = False
        for stat in ("max", "min", "mean"):
            va, vb = float(da[stat]), float(db[stat]) / (loss_scale if "grad" in key[1] else 1)
            row[f"{stat}_a"], row[f"{stat}_b"] = va, vb
            row[f"{stat}_diff"] = abs(va - vb)
            if abs(va - vb) > 1e-2 * max(1.0, abs(va)):
                bad = True
        row["nan_a"], row["inf_a"], row["nan_b"], row["inf_b"] = int(da["nan"]), int(da["inf"]), int(db["nan"]), int(db["inf"])
        bad = bad or (row["nan_b"] + row["inf_b"]) > (row["nan_a"] + row["inf_a"])
        row["suspicious"] = int(bad)
        table.append(row)
        if bad:
            suspicious.append(row)
    d = os.path.dirname(output_filename)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(output_filename, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(table[0].keys()) if table else ["op"])
        w.writeheader()
        w.writerows(table)
    return suspicious
