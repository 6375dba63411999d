"""dtypes. Parity: python/paddle/framework/dtype.py (reference).

paddle_b200 dtypes ARE torch dtypes; strings and numpy dtypes are accepted
everywhere through :func:`convert_dtype`.
"""
from __future__ import annotations

import builtins

import numpy as np
import torch

bool = torch.bool  # noqa: A001
uint8 = torch.uint8
int8 = torch.int8
int16 = torch.int16
int32 = torch.int32
int64 = torch.int64
float16 = torch.float16
bfloat16 = torch.bfloat16
float32 = torch.float32
float64 = torch.float64
complex64 = torch.complex64
complex128 = torch.complex128
float8_e4m3fn = torch.float8_e4m3fn
float8_e5m2 = torch.float8_e5m2
dtype = torch.dtype

_STR2DTYPE = {
    "bool": torch.bool, "uint8": torch.uint8, "int8": torch.int8, "int16": torch.int16,
    "int32": torch.int32, "int64": torch.int64, "float16": torch.float16, "half": torch.float16,
    "bfloat16": torch.bfloat16, "uint16": torch.bfloat16, "float32": torch.float32,
    "float": torch.float32, "float64": torch.float64, "double": torch.float64,
    "complex64": torch.complex64, "complex128": torch.complex128,
    "float8_e4m3fn": torch.float8_e4m3fn, "float8_e5m2": torch.float8_e5m2, "int": torch.int32,
}
_NP2DTYPE = {
    np.dtype("bool"): torch.bool, np.dtype("uint8"): torch.uint8, np.dtype("int8"): torch.int8,
    np.dtype("int16"): torch.int16, np.dtype("int32"): torch.int32, np.dtype("int64"): torch.int64,
    np.dtype("float16"): torch.float16, np.dtype("float32"): torch.float32,
    np.dtype("float64"): torch.float64, np.dtype("complex64"): torch.complex64,
    np.dtype("complex128"): torch.complex128, np.dtype("uint16"): torch.bfloat16,
}
_DTYPE2NP = {
    torch.bool: np.bool_, torch.uint8: np.uint8, torch.int8: np.int8, torch.int16: np.int16,
    torch.int32: np.int32, torch.int64: np.int64, torch.float16: np.float16,
    torch.float32: np.float32, torch.float64: np.float64, torch.complex64: np.complex64,
    torch.complex128: np.complex128, torch.bfloat16: np.uint16,
}
_default_dtype = torch.float32
_builtin_bool = builtins.bool


def convert_dtype(d):
    """Anything dtype-like -> torch.dtype (None passes through)."""
    if d is None or isinstance(d, torch.dtype):
        return d
    if isinstance(d, str):
        key = d.replace("paddle.", "").replace("torch.", "")
        if key in _STR2DTYPE:
            return _STR2DTYPE[key]
        raise TypeError(f"unsupported dtype string {d!r}")
    if d is float:
        return _default_dtype
    if d is int:
        return torch.int64
    if d is _builtin_bool:
        return torch.bool
    if d is complex:
        return torch.complex64
    try:
        return _NP2DTYPE[np.dtype(d)]
    except Exception as e:  # pragma: no cover
        raise TypeError(f"unsupported dtype {d!r}") from e


def dtype_name(d) -> str:
    return str(convert_dtype(d)).replace("torch.", "")


def to_numpy_dtype(d):
    return _DTYPE2NP[convert_dtype(d)]


def set_default_dtype(d):
    global _default_dtype
    d = convert_dtype(d)
    if d not in (torch.float16, torch.bfloat16, torch.float32, torch.float64):
        raise TypeError("set_default_dtype only supports floating dtypes")
    _default_dtype = d
    torch.set_default_dtype(d)


def get_default_dtype() -> str:
    return dtype_name(_default_dtype)


def default_dtype() -> torch.dtype:
    return _default_dtype


def is_floating(d):
    return convert_dtype(d).is_floating_point


class finfo:
    def __init__(self, d):
        i = torch.finfo(convert_dtype(d))
        self.min, self.max, self.eps = i.min, i.max, i.eps
        self.tiny = self.smallest_normal = i.tiny
        self.resolution, self.bits, self.dtype = i.resolution, i.bits, dtype_name(d)


class iinfo:
    def __init__(self, d):
        i = torch.iinfo(convert_dtype(d))
        self.min, self.max, self.bits, self.dtype = i.min, i.max, i.bits, dtype_name(d)
