"""paddle.utils.dlpack. Parity: python/paddle/utils/dlpack.py."""
import torch
import torch.utils.dlpack as _d

from ..tensor import Tensor


def to_dlpack(x):
    return _d.to_dlpack(x.as_subclass(torch.Tensor))


def from_dlpack(dlpack):
    t = _d.from_dlpack(dlpack)
    return t.as_subclass(Tensor)
