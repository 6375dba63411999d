"""einsum. Parity: python/paddle/tensor/einsum.py."""
from __future__ import annotations

import torch

from ._helpers import T


def einsum(equation, *operands):
    if len(operands) == 1 and isinstance(operands[0], (list, tuple)):
        operands = operands[0]
    return torch.einsum(equation, *[T(o) for o in operands])


__all__ = ["einsum"]
