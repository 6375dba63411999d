"""Places and device selection. Parity: python/paddle/device/__init__.py, paddle/phi/common/place.h."""
from __future__ import annotations

import torch


class Place:
    _kind = "undefined"

    def __init__(self, idx: int = 0):
        self._idx = int(idx)

    def get_device_id(self):
        return self._idx

    def is_gpu_place(self):
        return self._kind == "gpu"

    def is_cpu_place(self):
        return self._kind == "cpu"

    def is_cuda_pinned_place(self):
        return self._kind == "gpu_pinned"

    def is_custom_place(self):
        return False

    def to_torch(self) -> torch.device:
        if self._kind == "gpu":
            return torch.device("cuda", self._idx)
        return torch.device("cpu")

    def __eq__(self, other):
        return isinstance(other, Place) and (self._kind, self._idx) == (other._kind, other._idx)

    def __hash__(self):
        return hash((self._kind, self._idx))

    def __repr__(self):
        if self._kind == "gpu":
            return f"Place(gpu:{self._idx})"
        return f"Place({self._kind})"


class CPUPlace(Place):
    _kind = "cpu"


class CUDAPlace(Place):
    _kind = "gpu"


class CUDAPinnedPlace(Place):
    _kind = "gpu_pinned"


_current = None  # torch.device


def _default_device() -> torch.device:
    global _current
    if _current is None:
        if torch.cuda.is_available():
            _current = torch.device("cuda", torch.cuda.current_device())
        else:
            _current = torch.device("cpu")
    return _current


def to_torch_device(place) -> torch.device:
    if place is None:
        return _default_device()
    if isinstance(place, torch.device):
        return place
    if isinstance(place, Place):
        return place.to_torch()
    if isinstance(place, int):
        return torch.device("cuda", place)
    if isinstance(place, str):
        s = place.lower().replace("gpu", "cuda")
        if s == "cuda_pinned":
            return torch.device("cpu")
        return torch.device(s)
    raise TypeError(f"cannot interpret {place!r} as a place")


def place_of(t: torch.Tensor) -> Place:
    if t.device.type == "cuda":
        return CUDAPlace(t.device.index or 0)
    if t.device.type == "cpu" and t.is_pinned():
        return CUDAPinnedPlace()
    return CPUPlace()


def set_device(device):
    """paddle.set_device('gpu:0' | 'cpu' | 'gpu')."""
    global _current
    dev = to_torch_device(device)
    if dev.type == "cuda":
        if not torch.cuda.is_available():
            raise RuntimeError("set_device('gpu') but no CUDA device is visible")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        torch.cuda.set_device(idx)
        dev = torch.device("cuda", idx)
    _current = dev
    return place_of(torch.empty(0, device=dev))


def get_device() -> str:
    d = _default_device()
    return f"gpu:{d.index or 0}" if d.type == "cuda" else "cpu"


def is_compiled_with_cuda() -> bool:
    return True


def is_compiled_with_rocm() -> bool:
    return False


def is_compiled_with_xpu() -> bool:
    return False


def is_compiled_with_custom_device(device_type="") -> bool:
    from ..device import custom

    return custom.is_compiled_with_custom_device(device_type)


def is_compiled_with_cinn() -> bool:
    return False


def is_compiled_with_distribute() -> bool:
    return True
