"""paddle.device. Parity: python/paddle/device/__init__.py, device/cuda/."""
from __future__ import annotations

import torch

from ..framework.place import (CPUPlace, CUDAPinnedPlace, CUDAPlace, get_device, is_compiled_with_cinn, is_compiled_with_cuda,  # noqa: F401
                               is_compiled_with_custom_device, is_compiled_with_distribute, is_compiled_with_rocm,
                               is_compiled_with_xpu, set_device)
from . import cuda  # noqa: F401


def is_compiled_with_ipu():
    return False


def is_compiled_with_mlu():
    return False


def is_compiled_with_npu():
    return False


def get_cudnn_version():
    v = torch.backends.cudnn.version() if torch.backends.cudnn.is_available() else None
    return v


def get_all_device_type():
    return ["cpu"] + (["gpu"] if torch.cuda.is_available() else [])


def get_all_custom_device_type():
    from . import custom

    return custom.get_all_custom_device_type()      # plug-ins loaded through device.custom.load_custom_device / CUSTOM_DEVICE_ROOT


def get_available_device():
    return ["cpu"] + [f"gpu:{i}" for i in range(torch.cuda.device_count())]


def get_available_custom_device():
    from . import custom

    return custom.get_available_custom_device()


def device_count():
    return torch.cuda.device_count()


def synchronize(device=None):
    if torch.cuda.is_available():
        torch.cuda.synchronize()


Stream = cuda.Stream
Event = cuda.Event
current_stream = cuda.current_stream
set_stream = cuda.set_stream
stream_guard = cuda.stream_guard


class XPUPlace:
    def __init__(self, *a):
        raise RuntimeError("XPU is not supported by paddle_b200 (sm_100a only)")


IPUPlace = XPUPlace


def side_stream(device=None):
    """A stream for work that overlaps the compute stream (H2D prefetch, bucket reductions). With `FLAGS_b200_sync_debug`
    the current stream is returned instead, which serialises everything: a failure that disappears under the flag is a
    missing event / record_stream (docs/race_detection.md)."""
    import torch

    from ..framework.flags import get_flags

    if get_flags("FLAGS_b200_sync_debug")["FLAGS_b200_sync_debug"]:
        return torch.cuda.current_stream(device)
    return torch.cuda.Stream(device=device)
