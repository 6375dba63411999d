"""paddle.device.cuda. Parity: python/paddle/device/cuda/__init__.py."""
from __future__ import annotations

import contextlib

import torch


class Stream(torch.cuda.Stream):
    def __new__(cls, device=None, priority=2, **kw):
        return super().__new__(cls, priority=-1 if priority == 1 else 0)

    def wait_event(self, event):
        super().wait_event(event)

    def wait_stream(self, stream):
        super().wait_stream(stream)

    def record_event(self, event=None):
        return super().record_event(event)


class Event(torch.cuda.Event):
    def __new__(cls, enable_timing=False, blocking=False, interprocess=False):
        return super().__new__(cls, enable_timing=enable_timing, blocking=blocking, interprocess=interprocess)


def current_stream(device=None):
    return torch.cuda.current_stream()


def set_stream(stream):
    prev = torch.cuda.current_stream()
    torch.cuda.set_stream(stream)
    return prev


@contextlib.contextmanager
def stream_guard(stream):
    with torch.cuda.stream(stream):
        yield


def synchronize(device=None):
    torch.cuda.synchronize()


def device_count():
    return torch.cuda.device_count()


# ---- allocator.  Default: PyTorch's caching allocator.  FLAGS_allocator_strategy=auto_growth (environment, before the first CUDA
# allocation) or `use_auto_growth_allocator()` routes EVERY device allocation of the process through the native auto-growth best-fit
# allocator (csrc/runtime/allocator.cpp: chunked cudaMalloc, best-fit free map, neighbour coalescing, event-deferred cross-stream frees),
# loaded as a torch CUDAPluggableAllocator; the statistics below then come from it.
_auto_growth = [False]


def _dev_index(device):
    if device is None:
        return torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, int):
        return device
    s = str(device)
    return int(s.split(":")[1]) if ":" in s else 0


def use_auto_growth_allocator(chunk_mb=None):
    """Install the native auto-growth allocator for all CUDA memory of this process.  Must run before the first CUDA allocation
    (PyTorch cannot swap allocators afterwards); returns True when installed."""
    import os

    from ..._build import load

    if _auto_growth[0]:
        return True
    m = load()
    if m is None or not hasattr(m, "cuda_allocator_stats"):
        raise RuntimeError("use_auto_growth_allocator: the native extension is not built")
    if chunk_mb is not None:
        os.environ["B200_ALLOCATOR_CHUNK_MB"] = str(int(chunk_mb))
    alloc = torch.cuda.memory.CUDAPluggableAllocator(m.__file__, "b200_cuda_malloc", "b200_cuda_free")
    torch.cuda.memory.change_current_allocator(alloc)
    _auto_growth[0] = True
    return True


def auto_growth_allocator_active():
    return _auto_growth[0]


def allocator_stats(device=None):
    """Counters of the auto-growth allocator: allocated / reserved (+ peaks), chunks, backend (cudaMalloc) calls, splits, merges, frees
    deferred behind an event."""
    from ..._build import load

    return dict(load().cuda_allocator_stats(_dev_index(device)))


def empty_cache():
    if _auto_growth[0]:
        from ..._build import load

        torch.cuda.synchronize()
        load().cuda_allocator_release_idle(_dev_index(None))
        return
    torch.cuda.empty_cache()


def max_memory_allocated(device=None):
    return allocator_stats(device)["allocated_peak"] if _auto_growth[0] else torch.cuda.max_memory_allocated()


def max_memory_reserved(device=None):
    return allocator_stats(device)["reserved_peak"] if _auto_growth[0] else torch.cuda.max_memory_reserved()


def memory_allocated(device=None):
    return allocator_stats(device)["allocated"] if _auto_growth[0] else torch.cuda.memory_allocated()


def memory_reserved(device=None):
    return allocator_stats(device)["reserved"] if _auto_growth[0] else torch.cuda.memory_reserved()


def reset_max_memory_allocated(device=None):
    if _auto_growth[0]:
        from ..._build import load

        load().cuda_allocator_reset_peak(_dev_index(device))
    elif torch.cuda.is_available():
        torch.cuda.reset_peak_memory_stats()


def reset_max_memory_reserved(device=None):
    reset_max_memory_allocated(device)


def get_device_properties(device=None):
    return torch.cuda.get_device_properties(0 if device is None else device)


def get_device_name(device=None):
    return torch.cuda.get_device_name(0 if device is None else device)


def get_device_capability(device=None):
    return torch.cuda.get_device_capability(0 if device is None else device)
