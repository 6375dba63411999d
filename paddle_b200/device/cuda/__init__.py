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


def empty_cache():
    torch.cuda.empty_cache()


def max_memory_allocated(device=None):
    return torch.cuda.max_memory_allocated()


def max_memory_reserved(device=None):
    return torch.cuda.max_memory_reserved()


def memory_allocated(device=None):
    return torch.cuda.memory_allocated()


def memory_reserved(device=None):
    return torch.cuda.memory_reserved()


def reset_max_memory_allocated(device=None):
    if torch.cuda.is_available():
        torch.cuda.reset_peak_memory_stats()


def reset_max_memory_reserved(device=None):
    if torch.cuda.is_available():
        torch.cuda.reset_peak_memory_stats()


def get_device_properties(device=None):
    return torch.cuda.get_device_properties(0 if device is None else device)


def get_device_name(device=None):
    return torch.cuda.get_device_name(0 if device is None else device)


def get_device_capability(device=None):
    return torch.cuda.get_device_capability(0 if device is None else device)
