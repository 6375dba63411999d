"""Custom devices: load a vendor plug-in (C ABI `paddle_b200/include/b200_device_ext.h`) and use its memory, streams, events and kernels.

Parity: paddle.device.{get_all_custom_device_type, is_compiled_with_custom_device}, CustomPlace, the CustomDevice runtime
(paddle/phi/backends/custom/, paddle/phi/backends/device_manager.cc) and the custom-kernel fallback to the host.

    dev = load_custom_device("libmy_npu.so")            # or CUSTOM_DEVICE_ROOT=<dir> at import: every *.so inside is loaded
    x = to_device(np_or_tensor, CustomPlace("my_npu", 0))
    y = add(x, x)                                        # plug-in kernel when it has one, host fallback (copy - compute - copy) otherwise
    y.numpy()
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from .._build import load as _load

_REGISTRY = {}          # device type -> native CustomDevice
_NP2NAME = {np.dtype(k): k for k in ("float32", "float64", "float16", "int32", "int64", "int16", "int8", "uint8", "bool")}


def load_custom_device(path):
    """dlopen the plug-in, run B200InitPlugin, validate the interface, register its device type.  Returns the runtime object."""
    m = _load()
    if m is None or not hasattr(m, "CustomDevice"):
        raise RuntimeError("custom devices need the native extension (paddle_b200._C)")
    dev = m.CustomDevice(os.path.abspath(path))
    if dev.device_type in _REGISTRY:
        raise RuntimeError(f"custom device type '{dev.device_type}' is already registered (from {_REGISTRY[dev.device_type].path})")
    _REGISTRY[dev.device_type] = dev
    _register_kernels(dev.device_type)
    return dev


def _register_kernels(device_type):
    """The device type becomes a backend of the KernelFactory: every op with a host implementation gets a kernel that launches the plug-in's device
    kernel when it has one and falls back to the host otherwise; `register_kernel(op, backend=device_type)` adds more (custom kernels, reference:
    paddle/phi/core/custom_kernel.cc)."""
    from ..kernels.registry import KernelFactory

    f = KernelFactory.instance()
    dtypes = tuple(_NP2NAME.values())
    for op in _HOST_IMPL:
        if _HOST_IMPL[op] is None or any(k.backend == device_type for k in f.kernels(op)):
            continue
        f.register(op, device_type, dtypes, (lambda *ins, _op=op: run_op(_op, list(ins))), native=f"plug-in '{device_type}' kernel or host fallback")


def load_custom_device_dir(root):
    return [load_custom_device(p) for p in sorted(glob.glob(os.path.join(root, "*.so")))]


def unload_custom_device(device_type):
    _REGISTRY.pop(device_type, None)


def get_all_custom_device_type():
    return sorted(_REGISTRY)


def get_available_custom_device():
    return [f"{t}:{i}" for t in sorted(_REGISTRY) for i in range(_REGISTRY[t].device_count())]


def is_compiled_with_custom_device(device_type):
    return device_type in _REGISTRY


def _runtime(device_type):
    if device_type not in _REGISTRY:
        raise RuntimeError(f"custom device type '{device_type}' is not registered (loaded: {get_all_custom_device_type()})")
    return _REGISTRY[device_type]


class CustomPlace:
    def __init__(self, device_type, device_id=0):
        self._type, self._id = device_type, int(device_id)
        rt = _runtime(device_type)
        if not 0 <= self._id < rt.device_count():
            raise ValueError(f"{device_type} has {rt.device_count()} device(s); got id {device_id}")

    def get_device_type(self):
        return self._type

    def get_device_id(self):
        return self._id

    def __repr__(self):
        return f"Place({self._type}:{self._id})"

    def __eq__(self, o):
        return isinstance(o, CustomPlace) and (o._type, o._id) == (self._type, self._id)

    def __hash__(self):
        return hash((self._type, self._id))


class Stream:
    def __init__(self, place):
        self.place, self._rt = place, _runtime(place.get_device_type())
        self._h = self._rt.create_stream(place.get_device_id())

    def synchronize(self):
        self._rt.synchronize_stream(self.place.get_device_id(), self._h)

    def record_event(self, event=None):
        event = event or Event(self.place)
        self._rt.record_event(self.place.get_device_id(), self._h, event._h)
        return event

    def __del__(self):
        try:
            self._rt.destroy_stream(self.place.get_device_id(), self._h)
        except Exception:  # noqa: BLE001
            pass


class Event:
    def __init__(self, place):
        self.place, self._rt = place, _runtime(place.get_device_type())
        self._h = self._rt.create_event(place.get_device_id())

    def synchronize(self):
        self._rt.synchronize_event(self.place.get_device_id(), self._h)

    def __del__(self):
        try:
            self._rt.destroy_event(self.place.get_device_id(), self._h)
        except Exception:  # noqa: BLE001
            pass


class CustomTensor:
    """A tensor whose storage lives in plug-in memory."""

    def __init__(self, place, shape, dtype):
        self.place, self.shape, self.dtype = place, [int(s) for s in shape], np.dtype(dtype)
        self._rt = _runtime(place.get_device_type())
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._ptr = self._rt.malloc(place.get_device_id(), max(self.nbytes, 1))

    @property
    def data_ptr(self):
        return self._ptr

    def numpy(self):
        raw = self._rt.memcpy_d2h(self.place.get_device_id(), self._ptr, self.nbytes) if self.nbytes else b""
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.shape).copy()

    def to_tensor(self):
        from ..tensor import Tensor

        return torch.from_numpy(self.numpy()).as_subclass(Tensor)

    def cpu(self):
        return self.to_tensor()

    def copy_(self, other):
        if isinstance(other, CustomTensor):
            if other.nbytes != self.nbytes:
                raise ValueError("copy_: sizes differ")
            self._rt.memcpy_d2d(self.place.get_device_id(), self._ptr, other._ptr, self.nbytes)
        else:
            a = np.ascontiguousarray(_as_numpy(other), dtype=self.dtype).reshape(self.shape)
            self._rt.memcpy_h2d(self.place.get_device_id(), self._ptr, a)
        return self

    def _arg(self):
        return (self._ptr, _NP2NAME[self.dtype], self.shape)

    def __repr__(self):
        return f"CustomTensor(place={self.place!r}, shape={self.shape}, dtype={self.dtype.name})"

    def __del__(self):
        try:
            self._rt.free(self._ptr)
        except Exception:  # noqa: BLE001
            pass

    def __add__(self, o):
        return add(self, o)

    def __mul__(self, o):
        return multiply(self, o)

    def __matmul__(self, o):
        return matmul(self, o)


def _as_numpy(x):
    if isinstance(x, np.ndarray):
        return x
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().as_subclass(torch.Tensor).numpy()
    return np.asarray(x)


def to_device(x, place):
    a = np.ascontiguousarray(_as_numpy(x))
    t = CustomTensor(place, a.shape, a.dtype)
    if t.nbytes:
        t._rt.memcpy_h2d(place.get_device_id(), t._ptr, a)
    return t


def memory_stats(place):
    """(total, free, allocated_by_framework, peak_allocated) in bytes."""
    return tuple(_runtime(place.get_device_type()).memory_stats(place.get_device_id()))


def synchronize(place):
    _runtime(place.get_device_type()).synchronize(place.get_device_id())


_HOST_IMPL = {
    "add": lambda a, b: a + b, "subtract": lambda a, b: a - b, "multiply": lambda a, b: a * b, "matmul": lambda a, b: a @ b,
    "relu": lambda a: np.maximum(a, 0), "exp": np.exp, "scale": None,
}
stats = {"device_kernels": 0, "host_fallbacks": 0}


def run_op(op, inputs, out_shape=None, out_dtype=None, stream=None):
    """`op` over CustomTensors: the plug-in's kernel if it has one, otherwise the host implementation (results copied back to the device:
    the reference's 'fallback to CPU kernel' for custom devices)."""
    place = inputs[0].place
    rt = _runtime(place.get_device_type())
    host_in = None
    if out_shape is None:
        host_in = [t.numpy() for t in inputs]
        ref = _HOST_IMPL[op](*host_in)
        out_shape, out_dtype = ref.shape, ref.dtype
    out = CustomTensor(place, out_shape, out_dtype or inputs[0].dtype)
    ok = rt.launch(place.get_device_id(), 0 if stream is None else stream._h, op, [t._arg() for t in inputs] + [out._arg()], len(inputs))
    if ok:
        stats["device_kernels"] += 1
        return out
    stats["host_fallbacks"] += 1
    if host_in is None:
        host_in = [t.numpy() for t in inputs]
    out.copy_(np.asarray(_HOST_IMPL[op](*host_in)).astype(out.dtype))
    return out


def _bin(op, a, b):
    if not isinstance(b, CustomTensor):
        b = to_device(np.asarray(b, dtype=a.dtype), a.place)
    shape = list(np.broadcast_shapes(tuple(a.shape), tuple(b.shape)))
    return run_op(op, [a, b], shape if shape == a.shape == b.shape else None, a.dtype if shape == a.shape == b.shape else None)


def add(a, b):
    return _bin("add", a, b)


def subtract(a, b):
    return _bin("subtract", a, b)


def multiply(a, b):
    return _bin("multiply", a, b)


def matmul(a, b):
    return run_op("matmul", [a, b], [a.shape[0], b.shape[1]] if len(a.shape) == 2 and len(b.shape) == 2 else None, a.dtype)


def relu(a):
    return run_op("relu", [a], a.shape, a.dtype)


if os.environ.get("CUSTOM_DEVICE_ROOT"):
    try:
        load_custom_device_dir(os.environ["CUSTOM_DEVICE_ROOT"])
    except Exception as e:  # noqa: BLE001
        import warnings

        warnings.warn(f"CUSTOM_DEVICE_ROOT: {e}")
