"""Custom-device plug-in ABI (include/b200_device_ext.h): build the sample plug-in with the system C compiler, load it, and drive memory,
copies, streams / events, device kernels and the host fallback through paddle_b200.device.custom.
Parity model: test/custom_runtime/test_custom_cpu_plugin.py."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import paddle_b200 as paddle  # noqa: F401
from paddle_b200._build import load
from paddle_b200.device import custom as C

_m = load()
pytestmark = pytest.mark.skipif(_m is None or not hasattr(_m, "CustomDevice") or shutil.which("gcc") is None, reason="native extension or gcc missing")

HERE = os.path.dirname(os.path.abspath(__file__))
INC = os.path.join(os.path.dirname(HERE), "paddle_b200", "include")


def _build(tmp_path, name, defines=()):
    out = str(tmp_path / f"lib{name}.so")
    cmd = ["gcc", "-shared", "-fPIC", "-O1", "-I", INC, *[f"-D{d}" for d in defines], os.path.join(HERE, "plugins", "custom_cpu.c"), "-o", out]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


@pytest.fixture()
def plugin(tmp_path):
    C.unload_custom_device("custom_cpu")
    dev = C.load_custom_device(_build(tmp_path, "custom_cpu"))
    yield dev
    C.unload_custom_device("custom_cpu")


def test_load_register_and_query(plugin):
    assert plugin.device_type == "custom_cpu" and plugin.device_count() == 2
    assert C.get_all_custom_device_type() == ["custom_cpu"] and C.is_compiled_with_custom_device("custom_cpu") and not C.is_compiled_with_custom_device("npu")
    assert C.get_available_custom_device() == ["custom_cpu:0", "custom_cpu:1"]
    p = C.CustomPlace("custom_cpu", 1)
    assert p.get_device_type() == "custom_cpu" and p.get_device_id() == 1 and repr(p) == "Place(custom_cpu:1)"
    with pytest.raises(ValueError):
        C.CustomPlace("custom_cpu", 2)
    with pytest.raises(RuntimeError):
        C.CustomPlace("npu", 0)


def test_memory_copies_and_stats(plugin):
    place = C.CustomPlace("custom_cpu", 0)
    total, free0, alloc0, _ = C.memory_stats(place)
    a = np.arange(24, dtype=np.float32).reshape(4, 6)
    t = C.to_device(a, place)
    assert t.shape == [4, 6] and t.dtype == np.float32 and np.array_equal(t.numpy(), a)
    total1, free1, alloc1, peak1 = C.memory_stats(place)
    assert total1 == total and free0 - free1 == a.nbytes and alloc1 - alloc0 == a.nbytes and peak1 >= alloc1
    u = C.CustomTensor(place, [4, 6], "float32").copy_(t)              # device-to-device
    assert np.array_equal(u.numpy(), a)
    u.copy_(a * 2)
    assert np.array_equal(u.numpy(), a * 2)
    with pytest.raises(RuntimeError):                                  # copies are bounds-checked against the live allocation
        plugin.memcpy_d2h(0, t.data_ptr, a.nbytes + 4)
    with pytest.raises(RuntimeError):
        plugin.free(12345)
    del t, u
    assert C.memory_stats(place)[2] == alloc0
    with pytest.raises(RuntimeError):                                  # the plug-in reports out-of-memory as a status
        C.CustomTensor(place, [1 << 28], "float32")
    assert np.array_equal(C.to_device(paddle.to_tensor(a), place).to_tensor().numpy(), a)


def test_streams_and_events(plugin):
    place = C.CustomPlace("custom_cpu", 0)
    s = C.Stream(place)
    e = C.Event(place)
    with pytest.raises(RuntimeError):
        e.synchronize()                                                # never recorded: the plug-in returns a status, we raise
    s.record_event(e)
    e.synchronize()
    s.synchronize()
    C.synchronize(place)


def test_device_kernel_and_host_fallback(plugin):
    place = C.CustomPlace("custom_cpu", 0)
    a = C.to_device(np.random.RandomState(0).randn(8, 5).astype("float32"), place)
    b = C.to_device(np.random.RandomState(1).randn(8, 5).astype("float32"), place)
    C.stats.update(device_kernels=0, host_fallbacks=0)
    c = a + b                                                          # the plug-in implements float32 add
    assert C.stats == {"device_kernels": 1, "host_fallbacks": 0}
    np.testing.assert_allclose(c.numpy(), a.numpy() + b.numpy(), rtol=1e-6)
    d = a * b                                                          # no multiply kernel: host fallback, result lands back on the device
    w = C.to_device(np.random.RandomState(2).randn(5, 3).astype("float32"), place)
    e = d @ w
    r = C.relu(e)
    assert C.stats == {"device_kernels": 1, "host_fallbacks": 3}
    np.testing.assert_allclose(r.numpy(), np.maximum((a.numpy() * b.numpy()) @ w.numpy(), 0), rtol=1e-5)
    i = C.to_device(np.arange(6, dtype=np.int64), place)
    np.testing.assert_array_equal((i + i).numpy(), np.arange(6) * 2)   # int64 add is not in the plug-in either


def test_rejects_bad_plugins(tmp_path):
    C.unload_custom_device("custom_cpu")
    with pytest.raises(RuntimeError, match="ABI"):
        C.load_custom_device(_build(tmp_path, "bad_abi", ["PLUGIN_BAD_ABI"]))
    with pytest.raises(RuntimeError, match="memcpy_d2h"):
        C.load_custom_device(_build(tmp_path, "missing", ["PLUGIN_MISSING_REQUIRED"]))
    with pytest.raises(RuntimeError, match="cannot load"):
        C.load_custom_device(str(tmp_path / "nope.so"))
    dev = C.load_custom_device(_build(tmp_path, "nokern", ["PLUGIN_NO_KERNELS"]))
    try:
        place = C.CustomPlace("custom_cpu", 0)
        C.stats.update(device_kernels=0, host_fallbacks=0)
        x = C.to_device(np.ones(4, "float32"), place)
        assert np.array_equal((x + x).numpy(), np.full(4, 2, "float32")) and C.stats["host_fallbacks"] == 1
        with pytest.raises(RuntimeError, match="already registered"):
            C.load_custom_device(dev.path)
    finally:
        C.unload_custom_device("custom_cpu")


def test_plugin_device_type_is_a_kernel_factory_backend(plugin):
    """Loading a plug-in registers its device type as a backend: `dispatch(op, custom tensors)` selects its kernels, and a user kernel registered
    for that backend (reference: custom kernels for custom devices) takes part in the same selection."""
    from paddle_b200.kernels.registry import KernelFactory, KernelKey, dispatch, register_kernel

    f = KernelFactory.instance()
    assert f.has_kernel("add", KernelKey("custom_cpu", "float32")) and f.has_kernel("relu", KernelKey("custom_cpu", "float32"))
    place = C.CustomPlace("custom_cpu", 0)
    a = C.to_device(np.arange(6, dtype=np.float32).reshape(2, 3) - 2, place)
    b = C.to_device(np.ones((2, 3), np.float32), place)
    k = f.select("add", a, b)
    assert k.backend == "custom_cpu"
    np.testing.assert_allclose(dispatch("add", a, b).numpy(), a.numpy() + 1)
    np.testing.assert_allclose(dispatch("relu", a).numpy(), np.maximum(a.numpy(), 0))

    @register_kernel("negate", backend="custom_cpu", dtypes=("float32",))
    def negate(t):
        return C.to_device(-t.numpy(), t.place)

    np.testing.assert_allclose(dispatch("negate", a).numpy(), -a.numpy())
    with pytest.raises(NotImplementedError):
        dispatch("negate", paddle.ones([2]))                      # no kernel of this op for the CPU backend
