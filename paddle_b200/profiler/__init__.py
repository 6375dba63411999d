"""paddle.profiler. Parity: python/paddle/profiler/{profiler,utils,timer}.py.

Host ranges + CUDA-event device ranges, chrome-trace export, summary tables; RecordEvent also emits NVTX ranges so ncu /
nsys attribute kernels to framework scopes."""
from __future__ import annotations

import contextlib
import enum
import json
import os
import time
from collections import defaultdict

import torch


class ProfilerState(enum.Enum):
    CLOSED = 0
    READY = 1
    RECORD = 2
    RECORD_AND_RETURN = 3


class ProfilerTarget(enum.Enum):
    CPU = 0
    GPU = 1
    XPU = 2
    CUSTOM_DEVICE = 3


class SortedKeys(enum.Enum):
    CPUTotal = 0
    CPUAvg = 1
    CPUMax = 2
    CPUMin = 3
    GPUTotal = 4
    GPUAvg = 5
    GPUMax = 6
    GPUMin = 7


class SummaryView(enum.Enum):
    DeviceView = 0
    OverView = 1
    ModelView = 2
    DistributedView = 3
    KernelView = 4
    OperatorView = 5
    MemoryView = 6
    MemoryManipulationView = 7
    UDFView = 8


class TracerEventType(enum.Enum):
    Operator = 0
    Dataloader = 1
    ProfileStep = 2
    Forward = 10
    Backward = 11
    Optimization = 12
    Communication = 13
    PythonOp = 14
    PythonUserDefined = 15
    UserDefined = 16


def make_scheduler(*, closed, ready, record, repeat=0, skip_first=0):
    period = closed + ready + record

    def sched(step):
        if step < skip_first:
            return ProfilerState.CLOSED
        s = step - skip_first
        if repeat > 0 and s // period >= repeat:
            return ProfilerState.CLOSED
        m = s % period
        if m < closed:
            return ProfilerState.CLOSED
        if m < closed + ready:
            return ProfilerState.READY
        return ProfilerState.RECORD_AND_RETURN if m == period - 1 else ProfilerState.RECORD

    return sched


def export_chrome_tracing(dir_name, worker_name=None):
    def handler(prof):
        os.makedirs(dir_name, exist_ok=True)
        name = worker_name or f"host_{os.getpid()}"
        prof.export(os.path.join(dir_name, f"{name}_step{prof.step_num}.paddle_trace.json"))

    return handler


def export_protobuf(dir_name, worker_name=None):
    return export_chrome_tracing(dir_name, worker_name)


_active = [None]


def _native():
    """The C++ range tracer (csrc/runtime/tracer.cpp) when the extension is built; None otherwise."""
    try:
        from .. import _build

        return _build.load(required=False)
    except Exception:
        return None


def kernel_statistics(prof):
    """{kernel entry: (calls, host_ms, device_ms)} of the last recorded cycles."""
    out = {}
    for name, typ, tid, depth, t0, t1, d0, dd in prof._kernel_events:
        c, h, g = out.get(name, (0, 0.0, 0.0))
        out[name] = (c + 1, h + (t1 - t0) / 1e6, g + max(dd, 0.0) / 1e3)
    return out


class RecordEvent:
    """User range. Works as context manager / begin-end pair. Parity: profiler/utils.py:RecordEvent."""

    def __init__(self, name, event_type=TracerEventType.PythonUserDefined):
        self.name, self.event_type = name, event_type
        self._t0 = None

    def begin(self):
        self._t0 = time.perf_counter_ns()
        self._ev = None
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
            p = _active[0]
            if p is not None and p._recording and ProfilerTarget.GPU in p.targets:
                self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                self._ev[0].record()

    def end(self):
        if self._t0 is None:
            return
        t1 = time.perf_counter_ns()
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
            if self._ev is not None:
                self._ev[1].record()
        p = _active[0]
        if p is not None and p._recording:
            p._events.append((self.name, self.event_type.name, self._t0, t1, self._ev))
        self._t0 = None

    def __enter__(self):
        self.begin()
        return self

    def __exit__(self, *a):
        self.end()


class Profiler:
    def __init__(self, *, targets=None, scheduler=None, on_trace_ready=None, record_shapes=False, profile_memory=False, timer_only=False,
                 emit_nvtx=False, custom_device_types=None, with_flops=False):
        self.targets = list(targets) if targets else [ProfilerTarget.CPU] + ([ProfilerTarget.GPU] if torch.cuda.is_available() else [])
        if isinstance(scheduler, (tuple, list)):
            lo, hi = scheduler
            self._sched = make_scheduler(closed=max(lo - 1, 0), ready=1 if lo > 0 else 0, record=hi - lo, repeat=1)
        else:
            self._sched = scheduler or (lambda step: ProfilerState.RECORD)
        self.on_trace_ready = on_trace_ready
        self.timer_only = timer_only
        self.step_num = 0
        self._events, self._recording = [], False
        self._kernel_events = []   # (name, type, tid, depth, t0_ns, t1_ns, dev_t0_us, dev_dur_us) from the native tracer
        self._step_times, self._t_step = [], None
        self._torch_prof = None
        self._mem = profile_memory

    def start(self):
        _active[0] = self
        self._t_step = time.perf_counter()
        self._apply_state()

    def _apply_state(self):
        st = self._sched(self.step_num)
        rec = st in (ProfilerState.RECORD, ProfilerState.RECORD_AND_RETURN) and not self.timer_only
        if rec and not self._recording:
            acts = [torch.profiler.ProfilerActivity.CPU] + ([torch.profiler.ProfilerActivity.CUDA] if torch.cuda.is_available() and ProfilerTarget.GPU in self.targets else [])
            self._torch_prof = torch.profiler.profile(activities=acts, profile_memory=self._mem)
            self._torch_prof.__enter__()
        if rec and not self._recording:
            C = _native()
            if C is not None:
                C.tracer_collect()   # drop anything recorded outside a cycle
                C.tracer_enable(2 if (ProfilerTarget.GPU in self.targets and torch.cuda.is_available()) else 1)
        self._recording = rec
        self._state = st

    def step(self, num_samples=None):
        now = time.perf_counter()
        if self._t_step is not None:
            self._step_times.append((now - self._t_step, num_samples))
        self._t_step = now
        prev = getattr(self, "_state", ProfilerState.CLOSED)
        self.step_num += 1
        if prev == ProfilerState.RECORD_AND_RETURN:
            self._finish_cycle()
        self._apply_state()

    def _finish_cycle(self):
        if self._torch_prof is not None:
            self._torch_prof.__exit__(None, None, None)
            self._last_prof, self._torch_prof = self._torch_prof, None
        self._recording = False
        C = _native()
        if C is not None and C.tracer_mode():
            self._kernel_events.extend(C.tracer_collect())
            C.tracer_enable(0)
        if self.on_trace_ready:
            self.on_trace_ready(self)

    def stop(self):
        if self._recording or self._torch_prof is not None:
            self._finish_cycle()
        _active[0] = None

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, *a):
        self.stop()

    def step_info(self, unit=None):
        if not self._step_times:
            return ""
        ts = [t for t, _ in self._step_times[-10:]]
        avg = sum(ts) / len(ts)
        msg = f"reader_cost: 0.00000 s batch_cost: {avg:.5f} s"
        ns = [n for _, n in self._step_times[-10:] if n]
        if ns:
            msg += f" ips: {sum(ns) / sum(ts):.3f} {unit or 'samples'}/s"
        return msg

    def export(self, path, format="json"):
        evs = []
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for name, typ, t0, t1, ev in self._events:
            evs.append({"name": name, "cat": typ, "ph": "X", "ts": t0 / 1e3, "dur": (t1 - t0) / 1e3, "pid": os.getpid(), "tid": 0})
            if ev is not None:
                evs.append({"name": name, "cat": "gpu:" + typ, "ph": "X", "ts": t0 / 1e3, "dur": ev[0].elapsed_time(ev[1]) * 1e3, "pid": os.getpid(), "tid": 1})
        for name, typ, tid, depth, t0, t1, d0, dd in self._kernel_events:
            evs.append({"name": name, "cat": "kernel_launch", "ph": "X", "ts": t0 / 1e3, "dur": (t1 - t0) / 1e3, "pid": os.getpid(), "tid": 100 + tid})
            if dd >= 0:
                evs.append({"name": name, "cat": "kernel", "ph": "X", "ts": d0, "dur": dd, "pid": os.getpid(), "tid": "stream"})
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        lp = getattr(self, "_last_prof", None)
        if lp is not None:
            try:
                lp.export_chrome_trace(path)
                if evs:
                    data = json.load(open(path))
                    data["traceEvents"].extend(evs)
                    json.dump(data, open(path, "w"))
                return
            except Exception:
                pass
        json.dump({"traceEvents": evs}, open(path, "w"))

    def summary(self, sorted_by=SortedKeys.CPUTotal, op_detail=True, thread_sep=False, time_unit="ms", views=None):
        lp = getattr(self, "_last_prof", None)
        if lp is not None:
            key = "cuda_time_total" if "GPU" in sorted_by.name and torch.cuda.is_available() else "cpu_time_total"
            print(lp.key_averages().table(sort_by=key, row_limit=30))
        agg = defaultdict(lambda: [0, 0.0])
        for name, typ, t0, t1, ev in self._events:
            agg[name][0] += 1
            agg[name][1] += (t1 - t0) / 1e6
        kagg = defaultdict(lambda: [0, 0.0, 0.0])
        for name, typ, tid, depth, t0, t1, d0, dd in self._kernel_events:
            kagg[name][0] += 1
            kagg[name][1] += (t1 - t0) / 1e6
            kagg[name][2] += max(dd, 0.0) / 1e3
        if kagg:
            print(f"{'kernel entry (native tracer)':<40}{'calls':>8}{'host ms':>14}{'device ms':>14}")
            for k, (c, h, g) in sorted(kagg.items(), key=lambda kv: -(kv[1][2] or kv[1][1])):
                print(f"{k:<40}{c:>8}{h:>14.3f}{g:>14.3f}")
        if agg:
            print(f"{'user range':<40}{'calls':>8}{'total ms':>14}")
            for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                print(f"{k:<40}{c:>8}{t:>14.3f}")


def load_profiler_result(filename):
    return json.load(open(filename))


@contextlib.contextmanager
def profile_range(name):
    with RecordEvent(name):
        yield
