// Host/device range tracer. Per-thread append-only buffers (no lock on the record path), optional cudaEvent pairs on the
// launching stream so our own kernels get device durations without CUPTI.  Role of the reference's HostTracer +
// CudaTracer (paddle/fluid/platform/profiler/host_tracer.cc, cuda_tracer.cc), built for one process per GPU.
#pragma once
#include <atomic>
#include <cstdint>

namespace b200 {
namespace runtime {

extern std::atomic<int> g_trace_mode;   // 0 off, 1 host ranges, 2 host + device events

void trace_begin(const char* name, int type);
void trace_end();

struct TraceScope {
  bool on;
  explicit TraceScope(const char* name, int type = 1) : on(g_trace_mode.load(std::memory_order_relaxed) != 0) {
    if (on) trace_begin(name, type);
  }
  ~TraceScope() {
    if (on) trace_end();
  }
};

}  // namespace runtime
}  // namespace b200
