#include "tracer.h"

#include <ATen/cuda/CUDAContext.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "runtime.h"

namespace b200 {
namespace runtime {

std::atomic<int> g_trace_mode{0};

namespace {

inline int64_t now_ns() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Ev {
  std::string name;
  int type;
  int depth;
  int64_t t0, t1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t stream = nullptr;
};

struct Open {
  std::string name;
  int type;
  int64_t t0;
  cudaEvent_t e0;
  cudaStream_t stream;
};

struct ThreadBuf {
  uint64_t tid;
  std::vector<Ev> done;
  std::vector<Open> stack;
  std::mutex mu;   // only contended by collect(); the owning thread takes it uncontended
};

struct Global {
  std::mutex mu;
  std::vector<std::shared_ptr<ThreadBuf>> bufs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t base = nullptr;
  int64_t base_host_ns = 0;
  bool cuda_ok = false;
  uint64_t next_tid = 0;
};

Global& G() {
  static Global g;
  return g;
}

ThreadBuf& local() {
  thread_local std::shared_ptr<ThreadBuf> tb;
  if (!tb) {
    tb = std::make_shared<ThreadBuf>();
    std::lock_guard<std::mutex> l(G().mu);
    tb->tid = G().next_tid++;
    G().bufs.push_back(tb);
  }
  return *tb;
}

cudaEvent_t get_event() {
  {
    std::lock_guard<std::mutex> l(G().mu);
    if (!G().pool.empty()) {
      cudaEvent_t e = G().pool.back();
      G().pool.pop_back();
      return e;
    }
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return e;
}

void put_event(cudaEvent_t e) {
  if (!e) return;
  std::lock_guard<std::mutex> l(G().mu);
  G().pool.push_back(e);
}

void json_escape(const std::string& s, std::string& out) {
  for (char c : s) {
    if (c == '"' || c == '\\') {
      out.push_back('\\');
      out.push_back(c);
    } else if (static_cast<unsigned char>(c) < 0x20) {
      out.push_back(' ');
    } else {
      out.push_back(c);
    }
  }
}

}  // namespace

void trace_begin(const char* name, int type) {
  ThreadBuf& tb = local();
  Open o{name, type, 0, nullptr, nullptr};
  if (g_trace_mode.load(std::memory_order_relaxed) == 2 && G().cuda_ok) {
    o.stream = at::cuda::getCurrentCUDAStream().stream();
    o.e0 = get_event();
    if (o.e0) cudaEventRecord(o.e0, o.stream);
  }
  o.t0 = now_ns();
  std::lock_guard<std::mutex> l(tb.mu);
  tb.stack.push_back(std::move(o));
}

void trace_end() {
  int64_t t1 = now_ns();
  ThreadBuf& tb = local();
  std::lock_guard<std::mutex> l(tb.mu);
  if (tb.stack.empty()) return;
  Open o = std::move(tb.stack.back());
  tb.stack.pop_back();
  Ev e;
  e.name = std::move(o.name);
  e.type = o.type;
  e.depth = static_cast<int>(tb.stack.size());
  e.t0 = o.t0;
  e.t1 = t1;
  e.e0 = o.e0;
  e.stream = o.stream;
  if (o.e0) {
    e.e1 = get_event();
    if (e.e1) cudaEventRecord(e.e1, o.stream);
  }
  tb.done.push_back(std::move(e));
}

namespace {

// mode: 0 off, 1 host, 2 host+device (silently host-only when no CUDA device is usable)
void tracer_enable(int mode) {
  Global& g = G();
  if (mode == 2) {
    int n = 0;
    bool ok = cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
    if (!ok) cudaGetLastError();
    if (ok && !g.base) {
      if (cudaEventCreate(&g.base) == cudaSuccess) {
        cudaStream_t s = at::cuda::getCurrentCUDAStream().stream();
        cudaEventRecord(g.base, s);
        cudaEventSynchronize(g.base);
        g.base_host_ns = now_ns();
      } else {
        cudaGetLastError();
        ok = false;
      }
    }
    g.cuda_ok = ok;
  }
  g_trace_mode.store(mode);
}

struct Collected {
  std::string name;
  int type, depth;
  uint64_t tid;
  int64_t t0, t1;
  double dev_t0_us, dev_dur_us;   // < 0 when there is no device timing
};

std::vector<Collected> drain() {
  Global& g = G();
  std::vector<std::shared_ptr<ThreadBuf>> bufs;
  {
    std::lock_guard<std::mutex> l(g.mu);
    bufs = g.bufs;
  }
  std::vector<Collected> out;
  for (auto& tb : bufs) {
    std::vector<Ev> evs;
    {
      std::lock_guard<std::mutex> l(tb->mu);
      evs.swap(tb->done);
    }
    for (auto& e : evs) {
      Collected c{std::move(e.name), e.type, e.depth, tb->tid, e.t0, e.t1, -1.0, -1.0};
      if (e.e0 && e.e1) {
        float ms0 = 0.f, ms = 0.f;
        if (cudaEventSynchronize(e.e1) == cudaSuccess && cudaEventElapsedTime(&ms, e.e0, e.e1) == cudaSuccess &&
            cudaEventElapsedTime(&ms0, g.base, e.e0) == cudaSuccess) {
          c.dev_t0_us = g.base_host_ns / 1e3 + ms0 * 1e3;
          c.dev_dur_us = ms * 1e3;
        } else {
          cudaGetLastError();
        }
      }
      put_event(e.e0);
      put_event(e.e1);
      out.push_back(std::move(c));
    }
  }
  return out;
}

pybind11::list tracer_collect() {
  std::vector<Collected> evs;
  {
    pybind11::gil_scoped_release nogil;
    evs = drain();
  }
  pybind11::list out;
  for (auto& c : evs) out.append(pybind11::make_tuple(c.name, c.type, c.tid, c.depth, c.t0, c.t1, c.dev_t0_us, c.dev_dur_us));
  return out;
}

// Drains the buffers straight into a chrome://tracing file; returns the number of host ranges written.
int64_t tracer_export_chrome(const std::string& path, int64_t pid, const std::string& extra_events_json) {
  std::vector<Collected> evs;
  {
    pybind11::gil_scoped_release nogil;
    evs = drain();
  }
  FILE* f = std::fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("tracer: cannot open " + path);
  std::fputs("{\"traceEvents\":[", f);
  bool first = true;
  std::string esc;
  for (auto& c : evs) {
    esc.clear();
    json_escape(c.name, esc);
    std::fprintf(f, "%s{\"name\":\"%s\",\"cat\":\"host:%d\",\"ph\":\"X\",\"ts\":%.3f,\"dur\":%.3f,\"pid\":%lld,\"tid\":%llu}", first ? "" : ",", esc.c_str(),
                 c.type, c.t0 / 1e3, (c.t1 - c.t0) / 1e3, static_cast<long long>(pid), static_cast<unsigned long long>(c.tid));
    first = false;
    if (c.dev_dur_us >= 0)
      std::fprintf(f, ",{\"name\":\"%s\",\"cat\":\"device\",\"ph\":\"X\",\"ts\":%.3f,\"dur\":%.3f,\"pid\":%lld,\"tid\":\"stream\"}", esc.c_str(), c.dev_t0_us,
                   c.dev_dur_us, static_cast<long long>(pid));
  }
  if (!extra_events_json.empty()) std::fprintf(f, "%s%s", first ? "" : ",", extra_events_json.c_str());
  std::fputs("]}", f);
  std::fclose(f);
  return static_cast<int64_t>(evs.size());
}

void py_begin(const std::string& name, int type) {
  if (g_trace_mode.load(std::memory_order_relaxed)) trace_begin(name.c_str(), type);
}

void py_end() {
  if (g_trace_mode.load(std::memory_order_relaxed)) trace_end();
}

}  // namespace

void bind_tracer(pybind11::module_& m) {
  m.def("tracer_enable", &tracer_enable);
  m.def("tracer_mode", []() { return g_trace_mode.load(); });
  m.def("tracer_begin", &py_begin);
  m.def("tracer_end", &py_end);
  m.def("tracer_collect", &tracer_collect);
  m.def("tracer_export_chrome", &tracer_export_chrome, pybind11::arg("path"), pybind11::arg("pid") = 0, pybind11::arg("extra_events_json") = std::string());
}

}  // namespace runtime
}  // namespace b200
