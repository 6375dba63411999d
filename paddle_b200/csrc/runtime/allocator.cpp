// Auto-growth best-fit device allocator with stream-safe frees and statistics.
//
// Parity (role): paddle/phi/core/memory/allocation (AutoGrowthBestFitAllocator, StreamSafeCUDAAllocator, StatAllocator) behind
// FLAGS_allocator_strategy=auto_growth and paddle.device.cuda.{memory_allocated, memory_reserved, max_memory_*, empty_cache}.
// Design: memory comes from the backend (cudaMalloc, or malloc for the host instance the CPU tests drive) in CHUNKS of at least
// `chunk_bytes`; a request takes the smallest free block that fits (std::multimap keyed by size), splitting off the tail when it is
// worth a block; a free coalesces with its free neighbours inside the chunk; chunks that became entirely free go back to the backend on
// `release_idle()` (empty_cache) or when the backend runs out of memory.  A block freed on another stream than the one it was allocated
// on is parked behind a CUDA event and only re-enters the free map once that stream has passed the free point.  The process-wide CUDA
// instance is exported as `b200_cuda_malloc` / `b200_cuda_free`, the signature torch.cuda.memory.CUDAPluggableAllocator loads, so the
// whole framework (every torch allocation included) can run on it.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "runtime.h"

namespace py = pybind11;

namespace b200 {
namespace runtime {

class AutoGrowthAllocator {
 public:
  struct Stats {
    int64_t allocated = 0, allocated_peak = 0, reserved = 0, reserved_peak = 0;
    int64_t num_allocs = 0, num_frees = 0, num_chunks = 0, num_backend_allocs = 0, num_backend_frees = 0, num_splits = 0, num_merges = 0, deferred = 0;
  };

  AutoGrowthAllocator(bool cuda, int device, int64_t chunk_bytes, int64_t alignment)
      : cuda_(cuda), device_(device), chunk_bytes_(chunk_bytes > 0 ? chunk_bytes : (int64_t)64 << 20), align_(alignment > 0 ? alignment : 256) {
    if (align_ & (align_ - 1)) throw std::runtime_error("AutoGrowthAllocator: alignment must be a power of two");
  }
  ~AutoGrowthAllocator() {
    for (auto& c : chunks_) backend_free(c.second.base);
  }

  void* alloc(int64_t size, cudaStream_t stream) {
    std::lock_guard<std::mutex> g(mu_);
    if (size <= 0) size = 1;
    size = (size + align_ - 1) & ~(align_ - 1);
    reclaim_deferred();
    Block* b = take_free(size);
    if (!b) {
      b = grow(size);
      if (!b) {                                  // backend is out of memory: give idle chunks back and try once more
        wait_deferred();
        release_idle_locked();
        b = grow(size);
      }
      if (!b) throw std::bad_alloc();
    }
    if (b->size - size >= min_split()) split(b, size);
    b->free = false;
    b->stream = stream;
    live_[b->ptr] = b;
    st_.allocated += b->size;
    st_.allocated_peak = std::max(st_.allocated_peak, st_.allocated);
    ++st_.num_allocs;
    return b->ptr;
  }

  void free(void* ptr, cudaStream_t stream) {
    if (!ptr) return;
    std::lock_guard<std::mutex> g(mu_);
    auto it = live_.find(ptr);
    if (it == live_.end()) throw std::runtime_error("AutoGrowthAllocator: free of an address this allocator does not own");
    Block* b = it->second;
    live_.erase(it);
    st_.allocated -= b->size;
    ++st_.num_frees;
    if (cuda_ && stream != b->stream) {
      // freed from another stream than it was allocated on: work queued there may still use it - park it behind an event
      cudaEvent_t ev;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess && cudaEventRecord(ev, stream) == cudaSuccess) {
        deferred_.push_back({b, ev});
        ++st_.deferred;
        return;
      }
      cudaGetLastError();
      cudaStreamSynchronize(stream);
    }
    release_block(b);
  }

  // every chunk that is one free block goes back to the backend; returns the bytes released
  int64_t release_idle() {
    std::lock_guard<std::mutex> g(mu_);
    wait_deferred();
    return release_idle_locked();
  }

  Stats stats() {
    std::lock_guard<std::mutex> g(mu_);
    Stats s = st_;
    s.num_chunks = (int64_t)chunks_.size();
    return s;
  }
  void reset_peak() {
    std::lock_guard<std::mutex> g(mu_);
    st_.allocated_peak = st_.allocated;
    st_.reserved_peak = st_.reserved;
  }
  int64_t largest_free_block() {
    std::lock_guard<std::mutex> g(mu_);
    return free_.empty() ? 0 : free_.rbegin()->first;
  }
  bool owns(void* p) {
    std::lock_guard<std::mutex> g(mu_);
    return live_.count(p) != 0;
  }

 private:
  struct Block {
    char* ptr = nullptr;
    int64_t size = 0;
    bool free = true;
    cudaStream_t stream = nullptr;
    Block* prev = nullptr;      // neighbours inside the chunk (address order)
    Block* next = nullptr;
    char* chunk = nullptr;
    std::multimap<int64_t, Block*>::iterator it;
  };
  struct Chunk { char* base; int64_t size; };
  struct Deferred { Block* b; cudaEvent_t ev; };

  int64_t min_split() const { return std::max<int64_t>(align_, 512); }

  void* backend_alloc(int64_t size) {
    void* p = nullptr;
    if (cuda_) {
      int prev = -1;
      cudaGetDevice(&prev);
      if (prev != device_) cudaSetDevice(device_);
      cudaError_t e = cudaMalloc(&p, (size_t)size);
      if (prev != device_ && prev >= 0) cudaSetDevice(prev);
      if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    } else {
      if (host_limit_ > 0 && st_.reserved + size > host_limit_) return nullptr;
      p = std::aligned_alloc((size_t)std::max<int64_t>(align_, 64), (size_t)size);
    }
    if (p) ++st_.num_backend_allocs;
    return p;
  }
  void backend_free(char* p) {
    if (cuda_) cudaFree(p); else std::free(p);
    ++st_.num_backend_frees;
  }

  Block* take_free(int64_t size) {
    auto it = free_.lower_bound(size);
    if (it == free_.end()) return nullptr;
    Block* b = it->second;
    free_.erase(it);
    return b;
  }
  Block* grow(int64_t size) {
    int64_t csize = std::max(size, chunk_bytes_);
    char* base = (char*)backend_alloc(csize);
    if (!base && csize > size) {                 // a full chunk does not fit any more: take exactly what is needed
      csize = size;
      base = (char*)backend_alloc(csize);
    }
    if (!base) return nullptr;
    Block* b = new Block();
    b->ptr = base;
    b->size = csize;
    b->chunk = base;
    chunks_[base] = Chunk{base, csize};
    st_.reserved += csize;
    st_.reserved_peak = std::max(st_.reserved_peak, st_.reserved);
    return b;
  }
  void split(Block* b, int64_t size) {
    Block* tail = new Block();
    tail->ptr = b->ptr + size;
    tail->size = b->size - size;
    tail->chunk = b->chunk;
    tail->prev = b;
    tail->next = b->next;
    if (b->next) b->next->prev = tail;
    b->next = tail;
    b->size = size;
    tail->it = free_.emplace(tail->size, tail);
    ++st_.num_splits;
  }
  void release_block(Block* b) {
    b->free = true;
    if (b->next && b->next->free) {
      Block* n = b->next;
      free_.erase(n->it);
      b->size += n->size;
      b->next = n->next;
      if (n->next) n->next->prev = b;
      delete n;
      ++st_.num_merges;
    }
    if (b->prev && b->prev->free) {
      Block* p = b->prev;
      free_.erase(p->it);
      p->size += b->size;
      p->next = b->next;
      if (b->next) b->next->prev = p;
      delete b;
      b = p;
      ++st_.num_merges;
    }
    b->it = free_.emplace(b->size, b);
  }
  void reclaim_deferred() {
    while (!deferred_.empty()) {
      Deferred& d = deferred_.front();
      if (cudaEventQuery(d.ev) != cudaSuccess) { cudaGetLastError(); break; }
      cudaEventDestroy(d.ev);
      release_block(d.b);
      deferred_.pop_front();
    }
  }
  void wait_deferred() {
    while (!deferred_.empty()) {
      Deferred& d = deferred_.front();
      cudaEventSynchronize(d.ev);
      cudaEventDestroy(d.ev);
      release_block(d.b);
      deferred_.pop_front();
    }
  }
  int64_t release_idle_locked() {
    int64_t released = 0;
    for (auto it = free_.begin(); it != free_.end();) {
      Block* b = it->second;
      auto c = chunks_.find(b->ptr);
      if (b->prev == nullptr && b->next == nullptr && c != chunks_.end() && c->second.size == b->size) {     // the chunk is one free block
        it = free_.erase(it);
        backend_free(c->second.base);
        st_.reserved -= c->second.size;
        released += c->second.size;
        chunks_.erase(c);
        delete b;
      } else {
        ++it;
      }
    }
    return released;
  }

 public:
  void set_host_limit(int64_t bytes) { host_limit_ = bytes; }

 private:
  const bool cuda_;
  const int device_;
  const int64_t chunk_bytes_, align_;
  int64_t host_limit_ = 0;
  std::mutex mu_;
  std::multimap<int64_t, Block*> free_;
  std::map<void*, Block*> live_;
  std::map<char*, Chunk> chunks_;
  std::deque<Deferred> deferred_;
  Stats st_;
};

// ------------------------------------------------------------------------------------------------ process-wide CUDA instances
static std::mutex g_mu;
static std::map<int, std::unique_ptr<AutoGrowthAllocator>> g_cuda;
static int64_t g_chunk_bytes = (int64_t)256 << 20;

static AutoGrowthAllocator& cuda_instance(int device) {
  std::lock_guard<std::mutex> g(g_mu);
  auto& p = g_cuda[device];
  if (!p) {
    if (const char* e = std::getenv("B200_ALLOCATOR_CHUNK_MB")) g_chunk_bytes = (int64_t)std::atoll(e) << 20;
    p = std::make_unique<AutoGrowthAllocator>(true, device, g_chunk_bytes, 512);
  }
  return *p;
}

static py::dict stats_dict(const AutoGrowthAllocator::Stats& s) {
  py::dict d;
  d["allocated"] = s.allocated; d["allocated_peak"] = s.allocated_peak; d["reserved"] = s.reserved; d["reserved_peak"] = s.reserved_peak;
  d["num_allocs"] = s.num_allocs; d["num_frees"] = s.num_frees; d["num_chunks"] = s.num_chunks; d["num_backend_allocs"] = s.num_backend_allocs;
  d["num_backend_frees"] = s.num_backend_frees; d["num_splits"] = s.num_splits; d["num_merges"] = s.num_merges; d["deferred_frees"] = s.deferred;
  return d;
}

void bind_allocator(py::module_& m) {
  py::class_<AutoGrowthAllocator>(m, "AutoGrowthAllocator")
      .def(py::init([](const std::string& kind, int64_t chunk_bytes, int64_t alignment, int device) {
             if (kind != "host" && kind != "cuda") throw std::runtime_error("AutoGrowthAllocator: kind is 'host' or 'cuda'");
             return new AutoGrowthAllocator(kind == "cuda", device, chunk_bytes, alignment);
           }),
           py::arg("kind") = "host", py::arg("chunk_bytes") = (int64_t)1 << 20, py::arg("alignment") = 256, py::arg("device") = 0)
      .def("alloc", [](AutoGrowthAllocator& a, int64_t size, uint64_t stream) { return (uint64_t)(uintptr_t)a.alloc(size, (cudaStream_t)(uintptr_t)stream); }, py::arg("size"),
           py::arg("stream") = 0)
      .def("free", [](AutoGrowthAllocator& a, uint64_t ptr, uint64_t stream) { a.free((void*)(uintptr_t)ptr, (cudaStream_t)(uintptr_t)stream); }, py::arg("ptr"), py::arg("stream") = 0)
      .def("release_idle", &AutoGrowthAllocator::release_idle)
      .def("reset_peak", &AutoGrowthAllocator::reset_peak)
      .def("largest_free_block", &AutoGrowthAllocator::largest_free_block)
      .def("set_host_limit", &AutoGrowthAllocator::set_host_limit)
      .def("stats", [](AutoGrowthAllocator& a) { return stats_dict(a.stats()); });
  m.def("cuda_allocator_stats", [](int device) { return stats_dict(cuda_instance(device).stats()); }, py::arg("device") = 0);
  m.def("cuda_allocator_release_idle", [](int device) { return cuda_instance(device).release_idle(); }, py::arg("device") = 0);
  m.def("cuda_allocator_reset_peak", [](int device) { cuda_instance(device).reset_peak(); }, py::arg("device") = 0);
}

}  // namespace runtime
}  // namespace b200

// torch.cuda.memory.CUDAPluggableAllocator entry points
extern "C" {
__attribute__((visibility("default"))) void* b200_cuda_malloc(ssize_t size, int device, cudaStream_t stream) {
  try {
    return b200::runtime::cuda_instance(device).alloc((int64_t)size, stream);
  } catch (const std::bad_alloc&) {
    return nullptr;
  }
}
__attribute__((visibility("default"))) void b200_cuda_free(void* ptr, ssize_t /*size*/, int device, cudaStream_t stream) {
  b200::runtime::cuda_instance(device).free(ptr, stream);
}
}
