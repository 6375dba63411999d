// Symmetric peer heap: one cudaMalloc'ed slab per rank, IPC-mapped into every peer of the node so kernels can issue
// ld/st.global straight to peer HBM over NVLink/NVSwitch.  The reference has no direct equivalent (NCCL rings via
// ProcessGroupNCCL, paddle/fluid/distributed/collective/process_group_nccl.cc); closest is
// paddle/phi/core/memory/allocation/cuda_ipc_allocator.cc.
//
// Layout of the slab: [0, signal_bytes) = signal pad (uint32 flags, zero-initialised) | bump-allocated data region.
// Allocation is collective and deterministic (same sequence of alloc() calls on every rank => same offsets), which is
// what makes the heap "symmetric": peer address = peer_base + my_offset.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "runtime.h"
#include "../include/b200_comm.h"
#include "../include/b200_ops.h"
#include "offset_allocator.h"

namespace b200 {
namespace runtime {

namespace {
void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("symm_heap: ") + what + ": " + cudaGetErrorString(e));
}
}  // namespace

class SymmHeap {
 public:
  SymmHeap(int64_t bytes, int64_t signal_bytes, int device) : bytes_(bytes), signal_bytes_(signal_bytes), device_(device) {
    c10::cuda::CUDAGuard g(device_);
    ck(cudaMalloc(&base_, bytes_), "cudaMalloc");
    ck(cudaMemset(base_, 0, bytes_), "cudaMemset");
    ck(cudaDeviceSynchronize(), "sync");
    alloc_ = std::make_unique<BestFitAllocator>((signal_bytes_ + 1023) / 1024 * 1024, bytes_, 256);
  }
  ~SymmHeap() {
    for (size_t i = 0; i < peers_.size(); ++i)
      if (peers_[i] && (int)i != rank_) cudaIpcCloseMemHandle(peers_[i]);
    if (peer_table_dev_) cudaFree(peer_table_dev_);
    if (wait_stats_) cudaFree(wait_stats_);
    if (base_) cudaFree(base_);
  }

  pybind11::bytes ipc_handle() {
    cudaIpcMemHandle_t h;
    ck(cudaIpcGetMemHandle(&h, base_), "cudaIpcGetMemHandle");
    return pybind11::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
  }

  // handles[r] = ipc handle bytes of rank r (this rank's own entry is ignored)
  void open_peers(const std::vector<std::string>& handles, int rank) {
    c10::cuda::CUDAGuard g(device_);
    rank_ = rank;
    world_ = (int)handles.size();
    peers_.assign(world_, nullptr);
    for (int r = 0; r < world_; ++r) {
      if (r == rank) { peers_[r] = base_; continue; }
      if (handles[r].size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("symm_heap: bad ipc handle size");
      cudaIpcMemHandle_t h;
      memcpy(&h, handles[r].data(), sizeof(h));
      ck(cudaIpcOpenMemHandle(&peers_[r], h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    }
    std::vector<uint64_t> tab(world_);
    for (int r = 0; r < world_; ++r) tab[r] = reinterpret_cast<uint64_t>(peers_[r]);
    ck(cudaMalloc(&peer_table_dev_, sizeof(uint64_t) * world_), "cudaMalloc table");
    ck(cudaMemcpy(peer_table_dev_, tab.data(), sizeof(uint64_t) * world_, cudaMemcpyHostToDevice), "memcpy table");
  }

  // Single-process mode (world 1, or tests): the "peer" table just points at ourselves.
  void open_self() { open_peers(std::vector<std::string>(1), 0); }

  // Best-fit with coalescing; with no frees in between this degenerates to the bump order, and identical call sequences
  // give identical offsets on every rank (the heap stays symmetric).
  int64_t alloc(int64_t nbytes, int64_t align) {
    try {
      return alloc_->alloc(nbytes, align);
    } catch (const std::exception& e) {
      throw std::runtime_error(std::string("symm_heap: out of symmetric memory: ") + e.what());
    }
  }
  void free(int64_t off) { alloc_->free(off); }
  void reset_cursor(int64_t to) { alloc_->release_from(to); }
  // size() - cursor() = largest block that can still be allocated
  int64_t cursor() const { return bytes_ - alloc_->largest_free(); }
  pybind11::dict stats() const {
    pybind11::dict d;
    d["capacity"] = alloc_->capacity();
    d["allocated"] = alloc_->in_use();
    d["peak_allocated"] = alloc_->peak();
    d["largest_free_block"] = alloc_->largest_free();
    d["free_blocks"] = alloc_->num_free_blocks();
    d["live_blocks"] = alloc_->num_live();
    d["num_allocs"] = alloc_->num_allocs();
    d["num_frees"] = alloc_->num_frees();
    return d;
  }

  torch::Tensor tensor(int64_t offset, std::vector<int64_t> sizes, at::ScalarType dtype, int peer) {
    void* b = peer < 0 ? base_ : peers_.at(peer);
    auto opts = torch::TensorOptions().dtype(dtype).device(torch::kCUDA, device_);
    return torch::from_blob(static_cast<char*>(b) + offset, sizes, opts);
  }
  int64_t base_ptr() const { return reinterpret_cast<int64_t>(base_); }
  int64_t peer_ptr(int r) const { return reinterpret_cast<int64_t>(peers_.at(r)); }
  int64_t peer_table_ptr() const { return reinterpret_cast<int64_t>(peer_table_dev_); }
  int64_t size() const { return bytes_; }
  int64_t signal_bytes() const { return signal_bytes_; }
  int rank() const { return rank_; }
  int world() const { return world_; }

  // ---- peer-memory collectives on symmetric buffers (csrc/comm/p2p_collectives.cu) ----
  static int dcode(at::ScalarType st) {
    switch (st) {
      case at::kFloat: return 0;
      case at::kHalf: return 1;
      case at::kBFloat16: return 2;
      default: throw std::runtime_error("symm_heap: unsupported dtype for p2p collective");
    }
  }
  std::vector<int64_t> bases() const {
    std::vector<int64_t> b(world_);
    for (int r = 0; r < world_; ++r) b[r] = reinterpret_cast<int64_t>(peers_[r]);
    return b;
  }
  uint32_t* counter() {
    if (!counter_) {
      c10::cuda::CUDAGuard g(device_);
      ck(cudaMalloc(&counter_, 64), "cudaMalloc counter");
      ck(cudaMemset(counter_, 0, 64), "memset counter");
    }
    return static_cast<uint32_t*>(counter_);
  }
  void finish() {
    const char* e = b200::take_last_error();
    if (e[0]) throw std::runtime_error(std::string("paddle_b200 comm kernel error: ") + e);
  }
  void allreduce(int64_t off, int64_t n, at::ScalarType dt, int64_t epoch) {
    c10::cuda::CUDAGuard g(device_);
    auto b = bases();
    comm::p2p_allreduce(b.data(), off, n, dcode(dt), rank_, world_, (uint32_t)epoch, counter(), at::cuda::getCurrentCUDAStream().stream());
    finish();
  }
  void reduce_scatter(int64_t off, torch::Tensor out, int64_t n, int64_t epoch) {
    c10::cuda::CUDAGuard g(device_);
    auto b = bases();
    comm::p2p_reduce_scatter(b.data(), off, out.data_ptr(), n, dcode(out.scalar_type()), rank_, world_, (uint32_t)epoch, counter(),
                             at::cuda::getCurrentCUDAStream().stream());
    finish();
  }
  void reduce_slots(int64_t off, torch::Tensor out, int64_t n, int64_t epoch) {
    c10::cuda::CUDAGuard g(device_);
    auto b = bases();
    comm::p2p_reduce_slots(b.data(), off, out.data_ptr(), n, dcode(out.scalar_type()), rank_, world_, (uint32_t)epoch, counter(),
                           at::cuda::getCurrentCUDAStream().stream());
    finish();
  }
  void allgather(int64_t off, int64_t chunk_bytes, int64_t epoch) {
    c10::cuda::CUDAGuard g(device_);
    auto b = bases();
    comm::p2p_allgather(b.data(), off, chunk_bytes, rank_, world_, (uint32_t)epoch, counter(), at::cuda::getCurrentCUDAStream().stream());
    finish();
  }
  // src: local rows [*, H]; gather: optional int64 slot->row map; meta_off: symmetric int64[world] send counts; recv_off: receive buffer
  void a2av(torch::Tensor src, c10::optional<torch::Tensor> gather, int64_t meta_off, int64_t recv_off, int64_t rows_hint, int64_t epoch) {
    c10::cuda::CUDAGuard g(device_);
    auto b = bases();
    const int64_t row_bytes = src.size(-1) * src.element_size();
    comm::p2p_a2av(b.data(), src.data_ptr(), gather.has_value() && gather->defined() ? gather->data_ptr<int64_t>() : nullptr, meta_off, recv_off,
                   row_bytes, rows_hint, rank_, world_, (uint32_t)epoch, counter(), at::cuda::getCurrentCUDAStream().stream());
    finish();
  }
  void gather_pull(torch::Tensor dst, int64_t src_off, int64_t chunk_bytes, int64_t epoch) {
    c10::cuda::CUDAGuard g(device_);
    auto b = bases();
    comm::p2p_gather_pull(b.data(), src_off, dst.data_ptr(), chunk_bytes, rank_, world_, (uint32_t)epoch, counter(),
                          at::cuda::getCurrentCUDAStream().stream());
    finish();
  }
  // ---- mailbox primitives (pipeline p2p): flags live in the symmetric data region at byte offset `flag_off` ----
  void signal_flag(int peer, int64_t flag_off, int64_t value) {
    c10::cuda::CUDAGuard g(device_);
    comm::p2p_signal_flag(static_cast<char*>(peers_.at(peer)) + flag_off, (uint32_t)value, at::cuda::getCurrentCUDAStream().stream());
    finish();
  }
  void wait_flag(int64_t flag_off, int64_t value, double timeout_s) {
    c10::cuda::CUDAGuard g(device_);
    comm::p2p_wait_flag(static_cast<char*>(base_) + flag_off, (uint32_t)value, wait_stats_dev(), timeout_s,
                        at::cuda::getCurrentCUDAStream().stream());
    finish();
  }
  void* wait_stats_dev() {
    if (!wait_stats_) {
      c10::cuda::CUDAGuard g(device_);
      ck(cudaMalloc(&wait_stats_, 16), "cudaMalloc wait stats");
      ck(cudaMemset(wait_stats_, 0, 16), "memset wait stats");
    }
    return wait_stats_;
  }
  // (spun nanoseconds, number of waits) accumulated by wait_flag since the last reset; synchronises the device
  std::vector<int64_t> wait_stats(bool reset) {
    c10::cuda::CUDAGuard g(device_);
    unsigned long long h[2] = {0, 0};
    ck(cudaDeviceSynchronize(), "sync");
    ck(cudaMemcpy(h, wait_stats_dev(), 16, cudaMemcpyDeviceToHost), "memcpy wait stats");
    if (reset) ck(cudaMemset(wait_stats_, 0, 16), "memset wait stats");
    return {(int64_t)h[0], (int64_t)h[1]};
  }
  void alltoall(int64_t off_send, int64_t off_recv, int64_t chunk_bytes, int64_t epoch) {
    c10::cuda::CUDAGuard g(device_);
    auto b = bases();
    comm::p2p_alltoall(b.data(), off_send, off_recv, chunk_bytes, rank_, world_, (uint32_t)epoch, counter(),
                       at::cuda::getCurrentCUDAStream().stream());
    finish();
  }

 private:
  int64_t bytes_, signal_bytes_;
  int device_;
  void* base_ = nullptr;
  std::unique_ptr<BestFitAllocator> alloc_;
  int rank_ = 0, world_ = 1;
  std::vector<void*> peers_;
  void* peer_table_dev_ = nullptr;
  void* counter_ = nullptr;
  void* wait_stats_ = nullptr;
};

namespace {
int nvls_dcode(at::ScalarType st) {
  switch (st) {
    case at::kFloat: return 0;
    case at::kBFloat16: return 1;
    case at::kHalf: return 2;
    default: throw std::runtime_error("nvls: unsupported dtype (float32 / bfloat16 / float16)");
  }
}
uint32_t* counter_of(const at::Tensor& c) {
  if (!c.is_cuda() || c.scalar_type() != at::kInt || c.numel() < 1) throw std::runtime_error("nvls: counter must be a CUDA int32 tensor");
  return reinterpret_cast<uint32_t*>(c.data_ptr());
}
}  // namespace

void bind_symm(pybind11::module_& m) {
  // NVLS collectives over a torch.distributed._symmetric_memory buffer (parallel/nvls.py owns the handles)
  m.def("nvls_allreduce", [](const std::vector<int64_t>& pads, int64_t pad_off, int64_t mc, int64_t local, int64_t off, int64_t n, at::ScalarType dt, int rank,
                             int64_t epoch, const at::Tensor& counter) {
    comm::nvls_allreduce(pads.data(), pad_off, mc, local, off, n, nvls_dcode(dt), rank, (int)pads.size(), (uint32_t)epoch, counter_of(counter),
                         at::cuda::getCurrentCUDAStream().stream());
  });
  m.def("nvls_reduce_scatter", [](const std::vector<int64_t>& pads, int64_t pad_off, int64_t mc, int64_t local, int64_t off, at::Tensor out, int64_t n, int rank,
                                  int64_t epoch, const at::Tensor& counter) {
    comm::nvls_reduce_scatter(pads.data(), pad_off, mc, local, off, out.data_ptr(), n, nvls_dcode(out.scalar_type()), rank, (int)pads.size(), (uint32_t)epoch,
                              counter_of(counter), at::cuda::getCurrentCUDAStream().stream());
  });
  m.def("nvls_allgather", [](const std::vector<int64_t>& pads, int64_t pad_off, int64_t mc, int64_t local, int64_t off, const at::Tensor& src, int rank, int64_t epoch,
                             const at::Tensor& counter) {
    comm::nvls_allgather(pads.data(), pad_off, mc, local, off, src.data_ptr(), (int64_t)src.numel() * src.element_size(), rank, (int)pads.size(), (uint32_t)epoch,
                         counter_of(counter), at::cuda::getCurrentCUDAStream().stream());
  });
  pybind11::class_<BestFitAllocator>(m, "BestFitAllocator")
      .def(pybind11::init<int64_t, int64_t, int64_t>(), pybind11::arg("begin"), pybind11::arg("end"), pybind11::arg("min_align") = 256)
      .def("alloc", &BestFitAllocator::alloc, pybind11::arg("nbytes"), pybind11::arg("align") = 256)
      .def("free", &BestFitAllocator::free)
      .def("release_from", &BestFitAllocator::release_from)
      .def("block_size", &BestFitAllocator::block_size)
      .def("largest_free", &BestFitAllocator::largest_free)
      .def("free_bytes", &BestFitAllocator::free_bytes)
      .def("in_use", &BestFitAllocator::in_use)
      .def("peak", &BestFitAllocator::peak)
      .def("reset_peak", &BestFitAllocator::reset_peak)
      .def("num_live", &BestFitAllocator::num_live)
      .def("num_free_blocks", &BestFitAllocator::num_free_blocks);
  pybind11::class_<SymmHeap, std::shared_ptr<SymmHeap>>(m, "SymmHeap")
      .def(pybind11::init<int64_t, int64_t, int>())
      .def("ipc_handle", &SymmHeap::ipc_handle)
      .def("open_peers", &SymmHeap::open_peers)
      .def("open_self", &SymmHeap::open_self)
      .def("alloc", &SymmHeap::alloc)
      .def("free", &SymmHeap::free)
      .def("stats", &SymmHeap::stats)
      .def("reset_cursor", &SymmHeap::reset_cursor)
      .def("cursor", &SymmHeap::cursor)
      .def("tensor", &SymmHeap::tensor)
      .def("base_ptr", &SymmHeap::base_ptr)
      .def("peer_ptr", &SymmHeap::peer_ptr)
      .def("peer_table_ptr", &SymmHeap::peer_table_ptr)
      .def("size", &SymmHeap::size)
      .def("signal_bytes", &SymmHeap::signal_bytes)
      .def("rank", &SymmHeap::rank)
      .def("world", &SymmHeap::world)
      .def("allreduce", &SymmHeap::allreduce)
      .def("reduce_scatter", &SymmHeap::reduce_scatter)
      .def("reduce_slots", &SymmHeap::reduce_slots)
      .def("allgather", &SymmHeap::allgather)
      .def("a2av", &SymmHeap::a2av)
      .def("alltoall", &SymmHeap::alltoall)
      .def("gather_pull", &SymmHeap::gather_pull)
      .def("signal_flag", &SymmHeap::signal_flag)
      .def("wait_flag", &SymmHeap::wait_flag, pybind11::arg("flag_off"), pybind11::arg("value"), pybind11::arg("timeout_s") = 300.0)
      .def("wait_stats", &SymmHeap::wait_stats, pybind11::arg("reset") = true);
}

}  // namespace runtime
}  // namespace b200
