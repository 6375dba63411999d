// Native data-loader core: pinned staging ring + GIL-free multi-threaded collation + async H2D on a side stream.
// Parity (role): paddle/fluid/operators/reader/buffered_reader.cc (double-buffered pinned->device prefetch),
// paddle/fluid/framework/data_feed.cc, python/paddle/io/dataloader/dataloader_iter.py (_thread_loop / pin memory).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <thread>
#include <vector>

#include "runtime.h"

namespace b200 {
namespace runtime {

class ThreadPool {
 public:
  explicit ThreadPool(int n) : stop_(false), pending_(0) {
    if (n < 1) n = 1;
    for (int i = 0; i < n; ++i)
      workers_.emplace_back([this] {
        for (;;) {
          std::function<void()> job;
          {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
            if (stop_ && jobs_.empty()) return;
            job = std::move(jobs_.front());
            jobs_.pop();
          }
          job();
          {
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_cv_.notify_all();
          }
        }
      });
  }
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void submit(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      ++pending_;
      jobs_.push(std::move(f));
    }
    cv_.notify_one();
  }
  void wait_all() {
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
  }
  int size() const { return (int)workers_.size(); }

 private:
  std::vector<std::thread> workers_;
  std::queue<std::function<void()>> jobs_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  bool stop_;
  int pending_;
};

// Ring of pinned host slabs. A slot is handed to the collator, copied to the device on `stream`, and becomes reusable once
// its event has completed.
class PinnedRing {
 public:
  PinnedRing(int slots, int64_t bytes_per_slot, int threads) : bytes_(bytes_per_slot), pool_(threads) {
    const bool cuda = at::cuda::is_available();
    for (int i = 0; i < slots; ++i) {
      void* p = nullptr;
      if (cuda) {
        if (cudaHostAlloc(&p, bytes_, cudaHostAllocPortable) != cudaSuccess) throw std::runtime_error("PinnedRing: cudaHostAlloc failed");
      } else {
        p = ::operator new((size_t)bytes_);
      }
      slabs_.push_back(p);
      events_.push_back(nullptr);
      busy_.push_back(false);
    }
    pinned_ = cuda;
  }
  ~PinnedRing() {
    for (size_t i = 0; i < slabs_.size(); ++i) {
      if (events_[i]) cudaEventDestroy(events_[i]);
      if (pinned_) cudaFreeHost(slabs_[i]); else ::operator delete(slabs_[i]);
    }
  }
  int num_slots() const { return (int)slabs_.size(); }
  int64_t slot_bytes() const { return bytes_; }

  // Blocks until slot `i` is no longer in flight.
  void wait_slot(int i) {
    if (events_.at(i) && busy_[i]) {
      pybind11::gil_scoped_release nogil;
      cudaEventSynchronize(events_[i]);
    }
    busy_[i] = false;
  }

  // Stack `samples` (CPU tensors of identical shape/dtype) into slot `i` at byte offset `offset`; returns a [N, ...] view.
  torch::Tensor collate(int i, int64_t offset, const std::vector<torch::Tensor>& samples) {
    if (samples.empty()) throw std::runtime_error("collate: empty batch");
    const auto& s0 = samples[0];
    const int64_t each = s0.numel() * s0.element_size();
    const int64_t total = each * (int64_t)samples.size();
    if (offset + total > bytes_) throw std::runtime_error("collate: batch does not fit in the pinned slot");
    char* dst = static_cast<char*>(slabs_.at(i)) + offset;
    std::vector<const void*> srcs(samples.size());
    std::vector<torch::Tensor> keep(samples.size());
    for (size_t k = 0; k < samples.size(); ++k) {
      keep[k] = samples[k].is_contiguous() ? samples[k] : samples[k].contiguous();
      if (keep[k].numel() != s0.numel() || keep[k].scalar_type() != s0.scalar_type()) throw std::runtime_error("collate: ragged batch");
      srcs[k] = keep[k].data_ptr();
    }
    {
      pybind11::gil_scoped_release nogil;
      const int n = (int)samples.size();
      const int chunks = std::min(n, pool_.size() * 2);
      for (int c = 0; c < chunks; ++c) {
        const int lo = (int)((int64_t)n * c / chunks), hi = (int)((int64_t)n * (c + 1) / chunks);
        pool_.submit([=, &srcs] {
          for (int k = lo; k < hi; ++k) std::memcpy(dst + (int64_t)k * each, srcs[k], (size_t)each);
        });
      }
      pool_.wait_all();
    }
    std::vector<int64_t> sizes{(int64_t)samples.size()};
    for (auto d : s0.sizes()) sizes.push_back(d);
    return torch::from_blob(dst, sizes, torch::TensorOptions().dtype(s0.scalar_type()));
  }

  // Async copy of a host view living in slot `i` to `dst` (device) on the current stream of dst's device.
  void h2d(int i, const torch::Tensor& host_view, torch::Tensor dst) {
    if (!dst.is_cuda()) { dst.copy_(host_view); return; }
    c10::cuda::CUDAGuard g(dst.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    const int64_t n = host_view.numel() * host_view.element_size();
    if (cudaMemcpyAsync(dst.data_ptr(), host_view.data_ptr(), (size_t)n, cudaMemcpyHostToDevice, stream.stream()) != cudaSuccess)
      throw std::runtime_error("PinnedRing: cudaMemcpyAsync failed");
    if (!events_[i]) cudaEventCreateWithFlags(&events_[i], cudaEventDisableTiming);
    cudaEventRecord(events_[i], stream.stream());
    busy_[i] = true;
  }

 private:
  int64_t bytes_;
  bool pinned_ = false;
  std::vector<void*> slabs_;
  std::vector<cudaEvent_t> events_;
  std::vector<bool> busy_;
  ThreadPool pool_;
};

void bind_loader(pybind11::module_& m) {
  pybind11::class_<PinnedRing, std::shared_ptr<PinnedRing>>(m, "PinnedRing")
      .def(pybind11::init<int, int64_t, int>())
      .def("num_slots", &PinnedRing::num_slots)
      .def("slot_bytes", &PinnedRing::slot_bytes)
      .def("wait_slot", &PinnedRing::wait_slot)
      .def("collate", &PinnedRing::collate)
      .def("h2d", &PinnedRing::h2d);
}

}  // namespace runtime
}  // namespace b200
