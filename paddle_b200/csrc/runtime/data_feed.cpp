// Multi-slot text data feed held in host memory.  Role of the reference's MultiSlotInMemoryDataFeed
// (paddle/fluid/framework/data_feed.cc) + DatasetImpl::LoadIntoMemory / LocalShuffle (paddle/fluid/framework/data_set.cc).
//
// Line format (what MultiSlotDataGenerator emits): for every slot, in order, `<n> v_1 ... v_n`.
// Storage is columnar: per slot one flat value array plus per-record offsets, so a batch is a contiguous gather per slot and
// comes back as (values, lod) without any per-record Python object.  Files are parsed by a pool of threads with the GIL released.
#include <torch/extension.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "runtime.h"

namespace b200 {
namespace runtime {

namespace {

struct SlotColumn {
  bool is_float = false;
  std::vector<int64_t> ivals;
  std::vector<float> fvals;
  std::vector<int64_t> offsets{0};   // record r spans [offsets[r], offsets[r+1])
  int64_t count() const { return static_cast<int64_t>(offsets.size()) - 1; }
};

struct Shard {
  std::vector<SlotColumn> cols;
  int64_t records = 0;
  std::string error;
};

inline const char* skip_ws(const char* p, const char* e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
  return p;
}

bool parse_line(const char* p, const char* e, std::vector<SlotColumn>& cols, std::string& err) {
  for (size_t s = 0; s < cols.size(); ++s) {
    p = skip_ws(p, e);
    if (p >= e) {
      err = "line ends before slot " + std::to_string(s);
      return false;
    }
    char* q = nullptr;
    long long n = std::strtoll(p, &q, 10);
    if (q == p || n < 0) {
      err = "bad value count for slot " + std::to_string(s);
      return false;
    }
    p = q;
    SlotColumn& c = cols[s];
    for (long long i = 0; i < n; ++i) {
      p = skip_ws(p, e);
      if (p >= e) {
        err = "slot " + std::to_string(s) + " announces " + std::to_string(n) + " values, line is shorter";
        return false;
      }
      if (c.is_float) {
        float v = std::strtof(p, &q);
        if (q == p) { err = "bad float in slot " + std::to_string(s); return false; }
        c.fvals.push_back(v);
      } else {
        unsigned long long v = std::strtoull(p, &q, 10);
        if (q == p) { err = "bad integer in slot " + std::to_string(s); return false; }
        c.ivals.push_back(static_cast<int64_t>(v));
      }
      p = q;
    }
    c.offsets.push_back(static_cast<int64_t>(c.is_float ? c.fvals.size() : c.ivals.size()));
  }
  return true;
}

}  // namespace

class MultiSlotFeed {
 public:
  MultiSlotFeed(std::vector<std::string> slot_types, int threads) : threads_(std::max(1, threads)) {
    for (auto& t : slot_types) {
      SlotColumn c;
      if (t == "float" || t == "float32") c.is_float = true;
      else if (t != "uint64" && t != "int64") throw std::runtime_error("MultiSlotFeed: slot type must be uint64/int64/float, got " + t);
      cols_.push_back(std::move(c));
    }
    if (cols_.empty()) throw std::runtime_error("MultiSlotFeed: no slots");
  }

  // Parse `files` (appending). Returns the number of records added.
  int64_t load(const std::vector<std::string>& files) {
    std::vector<Shard> shards(files.size());
    {
      pybind11::gil_scoped_release nogil;
      std::atomic<size_t> next{0};
      auto work = [&]() {
        for (size_t i = next++; i < files.size(); i = next++) parse_file(files[i], shards[i]);
      };
      std::vector<std::thread> pool;
      const int n = std::min<int>(threads_, static_cast<int>(files.size()));
      for (int t = 1; t < n; ++t) pool.emplace_back(work);
      work();
      for (auto& t : pool) t.join();
    }
    int64_t added = 0;
    for (size_t i = 0; i < shards.size(); ++i) {
      if (!shards[i].error.empty()) throw std::runtime_error("MultiSlotFeed: " + files[i] + ": " + shards[i].error);
      for (size_t s = 0; s < cols_.size(); ++s) append(cols_[s], shards[i].cols[s]);
      added += shards[i].records;
    }
    const int64_t base = static_cast<int64_t>(order_.size());
    for (int64_t r = 0; r < added; ++r) order_.push_back(base + r);
    return added;
  }

  int64_t load_lines(const std::vector<std::string>& lines) {
    Shard sh;
    sh.cols = blank();
    for (auto& l : lines) {
      if (l.find_first_not_of(" \t\r\n") == std::string::npos) continue;
      std::string err;
      if (!parse_line(l.data(), l.data() + l.size(), sh.cols, err)) throw std::runtime_error("MultiSlotFeed: " + err);
      ++sh.records;
    }
    for (size_t s = 0; s < cols_.size(); ++s) append(cols_[s], sh.cols[s]);
    const int64_t base = static_cast<int64_t>(order_.size());
    for (int64_t r = 0; r < sh.records; ++r) order_.push_back(base + r);
    return sh.records;
  }

  int64_t size() const { return static_cast<int64_t>(order_.size()); }

  void shuffle(uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::shuffle(order_.begin(), order_.end(), rng);
  }

  // Keep only the records whose (position in the current order) % world == rank: the local part of a global shuffle that every
  // trainer performs with the same seed over the same file list.
  void keep_partition(int64_t rank, int64_t world) {
    std::vector<int64_t> kept;
    for (size_t i = 0; i < order_.size(); ++i)
      if (static_cast<int64_t>(i % world) == rank) kept.push_back(order_[i]);
    order_.swap(kept);
  }

  void clear() {
    cols_ = blank();
    order_.clear();
  }

  // Records [start, start+n) of the current order: per slot (values [total], lod [n+1]).
  std::vector<std::pair<torch::Tensor, torch::Tensor>> batch(int64_t start, int64_t n) {
    if (start < 0 || start > size()) throw std::runtime_error("MultiSlotFeed: batch start out of range");
    n = std::min(n, size() - start);
    std::vector<std::pair<torch::Tensor, torch::Tensor>> out;
    for (auto& c : cols_) {
      auto lod = torch::empty({n + 1}, torch::kInt64);
      int64_t* lp = lod.data_ptr<int64_t>();
      lp[0] = 0;
      for (int64_t i = 0; i < n; ++i) {
        const int64_t r = order_[start + i];
        lp[i + 1] = lp[i] + (c.offsets[r + 1] - c.offsets[r]);
      }
      auto vals = torch::empty({lp[n]}, c.is_float ? torch::kFloat32 : torch::kInt64);
      for (int64_t i = 0; i < n; ++i) {
        const int64_t r = order_[start + i], len = c.offsets[r + 1] - c.offsets[r];
        if (c.is_float) std::memcpy(vals.data_ptr<float>() + lp[i], c.fvals.data() + c.offsets[r], sizeof(float) * len);
        else std::memcpy(vals.data_ptr<int64_t>() + lp[i], c.ivals.data() + c.offsets[r], sizeof(int64_t) * len);
      }
      out.emplace_back(std::move(vals), std::move(lod));
    }
    return out;
  }

 private:
  std::vector<SlotColumn> blank() const {
    std::vector<SlotColumn> b(cols_.size());
    for (size_t s = 0; s < cols_.size(); ++s) b[s].is_float = cols_[s].is_float;
    return b;
  }

  void parse_file(const std::string& path, Shard& sh) const {
    sh.cols = blank();
    std::ifstream f(path);
    if (!f) {
      sh.error = "cannot open";
      return;
    }
    std::string line;
    int64_t lineno = 0;
    while (std::getline(f, line)) {
      ++lineno;
      if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
      std::string err;
      if (!parse_line(line.data(), line.data() + line.size(), sh.cols, err)) {
        sh.error = "line " + std::to_string(lineno) + ": " + err;
        return;
      }
      ++sh.records;
    }
  }

  static void append(SlotColumn& dst, const SlotColumn& src) {
    const int64_t base = dst.offsets.back();
    dst.ivals.insert(dst.ivals.end(), src.ivals.begin(), src.ivals.end());
    dst.fvals.insert(dst.fvals.end(), src.fvals.begin(), src.fvals.end());
    for (size_t i = 1; i < src.offsets.size(); ++i) dst.offsets.push_back(base + src.offsets[i]);
  }

  int threads_;
  std::vector<SlotColumn> cols_;
  std::vector<int64_t> order_;
};

void bind_data_feed(pybind11::module_& m) {
  pybind11::class_<MultiSlotFeed>(m, "MultiSlotFeed")
      .def(pybind11::init<std::vector<std::string>, int>(), pybind11::arg("slot_types"), pybind11::arg("threads") = 4)
      .def("load", &MultiSlotFeed::load)
      .def("load_lines", &MultiSlotFeed::load_lines)
      .def("size", &MultiSlotFeed::size)
      .def("shuffle", &MultiSlotFeed::shuffle)
      .def("keep_partition", &MultiSlotFeed::keep_partition)
      .def("clear", &MultiSlotFeed::clear)
      .def("batch", &MultiSlotFeed::batch);
}

}  // namespace runtime
}  // namespace b200
