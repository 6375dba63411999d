// Best-fit offset allocator with coalescing free list and statistics.  Manages a range of offsets, not pointers, so one
// instance per rank yields identical offsets on every rank for identical call sequences (symmetric heap), and the same
// class backs host-side pools.  Role of the reference's auto-growth best-fit allocator + stats
// (paddle/phi/core/memory/allocation/auto_growth_best_fit_allocator.cc, paddle/phi/core/memory/stats.h).
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>

namespace b200 {
namespace runtime {

class BestFitAllocator {
 public:
  BestFitAllocator(int64_t begin, int64_t end, int64_t min_align = 256) : begin_(begin), end_(end), min_align_(min_align) {
    if (end < begin) throw std::runtime_error("BestFitAllocator: empty range");
    if (end > begin) insert_free(begin, end - begin);
  }

  int64_t alloc(int64_t nbytes, int64_t align) {
    if (nbytes <= 0) nbytes = 1;
    if (align < min_align_) align = min_align_;
    nbytes = (nbytes + min_align_ - 1) / min_align_ * min_align_;
    // smallest free block that fits once its start is aligned (ties: lowest offset => deterministic)
    for (auto it = by_size_.lower_bound({nbytes, INT64_MIN}); it != by_size_.end(); ++it) {
      int64_t off = it->second, size = it->first;
      int64_t start = (off + align - 1) / align * align;
      if (start + nbytes > off + size) continue;
      erase_free(off, size);
      if (start > off) insert_free(off, start - off);
      if (start + nbytes < off + size) insert_free(start + nbytes, off + size - start - nbytes);
      live_[start] = nbytes;
      in_use_ += nbytes;
      if (in_use_ > peak_) peak_ = in_use_;
      if (start + nbytes > high_water_) high_water_ = start + nbytes;
      ++n_alloc_;
      return start;
    }
    throw std::runtime_error("BestFitAllocator: out of memory: need " + std::to_string(nbytes) + " B, in use " + std::to_string(in_use_) + " of " +
                             std::to_string(end_ - begin_) + " B, largest free block " + std::to_string(largest_free()) + " B");
  }

  void free(int64_t off) {
    auto it = live_.find(off);
    if (it == live_.end()) throw std::runtime_error("BestFitAllocator: free of unknown offset " + std::to_string(off));
    int64_t size = it->second;
    live_.erase(it);
    in_use_ -= size;
    ++n_free_;
    // coalesce with the neighbours
    auto next = by_off_.lower_bound(off);
    if (next != by_off_.end() && next->first == off + size) {
      int64_t nsz = next->second;
      erase_free(next->first, nsz);
      size += nsz;
    }
    auto prev = by_off_.lower_bound(off);
    if (prev != by_off_.begin()) {
      --prev;
      if (prev->first + prev->second == off) {
        int64_t poff = prev->first, psz = prev->second;
        erase_free(poff, psz);
        off = poff;
        size += psz;
      }
    }
    insert_free(off, size);
  }

  // Drop every allocation that starts at or after `to` (bump-style rollback used by scoped scratch regions).
  void release_from(int64_t to) {
    while (true) {
      auto it = live_.lower_bound(to);
      if (it == live_.end()) break;
      free(it->first);
    }
  }

  int64_t largest_free() const { return by_size_.empty() ? 0 : by_size_.rbegin()->first; }
  int64_t free_bytes() const { return end_ - begin_ - in_use_; }
  int64_t in_use() const { return in_use_; }
  int64_t peak() const { return peak_; }
  int64_t high_water() const { return high_water_ ? high_water_ : begin_; }
  int64_t num_live() const { return static_cast<int64_t>(live_.size()); }
  int64_t num_free_blocks() const { return static_cast<int64_t>(by_off_.size()); }
  int64_t num_allocs() const { return n_alloc_; }
  int64_t num_frees() const { return n_free_; }
  int64_t capacity() const { return end_ - begin_; }
  int64_t block_size(int64_t off) const {
    auto it = live_.find(off);
    return it == live_.end() ? -1 : it->second;
  }
  void reset_peak() { peak_ = in_use_; }

 private:
  void insert_free(int64_t off, int64_t size) {
    by_off_[off] = size;
    by_size_.insert({size, off});
  }
  void erase_free(int64_t off, int64_t size) {
    by_off_.erase(off);
    by_size_.erase({size, off});
  }

  int64_t begin_, end_, min_align_;
  std::map<int64_t, int64_t> by_off_;               // free blocks: offset -> size
  std::set<std::pair<int64_t, int64_t>> by_size_;   // free blocks: (size, offset)
  std::map<int64_t, int64_t> live_;                 // allocated: offset -> size
  int64_t in_use_ = 0, peak_ = 0, high_water_ = 0, n_alloc_ = 0, n_free_ = 0;
};

}  // namespace runtime
}  // namespace b200
