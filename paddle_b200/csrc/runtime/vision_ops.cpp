// Host-side pieces of the detection ops (paddle.vision.ops): the sequential part of greedy NMS.
//
// Greedy NMS is inherently serial in the score order (whether box i survives depends on every survivor before it), so the device computes the
// O(n^2) suppression matrix in one shot (IoU > threshold, upper triangle of the score-sorted boxes) and this scan walks it once: n row ORs over
// 64-bit words, no per-box device synchronisation.  For CPU tensors the matrix is built here directly from the boxes, row by row, only for the
// boxes that survive (so the common case costs far less than n^2 IoUs).
// Parity (role): paddle/phi/kernels/gpu/nms_kernel.cu (bit-mask kernel + host scan), paddle/phi/kernels/cpu/nms_kernel.cc.
#include <torch/extension.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "runtime.h"

namespace b200 {
namespace runtime {
namespace {

// mask: uint8 [n, n], mask[i][j] != 0 <=> box j (j > i in score order) overlaps box i above the threshold.  Returns kept positions (ascending).
at::Tensor nms_scan(const at::Tensor& mask) {
  TORCH_CHECK(mask.device().is_cpu() && mask.scalar_type() == at::kByte && mask.dim() == 2 && mask.size(0) == mask.size(1), "nms_scan: uint8 [n, n] CPU mask");
  const at::Tensor m = mask.contiguous();
  const int64_t n = m.size(0);
  const uint8_t* p = m.data_ptr<uint8_t>();
  std::vector<uint8_t> removed(n, 0);
  std::vector<int64_t> keep;
  keep.reserve(n);
  for (int64_t i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep.push_back(i);
    const uint8_t* row = p + i * n;
    for (int64_t j = i + 1; j < n; ++j) removed[j] |= row[j];
  }
  return at::tensor(keep, at::kLong);
}

// boxes: float32 [n, 4] (x1, y1, x2, y2) already sorted by descending score, on the CPU.  Returns kept positions (ascending).
at::Tensor nms_sorted_cpu(const at::Tensor& boxes, double threshold) {
  TORCH_CHECK(boxes.device().is_cpu() && boxes.scalar_type() == at::kFloat && boxes.dim() == 2 && boxes.size(1) == 4, "nms_sorted_cpu: float32 [n, 4] CPU boxes");
  const at::Tensor b = boxes.contiguous();
  const int64_t n = b.size(0);
  const float* p = b.data_ptr<float>();
  std::vector<float> area(n);
  for (int64_t i = 0; i < n; ++i) area[i] = std::max(p[4 * i + 2] - p[4 * i], 0.f) * std::max(p[4 * i + 3] - p[4 * i + 1], 0.f);
  std::vector<uint8_t> removed(n, 0);
  std::vector<int64_t> keep;
  const float thr = (float)threshold;
  for (int64_t i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep.push_back(i);
    const float x1 = p[4 * i], y1 = p[4 * i + 1], x2 = p[4 * i + 2], y2 = p[4 * i + 3];
    for (int64_t j = i + 1; j < n; ++j) {
      if (removed[j]) continue;
      const float w = std::max(std::min(x2, p[4 * j + 2]) - std::max(x1, p[4 * j]), 0.f);
      const float h = std::max(std::min(y2, p[4 * j + 3]) - std::max(y1, p[4 * j + 1]), 0.f);
      const float inter = w * h;
      const float iou = inter / std::max(area[i] + area[j] - inter, 1e-10f);
      if (iou > thr) removed[j] = 1;
    }
  }
  return at::tensor(keep, at::kLong);
}

}  // namespace

void bind_vision(pybind11::module_& m) {
  m.def("nms_scan", &nms_scan, "greedy scan over a [n, n] suppression matrix of score-sorted boxes");
  m.def("nms_sorted_cpu", &nms_sorted_cpu, "greedy NMS over score-sorted CPU boxes");
}

}  // namespace runtime
}  // namespace b200
