// Fused RMSNorm / LayerNorm forward+backward for sm_100a.
// Parity (behaviour): paddle/phi/kernels/fusion/gpu/fused_layernorm_kernel.cu, fused_rms_norm (reference).
// Design: one CTA per row, row cached in registers (16-byte vectors), fp32 statistics; the backward is a
// persistent grid that keeps per-CTA dW/dB partial sums in registers across rows and writes them once
// (two-stage reduction, no atomics).  HBM-bound: forward moves 2 x row bytes, backward 3 x row bytes.
#include "include/b200_common.cuh"
#include "include/b200_ops.h"

namespace b200 {

static constexpr int kNormThreads = 256;
static constexpr int kMaxVPT = 8;  // vectors per thread cached in registers (256 thr * 8 vec * 8 elem = 16384 bf16 cols)

template <typename T, int VPT, bool kResidual, bool kLayerNorm>
__global__ void __launch_bounds__(kNormThreads)
norm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ residual, const T* __restrict__ w,
                const T* __restrict__ b, T* __restrict__ y, T* __restrict__ res_out, float* __restrict__ mean_out,
                float* __restrict__ rstd_out, int cols, float eps) {
  constexpr int N = Vec16<T>::N;
  __shared__ float red[33];
  const int64_t row = blockIdx.x;
  const T* xr = x + row * cols;
  const int nvec = cols / N;
  Vec16<T> xv[VPT];
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    if (v < nvec) {
      xv[i] = ld16_stream(xr + v * N);
      if constexpr (kResidual) {
        Vec16<T> rv = ld16_stream(residual + row * cols + v * N);
#pragma unroll
        for (int j = 0; j < N; ++j) xv[i].v[j] = from_f<T>(to_f(xv[i].v[j]) + to_f(rv.v[j]));
        st16(res_out + row * cols + v * N, xv[i]);
      }
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float f = to_f(xv[i].v[j]);
        sum += f;
        sq += f * f;
      }
    }
  }
  float mean = 0.f, rstd;
  if constexpr (kLayerNorm) {
    mean = block_sum(sum, red) / cols;
    // second pass over registers for a numerically stable variance
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = threadIdx.x + i * kNormThreads;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float d = to_f(xv[i].v[j]) - mean;
          var += d * d;
        }
      }
    }
    rstd = rsqrtf(block_sum(var, red) / cols + eps);
  } else {
    rstd = rsqrtf(block_sum(sq, red) / cols + eps);
  }
  if (threadIdx.x == 0) {
    if (rstd_out) rstd_out[row] = rstd;
    if (kLayerNorm && mean_out) mean_out[row] = mean;
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    if (v < nvec) {
      Vec16<T> o;
      Vec16<T> wv, bv;
      if (w) wv = ld16(w + v * N);
      if (b) bv = ld16(b + v * N);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float f = (to_f(xv[i].v[j]) - mean) * rstd;
        if (w) f *= to_f(wv.v[j]);
        if (b) f += to_f(bv.v[j]);
        o.v[j] = from_f<T>(f);
      }
      st16_stream(y + row * cols + v * N, o);
    }
  }
}

template <typename T, bool kLayerNorm>
static void launch_norm_fwd(const void* x, const void* residual, const void* w, const void* b, void* y, void* res_out,
                            float* mean, float* rstd, int64_t rows, int cols, float eps, cudaStream_t s) {
  constexpr int N = Vec16<T>::N;
  if (cols % N != 0 || cols > kNormThreads * kMaxVPT * N) {
    set_last_error(__FILE__, __LINE__, "norm: cols must be a multiple of the 16B vector and <= 256*8 vectors");
    return;
  }
  const int nvec = cols / N;
  const int vpt = (nvec + kNormThreads - 1) / kNormThreads;
  dim3 grid((unsigned)rows), block(kNormThreads);
#define LAUNCH(V)                                                                                                   \
  if (residual)                                                                                                     \
    norm_fwd_kernel<T, V, true, kLayerNorm><<<grid, block, 0, s>>>((const T*)x, (const T*)residual, (const T*)w,   \
                                                                   (const T*)b, (T*)y, (T*)res_out, mean, rstd, cols, eps); \
  else                                                                                                              \
    norm_fwd_kernel<T, V, false, kLayerNorm><<<grid, block, 0, s>>>((const T*)x, nullptr, (const T*)w, (const T*)b, \
                                                                    (T*)y, nullptr, mean, rstd, cols, eps);
  if (vpt <= 1) { LAUNCH(1) } else if (vpt <= 2) { LAUNCH(2) } else if (vpt <= 4) { LAUNCH(4) } else { LAUNCH(8) }
#undef LAUNCH
  B200_CUDA_CHECK(cudaGetLastError());
}

void rms_norm_fwd(const void* x, const void* residual, const void* w, const void* b, void* y, void* res_out,
                  float* rstd, int64_t rows, int cols, float eps, int dtype, cudaStream_t s) {
  if (rows == 0) return;
  B200_DISPATCH_DTYPE(dtype, T, (launch_norm_fwd<T, false>(x, residual, w, b, y, res_out, nullptr, rstd, rows, cols, eps, s)));
}

void layer_norm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t rows,
                    int cols, float eps, int dtype, cudaStream_t s) {
  if (rows == 0) return;
  B200_DISPATCH_DTYPE(dtype, T, (launch_norm_fwd<T, true>(x, nullptr, w, b, y, nullptr, mean, rstd, rows, cols, eps, s)));
}

// ------------------------------------------------------------------------------------------------ backward
// Persistent: CTA c handles rows c, c+G, c+2G, ...; dW/dB partials for the thread's own columns live in registers.
template <typename T, int VPT, bool kLayerNorm>
__global__ void __launch_bounds__(kNormThreads)
norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ w,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in, T* __restrict__ dx,
                float* __restrict__ dw_partial, float* __restrict__ db_partial, int64_t rows, int cols) {
  constexpr int N = Vec16<T>::N;
  __shared__ float red[33];
  const int nvec = cols / N;
  float dw_acc[VPT][N];
  float db_acc[kLayerNorm ? VPT : 1][kLayerNorm ? N : 1];
  float wf[VPT][N];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    Vec16<T> wv;
    if (v < nvec && w) wv = ld16(w + v * N);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      dw_acc[i][j] = 0.f;
      if constexpr (kLayerNorm) db_acc[i][j] = 0.f;
      wf[i][j] = (v < nvec && w) ? to_f(wv.v[j]) : 1.f;
    }
  }
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float rstd = rstd_in[row];
    const float mean = kLayerNorm ? mean_in[row] : 0.f;
    Vec16<T> xv[VPT], gv[VPT];
    float s1 = 0.f, s2 = 0.f;  // s1 = sum(dy*w), s2 = sum(dy*w*xhat)
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = threadIdx.x + i * kNormThreads;
      if (v < nvec) {
        xv[i] = ld16_stream(x + row * cols + v * N);
        gv[i] = ld16_stream(dy + row * cols + v * N);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float xh = (to_f(xv[i].v[j]) - mean) * rstd;
          const float g = to_f(gv[i].v[j]);
          const float gw = g * wf[i][j];
          s1 += gw;
          s2 += gw * xh;
          dw_acc[i][j] += g * xh;
          if constexpr (kLayerNorm) db_acc[i][j] += g;
        }
      }
    }
    s2 = block_sum(s2, red) / cols;
    if constexpr (kLayerNorm) s1 = block_sum(s1, red) / cols; else s1 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = threadIdx.x + i * kNormThreads;
      if (v < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float xh = (to_f(xv[i].v[j]) - mean) * rstd;
          const float gw = to_f(gv[i].v[j]) * wf[i][j];
          o.v[j] = from_f<T>((gw - s1 - xh * s2) * rstd);
        }
        st16_stream(dx + row * cols + v * N, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        if (dw_partial) dw_partial[(int64_t)blockIdx.x * cols + v * N + j] = dw_acc[i][j];
        if constexpr (kLayerNorm) {
          if (db_partial) db_partial[(int64_t)blockIdx.x * cols + v * N + j] = db_acc[i][j];
        }
      }
    }
  }
}

int norm_bwd_num_partials(int64_t rows) {
  const int64_t g = (int64_t)sm_count() * 2;
  return (int)(rows < g ? rows : g);
}

template <typename T, bool kLayerNorm>
static void launch_norm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                            float* dwp, float* dbp, int64_t rows, int cols, int n_partial, cudaStream_t s) {
  constexpr int N = Vec16<T>::N;
  if (cols % N != 0 || cols > kNormThreads * kMaxVPT * N) {
    set_last_error(__FILE__, __LINE__, "norm bwd: unsupported cols");
    return;
  }
  const int nvec = cols / N;
  const int vpt = (nvec + kNormThreads - 1) / kNormThreads;
  dim3 grid((unsigned)n_partial), block(kNormThreads);
#define LAUNCH(V)                                                                                              \
  norm_bwd_kernel<T, V, kLayerNorm><<<grid, block, 0, s>>>((const T*)dy, (const T*)x, (const T*)w, mean, rstd, \
                                                           (T*)dx, dwp, dbp, rows, cols);
  if (vpt <= 1) { LAUNCH(1) } else if (vpt <= 2) { LAUNCH(2) } else if (vpt <= 4) { LAUNCH(4) } else { LAUNCH(8) }
#undef LAUNCH
  B200_CUDA_CHECK(cudaGetLastError());
}

void rms_norm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw_partial,
                  float* db_partial, int64_t rows, int cols, int dtype, int n_partial, cudaStream_t s) {
  if (rows == 0) return;
  B200_DISPATCH_DTYPE(dtype, T, (launch_norm_bwd<T, false>(dy, x, w, nullptr, rstd, dx, dw_partial, db_partial, rows, cols, n_partial, s)));
}

void layer_norm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                    float* dw_partial, float* db_partial, int64_t rows, int cols, int dtype, int n_partial,
                    cudaStream_t s) {
  if (rows == 0) return;
  B200_DISPATCH_DTYPE(dtype, T, (launch_norm_bwd<T, true>(dy, x, w, mean, rstd, dx, dw_partial, db_partial, rows, cols, n_partial, s)));
}

// out[c] = sum_p partial[p][c]; one thread per column chunk, coalesced over c.
template <typename T>
__global__ void reduce_partials_kernel(const float* __restrict__ partial, T* __restrict__ out, int n_partial, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int p = 0;
  for (; p + 3 < n_partial; p += 4) {
    acc0 += partial[(int64_t)p * cols + c];
    acc1 += partial[(int64_t)(p + 1) * cols + c];
    acc2 += partial[(int64_t)(p + 2) * cols + c];
    acc3 += partial[(int64_t)(p + 3) * cols + c];
  }
  for (; p < n_partial; ++p) acc0 += partial[(int64_t)p * cols + c];
  out[c] = from_f<T>((acc0 + acc1) + (acc2 + acc3));
}

void reduce_partials(const float* partial, void* out, int n_partial, int cols, int dtype, cudaStream_t s) {
  const int threads = 128;
  B200_DISPATCH_DTYPE(dtype, T, (reduce_partials_kernel<T><<<(cols + threads - 1) / threads, threads, 0, s>>>(partial, (T*)out, n_partial, cols)));
  B200_CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200
