// Weight-only quantised GEMM for sm_100a: out[M,N] = x[M,K] (bf16 / fp16) @ dequant(Wq[N,K] int8 | int4) * scale[N] (+ bias).
//
// The int8 / int4 weights never exist in HBM as 16-bit values: TMA streams the raw quantised tile (one or half a byte per weight)
// into a deep shared-memory ring, four dequantise warps expand it IN THE SM into the 128B-swizzled K-major operand layout, and
// tcgen05 multiplies.  The problem is computed transposed (D^T[N, M] = W[N, K] x^T) so that the weight tile is the 128-row A operand
// at full UMMA height even when M is a handful of decode tokens, the per-channel scale is a per-thread scalar in the epilogue
// (TMEM lane = output channel), and the token tile (UMMA N = 16 / 64 / 128) only costs what the batch needs.
// Decode is weight-bandwidth bound: 12 stages x 8 KB of raw weights in flight per SM cover the HBM latency; split-K spreads narrow
// layers over all SMs (fp32 atomics into a workspace, finalised by a tiny kernel).
//
// Parity: paddle/phi/kernels/gpu/weight_only_linear_kernel.cu:27, python/paddle/nn/quant/quantized_linear.py:183.
#include <cuda.h>
#include <cstdio>
#include <string>

#include "include/b200_common.cuh"
#include "include/b200_ops.h"

namespace b200 {
namespace gemm {
bool make_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t ld, uint64_t bstride,
              uint32_t box_inner, uint32_t box_rows, int dtype);
}
namespace wo {

constexpr int BLOCK_N = 128;     // output channels per CTA = UMMA M (TMEM lanes)
constexpr int BLOCK_K = 64;
constexpr int kBStages = 3;      // dequantised A-operand ring
constexpr int kDqWarps = 8;      // dequantise warps: two per SM sub-partition so that their ld.shared / convert / st.shared chains overlap
constexpr int kThreads = 64 + kDqWarps * 32;    // warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2-9: dequantise (2-5 also epilogue)
constexpr uint32_t A_TILE_BYTES = BLOCK_N * BLOCK_K * 2;   // 16 KB dequantised weight tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t gtimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins == 2048) t0 = gtimer();
    if (spins > 2048 && (spins & 1023) == 0 && gtimer() - t0 > 4000000000ull) {
      printf("b200 weight-only gemm: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {   // SWIZZLE_128B, K-major
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

struct Params {
  int m, n, k;                 // tokens, output channels, reduction
  int kb_per_split;            // 64-wide k-blocks handled by one blockIdx.y
  const float* scale;          // [n] per-channel dequantisation factors
  const void* bias;            // [n] in the activation dtype or nullptr
  void* out;                   // [m, n] activation dtype (splits == 1)
  float* ws;                   // [m, n] fp32 workspace (splits > 1, zeroed)
  int splits;
  int bf16;                    // activation dtype: 1 bf16, 0 fp16
  uint32_t idesc;
};

template <bool BF16> __device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (BF16) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
  else { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
}

// four signed int8 in `w` -> four floats (exact): byte ^ 0x80 is 0..255; 0x4B0000xx is 8388608 + xx
__device__ __forceinline__ void int8x4_to_f(uint32_t w, float (&f)[4]) {
  const uint32_t u = w ^ 0x80808080u;
  f[0] = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7650)) - 8388736.f;
  f[1] = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7651)) - 8388736.f;
  f[2] = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7652)) - 8388736.f;
  f[3] = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7653)) - 8388736.f;
}

// NTOK: token tile (UMMA N).  INT4: two weights per byte (low nibble = even k).
template <int NTOK, int STAGES, bool INT4, bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
wo_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Params p) {
  constexpr uint32_t RAW_BYTES = BLOCK_N * (INT4 ? BLOCK_K / 2 : BLOCK_K);    // 8 KB (int8) / 4 KB (int4) raw weights per stage
  constexpr uint32_t X_BYTES = NTOK * BLOCK_K * 2;
  constexpr uint32_t STAGE_BYTES = ((RAW_BYTES + X_BYTES + 1023) / 1024) * 1024;
  constexpr uint32_t TMEM_COLS = NTOK < 32 ? 32 : NTOK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base;                                   // [STAGES] x {raw W | x tile}
  const uint32_t aring = base + STAGES * STAGE_BYTES;           // [kBStages] dequantised weight tiles
  const uint32_t bars = aring + kBStages * A_TILE_BYTES;
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto raw_empty = [&](int s) { return bars + 8u * (STAGES + s); };
  auto a_ready = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto a_empty = [&](int s) { return bars + 8u * (2 * STAGES + kBStages + s); };
  const uint32_t tfull = bars + 8u * (2 * STAGES + 2 * kBStages);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + STAGES * STAGE_BYTES + kBStages * A_TILE_BYTES + 8 * (2 * STAGES + 2 * kBStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N;
  const int tok0 = blockIdx.z * NTOK;
  const int num_kb_total = (p.k + BLOCK_K - 1) / BLOCK_K;
  const int kb0 = blockIdx.y * p.kb_per_split;
  const int num_kb = max(0, min(p.kb_per_split, num_kb_total - kb0));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), 1); }
    for (int s = 0; s < kBStages; ++s) { mbar_init(a_ready(s), kDqWarps); mbar_init(a_empty(s), 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
    fence_proxy_async();
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer: raw quantised weights + the activation tile =================
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        mbar_wait(raw_empty(s), ((i / STAGES) & 1) ^ 1);
        mbar_expect_tx(raw_full(s), RAW_BYTES + X_BYTES);
        const int kb = kb0 + i;
        tma_load_2d(ring + s * STAGE_BYTES, &map_w, raw_full(s), INT4 ? kb * (BLOCK_K / 2) : kb * BLOCK_K, n0);
        tma_load_3d(ring + s * STAGE_BYTES + RAW_BYTES, &map_x, raw_full(s), kb * BLOCK_K, tok0, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer: D^T[128 channels, NTOK tokens] += W_tile[128, 64] x_tile[NTOK, 64]^T =================
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES, t = i % kBStages;
        mbar_wait(raw_full(s), (i / STAGES) & 1);          // the activation tile of this k-block has landed
        mbar_wait(a_ready(t), (i / kBStages) & 1);         // the dequantise warps have written the weight tile
        tc_fence_after();
        const uint32_t sa = aring + t * A_TILE_BYTES;
        const uint32_t sb = ring + s * STAGE_BYTES + RAW_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k)
          umma_f16(tmem_base, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), p.idesc, (i | k) != 0);
        umma_commit(raw_empty(s));
        umma_commit(a_empty(t));
      }
      umma_commit(tfull);
    }
  } else {
    // ================= dequantise warps (then epilogue): thread = (row, 16-byte piece) work items =================
    const int tid = threadIdx.x - 64;                     // 0..kDqWarps*32-1
    constexpr int DQ = kDqWarps * 32;
    const int ew = warp & 3;                              // TMEM lane quadrant this warp may read (hardware: warp id % 4)
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % STAGES, t = i % kBStages;
      mbar_wait(raw_full(s), (i / STAGES) & 1);
      mbar_wait(a_empty(t), ((i / kBStages) & 1) ^ 1);
      const uint32_t raw = ring + s * STAGE_BYTES;
      const uint32_t dst = aring + t * A_TILE_BYTES;
      if constexpr (!INT4) {
        // raw tile: 128 rows x 64 B, SWIZZLE_64B (16-byte piece q of row r sits at piece q ^ ((r >> 1) & 3))
#pragma unroll
        for (int it = 0; it < 512 / DQ; ++it) {
          const int c = tid + it * DQ, r = c >> 2, q = c & 3;
          uint4 v;
          const uint32_t src = raw + r * 64 + ((q ^ ((r >> 1) & 3)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src));
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[4];
            int8x4_to_f(w4[j], f);
            o[2 * j] = pack2<BF16>(f[0], f[1]);
            o[2 * j + 1] = pack2<BF16>(f[2], f[3]);
          }
          // dequantised tile: K-major SWIZZLE_128B, row r at r * 128 B, 16-byte chunk j at (j ^ (r & 7)); int8 piece q -> chunks 2q, 2q+1
          const uint32_t row = dst + r * 128;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((2 * q) ^ (r & 7)) << 4)), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((2 * q + 1) ^ (r & 7)) << 4)), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
        }
      } else {
        // raw tile: 128 rows x 32 B, SWIZZLE_32B (piece q of row r at q ^ ((r >> 2) & 1)); one 16-byte piece = 32 weights = 4 output chunks
#pragma unroll
        for (int it = 0; it < 256 / DQ; ++it) {
          const int c = tid + it * DQ, r = c >> 1, q = c & 1;
          uint4 v;
          const uint32_t src = raw + r * 32 + ((q ^ ((r >> 2) & 1)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src));
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
          const uint32_t row = dst + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {                    // 4 bytes = 8 weights = one 16-byte output chunk
            const uint32_t lo = (w4[j] & 0x0F0F0F0Fu) ^ 0x08080808u, hi = ((w4[j] >> 4) & 0x0F0F0F0Fu) ^ 0x08080808u;   // nibble ^ 8 = value + 8
            float a[4], b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              a[e] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7650 + e)) - 8388616.f;
              b[e] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7650 + e)) - 8388616.f;
            }
            const uint32_t o0 = pack2<BF16>(a[0], b[0]), o1 = pack2<BF16>(a[1], b[1]), o2 = pack2<BF16>(a[2], b[2]), o3 = pack2<BF16>(a[3], b[3]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((4 * q + j) ^ (r & 7)) << 4)), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(t));
    }
    // ---- epilogue (warps 2-5: one per TMEM lane quadrant): TMEM lane = output channel, column = token ----
    if (num_kb > 0 && warp < 6) {
      mbar_wait(tfull, 0);
      tc_fence_after();
      const int ch = n0 + ew * 32 + lane;
      const bool ch_ok = ch < p.n;
      const float sc = ch_ok ? p.scale[ch] : 0.f;
      float bv = 0.f;
      if (ch_ok && p.bias && p.splits == 1) bv = BF16 ? __bfloat162float(((const __nv_bfloat16*)p.bias)[ch]) : __half2float(((const __half*)p.bias)[ch]);
#pragma unroll 1
      for (int c = 0; c < NTOK / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + c * 16, r);
        if (!ch_ok) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int tok = tok0 + c * 16 + j;
          if (tok >= p.m) break;
          const float v = __uint_as_float(r[j]);
          if (p.splits == 1) {
            const float y = v * sc + bv;
            if constexpr (BF16) ((__nv_bfloat16*)p.out)[(int64_t)tok * p.n + ch] = __float2bfloat16_rn(y);
            else ((__half*)p.out)[(int64_t)tok * p.n + ch] = __float2half_rn(y);
          } else {
            atomicAdd(p.ws + (int64_t)tok * p.n + ch, v);
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <bool BF16>
__global__ void wo_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ scale, const void* __restrict__ bias, void* __restrict__ out,
                                   int64_t total, int n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % n);
    float y = ws[i] * scale[ch];
    if (bias) y += BF16 ? __bfloat162float(((const __nv_bfloat16*)bias)[ch]) : __half2float(((const __half*)bias)[ch]);
    if constexpr (BF16) ((__nv_bfloat16*)out)[i] = __float2bfloat16_rn(y);
    else ((__half*)out)[i] = __float2half_rn(y);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}
// raw weights as bytes: [n rows, row_bytes], box {64 | 32 bytes, 128 rows}, swizzle = box width
static bool make_w_map(CUtensorMap* out, const void* w, int n, int row_bytes, int box_bytes) {
  bind_primary_context();
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)n};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_bytes, (cuuint32_t)BLOCK_N};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error(__FILE__, __LINE__, ("weight-only gemm: cuTensorMapEncodeTiled failed: " + std::to_string((int)r)).c_str()); return false; }
  return true;
}

static uint32_t make_idesc(bool bf16, int ntok) {
  uint32_t d = 0;
  d |= 1u << 4;                                  // fp32 accumulate
  const uint32_t f = bf16 ? 1u : 0u;
  d |= f << 7;                                   // A (weights) format
  d |= f << 10;                                  // B (activations) format
  d |= (uint32_t)(ntok >> 3) << 17;              // UMMA N = tokens
  d |= (uint32_t)(BLOCK_N >> 4) << 24;           // UMMA M = 128 channels
  return d;
}

template <int NTOK, int STAGES, bool INT4, bool BF16>
static int launch(const WoGemmArgs& g, const CUtensorMap& mw, cudaStream_t s) {
  constexpr uint32_t RAW_BYTES = BLOCK_N * (INT4 ? BLOCK_K / 2 : BLOCK_K);
  constexpr uint32_t STAGE_BYTES = ((RAW_BYTES + NTOK * BLOCK_K * 2 + 1023) / 1024) * 1024;
  constexpr uint32_t SMEM = STAGES * STAGE_BYTES + kBStages * A_TILE_BYTES + 1024 + 512;
  static_assert(SMEM <= 232448, "weight-only gemm: shared memory budget");
  CUtensorMap mx;
  if (!gemm::make_map(&mx, g.x, g.k, g.m, 1, g.k, 0, BLOCK_K, NTOK, g.bf16 ? kBF16 : kF16)) return 2;
  Params p;
  p.m = g.m; p.n = g.n; p.k = g.k;
  const int num_kb = (g.k + BLOCK_K - 1) / BLOCK_K;
  const int n_tiles = (g.n + BLOCK_N - 1) / BLOCK_N, t_tiles = (g.m + NTOK - 1) / NTOK;
  int splits = 1;
  if (g.ws != nullptr) {      // decode: one wave of CTAs over all SMs - split the reduction as far as the SM count and K allow
    const int ctas = n_tiles * t_tiles;
    splits = sm_count() / (ctas > 0 ? ctas : 1);
    if (splits > num_kb / 8) splits = num_kb / 8;
    if (splits > 16) splits = 16;
    if (splits < 1) splits = 1;
  }
  p.splits = splits;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.scale = g.scale; p.bias = g.bias; p.out = g.out; p.ws = g.ws; p.bf16 = g.bf16;
  p.idesc = make_idesc(g.bf16, NTOK);
  auto kern = wo_gemm_kernel<NTOK, STAGES, INT4, BF16>;
  static bool attr_set = false;
  if (!attr_set) { B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr_set = true; }
  if (splits > 1) B200_CUDA_CHECK(cudaMemsetAsync(g.ws, 0, (size_t)g.m * g.n * sizeof(float), s));
  dim3 grid(n_tiles, splits, t_tiles);
  kern<<<grid, kThreads, SMEM, s>>>(mw, mx, p);
  if (splits > 1) {
    const int64_t total = (int64_t)g.m * g.n;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    wo_finalize_kernel<BF16><<<blocks, 256, 0, s>>>(g.ws, g.scale, g.bias, g.out, total, g.n);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); return 3; }
  return 0;
}

template <bool INT4, bool BF16>
static int dispatch_tok(const WoGemmArgs& g, const CUtensorMap& mw, cudaStream_t s) {
  if (g.m <= 16) return launch<16, 12, INT4, BF16>(g, mw, s);
  if (g.m <= 64) return launch<64, 8, INT4, BF16>(g, mw, s);
  return launch<128, 6, INT4, BF16>(g, mw, s);
}

}  // namespace wo

int gemm_weight_only(const WoGemmArgs& g, cudaStream_t s) {
  using namespace wo;
  if (g.k % 64 || g.n % 8 || g.m <= 0) return 1;
  if ((reinterpret_cast<uintptr_t>(g.x) | reinterpret_cast<uintptr_t>(g.w) | reinterpret_cast<uintptr_t>(g.out)) & 15) return 1;
  CUtensorMap mw;
  const int row_bytes = g.int4 ? g.k / 2 : g.k;
  if (!make_w_map(&mw, g.w, g.n, row_bytes, g.int4 ? 32 : 64)) return 2;
  if (g.int4) return g.bf16 ? dispatch_tok<true, true>(g, mw, s) : dispatch_tok<true, false>(g, mw, s);
  return g.bf16 ? dispatch_tok<false, true>(g, mw, s) : dispatch_tok<false, false>(g, mw, s);
}

}  // namespace b200
