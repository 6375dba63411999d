// Weight-only quantised GEMM for sm_100a: out[M,N] = x[M,K] (bf16 / fp16) @ dequant(Wq[N,K] int8 | int4) * scale[N] (+ bias).
//
// The int8 / int4 weights never exist in HBM as 16-bit values: TMA streams the raw quantised tile (one or half a byte per weight)
// into a deep shared-memory ring, eight dequantise warps (four groups, one k-block each) expand it IN THE SM into the 128B-swizzled K-major operand layout, and
// tcgen05 multiplies.  The problem is computed transposed (D^T[N, M] = W[N, K] x^T) so that the weight tile is the 128-row A operand
// at full UMMA height even when M is a handful of decode tokens, the per-channel scale is a per-thread scalar in the epilogue
// (TMEM lane = output channel), and the token tile (UMMA N = 16 / 64 / 128) only costs what the batch needs.
// Decode is weight-bandwidth bound: 5-7 raw boxes of 16 KB (128 channels x one full 128-byte line) in flight per SM cover the HBM
// latency; split-K spreads narrow
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
// Warp roles for G dequantise groups (G = 4 or 6): warp 0 TMA producer; warp 1 TMEM alloc + MMA issuer 0; warps 2 .. 2G+1 dequantise
// (two per group; 2-5 also run the epilogue); warps 2G+2 .. 3G issuers 1 .. G-1.  Group g converts k-blocks g, g + G, ... into ring slot g
// and issuer g multiplies them, so the wait -> ld.shared -> convert -> st.shared -> fence -> arrive -> issue -> commit chains of G
// consecutive k-blocks overlap.  The MMAs are tiny (NTOK columns): a k-block costs what its ISSUE sequence costs, and one issuing
// thread caps at ~800 cycles per k-block - hence one issuer per slot.
constexpr int kDqGroupWarps = 2;
__host__ __device__ constexpr int wo_threads(int G) { return 32 * (3 * G + 1); }
__host__ __device__ constexpr uint32_t pow2_at_least(uint32_t v) { uint32_t r = 32; while (r < v) r *= 2; return r; }
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
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
  int splits;                  // blockIdx.y extent = cluster size: the CTAs of one output tile
  int bf16;                    // activation dtype: 1 bf16, 0 fp16
  uint32_t idesc;
};

template <bool BF16> __device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (BF16) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
  else { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
}

// Pack two SMALL INTEGER floats (|v| <= 256: 8 significant bits) into the 16-bit pair.  bf16 keeps 8 significand bits, so taking the
// high halves is exact and costs one full-rate PRMT instead of a quarter-rate F2FP; fp16 goes through the converter.
template <bool BF16> __device__ __forceinline__ uint32_t pack2_exact(float a, float b) {
  if constexpr (BF16) return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);
  else return pack2<false>(a, b);
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
// The raw weights arrive as [128 channels x 128 bytes] boxes (one TMA load = 2 k-blocks of int8 / 4 k-blocks of int4): full 128-byte
// lines per channel row keep the L2 / DRAM request count at a quarter (int4) / half (int8) of one-k-block boxes, which is what bounded
// the first version of this kernel.  The activation tile and the dequantised operand stay per 64-wide k-block.
template <int NTOK, int G, int WST, int XPER, bool INT4, bool BF16>
__global__ void __launch_bounds__(wo_threads(G), 1)
wo_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Params p) {
  constexpr int kBStages = G, kDqWarps = kDqGroupWarps * G, kDqGroups = G, kIssuers = G;
  // Activation tiles: XPER private stages per issuer (k-block i -> stage (i % G) * XPER + (i / G) % XPER), so every x_full barrier has ONE
  // waiter that sees its phases in order.  (A ring shared by all issuers lets one issuer run a whole phase ahead of another and read
  // the parity of the previous phase as "done".)
  constexpr int XST = G * XPER;
  constexpr int KB_PER_W = INT4 ? 4 : 2;                        // k-blocks covered by one raw box
  constexpr uint32_t W_BYTES = BLOCK_N * 128;                   // 16 KB raw box
  constexpr uint32_t X_BYTES = NTOK * BLOCK_K * 2;
  constexpr int ACC_PER = NTOK <= 16 ? 2 : 1;                   // independent TMEM accumulators per issuer (summed in the epilogue)
  constexpr int NACC = G * ACC_PER;
  constexpr uint32_t TMEM_COLS = pow2_at_least(NACC * NTOK);
  static_assert(TMEM_COLS <= 512, "weight-only gemm: TMEM budget");
  static_assert(NTOK * 512 <= WST * W_BYTES, "weight-only gemm: the split-K partial tile is staged in the raw weight ring");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t wring = base;                                  // [WST] raw weight boxes
  const uint32_t xring = wring + WST * W_BYTES;                 // [XST] activation tiles
  const uint32_t aring = xring + XST * X_BYTES;                 // [kBStages] dequantised weight tiles
  const uint32_t bars = aring + kBStages * A_TILE_BYTES;
  auto w_full = [&](int s) { return bars + 8u * s; };
  auto w_empty = [&](int s) { return bars + 8u * (WST + s); };
  auto x_full = [&](int s) { return bars + 8u * (2 * WST + s); };
  auto x_empty = [&](int s) { return bars + 8u * (2 * WST + XST + s); };
  auto a_ready = [&](int s) { return bars + 8u * (2 * WST + 2 * XST + s); };
  auto a_empty = [&](int s) { return bars + 8u * (2 * WST + 2 * XST + kBStages + s); };
  const uint32_t tfull = bars + 8u * (2 * WST + 2 * XST + 2 * kBStages);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + WST * W_BYTES + XST * X_BYTES + kBStages * A_TILE_BYTES + 8 * (2 * WST + 2 * XST + 2 * kBStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N;
  const int tok0 = blockIdx.z * NTOK;
  const int num_kb_total = (p.k + BLOCK_K - 1) / BLOCK_K;
  const int kb0 = blockIdx.y * p.kb_per_split;
  const int num_kb = max(0, min(p.kb_per_split, num_kb_total - kb0));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < WST; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), KB_PER_W * kDqGroupWarps); }
    for (int s = 0; s < XST; ++s) { mbar_init(x_full(s), 1); mbar_init(x_empty(s), 1); }
    for (int s = 0; s < kBStages; ++s) { mbar_init(a_ready(s), kDqGroupWarps); mbar_init(a_empty(s), 1); }
    mbar_init(tfull, kIssuers);
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
      // ================= TMA producer: raw quantised weights (one box per KB_PER_W k-blocks) + the activation tile =================
      for (int i = 0; i < num_kb; ++i) {
        const int kb = kb0 + i;
        if (i % KB_PER_W == 0) {                          // kb0 is a multiple of KB_PER_W (launcher), so boxes start on 128-byte columns
          const int j = i / KB_PER_W, ws = j % WST;
          mbar_wait(w_empty(ws), ((j / WST) & 1) ^ 1);
          mbar_expect_tx(w_full(ws), W_BYTES);
          tma_load_2d(wring + ws * W_BYTES, &map_w, w_full(ws), kb * (INT4 ? BLOCK_K / 2 : BLOCK_K), n0);
        }
        const int s = (i % G) * XPER + (i / G) % XPER;
        mbar_wait(x_empty(s), (((i / G) / XPER) & 1) ^ 1);
        mbar_expect_tx(x_full(s), X_BYTES);
        tma_load_3d(xring + s * X_BYTES, &map_x, x_full(s), kb * BLOCK_K, tok0, 0);
      }
    }
  } else if (warp == 1 || warp >= 2 + kDqWarps) {
    if (lane == 0) {
      // ================= MMA issuers: D^T[128 channels, NTOK tokens] += W_tile[128, 64] x_tile[NTOK, 64]^T =================
      // Issuer g takes k-blocks g, g + 4, ... (ring slot g, its own accumulators).  Descriptors are built once and advanced by adding to
      // the 14-bit (address >> 4) field.
      const int g = warp == 1 ? 0 : warp - (1 + kDqWarps);
      const uint64_t ad = make_desc(aring + g * A_TILE_BYTES, 16, 1024), bdesc0 = make_desc(xring, 16, 1024);
      const uint32_t tacc = tmem_base + (uint32_t)(g * ACC_PER * NTOK);
      int aph = 0;
      for (int i = g; i < num_kb; i += kIssuers) {
        const int s = g * XPER + (i / G) % XPER;
        mbar_wait(x_full(s), ((i / G) / XPER) & 1);        // the activation tile of this k-block has landed
        mbar_wait(a_ready(g), aph);                        // dequantise group g has written the weight tile
        aph ^= 1;
        tc_fence_after();
        const uint64_t bd = bdesc0 + (uint64_t)((s * X_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k)
          umma_f16(tacc + (uint32_t)((k % ACC_PER) * NTOK), ad + 2 * k, bd + 2 * k, p.idesc, (i != g || k >= ACC_PER) ? 1u : 0u);
        umma_commit(x_empty(s));
        umma_commit(a_empty(g));
      }
      umma_commit(tfull);
    }
  } else {
    // ================= dequantise warps (then epilogue): thread = (row, 16-byte piece) work items =================
    constexpr int DQ = kDqGroupWarps * 32;                // threads that share one k-block
    const int grp = (warp - 2) / kDqGroupWarps;
    const int tid = (threadIdx.x - 64) % DQ;
    const int ew = warp & 3;                              // TMEM lane quadrant this warp may read (hardware: warp id % 4)
    for (int i = grp; i < num_kb; i += kDqGroups) {
      const int t = grp, j = i / KB_PER_W, ws = j % WST, sub = i % KB_PER_W;
      mbar_wait(w_full(ws), (j / WST) & 1);
      mbar_wait(a_empty(t), ((i / kBStages) & 1) ^ 1);
      const uint32_t raw = wring + ws * W_BYTES;             // 128 rows x 128 B, SWIZZLE_128B: 16-byte piece c of row r sits at piece c ^ (r & 7)
      const uint32_t dst = aring + t * A_TILE_BYTES;
      if constexpr (!INT4) {
        // this k-block = pieces 4 sub .. 4 sub + 3 of every row.  All loads are issued before the first conversion (the asm statements are
        // volatile, i.e. kept in program order: interleaving load / convert / store per item would expose every shared-memory latency).
        constexpr int ITEMS = 512 / DQ;
        uint4 v[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int c = tid + it * DQ, r = c >> 2, q = c & 3;
          const uint32_t src = raw + r * 128 + (((sub * 4 + q) ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[it].x), "=r"(v[it].y), "=r"(v[it].z), "=r"(v[it].w) : "r"(src));
        }
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int c = tid + it * DQ, r = c >> 2, q = c & 3;
          const uint32_t w4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[4];
            int8x4_to_f(w4[j], f);
            o[2 * j] = pack2_exact<BF16>(f[0], f[1]);
            o[2 * j + 1] = pack2_exact<BF16>(f[2], f[3]);
          }
          // dequantised tile: K-major SWIZZLE_128B, row r at r * 128 B, 16-byte chunk j at (j ^ (r & 7)); int8 piece q -> chunks 2q, 2q+1
          const uint32_t row = dst + r * 128;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((2 * q) ^ (r & 7)) << 4)), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((2 * q + 1) ^ (r & 7)) << 4)), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
        }
      } else {
        // this k-block = pieces 2 sub, 2 sub + 1 of every row; one 16-byte piece = 32 weights = 4 output chunks
        constexpr int ITEMS = 256 / DQ;
        uint4 v[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int c = tid + it * DQ, r = c >> 1, q = c & 1;
          const uint32_t src = raw + r * 128 + (((sub * 2 + q) ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[it].x), "=r"(v[it].y), "=r"(v[it].z), "=r"(v[it].w) : "r"(src));
        }
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int c = tid + it * DQ, r = c >> 1, q = c & 1;
          const uint32_t w4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
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
            const uint32_t o0 = pack2_exact<BF16>(a[0], b[0]), o1 = pack2_exact<BF16>(a[1], b[1]), o2 = pack2_exact<BF16>(a[2], b[2]), o3 = pack2_exact<BF16>(a[3], b[3]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((4 * q + j) ^ (r & 7)) << 4)), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(a_ready(t)); mbar_arrive(w_empty(ws)); }   // the raw pieces are in registers / converted: the box slot may be refilled
    }
    // ---- epilogue (warps 2-5: one per TMEM lane quadrant): TMEM lane = output channel, column = token ----
    if (warp < 6) {
      if (num_kb > 0) {
        mbar_wait(tfull, 0);
        tc_fence_after();
      }
      const int chl = ew * 32 + lane, ch = n0 + chl;
      const bool ch_ok = ch < p.n;
      const float sc = ch_ok ? p.scale[ch] : 0.f;
      float bv = 0.f;
      if (ch_ok && p.bias && p.splits == 1) bv = BF16 ? __bfloat162float(((const __nv_bfloat16*)p.bias)[ch]) : __half2float(((const __half*)p.bias)[ch]);
      const int nacc = min(kIssuers, num_kb) * ACC_PER;    // accumulators that received at least one MMA (issuer g has work iff num_kb > g)
#pragma unroll 1
      for (int c = 0; c < NTOK / 16; ++c) {
        uint32_t r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
        for (int a = 0; a < nacc; ++a) {
          uint32_t r2[16];
          tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + a * NTOK + c * 16, r2);
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
        if (p.splits > 1) {
          // split-K inside a cluster: park the partial tile [token][channel] in this CTA's shared memory (the raw weight ring is idle
          // now: every box was consumed before the last MMA could be issued); the cluster reduces it below through DSMEM
#pragma unroll
          for (int j = 0; j < 16; ++j)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(wring + (uint32_t)(((c * 16 + j) * BLOCK_N + chl) * 4)), "r"(r[j]) : "memory");
          continue;
        }
        if (!ch_ok) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int tok = tok0 + c * 16 + j;
          if (tok >= p.m) break;
          const float y = __uint_as_float(r[j]) * sc + bv;
          if constexpr (BF16) ((__nv_bfloat16*)p.out)[(int64_t)tok * p.n + ch] = __float2bfloat16_rn(y);
          else ((__half*)p.out)[(int64_t)tok * p.n + ch] = __float2half_rn(y);
        }
      }
      tc_fence_before();
    }
  }
  if (p.splits > 1) {
    // ---- cluster reduction: the p.splits CTAs of a cluster hold the partial tiles of ONE output tile; CTA r sums 32-channel chunks
    // r, r + splits, ... over all ranks (ld.shared::cluster) and writes the scaled result.  No workspace, no atomics, no second kernel.
    cluster_sync();
    if (warp >= 2 && warp < 6) {
      uint32_t rank;
      asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
      const int tok_n = min(NTOK, p.m - tok0);
      for (int c = (int)rank * 4 + (warp - 2); c < tok_n * 4; c += p.splits * 4) {
        const int tok = c >> 2, chl = (c & 3) * 32 + lane, ch = n0 + chl;
        float acc = 0.f;
        const uint32_t local = wring + (uint32_t)((tok * BLOCK_N + chl) * 4);
        for (int q = 0; q < p.splits; ++q) {
          uint32_t remote;
          float v;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(q));
          asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
          acc += v;
        }
        if (ch < p.n) {
          float y = acc * p.scale[ch];
          if (p.bias) y += BF16 ? __bfloat162float(((const __nv_bfloat16*)p.bias)[ch]) : __half2float(((const __half*)p.bias)[ch]);
          if constexpr (BF16) ((__nv_bfloat16*)p.out)[(int64_t)(tok0 + tok) * p.n + ch] = __float2bfloat16_rn(y);
          else ((__half*)p.out)[(int64_t)(tok0 + tok) * p.n + ch] = __float2half_rn(y);
        }
      }
    }
    cluster_sync();                                        // nobody leaves while a peer may still read its partial tile
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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
// raw weights as bytes: [n rows, row_bytes], box {128 bytes, 128 rows}, SWIZZLE_128B
static bool make_w_map(CUtensorMap* out, const void* w, int n, int row_bytes) {
  bind_primary_context();
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)n};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {128u, (cuuint32_t)BLOCK_N};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

template <int NTOK, int G, int WST, int XPER, bool INT4, bool BF16>
static int launch(const WoGemmArgs& g, const CUtensorMap& mw, cudaStream_t s) {
  constexpr int XST = G * XPER;
  constexpr uint32_t SMEM = WST * BLOCK_N * 128 + XST * NTOK * BLOCK_K * 2 + G * A_TILE_BYTES + 1024 + 512;
  static_assert(SMEM <= 232448, "weight-only gemm: shared memory budget");
  static_assert(8 * (2 * WST + 2 * XST + 2 * G + 1) + 8 <= 512, "weight-only gemm: barrier area");
  constexpr int kMaxSplits = 8;                 // portable cluster size
  CUtensorMap mx;
  if (!gemm::make_map(&mx, g.x, g.k, g.m, 1, g.k, 0, BLOCK_K, NTOK, g.bf16 ? kBF16 : kF16)) return 2;
  auto kern = wo_gemm_kernel<NTOK, G, WST, XPER, INT4, BF16>;
  static bool attr_set = false;
  static int wave_ctas[kMaxSplits + 1];         // CTAs the device holds at once when they come in clusters of `splits`
  if (!attr_set) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    for (int c = 1; c <= kMaxSplits; ++c) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1, c, 1); cfg.blockDim = dim3(wo_threads(G)); cfg.dynamicSmemBytes = SMEM;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = c; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n_clusters = 0; }
      wave_ctas[c] = n_clusters * c;
    }
    if (wave_ctas[1] <= 0) wave_ctas[1] = sm_count();
    attr_set = true;
  }
  Params p;
  p.m = g.m; p.n = g.n; p.k = g.k;
  const int num_kb = (g.k + BLOCK_K - 1) / BLOCK_K;
  const int n_tiles = (g.n + BLOCK_N - 1) / BLOCK_N, t_tiles = (g.m + NTOK - 1) / NTOK;
  // Split-K (a cluster of `splits` CTAs per output tile) when the output tiles alone leave SMs idle or the last wave ragged.  Cost in
  // k-block units: waves x (k-blocks per CTA + fixed prologue / epilogue) + the cluster reduction.
  const int ctas = n_tiles * t_tiles;
  int splits = 1, kb_per = (num_kb + 3) / 4 * 4;
  long best = -1;
  for (int c = 1; c <= kMaxSplits; ++c) {
    if (wave_ctas[c] <= 0) continue;
    const int per = ((num_kb + c - 1) / c + 3) / 4 * 4;           // raw boxes span up to 4 k-blocks: split boundaries stay on box boundaries
    if (c > 1 && (c - 1) * per >= num_kb) continue;               // every rank of the cluster gets work
    const long waves = ((long)ctas * c + wave_ctas[c] - 1) / wave_ctas[c];
    const long cost = waves * (per + 24) + (c > 1 ? 6 : 0);
    if (best < 0 || cost < best) { best = cost; splits = c; kb_per = per; }
  }
  p.splits = splits;
  p.kb_per_split = kb_per;
  p.scale = g.scale; p.bias = g.bias; p.out = g.out; p.bf16 = g.bf16;
  p.idesc = make_idesc(g.bf16, NTOK);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_tiles, splits, t_tiles); cfg.blockDim = dim3(wo_threads(G)); cfg.dynamicSmemBytes = SMEM; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = splits; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mw, mx, p);
  if (e != cudaSuccess) { cudaGetLastError(); set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); return 3; }
  return 0;
}

template <bool INT4, bool BF16>
static int dispatch_tok(const WoGemmArgs& g, const CUtensorMap& mw, cudaStream_t s) {
  if (g.m <= 16) return launch<16, 6, 5, 2, INT4, BF16>(g, mw, s);       // decode: 6 groups, 80 KB of raw weights in flight
  if (g.m <= 64) return launch<64, 4, 5, 2, INT4, BF16>(g, mw, s);
  return launch<128, 4, 5, 1, INT4, BF16>(g, mw, s);
}

}  // namespace wo

int gemm_weight_only(const WoGemmArgs& g, cudaStream_t s) {
  using namespace wo;
  if (g.k % 64 || g.n % 8 || g.m <= 0) return 1;
  if ((reinterpret_cast<uintptr_t>(g.x) | reinterpret_cast<uintptr_t>(g.w) | reinterpret_cast<uintptr_t>(g.out)) & 15) return 1;
  CUtensorMap mw;
  const int row_bytes = g.int4 ? g.k / 2 : g.k;
  if (!make_w_map(&mw, g.w, g.n, row_bytes)) return 2;
  if (g.int4) return g.bf16 ? dispatch_tok<true, true>(g, mw, s) : dispatch_tok<true, false>(g, mw, s);
  return g.bf16 ? dispatch_tok<false, true>(g, mw, s) : dispatch_tok<false, false>(g, mw, s);
}

}  // namespace b200
