// Flash-attention forward, ping-pong variant (two query tiles per CTA): see fwd2_kernel.  Same operands / outputs as
// attention_sm100.cu.  Experimental (B200_ATTN_FWD=2): measured slower than the single-tile kernel on B200 because the single
// V stage puts the TMA latency of V(j+1) on the critical path; kept for the next iteration (P in TMEM frees the smem for it).
#include <cuda.h>
#include <cstdio>
#include <string>

#include "include/b200_common.cuh"
#include "include/b200_ops.h"

namespace b200 {
namespace attn2 {

constexpr int BM = 128, BN = 128, HD = 128;
constexpr int kThreads2 = 320;
constexpr uint32_t TILE_BYTES = 128 * 128 * 2;
constexpr uint32_t HALF_BYTES = TILE_BYTES / 2;
constexpr uint32_t SMEM_BYTES = 7 * TILE_BYTES + 1024 + 256;
constexpr uint32_t TMEM_COLS = 512;

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
    if (spins > 2048 && (spins & 1023) == 0 && gtimer() - t0 > 4000000000ull) {  // protocol bug: trap, never hang the GPU
      printf("b200 attention: mbarrier timeout (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {   // SWIZZLE_128B smem descriptor
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n\t"
      "tcgen05.wait::st.sync.aligned;"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}


struct Params {
  int b, sq, sk, h, hk;
  float scale_log2;
  int causal, causal_off;
  void* o;
  float* lse;
  int64_t o_sb, o_ss, o_sh;
  uint32_t idesc_qk, idesc_pv;
};

// Ping-pong forward: one CTA = TWO 128-row query tiles (A, B) of one (batch, head); warpgroup X (4 warps, thread = row) owns
// tile X's softmax while the tensor core runs the other tile's MMAs.  MMA issue order in steady state:
//   P_A V(j) | S_A(j+1) = Q_A K(j+1)^T | P_B V(j) | S_B(j+1) = Q_B K(j+1)^T
// so each warpgroup's softmax latency is covered by three MMAs of work.
// smem (32 KB tiles): Q_A, Q_B, K x2, V, P_A, P_B.  TMEM: S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512).
template <typename T>
__global__ void __launch_bounds__(kThreads2, 1)
fwd2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
            const __grid_constant__ CUtensorMap map_v, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  auto sQ = [&](int x) { return base + x * TILE_BYTES; };
  auto sK = [&](int s) { return base + (2 + s) * TILE_BYTES; };
  const uint32_t sV = base + 4 * TILE_BYTES;
  auto sP = [&](int x) { return base + (5 + x) * TILE_BYTES; };
  const uint32_t bars = base + 7 * TILE_BYTES;
  const uint32_t q_full = bars, v_full = bars + 8, v_empty = bars + 16;
  auto k_full = [&](int s) { return bars + 8u * (3 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (5 + s); };
  auto s_full = [&](int x) { return bars + 8u * (7 + x); };
  auto s_free = [&](int x) { return bars + 8u * (9 + x); };
  auto p_full = [&](int x) { return bars + 8u * (11 + x); };
  auto pv_done = [&](int x) { return bars + 8u * (13 + x); };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 7 * TILE_BYTES + 8 * 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = (int)gridDim.x - 1 - (int)blockIdx.x;   // long rows first
  const int head = blockIdx.y, batch = blockIdx.z;
  const int kv_head = head / (p.h / p.hk);
  const int m0 = pair * 2 * BM;
  auto tiles_for = [&](int mrow0) {
    if (mrow0 >= p.sq) return 0;
    int n = (p.sk + BN - 1) / BN;
    if (p.causal) {
      const int last_key = min(p.sk - 1, mrow0 + BM - 1 + p.causal_off);
      n = last_key < 0 ? 0 : min(n, last_key / BN + 1);
    }
    return n;
  };
  const int nA = tiles_for(m0), nB = tiles_for(m0 + BM);
  const int n_tiles = max(nA, nB);

  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    mbar_init(q_full, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
      mbar_init(s_full(s), 1); mbar_init(s_free(s), 4); mbar_init(p_full(s), 4); mbar_init(pv_done(s), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  } else if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    if (lane == 0 && n_tiles > 0) {
      // ================= TMA producer =================
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
#pragma unroll
      for (int x = 0; x < 2; ++x) {       // rows beyond sq are zero-filled by TMA
        tma_load_4d(sQ(x), &map_q, q_full, 0, m0 + x * BM, head, batch);
        tma_load_4d(sQ(x) + HALF_BYTES, &map_q, q_full, 64, m0 + x * BM, head, batch);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1, n0 = j * BN;
        mbar_wait(k_empty(s), ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(k_full(s), TILE_BYTES);
        tma_load_4d(sK(s), &map_k, k_full(s), 0, n0, kv_head, batch);
        tma_load_4d(sK(s) + HALF_BYTES, &map_k, k_full(s), 64, n0, kv_head, batch);
        mbar_wait(v_empty, (j & 1) ^ 1);
        mbar_expect_tx(v_full, TILE_BYTES);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int i = 0; i < 2; ++i)
            tma_load_4d(sV + kb * HALF_BYTES + i * 8192, &map_v, v_full, i * 64, n0 + kb * 64, kv_head, batch);
      }
    }
  } else if (warp == 9) {
    if (lane == 0 && n_tiles > 0) {
      // ================= MMA issuer =================
      mbar_wait(q_full, 0);
      auto issue_s = [&](int x, int j) {          // S_x(j) = Q_x K(j)^T  (caller has waited for K(j))
        mbar_wait(s_free(x), (j & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + x * BN, make_desc(sQ(x) + kb * HALF_BYTES + k * 32, 16, 1024),
                     make_desc(sK(j & 1) + kb * HALF_BYTES + k * 32, 16, 1024), p.idesc_qk, (kb | k) != 0);
        umma_commit(s_full(x));
      };
      auto issue_pv = [&](int x, int j) {         // O_x += P_x(j) V(j)  (caller has waited for V(j))
        mbar_wait(p_full(x), j & 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + 2 * BN + x * HD, make_desc(sP(x) + kb * HALF_BYTES + k * 32, 16, 1024),
                     make_desc(sV + kb * HALF_BYTES + k * 2048, 8192, 1024), p.idesc_pv, (j | kb | k) != 0);
        umma_commit(pv_done(x));
      };
      mbar_wait(k_full(0), 0);
      if (nA > 0) issue_s(0, 0);
      if (nB > 0) issue_s(1, 0);
      umma_commit(k_empty(0));
      for (int j = 0; j < n_tiles; ++j) {
        const bool more = j + 1 < n_tiles;
        mbar_wait(v_full, j & 1);
        if (j < nA) issue_pv(0, j);
        if (more) mbar_wait(k_full((j + 1) & 1), ((j + 1) >> 1) & 1);
        if (more && j + 1 < nA) issue_s(0, j + 1);
        if (j < nB) issue_pv(1, j);
        umma_commit(v_empty);
        if (more && j + 1 < nB) issue_s(1, j + 1);
        if (more) umma_commit(k_empty((j + 1) & 1));
      }
    }
  } else {
    // ================= softmax + epilogue: warpgroup x owns query tile x (thread == row == TMEM lane) =================
    const int x = warp >> 2;
    const int rl = (warp & 3) * 32 + lane;
    const int row = m0 + x * BM + rl;
    const int nX = x == 0 ? nA : nB;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s_col = tmem_base + lane_off + x * BN, o_col = tmem_base + lane_off + 2 * BN + x * HD;
    float m_i = -INFINITY, l_i = 0.f;
    for (int j = 0; j < nX; ++j) {
      const int n0 = j * BN;
      mbar_wait(s_full(x), j & 1);
      tc_fence_after();
      const bool edge = (n0 + BN > p.sk) || (p.causal && n0 + BN - 1 > m0 + x * BM + p.causal_off);
      const int lim = p.causal ? min(p.sk - 1, row + p.causal_off) : p.sk - 1;
      // pass 1: row maximum (values are re-read from TMEM in pass 2 to keep the register footprint of 8 softmax warps small)
      float mxp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(s_col + c * 32, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = __uint_as_float(r[i]);
          if (edge && n0 + c * 32 + i > lim) v = -INFINITY;
          mxp[i & 3] = fmaxf(mxp[i & 3], v);
        }
      }
      const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3])) * p.scale_log2;
      float m_new = fmaxf(m_i, mx);
      if (m_new == -INFINITY) m_new = 0.f;
      if (j > 0) {
        mbar_wait(pv_done(x), (j - 1) & 1);       // P_x buffer free and O_x final (already true in steady state: S_x(j) was issued after P_x V(j-1))
        tc_fence_after();
        const bool need = (m_new - m_i) > 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2(m_i - m_new) : 1.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld32(o_col + c * 32, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(o_col + c * 32, r);
          }
          l_i *= alpha;
          if (need) m_i = m_new;
        }
      } else {
        m_i = m_new;
      }
      // pass 2: P = exp2(S * scale - m), written K-major / 128B-swizzled for the P V MMA
      float sump[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(s_col + c * 32, r);
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          uint32_t u[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i0 = q8 * 8 + 2 * e;
            float p0 = ex2(fmaf(__uint_as_float(r[i0]), p.scale_log2, -m_i));
            float p1 = ex2(fmaf(__uint_as_float(r[i0 + 1]), p.scale_log2, -m_i));
            if (edge) {
              if (n0 + c * 32 + i0 > lim) p0 = 0.f;
              if (n0 + c * 32 + i0 + 1 > lim) p1 = 0.f;
            }
            sump[e] += p0 + p1;
            u[e] = pack2<T>(p0, p1);
          }
          const int col = c * 32 + q8 * 8;
          const uint32_t addr = sP(x) + (col >> 6) * HALF_BYTES + rl * 128 + ((((col & 63) >> 3) ^ (rl & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]) : "memory");
        }
      }
      l_i += (sump[0] + sump[1]) + (sump[2] + sump[3]);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(s_free(x)); mbar_arrive(p_full(x)); }
    }
    if (nX > 0) {
      mbar_wait(pv_done(x), (nX - 1) & 1);
      tc_fence_after();
    }
    const float inv = l_i > 0.f ? 1.f / l_i : 0.f;
    T* orow = reinterpret_cast<T*>(p.o) + (int64_t)batch * p.o_sb + (int64_t)row * p.o_ss + (int64_t)head * p.o_sh;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      if (nX > 0) {
        tmem_ld32(o_col + c * 32, r);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (row < p.sq) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack2<T>(__uint_as_float(r[q * 8 + 0]) * inv, __uint_as_float(r[q * 8 + 1]) * inv);
          o.y = pack2<T>(__uint_as_float(r[q * 8 + 2]) * inv, __uint_as_float(r[q * 8 + 3]) * inv);
          o.z = pack2<T>(__uint_as_float(r[q * 8 + 4]) * inv, __uint_as_float(r[q * 8 + 5]) * inv);
          o.w = pack2<T>(__uint_as_float(r[q * 8 + 6]) * inv, __uint_as_float(r[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = o;
        }
      }
    }
    if (row < p.sq && p.lse) p.lse[((int64_t)batch * p.h + head) * p.sq + row] = l_i > 0.f ? (m_i + log2f(l_i)) * 0.69314718055994531f : -INFINITY;
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 4-D map {d, s, h, b} over a strided [B,S,H,D] view (strides in elements), box {64, rows, 1, 1}, 128B swizzle
static bool make_map4(CUtensorMap* out, const void* ptr, int d, int s, int h, int b, int64_t ss, int64_t sh, int64_t sb, uint32_t box_rows, int dtype) {
  bind_primary_context();
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)s, (cuuint64_t)h, (cuuint64_t)b};
  cuuint64_t strides[3] = {(cuuint64_t)ss * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(__FILE__, __LINE__, ("cuTensorMapEncodeTiled (attention) failed: " + std::to_string((int)r)).c_str());
    return false;
  }
  return true;
}

static uint32_t make_idesc(int dtype, int n, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                                  // fp32 accumulate
  const uint32_t f = dtype == kBF16 ? 1u : 0u;
  d |= f << 7;
  d |= f << 10;
  d |= (b_mn ? 1u : 0u) << 16;                   // B operand MN-major (V: head_dim contiguous)
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

}  // namespace attn2

int attention_fwd2(const AttnArgs& a, cudaStream_t s) {
  using namespace attn2;
  if (!attention_fwd_supported(a)) return 1;
  CUtensorMap mq, mk, mv;
  if (!make_map4(&mq, a.q, a.d, a.sq, a.h, a.b, a.q_strides[1], a.q_strides[2], a.q_strides[0], BM, a.dtype)) return 2;
  if (!make_map4(&mk, a.k, a.d, a.sk, a.hk, a.b, a.k_strides[1], a.k_strides[2], a.k_strides[0], BN, a.dtype)) return 2;
  if (!make_map4(&mv, a.v, a.d, a.sk, a.hk, a.b, a.v_strides[1], a.v_strides[2], a.v_strides[0], 64, a.dtype)) return 2;
  Params p;
  p.b = a.b; p.sq = a.sq; p.sk = a.sk; p.h = a.h; p.hk = a.hk;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.causal = a.causal; p.causal_off = a.sk - a.sq;
  p.o = a.o; p.lse = a.lse;
  p.o_sb = a.o_strides[0]; p.o_ss = a.o_strides[1]; p.o_sh = a.o_strides[2];
  p.idesc_qk = make_idesc(a.dtype, BN, false);
  p.idesc_pv = make_idesc(a.dtype, HD, true);
  dim3 grid((a.sq + 2 * BM - 1) / (2 * BM), a.h, a.b);
  static bool attr_bf = false, attr_h = false;
  if (a.dtype == kBF16) {
    auto kern = fwd2_kernel<__nv_bfloat16>;
    if (!attr_bf) { B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr_bf = true; }
    kern<<<grid, kThreads2, SMEM_BYTES, s>>>(mq, mk, mv, p);
  } else {
    auto kern = fwd2_kernel<__half>;
    if (!attr_h) { B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr_h = true; }
    kern<<<grid, kThreads2, SMEM_BYTES, s>>>(mq, mk, mv, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); return 3; }
  return 0;
}

}  // namespace b200
