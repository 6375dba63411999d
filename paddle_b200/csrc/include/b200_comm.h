// Host launch API of the peer-memory collectives (csrc/comm/*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {
namespace comm {
// bases[r] = address of rank r's heap slab as mapped in THIS process; off = byte offset of the symmetric buffer.
void p2p_allreduce(const int64_t* bases, int64_t off, int64_t n, int dtype, int rank, int world, uint32_t epoch, uint32_t* counter,
                   cudaStream_t s);
void p2p_reduce_scatter(const int64_t* bases, int64_t off, void* out, int64_t n, int dtype, int rank, int world, uint32_t epoch,
                        uint32_t* counter, cudaStream_t s);
// out[i] = sum_s staging[s][i]: `world` slots of n elements at byte offset `off` of THIS rank's heap, filled by the peers'
// GEMM epilogues (GemmArgs::rs_dst). Cross-rank barriers before (all pushes landed) and after (slots may be reused).
void p2p_reduce_slots(const int64_t* bases, int64_t off, void* out, int64_t n, int dtype, int rank, int world, uint32_t epoch,
                      uint32_t* counter, cudaStream_t s);
void p2p_allgather(const int64_t* bases, int64_t off, int64_t chunk_bytes, int rank, int world, uint32_t epoch, uint32_t* counter,
                   cudaStream_t s);
void p2p_alltoall(const int64_t* bases, int64_t off_send, int64_t off_recv, int64_t chunk_bytes, int rank, int world, uint32_t epoch,
                  uint32_t* counter, cudaStream_t s);
// MoE dispatch / combine: variable all-to-all push with optional row gather (see a2av_kernel). rows_hint sizes the grid.
void p2p_a2av(const int64_t* bases, const void* src, const int64_t* gather, int64_t meta_off, int64_t recv_off, int64_t row_bytes,
              int64_t rows_hint, int rank, int world, uint32_t epoch, uint32_t* counter, cudaStream_t s);
// dst[r*chunk_bytes ...] (ordinary local memory) <- chunk_bytes at byte offset src_off of rank r's heap, for every r
void p2p_gather_pull(const int64_t* bases, int64_t src_off, void* dst, int64_t chunk_bytes, int rank, int world, uint32_t epoch,
                     uint32_t* counter, cudaStream_t s);
// pipeline mailbox: release-store `value` into a flag in a peer's heap / spin (one thread) until the local flag reaches `value`;
// stats (optional, device uint64[2]) accumulates the spun nanoseconds and the number of waits.
void p2p_signal_flag(void* remote_flag, uint32_t value, cudaStream_t s);
void p2p_wait_flag(const void* flag, uint32_t value, void* stats, double timeout_s, cudaStream_t s);
// NVLS (multicast) collectives, csrc/comm/nvls_collectives.cu.  pads[r] = rank r's signal pad as mapped here, pad_off = byte offset of our
// barrier words inside it; mc / local = multicast and local address of the symmetric buffer; dtype 0 fp32, 1 bf16, 2 fp16.
void nvls_allreduce(const int64_t* pads, int64_t pad_off, int64_t mc, int64_t local, int64_t off, int64_t n, int dtype, int rank, int world, uint32_t epoch,
                    uint32_t* counter, cudaStream_t s);
void nvls_reduce_scatter(const int64_t* pads, int64_t pad_off, int64_t mc, int64_t local, int64_t off, void* out, int64_t n, int dtype, int rank, int world,
                         uint32_t epoch, uint32_t* counter, cudaStream_t s);
void nvls_allgather(const int64_t* pads, int64_t pad_off, int64_t mc, int64_t local, int64_t off, const void* src, int64_t chunk_bytes, int rank, int world,
                    uint32_t epoch, uint32_t* counter, cudaStream_t s);
}  // namespace comm
}  // namespace b200
