// ks_xgpu.cuh -- sum of the per-GPU count tables over NVLink / NVSwitch peer memory, inside one kernel.
//
// The one real exchange of the multi-GPU path is the sum of the int32[4^k] count tables
// of all ranks (the reference counts all sequences into one table, /root/reference/src/kmer_spans.c:475-483).
// With the tables in symmetric memory every rank owns one slice and
//   * NVSwitch multicast present: multimem.ld_reduce adds the slice across all GPUs INSIDE the switch and
//     multimem.st writes the sum back into every GPU's table -- each rank moves 1/N of the table once;
//   * otherwise: plain peer loads from every rank's table and peer stores into every rank's table.
// Two packed int32 counters travel as one u64 (no carry can cross: totals fit int32, as in the reference),
// and the trailing u64 of the buffer is the word count, summed by the same instruction.
// The caller brackets the kernel with two cross-GPU barriers on the same stream (dist.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

constexpr int XSUM_MAX_RANKS = 16;
struct XsumPeers { uint64_t *table[XSUM_MAX_RANKS]; };

__global__ void __launch_bounds__(256) xsum_multicast_kernel(uint64_t *__restrict__ mc, size_t first, size_t count) {
  // four independent switch reductions in flight per thread: the loop is latency bound otherwise
  constexpr int U = 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < count; i0 += U * stride) {
    uint64_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      if (i < count)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v[u]) : "l"(mc + first + i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      if (i < count)
        asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(mc + first + i), "l"(v[u]) : "memory");
    }
  }
}

__global__ void __launch_bounds__(256) xsum_peer_kernel(XsumPeers P, int nranks, size_t first, size_t count) {
  // 16-byte vectors where the slice allows it; the slice bounds are multiples of 2 u64 except at the tail
  const size_t nvec = count / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    ulonglong2 acc = make_ulonglong2(0, 0);
    for (int r = 0; r < nranks; ++r) {
      unsigned long long vx, vy;  // peer data written by another GPU: not through the read-only path
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(vx), "=l"(vy) : "l"(P.table[r] + first + 2 * i) : "memory");
      acc.x += vx;
      acc.y += vy;
    }
    for (int r = 0; r < nranks; ++r) *reinterpret_cast<ulonglong2 *>(P.table[r] + first + 2 * i) = acc;
  }
  if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const size_t i = first + count - 1;
    uint64_t acc = 0;
    for (int r = 0; r < nranks; ++r) acc += *reinterpret_cast<const volatile uint64_t *>(P.table[r] + i);
    for (int r = 0; r < nranks; ++r) P.table[r][i] = acc;
  }
}

}  // namespace ks
