// ks_hash.cuh -- device helpers of the large-k path shared by the counting kernels (ks_large.cuh) and the scan
// kernels (ks_kernels.cuh): 64-bit k-mer codes from the packed stream and the open-addressing hash table.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ks_chunk.cuh"

namespace ks {

struct __align__(16) HashSlot {
  unsigned long long key;  // code + 1; 0 = empty
  long long val;           // count while counting, then the fixed-point score
};

__device__ __forceinline__ uint64_t hash_mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// run masks over 64 positions: bit b of the result is set iff bits b-len+1 .. b of ok are all set (len 1..32)
__device__ __forceinline__ uint64_t run_ending64(uint64_t ok, int len) {
  uint64_t r = ok;
  int have = 1;
  while (have * 2 <= len) { r &= r << have; have *= 2; }
  if (have < len) r &= r << (len - have);
  return r;
}

// packed codes and break bits of positions [p0 - 32, p0 + 16): hi32 = the first 16 of them, lo64 the other 32
// (first base most significant), brk48 bit b = position p0 - 32 + b.  Chunks in front of `first_chunk` do not
// exist (front of the buffer): they read as breaks.
__device__ __forceinline__ void load_window_wide(const uint32_t *__restrict__ pk, const uint16_t *__restrict__ brk,
                                                 int64_t first_chunk, int64_t p0, uint32_t &hi32, uint64_t &lo64,
                                                 uint64_t &brk48) {
  const int64_t wq = p0 >> 4;
  const int r = (int)(p0 & 15);
  uint32_t w0 = 0, b0 = 0xffffu;
  if (wq - 2 >= first_chunk) { w0 = __ldg(&pk[wq - 2]); b0 = __ldg(&brk[wq - 2]); }
  uint32_t w1 = 0, b1 = 0xffffu;
  if (wq - 1 >= first_chunk) { w1 = __ldg(&pk[wq - 1]); b1 = __ldg(&brk[wq - 1]); }
  const uint32_t w2 = __ldg(&pk[wq]), b2 = __ldg(&brk[wq]);
  if (r == 0) {
    hi32 = w0;
    lo64 = ((uint64_t)w1 << 32) | w2;
    brk48 = (uint64_t)b0 | ((uint64_t)b1 << 16) | ((uint64_t)b2 << 32);
  } else {
    const uint32_t w3 = __ldg(&pk[wq + 1]), b3 = __ldg(&brk[wq + 1]);
    hi32 = __funnelshift_l(w1, w0, 2 * r);
    lo64 = ((uint64_t)__funnelshift_l(w2, w1, 2 * r) << 32) | __funnelshift_l(w3, w2, 2 * r);
    const uint64_t b64 = (uint64_t)b0 | ((uint64_t)b1 << 16) | ((uint64_t)b2 << 32) | ((uint64_t)b3 << 48);
    brk48 = (b64 >> r) & 0xffffffffffffull;
  }
}
// 2k-bit code whose last base sits `s` bits above the low end of the 96-bit window hi32:lo64 (s even, 0..32)
__device__ __forceinline__ uint64_t wide_code(uint32_t hi32, uint64_t lo64, int s, uint64_t kmask) {
  const uint64_t v = s ? ((lo64 >> s) | ((uint64_t)hi32 << (64 - s))) : lo64;
  return v & kmask;
}

// score of the k-mer `code` (it occurs: every scored k-mer was counted); WFX_KILL if the table does not hold it
__device__ __forceinline__ int64_t hash_lookup(const HashSlot *__restrict__ slots, uint64_t mask, uint64_t code) {
  const unsigned long long key = code + 1;
  uint64_t h = hash_mix(code) & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes) {
    const ulonglong2 s = *reinterpret_cast<const ulonglong2 *>(&slots[h]);
    if (s.x == key) return (int64_t)s.y;
    if (s.x == 0ull) break;
    h = (h + 1) & mask;
  }
  return WFX_KILL;
}

}  // namespace ks
