// ks_large.cuh -- the large-k path (k = 16 .. 31; BASELINE.json configs[3]: k = 21), where a direct 4^k table
// cannot exist: k-mer codes are 64-bit and the count table is an open-addressing hash table in HBM.
//
// Outside the reference's domain (/root/reference/src/kmer_spans.c:37 MAX_K 16, :139 int shift, :504): defined by
// extension, checked against oracle/ks_oracle_large.c:
//   hash_count_kernel     packed codes -> 64-bit rolling codes -> linear probing with atomicCAS on the key,
//                         atomicAdd on the count; the counting rules of sequence_kmer_count (:135-155) unchanged
//   hash_compact_kernel   occupied slots -> composite sort keys (count << 2k | code) + slot index
//   (ks_sort.cuh)         stable LSD radix sort: the order (count, code) of rank_kmers_w (:189-202) over the
//                         k-mers that occur (absent ones add 0 to the reference's running sum)
//   large_heads_kernel    run heads of the count part -> (count, first position) per distinct count
//   large_eval_kernel     rank from the linear pieces of ks_rankseg.h, written back into the k-mer's slot as the
//                         exact fixed-point score (rank - thr, or +-1 - thr) the scan gathers
//   scan_gather_kernel<4> (ks_kernels.cuh) probes the table per position instead of indexing a 4^k array
#pragma once
#include "ks_kernels.cuh"  // includes ks_hash.cuh

namespace ks {

struct HashArgs {
  HashSlot *slots;
  uint64_t mask;              // capacity - 1 (power of two)
  unsigned long long *stats;  // [0] words counted, [1] distinct k-mers, [2] largest count, [3] error flags
};

__global__ void __launch_bounds__(256) hash_count_kernel(const uint32_t *__restrict__ pk, const uint16_t *__restrict__ brk,
                                                         const uint8_t *__restrict__ buf, int64_t first_chunk,
                                                         int64_t nchunks, int k, HashArgs H) {
  const uint64_t kmask = (((uint64_t)1) << (2 * k)) - 1;
  unsigned long long local = 0;
  for (int64_t c = first_chunk + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < first_chunk + nchunks;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p0 = 16 * c;  // this thread counts the k-mers ENDING at positions p0 .. p0 + 15
    uint32_t hi32;
    uint64_t lo64, brk48;
    load_window_wide(pk, brk, 0, p0, hi32, lo64, brk48);
    const uint64_t runk = run_ending64(~brk48, k);          // bit b: positions b-k+1 .. b hold no break
    const uint64_t first = runk & (brk48 << k);             // ... and the run starts exactly k positions back
    uint32_t counted = (uint32_t)(runk >> 32) & 0xffffu;
    // a run of exactly k bases that ends at the terminator counts nothing (:143-144): first k-mers of a run are
    // rare, so the one ASCII byte behind them is read only for those
    uint32_t f = (uint32_t)(first >> 32) & 0xffffu;
    while (f) {
      const int j = __ffs(f) - 1;
      f &= f - 1;
      if (buf[p0 + j + 1] == 0) counted &= ~(1u << j);
    }
#pragma unroll 4
    for (int j = 0; j < CHUNK; ++j) {
      if (!(counted & (1u << j))) continue;
      const uint64_t code = wide_code(hi32, lo64, 2 * (15 - j), kmask);
      const unsigned long long key = code + 1;
      uint64_t h = hash_mix(code) & H.mask;
      for (uint64_t probes = 0;; ++probes) {
        const unsigned long long prev = atomicCAS(&H.slots[h].key, 0ull, key);
        if (prev == 0ull || prev == key) {
          atomicAdd(reinterpret_cast<unsigned long long *>(&H.slots[h].val), 1ull);
          break;
        }
        if (probes > H.mask) { atomicOr(&H.stats[3], 1ull); break; }  // table full: reported, never silent
        h = (h + 1) & H.mask;
      }
    }
    local += __popc(counted);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&H.stats[0], local);
}

// number of occupied slots and the largest count: decide the buffer sizes and whether (count << 2k | code) fits 64 bits
__global__ void __launch_bounds__(256) hash_stats_kernel(HashArgs H) {
  unsigned long long nd = 0, mx = 0;
  for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h <= H.mask; h += (uint64_t)gridDim.x * blockDim.x) {
    if (H.slots[h].key == 0ull) continue;
    ++nd;
    const unsigned long long c = (unsigned long long)H.slots[h].val;
    mx = c > mx ? c : mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nd += __shfl_down_sync(0xffffffffu, nd, o);
    const unsigned long long y = __shfl_down_sync(0xffffffffu, mx, o);
    mx = y > mx ? y : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    if (nd) atomicAdd(&H.stats[1], nd);
    if (mx) atomicMax(&H.stats[2], mx);
  }
}
// occupied slots -> sort key and slot index, in arbitrary order (the sort fixes it).  composite: key = count << 2k |
// code, one sort gives the order (count, code); otherwise key = code and the count is sorted in a second pass.
__global__ void __launch_bounds__(256) hash_compact_kernel(HashArgs H, int k, int composite, uint64_t *__restrict__ keys,
                                                           uint32_t *__restrict__ slot_of, uint64_t cap_out,
                                                           unsigned long long *cursor) {
  for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h <= H.mask; h += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = H.slots[h].key;
    if (key == 0ull) continue;
    const unsigned long long cnt = (unsigned long long)H.slots[h].val;
    const unsigned long long at = atomicAdd(cursor, 1ull);
    if (at < cap_out) {
      keys[at] = composite ? ((cnt << (2 * k)) | (key - 1)) : (key - 1);
      slot_of[at] = (uint32_t)h;
    }
  }
}
// counts in code order (second sort of the two-pass order)
__global__ void __launch_bounds__(256) large_counts_kernel(const HashSlot *__restrict__ slots, const uint32_t *__restrict__ slot_of,
                                                           uint64_t n, uint32_t *__restrict__ cnt) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = (uint32_t)slots[slot_of[i]].val;
}

// run heads of the count part of the sorted composite keys -> (count, first position), appended unordered
__global__ void __launch_bounds__(256) large_heads_kernel(const uint64_t *__restrict__ comp, uint64_t n, int k,
                                                          uint32_t *__restrict__ gcount, uint32_t *__restrict__ gstart,
                                                          uint32_t *ngroups, uint32_t cap) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t c = comp[i] >> (2 * k);
  if (i == 0 || (comp[i - 1] >> (2 * k)) != c) {
    const uint32_t slot = atomicAdd(ngroups, 1u);
    if (slot < cap) { gcount[slot] = (uint32_t)c; gstart[slot] = (uint32_t)i; }
  }
}

// position p of the (count, code) order: rank from the linear pieces (the same fma as rank_eval_kernel), then the
// slot of the k-mer receives the fixed-point score the scan adds.  mode 0: rank - thr; mode 2: (count / T >= f_t ?
// +1 : -1) - thr.  ranks_out (optional): the rank doubles in order, for inspection.
__global__ void __launch_bounds__(256) large_eval_kernel(const uint64_t *__restrict__ comp, const uint32_t *__restrict__ slot_of,
                                                         const uint32_t *__restrict__ idx /* or NULL */,
                                                         const uint32_t *__restrict__ cnt32 /* or NULL */,
                                                         uint64_t n, int k, const uint32_t *__restrict__ gstart,
                                                         uint32_t ngroups, const uint32_t *__restrict__ seg_first,
                                                         const unsigned long long *__restrict__ seg_j0,
                                                         const double *__restrict__ seg_x0,
                                                         const double *__restrict__ seg_inc, int mode, double total,
                                                         double param, double thr, int qs, HashSlot *__restrict__ slots,
                                                         double *__restrict__ ranks_out) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t lo = 0, hi = ngroups;  // largest g with gstart[g] <= p
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(&gstart[mid]) <= p) lo = mid; else hi = mid;
  }
  const unsigned long long j = p - __ldg(&gstart[lo]);
  uint32_t a = __ldg(&seg_first[lo]), b = __ldg(&seg_first[lo + 1]);
  while (b - a > 1) {
    const uint32_t mid = (a + b) >> 1;
    if (__ldg(&seg_j0[mid]) <= j) a = mid; else b = mid;
  }
  const double r = fma((double)(j - __ldg(&seg_j0[a])), __ldg(&seg_inc[a]), __ldg(&seg_x0[a]));
  if (ranks_out) ranks_out[p] = r;
  double w;
  if (mode == 0) {
    w = r - thr;
  } else {
    const uint32_t c = cnt32 ? cnt32[p] : (uint32_t)(comp[p] >> (2 * k));
    const double f = (double)(int32_t)c / total;
    w = (f >= param ? 1.0 : -1.0) - thr;
  }
  slots[slot_of[idx ? idx[p] : p]].val = wfx_from_double(w, qs);
}

}  // namespace ks
