// ks_count.cuh -- shared-memory-privatised k-mer counting (K2), replacing sequence_kmer_count
// (/root/reference/src/kmer_spans.c:135-155) where one global reduction per base is not the fastest way.
//
// Measured on a B200 (tools/unit_peaks2.cu, profiles/r02_unit_peaks2.json): random 32-bit reductions into an
// L2-resident table run at 178 G/s chip-wide (the rate of the L2 atomic units), shared-memory atomics over a
// 64 KiB table at 1 630 G/s.  So:
//   k <= 7   pack_count_smem_kernel: the whole 4^k table lives in the shared memory of every CTA (<= 64 KiB),
//            one shared-memory atomic per base, one flush of the non-zero counters per CTA at the end.
//   8..12    two phases over 1024 buckets = the leading 10 bits of the 2k-bit code:
//            bucket_scatter_kernel (fused with the 2-bit packing of K1) ranks every k-mer of a 12 288-position
//            tile inside its bucket with one shared-memory atomic, stages the remaining 2k-10 bits (<= 14, a
//            uint16) in shared memory and appends each bucket's segment to the bucket's region in HBM with
//            8-byte stores (2 B written + 2 B read per base instead of a 32-byte-sector L2 atomic);
//            bucket_count_kernel then counts one bucket per CTA in a shared-memory table of 4^k / 1024 entries
//            and adds it to the bucket's slice of the count table with plain coalesced stores.
//            A k-mer that does not fit its staging row (24 per bucket and tile: repeats, skewed spectra) or its
//            bucket's region falls back to the direct global reduction, so every input stays exact.
//   >= 13    the direct kernel of ks_kernels.cuh in slices of the table (the table exceeds L2).
#pragma once
#include "ks_kernels.cuh"

namespace ks {

constexpr int BK_LOG = 10;
constexpr int BK_BUCKETS = 1 << BK_LOG;
constexpr int BK_CAP = 24;      // staged sub-keys per bucket and tile; rows of 48 bytes (8-byte aligned)
constexpr int BK_ROUNDS = 3;    // chunks per thread and tile
constexpr int BK_THREADS = 256;
constexpr int BK_TILE_CHUNKS = BK_THREADS * BK_ROUNDS;  // 768 chunks = 12 288 positions, 12 per bucket on average
constexpr uint32_t BK_PAD = 0xffffu;                    // filler of the 4-entry granules (a sub-key has <= 14 bits)
constexpr size_t BK_SCATTER_SMEM = BK_BUCKETS * 4 + (size_t)BK_BUCKETS * BK_CAP * 2;  // counters + rows

__device__ __forceinline__ unsigned long long block_sum_to(unsigned long long local, unsigned long long *dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  __shared__ unsigned long long sm_total[8];
  if ((threadIdx.x & 31) == 0) sm_total[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm_total[i];
    if (t) atomicAdd(dst, t);
  }
  return local;
}

// front half shared with pack_count_kernel, in two steps so that callers can keep several chunks' loads in flight:
// raw bytes of chunk ci (16 bytes before it, the chunk, first byte behind it) -> packed output + the codes
// sequence_kmer_count counts
struct RawChunk { uint4 a, b; uint32_t next; };
__device__ __forceinline__ RawChunk load_chunk_raw(const uint8_t *__restrict__ buf, int64_t ci) {
  const uint8_t *p = buf + 16 * ci;  // chunk ci+1 of the buffer starts at p + 16
  const uint4 *v = reinterpret_cast<const uint4 *>(p);
  RawChunk r;
  r.a = ld_stream_u4(v);
  r.b = ld_stream_u4(v + 1);
  r.next = __ldg(p + 32);
  return r;
}
__device__ __forceinline__ void pack_decode_raw(const RawChunk &r, int64_t ci, int64_t first, int k, uint32_t kmask,
                                                uint32_t *__restrict__ pk_out, uint16_t *__restrict__ brk_out,
                                                uint32_t code[CHUNK], uint32_t &counted) {
  uint32_t wp[4] = {r.a.x, r.a.y, r.a.z, r.a.w}, wc[4] = {r.b.x, r.b.y, r.b.z, r.b.w};
  uint32_t pkp, bp, np, pkc, bc, nc;
  pack16(wp, pkp, bp, np);
  pack16(wc, pkc, bc, nc);
  pk_out[ci + 1] = pkc;
  brk_out[ci + 1] = (uint16_t)bc;
  if (ci == first) { pk_out[ci] = pkp; brk_out[ci] = (uint16_t)bp; }  // no thread of its own (ks_kernels.cuh)
  decode_count(((uint64_t)pkp << 32) | pkc, bp | (bc << 16), np | (nc << 16), r.next == 0u, k, kmask, code, counted);
}
__device__ __forceinline__ void pack_decode_chunk(const uint8_t *__restrict__ buf, int64_t ci, int64_t first, int k,
                                                  uint32_t kmask,
                                                  uint32_t *__restrict__ pk_out, uint16_t *__restrict__ brk_out,
                                                  uint32_t code[CHUNK], uint32_t &counted) {
  pack_decode_raw(load_chunk_raw(buf, ci), ci, first, k, kmask, pk_out, brk_out, code, counted);
}

// ------------------------------------------------------------------------------------------------------------
// k <= 7: the table in shared memory (dynamic, 4 << 2k bytes)
__global__ void __launch_bounds__(256) pack_count_smem_kernel(const uint8_t *__restrict__ buf, int64_t first,
                                                              int64_t nchunks, int k, uint32_t kmask,
                                                              uint32_t *__restrict__ pk_out,
                                                              uint16_t *__restrict__ brk_out,
                                                              int32_t *__restrict__ counts,
                                                              unsigned long long *__restrict__ nwords) {
  extern __shared__ __align__(16) unsigned char ks_dyn_smem[];
  uint32_t *s_tab = reinterpret_cast<uint32_t *>(ks_dyn_smem);
  const uint32_t entries = kmask + 1u;
  for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) s_tab[i] = 0;
  __syncthreads();
  unsigned long long local = 0;
  for (int64_t ci = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ci < first + nchunks;
       ci += (int64_t)gridDim.x * blockDim.x) {
    uint32_t code[CHUNK], counted;
    pack_decode_chunk(buf, ci, first, k, kmask, pk_out, brk_out, code, counted);
#pragma unroll
    for (int j = 0; j < CHUNK; ++j)
      if (counted & (1u << j)) atomicAdd(&s_tab[code[j]], 1u);
    local += __popc(counted);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) {
    const uint32_t c = s_tab[i];
    if (c) atomicAdd(reinterpret_cast<uint32_t *>(counts) + i, c);
  }
  block_sum_to(local, nwords);
}

// ------------------------------------------------------------------------------------------------------------
// 8 <= k <= 12, phase 1: pack + scatter the sub-keys into their buckets
__global__ void __launch_bounds__(BK_THREADS, 4) bucket_scatter_kernel(const uint8_t *__restrict__ buf, int64_t first,
                                                                      int64_t nchunks, int k, uint32_t kmask,
                                                                      uint32_t *__restrict__ pk_out,
                                                                      uint16_t *__restrict__ brk_out,
                                                                      int32_t *__restrict__ counts,
                                                                      unsigned long long *__restrict__ nwords,
                                                                      uint16_t *__restrict__ bk_buf,
                                                                      uint32_t *__restrict__ bk_cursor, uint32_t gcap,
                                                                      int sub_bits) {
  extern __shared__ __align__(16) unsigned char ks_dyn_smem[];
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(ks_dyn_smem);
  uint16_t *s_stage = reinterpret_cast<uint16_t *>(ks_dyn_smem + BK_BUCKETS * 4);
  const int tid = threadIdx.x;
  const uint32_t submask = (1u << sub_bits) - 1u;
  const uint64_t keep = l2_policy_evict_last();
  unsigned long long local = 0;
  const int64_t ntiles = (nchunks + BK_TILE_CHUNKS - 1) / BK_TILE_CHUNKS;
  constexpr int BPT = BK_BUCKETS / BK_THREADS;  // buckets per thread (flush)
  for (int i = tid; i < BK_BUCKETS; i += BK_THREADS) s_cnt[i] = 0;
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    RawChunk raw[BK_ROUNDS];  // all loads of the tile in flight before the first one is used
#pragma unroll
    for (int r = 0; r < BK_ROUNDS; ++r) {
      const int64_t ci = first + tile * BK_TILE_CHUNKS + r * BK_THREADS + tid;
      if (ci < first + nchunks) raw[r] = load_chunk_raw(buf, ci);
    }
#pragma unroll
    for (int r = 0; r < BK_ROUNDS; ++r) {
      const int64_t ci = first + tile * BK_TILE_CHUNKS + r * BK_THREADS + tid;
      if (ci < first + nchunks) {
        uint32_t code[CHUNK], counted;
        pack_decode_raw(raw[r], ci, first, k, kmask, pk_out, brk_out, code, counted);
        // the 16 rank requests first, the 16 stores after: the latency of a shared-memory atomic with a result is
        // paid once per chunk, not once per position
        uint32_t slot[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; ++j)
          slot[j] = (counted & (1u << j)) ? atomicAdd(&s_cnt[code[j] >> sub_bits], 1u) : 0u;
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
          if (counted & (1u << j)) {
            if (slot[j] < (uint32_t)BK_CAP) s_stage[(code[j] >> sub_bits) * BK_CAP + slot[j]] = (uint16_t)(code[j] & submask);
            else red_add_u32_keep(&counts[code[j]], 1u, keep);  // the row is full: direct reduction
          }
        }
        local += __popc(counted);
      }
    }
    __syncthreads();
    // every thread appends the FULL 4-entry granules (8 bytes) of four buckets' rows to the buckets' regions; the up
    // to three sub-keys left over move to the front of the row and wait for the next tile, so nothing is padded
    // until the CTA's last tile
    const bool last_tile = tile + gridDim.x >= ntiles;
    uint32_t fn[BPT], fg[BPT];
#pragma unroll
    for (int q = 0; q < BPT; ++q) {  // the reservations first: BPT independent atomics in flight
      const uint32_t b = (uint32_t)tid + (uint32_t)BK_THREADS * q;
      uint32_t n = s_cnt[b];
      if (n > (uint32_t)BK_CAP) n = BK_CAP;
      fn[q] = n;
      const uint32_t take = last_tile ? ((n + 3u) & ~3u) : (n & ~3u);
      fg[q] = take ? atomicAdd(&bk_cursor[b], take) : 0u;
    }
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
      const uint32_t b = (uint32_t)tid + (uint32_t)BK_THREADS * q;
      const uint32_t n = fn[q], g = fg[q];
      uint16_t *row = s_stage + b * BK_CAP;
      if (last_tile)
        for (uint32_t i = n; i < ((n + 3u) & ~3u); ++i) row[i] = (uint16_t)BK_PAD;
      const uint32_t take = last_tile ? ((n + 3u) & ~3u) : (n & ~3u);
      uint2 *rowv = reinterpret_cast<uint2 *>(row);
      if (take) {
        uint2 *dst = reinterpret_cast<uint2 *>(bk_buf + (size_t)b * gcap + g);
        if (g + take <= gcap) {
#pragma unroll
          for (int i = 0; i < BK_CAP / 4; ++i)
            if ((uint32_t)(4 * i) < take) dst[i] = rowv[i];
        } else {
          // the bucket's region is full (skewed spectrum): pad what is left of it, count these sub-keys directly
          if (g < gcap)
            for (uint32_t i = 0; i < (gcap - g) / 4; ++i) dst[i] = make_uint2(0xffffffffu, 0xffffffffu);
          for (uint32_t i = 0; i < take; ++i)
            if (row[i] != (uint16_t)BK_PAD) red_add_u32_keep(&counts[(b << sub_bits) | row[i]], 1u, keep);
        }
      }
      // leftover (n - take, only before the last tile) to the front of the row
      const uint32_t left = n - (take < n ? take : n);
      if (left && take) rowv[0] = rowv[take / 4];
      s_cnt[b] = left;
    }
    __syncthreads();
  }
  block_sum_to(local, nwords);
}

// phase 2: one bucket per CTA, counted in shared memory (dynamic, 4 << sub_bits bytes), added to the table slice
constexpr int BK_COUNT_THREADS = 512;
__global__ void __launch_bounds__(BK_COUNT_THREADS, 3) bucket_count_kernel(const uint16_t *__restrict__ bk_buf,
                                                                        const uint32_t *__restrict__ bk_cursor,
                                                                        uint32_t gcap, int sub_bits,
                                                                        int32_t *__restrict__ counts) {
  extern __shared__ __align__(16) unsigned char ks_dyn_smem[];
  uint32_t *s_tab = reinterpret_cast<uint32_t *>(ks_dyn_smem);
  const uint32_t b = blockIdx.x;
  const uint32_t entries = 1u << sub_bits;
  uint32_t n = bk_cursor[b];
  if (n > gcap) n = gcap;
  if (n == 0) return;
  const uint16_t *base = bk_buf + (size_t)b * gcap;  // gcap is a multiple of 8: 16-byte aligned
  const uint4 *src = reinterpret_cast<const uint4 *>(base);
  const uint32_t nv = n / 8;
  const uint4 filler = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  constexpr int UNR = 4;
  uint4 v[UNR];
  uint32_t i0 = threadIdx.x;
#pragma unroll
  for (int u = 0; u < UNR; ++u) v[u] = (i0 + u * BK_COUNT_THREADS < nv) ? __ldcs(src + i0 + u * BK_COUNT_THREADS) : filler;
  for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) s_tab[i] = 0;
  __syncthreads();
  auto add8 = [&](const uint4 &x) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const uint32_t lo = w[h] & 0xffffu, hi = w[h] >> 16;
      if (lo != BK_PAD) atomicAdd(&s_tab[lo], 1u);
      if (hi != BK_PAD) atomicAdd(&s_tab[hi], 1u);
    }
  };
  for (; i0 < nv; i0 += UNR * BK_COUNT_THREADS) {
    uint4 nx[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {  // the next batch is on its way while this one is counted
      const uint32_t j = i0 + (UNR + u) * BK_COUNT_THREADS;
      nx[u] = j < nv ? __ldcs(src + j) : filler;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) add8(v[u]);
#pragma unroll
    for (int u = 0; u < UNR; ++u) v[u] = nx[u];
  }
  if ((n & 4u) && threadIdx.x < 4) {  // n is a multiple of 4: one trailing granule
    const uint32_t s = base[nv * 8 + threadIdx.x];
    if (s != BK_PAD) atomicAdd(&s_tab[s], 1u);
  }
  __syncthreads();
  uint32_t *slice = reinterpret_cast<uint32_t *>(counts) + ((size_t)b << sub_bits);
  for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) {
    const uint32_t c = s_tab[i];
    if (c) slice[i] += c;  // the slice belongs to this CTA; direct reductions of phase 1 are already in it
  }
}

}  // namespace ks
