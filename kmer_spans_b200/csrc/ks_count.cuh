// ks_count.cuh -- shared-memory-privatised k-mer counting (K2), replacing sequence_kmer_count
// (/root/reference/src/kmer_spans.c:135-155) where one global reduction per base is not the fastest way.
//
// Measured on a B200 (tools/unit_peaks2.cu, profiles/r02_unit_peaks2.json): random 32-bit reductions into an
// L2-resident table run at 178 G/s chip-wide (the rate of the L2 atomic units), shared-memory atomics over a
// 64 KiB table at 1 630 G/s.  So:
//   k <= 7   pack_count_smem_kernel: the whole 4^k table lives in the shared memory of every CTA (<= 64 KiB),
//            one shared-memory atomic per base, one flush of the non-zero counters per CTA at the end.
//   8..12    two phases over 1024 buckets, ONE SUB-KEY PER TWO K-MERS: the k-mers ending at positions 2i and 2i+1
//            are a.c and c.b around the same (k-1)-mer c (the pairing the scan's core records use), so the pair is
//            filed under the leading 10 bits of c and travels as one uint16 = (rest of c, a, b) = 2k - 8 bits.
//            bucket_scatter_kernel (fused with the 2-bit packing of K1) ranks every pair of a 24 576-position tile
//            inside its bucket with one shared-memory atomic, stages the sub-key in shared memory and appends each
//            bucket's full 8-byte granules to the bucket's region in HBM (1 B written + 1 B read per base instead of
//            a 32-byte-sector L2 atomic per base); bucket_count_kernel then counts one bucket per CTA in two
//            shared-memory tables of 4^k / 1024 entries -- c.b lands in the bucket's own slice of the count table,
//            a.c in a second table owned by the same CTA -- and bucket_fold_kernel adds the second table in.
//            A pair that does not fit its staging row (repeats, skewed spectra) or its bucket's region, and a
//            k-mer without a partner (run boundaries), falls back to the direct global reduction: every input stays exact.
//   >= 13    the direct kernel of ks_kernels.cuh in slices of the table (the table exceeds L2).
#pragma once
#include "ks_kernels.cuh"

namespace ks {

constexpr int BK_LOG = 10;
constexpr int BK_BUCKETS = 1 << BK_LOG;
constexpr int BK_CAP = 24;      // staged sub-keys per bucket and tile; rows of 48 bytes (8-byte aligned)
constexpr int BK_ROUNDS = 6;    // chunks per thread and tile, loaded in two batches of three
constexpr int BK_BATCH = 3;
constexpr int BK_THREADS = 256;
constexpr int BK_TILE_CHUNKS = BK_THREADS * BK_ROUNDS;  // 1536 chunks = 24 576 positions = 12 288 pairs, 12 per bucket
constexpr uint32_t BK_PAD = 0xffffu;  // filler of the last granule of a row; never a sub-key: an all-ones sub-key would
                                      // need k = 12 with rest, a, b all ones, which is filed as two direct reductions
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
// 8 <= k <= 12, phase 1: pack + scatter one sub-key per pair of consecutive k-mers into the buckets
struct PairGeom {   // geometry of a sub-key for this k
  int k;
  int core_bits;    // 2k - 2
  int rest_bits;    // core_bits - BK_LOG
  __host__ __device__ explicit PairGeom(int k_) : k(k_), core_bits(2 * k_ - 2), rest_bits(2 * k_ - 2 - BK_LOG) {}
};
__global__ void __launch_bounds__(BK_THREADS, 4) bucket_scatter_kernel(const uint8_t *__restrict__ buf, int64_t first,
                                                                      int64_t nchunks, int k, uint32_t kmask,
                                                                      uint32_t *__restrict__ pk_out,
                                                                      uint16_t *__restrict__ brk_out,
                                                                      int32_t *__restrict__ counts,
                                                                      unsigned long long *__restrict__ nwords,
                                                                      uint16_t *__restrict__ bk_buf,
                                                                      uint32_t *__restrict__ bk_cursor, uint32_t gcap) {
  extern __shared__ __align__(16) unsigned char ks_dyn_smem[];
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(ks_dyn_smem);
  uint16_t *s_stage = reinterpret_cast<uint16_t *>(ks_dyn_smem + BK_BUCKETS * 4);
  const int tid = threadIdx.x;
  const PairGeom G(k);
  const uint32_t cmask = kmask >> 2;                        // (k-1)-mer
  const uint32_t restmask = (1u << G.rest_bits) - 1u;
  const uint64_t keep = l2_policy_evict_last();
  unsigned long long local = 0;
  const int64_t ntiles = (nchunks + BK_TILE_CHUNKS - 1) / BK_TILE_CHUNKS;
  constexpr int BPT = BK_BUCKETS / BK_THREADS;  // buckets per thread (flush)
  for (int i = tid; i < BK_BUCKETS; i += BK_THREADS) s_cnt[i] = 0;
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
    for (int batch = 0; batch < BK_ROUNDS / BK_BATCH; ++batch) {
      RawChunk raw[BK_BATCH];  // all loads of the batch in flight before the first one is used
#pragma unroll
      for (int r = 0; r < BK_BATCH; ++r) {
        const int64_t ci = first + tile * BK_TILE_CHUNKS + (batch * BK_BATCH + r) * BK_THREADS + tid;
        if (ci < first + nchunks) raw[r] = load_chunk_raw(buf, ci);
      }
#pragma unroll
      for (int r = 0; r < BK_BATCH; ++r) {
        const int64_t ci = first + tile * BK_TILE_CHUNKS + (batch * BK_BATCH + r) * BK_THREADS + tid;
        if (ci >= first + nchunks) continue;
        uint32_t code[CHUNK], counted;
        pack_decode_raw(raw[r], ci, first, k, kmask, pk_out, brk_out, code, counted);
        // pairs (2i, 2i+1): code[2i] = a.c, code[2i+1] = c.b.  The 8 rank requests first, the 8 stores after: the
        // latency of a shared-memory atomic with a result is paid once per chunk, not once per pair
        uint32_t slot[CHUNK / 2];
#pragma unroll
        for (int i = 0; i < CHUNK / 2; ++i) {
          const bool both = ((counted >> (2 * i)) & 3u) == 3u;
          slot[i] = both ? atomicAdd(&s_cnt[(code[2 * i] & cmask) >> G.rest_bits], 1u) : 0u;
        }
#pragma unroll
        for (int i = 0; i < CHUNK / 2; ++i) {
          const uint32_t m2 = (counted >> (2 * i)) & 3u;
          if (m2 == 3u) {
            const uint32_t c = code[2 * i] & cmask;
            const uint32_t sub = ((c & restmask) << 4) | ((code[2 * i] >> G.core_bits) << 2) | (code[2 * i + 1] & 3u);
            if (slot[i] < (uint32_t)BK_CAP && sub != BK_PAD) {
              s_stage[(c >> G.rest_bits) * BK_CAP + slot[i]] = (uint16_t)sub;
            } else {  // the row is full (or the one sub-key that looks like filler): two direct reductions.  A slot
                      // taken for the filler look-alike stays unwritten only if sub == BK_PAD: write filler there
              if (slot[i] < (uint32_t)BK_CAP) s_stage[(c >> G.rest_bits) * BK_CAP + slot[i]] = (uint16_t)BK_PAD;
              red_add_u32_keep(&counts[code[2 * i]], 1u, keep);
              red_add_u32_keep(&counts[code[2 * i + 1]], 1u, keep);
            }
          } else {  // a k-mer without its partner (run boundary): direct
            if (m2 & 1u) red_add_u32_keep(&counts[code[2 * i]], 1u, keep);
            if (m2 & 2u) red_add_u32_keep(&counts[code[2 * i + 1]], 1u, keep);
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
          // the bucket's region is full (skewed spectrum): pad what is left of it, count these pairs directly
          if (g < gcap)
            for (uint32_t i = 0; i < (gcap - g) / 4; ++i) dst[i] = make_uint2(0xffffffffu, 0xffffffffu);
          for (uint32_t i = 0; i < take; ++i) {
            const uint32_t sub = row[i];
            if (sub == BK_PAD) continue;
            const uint32_t c = (b << G.rest_bits) | (sub >> 4);
            red_add_u32_keep(&counts[(((sub >> 2) & 3u) << G.core_bits) | c], 1u, keep);  // a.c
            red_add_u32_keep(&counts[(c << 2) | (sub & 3u)], 1u, keep);                    // c.b
          }
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

// phase 2: one bucket per CTA.  Two shared-memory tables of 2^(rest_bits + 2) entries: tabB[(rest, b)] counts the
// k-mers c.b -- the bucket's own contiguous slice of the count table -- and tabA[(a, rest)] the k-mers a.c, which
// belong to four other slices: they go, plainly stored, to the CTA's part of a second table that bucket_fold_kernel
// adds in afterwards (an entry of the count table would otherwise have two writers).
constexpr int BK_COUNT_THREADS = 1024;
__global__ void __launch_bounds__(BK_COUNT_THREADS, 1) bucket_count_kernel(const uint16_t *__restrict__ bk_buf,
                                                                        const uint32_t *__restrict__ bk_cursor,
                                                                        uint32_t gcap, int k,
                                                                        int32_t *__restrict__ counts,
                                                                        uint32_t *__restrict__ table_a) {
  extern __shared__ __align__(16) unsigned char ks_dyn_smem[];
  const PairGeom G(k);
  const uint32_t entries = 1u << (G.rest_bits + 2);
  uint32_t *tabB = reinterpret_cast<uint32_t *>(ks_dyn_smem);
  uint32_t *tabA = tabB + entries;
  const uint32_t b = blockIdx.x;
  uint32_t n = bk_cursor[b];
  if (n > gcap) n = gcap;
  const uint16_t *base = bk_buf + (size_t)b * gcap;  // gcap is a multiple of 8: 16-byte aligned
  const uint4 *src = reinterpret_cast<const uint4 *>(base);
  const uint32_t nv = n / 8;
  const uint4 filler = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  constexpr int UNR = 2;
  uint4 v[UNR];
  uint32_t i0 = threadIdx.x;
#pragma unroll
  for (int u = 0; u < UNR; ++u) v[u] = (i0 + u * BK_COUNT_THREADS < nv) ? __ldcs(src + i0 + u * BK_COUNT_THREADS) : filler;
  for (uint32_t i = threadIdx.x; i < 2 * entries; i += blockDim.x) tabB[i] = 0;
  __syncthreads();
  const uint32_t restmask = (1u << G.rest_bits) - 1u;
  auto add1 = [&](uint32_t sub) {
    if (sub == BK_PAD) return;
    const uint32_t rest = (sub >> 4) & restmask;
    atomicAdd(&tabB[(rest << 2) | (sub & 3u)], 1u);
    atomicAdd(&tabA[(((sub >> 2) & 3u) << G.rest_bits) | rest], 1u);
  };
  auto add8 = [&](const uint4 &x) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int h = 0; h < 4; ++h) { add1(w[h] & 0xffffu); add1(w[h] >> 16); }
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
  if ((n & 4u) && threadIdx.x < 4) add1(base[nv * 8 + threadIdx.x]);  // n is a multiple of 4: one trailing granule
  __syncthreads();
  uint32_t *slice = reinterpret_cast<uint32_t *>(counts) + ((size_t)b << (G.rest_bits + 2));
  uint32_t *mineA = table_a + ((size_t)b << (G.rest_bits + 2));
  for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) {
    const uint32_t c = tabB[i];
    if (c) slice[i] += c;   // the slice belongs to this CTA; direct reductions of phase 1 are already in it
    mineA[i] = tabA[i];     // [a][rest] of this bucket
  }
}

// counts[a.c] += tableA[bucket(c)][a][rest(c)]
__global__ void __launch_bounds__(256) bucket_fold_kernel(int32_t *__restrict__ counts, const uint32_t *__restrict__ table_a,
                                                          int k) {
  const PairGeom G(k);
  const size_t n = (size_t)1 << (2 * k);
  const uint32_t cmask = (uint32_t)(n >> 2) - 1u, restmask = (1u << G.rest_bits) - 1u;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) {
    const uint32_t a = (uint32_t)(x >> G.core_bits), c = (uint32_t)x & cmask;
    const uint32_t add = __ldcs(&table_a[((size_t)(c >> G.rest_bits) << (G.rest_bits + 2)) | ((size_t)a << G.rest_bits) | (c & restmask)]);
    if (add) counts[x] += (int32_t)add;
  }
}

}  // namespace ks
