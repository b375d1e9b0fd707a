// ks_count.cuh -- shared-memory-privatised k-mer counting (K2), replacing sequence_kmer_count
// (/root/reference/src/kmer_spans.c:135-155) where one global reduction per base is not the fastest way.
//
// Measured on a B200 (tools/unit_peaks2.cu, profiles/r02_unit_peaks2.json): random 32-bit reductions into an
// L2-resident table run at 178 G/s chip-wide (the rate of the L2 atomic units), shared-memory atomics over a
// 64 KiB table at 1 630 G/s.  So:
//   k <= 7   pack_count_smem_kernel: the whole 4^k table lives in the shared memory of every CTA (<= 64 KiB),
//            one shared-memory atomic per base, one flush of the non-zero counters per CTA at the end.
//   8..13    two phases over 1024 buckets (4096 at k = 13), ONE SUB-KEY PER TWO K-MERS: the k-mers ending at
//            positions 2i and 2i+1 are a.c and c.b around the same (k-1)-mer c (the pairing the scan's core records
//            use), so the pair is filed under the leading 10 (12) bits of c and travels as one uint16 =
//            (a, rest of c, b) = 2k - 8 (2k - 10) bits.
//            bucket_scatter_kernel (fused with the 2-bit packing of K1) ranks every pair of a 24 576-position tile
//            inside its bucket with one shared-memory atomic, stages the sub-key in shared memory and appends each
//            bucket's full 8-byte granules to the bucket's region in HBM (1 B written + 1 B read per base instead of
//            a 32-byte-sector L2 atomic per base); bucket_count_kernel then counts one bucket per CTA in two
//            shared-memory tables of 4^k / buckets entries -- c.b lands in the bucket's own slice of the count table,
//            a.c in a second table owned by the same CTA -- and bucket_fold_kernel adds the second table in.
//            A pair that does not fit its staging row (repeats, skewed spectra) or its bucket's region, and a
//            k-mer without a partner (run boundaries), falls back to the direct global reduction: every input stays exact.
//   >= 14    the direct kernel of ks_kernels.cuh in slices of the table (the table exceeds L2).
#pragma once
#include "ks_kernels.cuh"
#include "ks_pairgeom.h"

namespace ks {

constexpr int BK_CAP = 24;      // staged sub-keys per bucket and tile; rows of 48 bytes (8-byte aligned)
constexpr int BK_BATCH = 3;
constexpr uint32_t BK_PAD = 0xffffu;  // filler of the last granule of a row; never a sub-key: an all-ones sub-key would
                                      // need k = 12 with rest, a, b all ones, which is filed as two direct reductions
constexpr size_t bk_scatter_smem(int k) { return ((size_t)4 + (size_t)BK_CAP * 2) << bk_log(k); }  // counters + rows

__device__ __forceinline__ unsigned long long block_sum_to(unsigned long long local, unsigned long long *dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  __shared__ unsigned long long sm_total[32];
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
// 8 <= k <= 12, phase 1: pack + scatter one sub-key per pair of consecutive k-mers into the buckets.
//
// "does any of the 16 bytes break a run" (N, n, terminator): exact as a yes / no, a third of the work of the flags
__device__ __forceinline__ uint32_t any_break16(const uint32_t w[4]) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t v = w[i], x = (v | 0x20202020u) ^ 0x6e6e6e6eu;
    acc |= ((v - 0x01010101u) & ~v) | ((x - 0x01010101u) & ~x);
  }
  return acc & 0x80808080u;
}
__device__ __forceinline__ uint32_t pack16_codes(const uint32_t w[4]) {
  uint32_t pk = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) pk |= ((((w[i] >> 1) & 0x03030303u) * 0x40100401u) >> 24) << (24 - 8 * i);
  return pk;
}
// packed codes + (brk | nul << 16) of 16 bytes; the flags only where any_break16 saw one
__device__ __forceinline__ void pack16_lazy(const uint4 &raw, uint32_t &pk, uint32_t &flags) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  pk = pack16_codes(w);
  flags = 0;
  if (any_break16(w)) {
    uint32_t pk2, brk, nul;
    pack16(w, pk2, brk, nul);
    flags = brk | (nul << 16);
  }
}

template <int K>
__global__ void __launch_bounds__(bk_threads(K), K >= 13 ? 1 : 4) bucket_scatter_kernel(const uint8_t *__restrict__ buf, int64_t first,
                                                                      int64_t nchunks, uint32_t *__restrict__ pk_out,
                                                                      uint16_t *__restrict__ brk_out,
                                                                      int32_t *__restrict__ counts,
                                                                      unsigned long long *__restrict__ nwords,
                                                                      uint16_t *__restrict__ bk_buf,
                                                                      uint32_t *__restrict__ bk_cursor, uint32_t gcap) {
  typedef PairGeom<K> G;
  extern __shared__ __align__(16) unsigned char ks_dyn_smem[];
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(ks_dyn_smem);
  unsigned char *s_cnt_b = ks_dyn_smem;
  uint16_t *s_stage = reinterpret_cast<uint16_t *>(ks_dyn_smem + G::NB * 4);
  unsigned char *s_stage_b = ks_dyn_smem + G::NB * 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t keep = l2_policy_evict_last();
  const int64_t end = first + nchunks;
  unsigned long long local = 0;
  const int64_t ntiles = (nchunks + G::TILE_CHUNKS - 1) / G::TILE_CHUNKS;
  constexpr int BPT = G::NB / G::THREADS;  // buckets per thread (flush)
  for (int i = tid; i < G::NB; i += G::THREADS) s_cnt[i] = 0;
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // a warp owns 32 * BK_ROUNDS consecutive chunks of the tile, 32 per round: the 16 positions before a chunk are
    // the neighbouring lane's chunk (one shuffle), those before lane 0 the last lane's of the round before, and only
    // the very first of the warp is packed a second time
    const int64_t wbase = first + tile * G::TILE_CHUNKS + (int64_t)warp * (32 * BK_ROUNDS);
    uint32_t carry_pk = 0, carry_fl = 0;
    if (wbase < end) pack16_lazy(ld_stream_u4(reinterpret_cast<const uint4 *>(buf + 16 * wbase)), carry_pk, carry_fl);
#pragma unroll 1
    for (int batch = 0; batch < BK_ROUNDS / BK_BATCH; ++batch) {
      uint4 raw[BK_BATCH];  // all loads of the batch in flight before the first one is used
#pragma unroll
      for (int r = 0; r < BK_BATCH; ++r) {
        const int64_t ci = wbase + (batch * BK_BATCH + r) * 32 + lane;
        raw[r] = ci < end ? ld_stream_u4(reinterpret_cast<const uint4 *>(buf + 16 * (ci + 1))) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int r = 0; r < BK_BATCH; ++r) {
        const int64_t ci = wbase + (batch * BK_BATCH + r) * 32 + lane;
        const bool live = ci < end;
        uint32_t pkc, flc;
        pack16_lazy(raw[r], pkc, flc);
        uint32_t pkp = __shfl_up_sync(0xffffffffu, pkc, 1), flp = __shfl_up_sync(0xffffffffu, flc, 1);
        if (lane == 0) { pkp = carry_pk; flp = carry_fl; }
        carry_pk = __shfl_sync(0xffffffffu, pkc, 31);
        carry_fl = __shfl_sync(0xffffffffu, flc, 31);
        if (!live) continue;  // the lanes behind the end of the set: shuffled with, nothing else
        pk_out[ci + 1] = pkc;
        brk_out[ci + 1] = (uint16_t)flc;
        if (ci == first) { pk_out[ci] = pkp; brk_out[ci] = (uint16_t)flp; }  // no thread of its own (ks_kernels.cuh)
        uint32_t counted = 0xffffu;
        if (flp | flc) {  // a break in sight: the full rule of sequence_kmer_count (decode_count)
          const uint32_t brk32 = (flp & 0xffffu) | (flc << 16), nul32 = (flp >> 16) | (flc & 0xffff0000u);
          const uint32_t runk = run_ending(~brk32, K);
          const uint32_t head = runk & (brk32 << K);
          const uint32_t nulnext = (nul32 >> 1) | (__ldg(buf + 16 * ci + 32) == 0u ? 0x80000000u : 0u);
          counted = (runk & ~(head & nulnext)) >> 16;
        }
        local += __popc(counted);
        // pair i = the k-mers ending at positions 2i and 2i + 1 of the chunk = a.c and c.b inside Y_i
        const uint32_t paired = counted & (counted >> 1) & 0x5555u;
        uint32_t slot[CHUNK / 2];
        // the 8 rank requests first, the 8 stores after: the latency of a shared-memory atomic with a result is
        // paid once per chunk, not once per pair
#pragma unroll
        for (int i = 0; i < CHUNK / 2; ++i) {
          const uint32_t y = __funnelshift_r(pkc, pkp, 28 - 4 * i);
          slot[i] = 0;
          if (paired & (1u << (2 * i))) slot[i] = atomicAdd(reinterpret_cast<uint32_t *>(s_cnt_b + G::bucket4(y)), 1u);
        }
#pragma unroll
        for (int i = 0; i < CHUNK / 2; ++i) {
          if (!(paired & (1u << (2 * i)))) continue;
          const uint32_t y = __funnelshift_r(pkc, pkp, 28 - 4 * i);
          const uint32_t sub = G::sub(y);
          const bool staged = slot[i] < (uint32_t)BK_CAP;
          const bool filler_like = (2 * K - 2 - G::LOG + 4 == 16) && sub == BK_PAD;
          if (staged)
            *reinterpret_cast<uint16_t *>(s_stage_b + G::bucket4(y) * (BK_CAP / 2) + slot[i] * 2) =
                (uint16_t)(filler_like ? BK_PAD : sub);
          if (!staged || filler_like) {  // the row is full, or the one sub-key that reads as filler: directly
            red_add_u32_keep(&counts[(y >> 2) & G::KMASK], 1u, keep);
            red_add_u32_keep(&counts[y & G::KMASK], 1u, keep);
          }
        }
        if (counted != (paired | (paired << 1))) {  // k-mers without a partner (run boundaries): directly
          const uint32_t single = counted & ~(paired | (paired << 1));
#pragma unroll 1
          for (uint32_t m = single; m; m &= m - 1) {
            const int j = __ffs(m) - 1;
            red_add_u32_keep(&counts[__funnelshift_r(pkc, pkp, 30 - 2 * j) & G::KMASK], 1u, keep);
          }
        }
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
      const uint32_t b = (uint32_t)tid + (uint32_t)G::THREADS * q;
      uint32_t n = s_cnt[b];
      if (n > (uint32_t)BK_CAP) n = BK_CAP;
      fn[q] = n;
      const uint32_t take = last_tile ? ((n + 3u) & ~3u) : (n & ~3u);
      fg[q] = take ? atomicAdd(&bk_cursor[b], take) : 0u;
    }
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
      const uint32_t b = (uint32_t)tid + (uint32_t)G::THREADS * q;
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
            red_add_u32_keep(&counts[G::code_ac(b, sub)], 1u, keep);
            red_add_u32_keep(&counts[G::code_cb(b, sub)], 1u, keep);
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

// phase 2: one bucket per CTA.  Two shared-memory tables of 4^k / 1024 entries: tabB[sub & LOW] counts the k-mers
// c.b -- the bucket's own contiguous slice of the count table -- and tabA[sub >> 2] = [a][rest of c] the k-mers a.c,
// which belong to four other slices: they go, plainly stored, to the CTA's part of a second table that
// bucket_fold_kernel adds in afterwards (an entry of the count table would otherwise have two writers).
constexpr int BK_COUNT_THREADS = 1024;
template <int K>
__global__ void __launch_bounds__(BK_COUNT_THREADS, 1) bucket_count_kernel(const uint16_t *__restrict__ bk_buf,
                                                                        const uint32_t *__restrict__ bk_cursor,
                                                                        uint32_t gcap, int32_t *__restrict__ counts,
                                                                        uint32_t *__restrict__ table_a) {
  typedef PairGeom<K> G;
  extern __shared__ __align__(16) unsigned char ks_dyn_smem[];
  uint32_t *tabB = reinterpret_cast<uint32_t *>(ks_dyn_smem);
  uint32_t *tabA = tabB + G::ENTRIES;
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
  for (uint32_t i = threadIdx.x; i < 2 * G::ENTRIES; i += blockDim.x) tabB[i] = 0;
  __syncthreads();
  auto add1 = [&](uint32_t sub) {
    if (sub == BK_PAD) return;
    atomicAdd(&tabB[sub & G::LOW], 1u);
    atomicAdd(&tabA[sub >> 2], 1u);
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
  uint32_t *slice = reinterpret_cast<uint32_t *>(counts) + (size_t)b * G::ENTRIES;
  uint32_t *mineA = table_a + (size_t)b * G::ENTRIES;
  for (uint32_t i = threadIdx.x; i < G::ENTRIES; i += blockDim.x) {
    const uint32_t c = tabB[i];
    if (c) slice[i] += c;   // the slice belongs to this CTA; direct reductions of phase 1 are already in it
    mineA[i] = tabA[i];     // [a][rest] of this bucket
  }
}

// counts[a.c] += tableA[bucket(c)][a][rest(c)], four consecutive c per thread (same a, same bucket for k >= 8)
template <int K>
__global__ void __launch_bounds__(256) bucket_fold_kernel(int32_t *__restrict__ counts, const uint32_t *__restrict__ table_a) {
  typedef PairGeom<K> G;
  const size_t n4 = ((size_t)1 << (2 * K)) / 4;
  int4 *c4 = reinterpret_cast<int4 *>(counts);
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
    const uint4 add = __ldcs(reinterpret_cast<const uint4 *>(&table_a[G::fold_index((uint32_t)(q * 4))]));
    if (add.x | add.y | add.z | add.w) {
      int4 v = c4[q];
      v.x += (int)add.x; v.y += (int)add.y; v.z += (int)add.z; v.w += (int)add.w;
      c4[q] = v;
    }
  }
}

}  // namespace ks
