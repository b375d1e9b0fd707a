// ks_kernels.cuh -- hand-written sm_100a kernels of the kmer_spans hot path.
//
//   count_kernel        K1+K2: ASCII stream -> rolling 2-bit codes -> red.global.add into int32[4^k]
//                       (replaces sequence_kmer_count, /root/reference/src/kmer_spans.c:135-155)
//   wmax/wfx kernels    score table double[4^k] -> exact fixed-point table int64[4^k] (W - thr, :268)
//   scan_level_kernel   K4+K5+K6: per-position gather, max-plus scan with a decoupled look-back,
//                       excursion (start, leftmost peak, close) extraction, qualification and
//                       compacted emission (replaces kmer_regions, :243-307, one restart level per
//                       launch; level 0 is the dense pass over the whole buffer)
//   seg_build_kernel    qualifying excursions -> child segments [peak+1, close] (restart at peak, :281-283)
//   finalize_kernel     sorted records -> reference layout (seq_id, start, end | score, 0) (:95-99)
//   rank / lut kernels  K3: score-table derivation (rank_kmers_w :189-202, README.md:27-42 modes)
//
// No tensor-core work exists on this path (integer / byte / gather / atomic work); see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ks_chunk.cuh"

namespace ks {

#ifndef KS_TILE_THREADS
#define KS_TILE_THREADS 256
#endif
#ifndef KS_SCAN_MINBLOCKS
#define KS_SCAN_MINBLOCKS 2
#endif
constexpr int TILE_THREADS = KS_TILE_THREADS;
constexpr int TILE_WARPS = TILE_THREADS / 32;

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// streaming (evict-first) 128-bit load for the sequence: it is read once per pass and must not push
// the count / score table out of L2
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) { return __ldcs(p); }
// L2 evict-last policy for the tables that are gathered / reduced at random
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_add_u32_keep(int32_t *addr, uint32_t v, uint64_t pol) {
  asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(addr), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ldg_u32_keep(const uint32_t *addr, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(addr), "l"(pol));
  return v;
}
__device__ __forceinline__ int64_t ldg_s64_keep(const int64_t *addr, uint64_t pol) {
  int64_t v;
  asm volatile("ld.global.nc.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(addr), "l"(pol));
  return v;
}

__device__ __forceinline__ uint64_t fx_lo(fx_t v) { return (uint64_t)(unsigned __int128)v; }
__device__ __forceinline__ uint64_t fx_hi(fx_t v) { return (uint64_t)(((unsigned __int128)v) >> 64); }
__device__ __forceinline__ fx_t fx_make(uint64_t hi, uint64_t lo) {
  return (fx_t)((((unsigned __int128)hi) << 64) | (unsigned __int128)lo);
}

__device__ __forceinline__ fx_t shfl_fx(fx_t v, int src) {
  uint64_t lo = __shfl_sync(0xffffffffu, (unsigned long long)fx_lo(v), src);
  uint64_t hi = __shfl_sync(0xffffffffu, (unsigned long long)fx_hi(v), src);
  return fx_make(hi, lo);
}
__device__ __forceinline__ Xf shfl_xf(const Xf &f, int src) {
  Xf r;
  r.a = shfl_fx(f.a, src);
  r.b = shfl_fx(f.b, src);
  r.kill = __shfl_sync(0xffffffffu, f.kill, src);
  return r;
}
__device__ __forceinline__ Ex shfl_ex(const Ex &e, int src) {
  Ex r;
  r.M = shfl_fx(e.M, src);
  r.beg = __shfl_sync(0xffffffffu, (long long)e.beg, src);
  r.pk = __shfl_sync(0xffffffffu, (long long)e.pk, src);
  uint32_t fl = __shfl_sync(0xffffffffu, e.reset | (e.open << 1), src);
  r.reset = fl & 1u;
  r.open = (fl >> 1) & 1u;
  return r;
}

// ------------------------------------------------------------------------------------------
// K1 + K2: pack + count.  One thread = one 16-byte chunk of the ASCII buffer; grid-stride, consecutive
// threads read consecutive 16-byte vectors (512 B per warp request).  Writes the 2-bit packed codes
// (4 B per 16 bases) and the break mask (2 B per 16 bases) that every scan pass reads instead of the
// ASCII, and (kCount) reduces the k-mer ending at every position into the int32[4^k] table.
template <bool kCount>
__global__ void __launch_bounds__(256) pack_count_kernel(const uint8_t *__restrict__ buf, int64_t first,
                                                         int64_t nchunks, int k, uint32_t kmask,
                                                         uint32_t *__restrict__ pk_out,
                                                         uint16_t *__restrict__ brk_out,
                                                         int32_t *__restrict__ counts,
                                                         unsigned long long *__restrict__ nwords) {
  unsigned long long local = 0;
  const uint64_t keep = l2_policy_evict_last();
  for (int64_t ci = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ci < first + nchunks;
       ci += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t *p = buf + 16 * ci;  // chunk ci+1 of the buffer starts at p + 16
    const uint4 *v = reinterpret_cast<const uint4 *>(p);
    uint4 a = ld_stream_u4(v), b = ld_stream_u4(v + 1);
    uint32_t wp[4] = {a.x, a.y, a.z, a.w}, wc[4] = {b.x, b.y, b.z, b.w};
    uint32_t pkp, bp, np, pkc, bc, nc;
    pack16(wp, pkp, bp, np);
    pack16(wc, pkc, bc, nc);
    pk_out[ci + 1] = pkc;
    brk_out[ci + 1] = (uint16_t)bc;
    if (ci == 0) { pk_out[0] = pkp; brk_out[0] = (uint16_t)bp; }
    if (kCount) {
      uint32_t next = __ldg(p + 32);
      uint32_t code[CHUNK], counted;
      decode_count(((uint64_t)pkp << 32) | pkc, bp | (bc << 16), np | (nc << 16), next == 0u, k, kmask, code,
                   counted);
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (counted & (1u << j)) red_add_u32_keep(&counts[code[j]], 1u, keep);
      local += __popc(counted);
    }
  }
  if (kCount) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    __shared__ unsigned long long sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
      if (t) atomicAdd(nwords, t);
    }
  }
}

// ------------------------------------------------------------------------------------------
// scan parameters living in device memory (written by wfx_kernel, read by every scan launch)
struct DevScanParams {
  uint64_t min_width;
  uint64_t min_lo;
  int64_t min_hi;
  int32_t qs;
  int32_t err;             // 1: a weight is +inf or >= 2^40
  unsigned long long wmax_bits;  // bits of max finite |W - thr|
};

__global__ void __launch_bounds__(256) wmax_kernel(const double *__restrict__ W, size_t n, double thr,
                                                   DevScanParams *prm) {
  double m = 0.0;
  int err = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double w = W[i] - thr;
    if (w >= 0x1p40) err = 1;
    else if (w == w && w > -0x1p40) m = fmax(m, fabs(w));
  }
  unsigned long long bits = (unsigned long long)__double_as_longlong(m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long y = __shfl_down_sync(0xffffffffu, bits, o);
    bits = y > bits ? y : bits;
    err |= __shfl_down_sync(0xffffffffu, err, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (bits) atomicMax(&prm->wmax_bits, bits);
    if (err) atomicOr(&prm->err, 1);
  }
}

__global__ void __launch_bounds__(256) wfx_kernel(const double *__restrict__ W, size_t n, double thr,
                                                  int64_t *__restrict__ wfx, DevScanParams *prm,
                                                  uint64_t min_width, double min_score) {
  const int qs = qs_for_max(__longlong_as_double((long long)prm->wmax_bits));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    wfx[i] = wfx_from_double(W[i] - thr, qs);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    fx_t mu = fx_ceil_units(min_score, qs);
    prm->min_width = min_width;
    prm->min_lo = fx_lo(mu);
    prm->min_hi = (int64_t)fx_hi(mu);
    prm->qs = qs;
  }
}

// ------------------------------------------------------------------------------------------
// look-back descriptors: 16-byte self-validating words {tag, 96-bit payload}, written and read with
// single 128-bit accesses, so a reader needs ONE L2 round trip per window and no fences.
//   tag = epoch << 4 | open << 3 | kill/reset << 2 | state      state 1 = aggregate of this tile alone,
//                                                               2 = inclusive (everything to the left folded in)
// Buffers are zeroed once; a new epoch per launch makes older words read as "not ready".
// 96-bit payloads bound the running sums to |S| < 2^95 units, i.e. < 2^32 positions per scan.
struct TileState {
  uint4 *xfA;  // aggregate: a
  uint4 *xfB;  // aggregate: b (+kill in the tag)   | inclusive: S at the tile end
  uint4 *gA;   // the same pair per GROUP of 32 consecutive tiles (second look-back level)
  uint4 *gB;
  uint32_t *gdone;  // per group: tiles whose aggregate is published (zeroed before every launch)
  uint4 *exA;  // open-excursion aggregate of the tile: M
  uint4 *exB;  //                                       beg (48 bit) | pk (48 bit)
};
constexpr int GROUP_TILES = 32;

__device__ __forceinline__ uint4 ld_desc(const uint4 *p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_desc(uint4 *p, uint32_t tag, uint32_t hi, uint64_t lo) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(tag), "r"(hi),
               "r"((uint32_t)lo), "r"((uint32_t)(lo >> 32)) : "memory");
}
__device__ __forceinline__ void st_desc_fx(uint4 *p, uint32_t tag, fx_t v) {
  st_desc(p, tag, (uint32_t)fx_hi(v), fx_lo(v));
}
__device__ __forceinline__ fx_t desc_fx(const uint4 &d) {  // sign-extend the 96-bit payload
  return fx_make((uint64_t)(int64_t)(int32_t)d.y, ((uint64_t)d.w << 32) | d.z);
}
constexpr uint32_t TAG_AGG = 1u, TAG_INC = 2u, TAG_KILL = 4u, TAG_OPEN = 8u;
constexpr int64_t POS48_NONE = (1ll << 48) - 1;

// An excursion that entered a tile from the left and closes inside it needs the open-excursion state
// at the tile start.  The scan kernel never waits for it: the one thread per tile that needs it
// leaves this record, and ex_fixup_kernel resolves it after the launch from the per-tile aggregates.
struct __align__(16) ExPending {
  uint64_t m_lo;
  int64_t m_hi;   // max over the tile's positions before the close (fixed point)
  int64_t pk;     // its leftmost position
  int64_t c;      // close position
  uint32_t epoch; // valid iff == launch epoch
  uint32_t pad[3];
};

struct LevelArgs {
  ExPending *pending;  // one slot per tile
  const uint32_t *pk;   // packed 2-bit codes, one word per 16 positions (chunk c = positions [16c, 16c+16))
  const uint16_t *brk;  // break masks, one half-word per 16 positions
  int64_t ntiles;
  const int64_t *wfx;         // table mode: exact fixed-point score per k-mer (W - thr)
  // LUT mode (score is a function of the count): gather the int32 count (4 B/entry, L2 resident at
  // k <= 12), then the score from a dense count -> score table; counts >= lut_size use the sorted
  // sparse list (sp_count, sp_val)
  const uint32_t *counts;
  const int64_t *lut;
  uint32_t lut_size;
  const uint32_t *sp_count;
  const int64_t *sp_val;
  uint32_t sp_n;
  const DevScanParams *prm;
  int k;
  uint32_t kmask;
  // level >= 1: nseg segments, seg_chunk0 = exclusive prefix of their chunk counts (nseg + 1)
  int64_t nseg;
  const int64_t *seg_start;
  const int64_t *seg_len;
  const uint64_t *seg_chunk0;
  // level 0 (nseg == 0): dense pass over [dense_start, dense_start + 16 * total_chunks)
  int64_t dense_start;
  int64_t total_chunks;
  int32_t *inscan;  // or NULL
  TileState ts;
  uint32_t epoch;
  unsigned int *tile_counter;
  unsigned int tile_base;
  // emitted records (SoA), appended across levels
  int64_t *rec_beg, *rec_pk, *rec_c, *rec_mhi;
  uint64_t *rec_mlo;
  unsigned long long *rec_count;
  unsigned long long rec_cap;
  unsigned long long *dbg;  // KS_EXP_TIMING builds: per-phase cycle sums (thread 0 of every tile)
};

struct DevEmit {
  const LevelArgs *A;
  __device__ __forceinline__ void operator()(int64_t beg, int64_t pk, int64_t c, fx_t M) const {
    unsigned long long slot = atomicAdd(A->rec_count, 1ull);
    if (slot < A->rec_cap) {
      A->rec_beg[slot] = beg;
      A->rec_pk[slot] = pk;
      A->rec_c[slot] = c;
      A->rec_mhi[slot] = (int64_t)fx_hi(M);
      A->rec_mlo[slot] = fx_lo(M);
    }
  }
};

__device__ __forceinline__ Xf poll_xf(const uint4 *dA, const uint4 *dB, int64_t idx, uint32_t epoch) {
  Xf x;
  x.kill = 1; x.a = 0; x.b = 0;
  for (;;) {
    uint4 B = ld_desc(&dB[idx]);
    uint4 Aw = ld_desc(&dA[idx]);
    if ((B.x >> 4) != epoch) continue;
    if ((B.x & 3u) == TAG_INC) { x.b = desc_fx(B); break; }
    if ((B.x & 3u) == TAG_AGG && (Aw.x >> 4) == epoch && (Aw.x & 3u) == TAG_AGG) {
      x.kill = (B.x >> 2) & 1u;
      x.a = desc_fx(Aw);
      x.b = desc_fx(B);
      break;
    }
  }
  return x;
}
__device__ __forceinline__ void publish_xf(uint4 *dA, uint4 *dB, int64_t idx, uint32_t epoch, const Xf &f) {
  const uint32_t tg = epoch << 4;
  if (f.kill) {  // does not depend on anything to the left: final
    st_desc_fx(&dB[idx], tg | TAG_INC, f.b);
  } else {
    st_desc_fx(&dA[idx], tg | TAG_AGG, f.a);
    st_desc_fx(&dB[idx], tg | TAG_AGG, f.b);
  }
}
// ordered warp reduction: lane L holds the transform of an EARLIER range than lane L-1
__device__ __forceinline__ Xf warp_fold_xf(Xf x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Xf y = shfl_xf(x, (lane + o) & 31);
    if (lane + o < 32) x = xf_compose(y, x);
  }
  return shfl_xf(x, 0);
}

// Two-level decoupled look-back for the max-plus transform, run by warp 0: returns the state S at
// the start of `tile`.  Window 1 covers the earlier tiles of the own 32-tile group; if that is not
// enough, window 2+ walks whole groups, 32 per step, so ~1000 tiles of lag cost two windows.
// Group aggregates are published EARLY, by whichever tile of the group finishes its local phase last
// (publish_group_if_last); the last tile of a group later upgrades it to the inclusive value.
// A transform with kill set (an inclusive value, or an aggregate containing a reset) ends the walk
// because composition ignores everything left of it.
__device__ __forceinline__ fx_t lookback_xf(const TileState &ts, int64_t tile, uint32_t epoch,
                                            const Xf &agg, int lane) {
  const int64_t g = tile / GROUP_TILES;
  const int l = (int)(tile % GROUP_TILES);
  Xf x = xf_identity();  // (the tile's own aggregate was published at the end of its local phase)
  if (lane < l) x = poll_xf(ts.xfA, ts.xfB, tile - 1 - lane, epoch);
  Xf acc = warp_fold_xf(x, lane);  // tiles [32 g, tile)
  const Xf grp = xf_compose(acc, agg);  // tiles [32 g, tile]
  const bool last = (l == GROUP_TILES - 1);
  int64_t gbase = g - 1;
  while (!acc.kill) {
    int64_t idx = gbase - lane;
    Xf y;
    y.kill = 1; y.a = 0; y.b = 0;  // before the first tile the state is 0
    if (idx >= 0) y = poll_xf(ts.gA, ts.gB, idx, epoch);
    acc = xf_compose(warp_fold_xf(y, lane), acc);
    gbase -= 32;
  }
  const fx_t S_tile = acc.b;
  if (lane == 0) {
    const fx_t S_end = xf_apply(agg, S_tile);
    if (!agg.kill) st_desc_fx(&ts.xfB[tile], (epoch << 4) | TAG_INC, S_end);
    if (last && !grp.kill) st_desc_fx(&ts.gB[g], (epoch << 4) | TAG_INC, S_end);
  }
  return S_tile;
}

__device__ __forceinline__ void publish_ex(const TileState &ts, int64_t tile, uint32_t epoch, const Ex &e) {
  uint32_t tag = (epoch << 4) | (e.open ? TAG_OPEN : 0u) | (e.reset ? TAG_KILL : 0u) | TAG_AGG;
  fx_t M = e.M;
  const fx_t lo_lim = -(((fx_t)1) << 94);
  if (M < lo_lim) M = lo_lim;
  uint64_t b48 = (uint64_t)(e.beg < 0 ? POS48_NONE : e.beg) & 0xffffffffffffull;
  uint64_t p48 = (uint64_t)(e.pk < 0 ? POS48_NONE : e.pk) & 0xffffffffffffull;
  st_desc_fx(&ts.exA[tile], tag, M);
  st_desc(&ts.exB[tile], tag, (uint32_t)(b48 >> 16), (b48 << 48) | p48);
}

// Open-excursion state at the start of `tile`, resolved LAZILY: every tile publishes the aggregate
// of its own chunks (no chain), and only a tile in which an excursion that entered from the left
// closes walks back (warp 0, 32 tiles per step) to the tile that holds the excursion's start.
__device__ __forceinline__ Ex lookback_ex(const TileState &ts, int64_t tile, uint32_t epoch, int lane) {
  Ex acc = ex_identity();
  int64_t base = tile - 1;
  for (;;) {
    int64_t idx = base - lane;
    Ex x = ex_identity();
    x.reset = 1; x.open = 0;
    if (idx >= 0) {
      for (;;) {
        uint4 Aw = ld_desc(&ts.exA[idx]);
        uint4 B = ld_desc(&ts.exB[idx]);
        if ((Aw.x >> 4) != epoch || (Aw.x & 3u) == 0u || Aw.x != B.x) continue;
        x.M = desc_fx(Aw);
        uint64_t lo = ((uint64_t)B.w << 32) | B.z;
        int64_t b48 = (int64_t)(((uint64_t)B.y << 16) | (lo >> 48));
        int64_t p48 = (int64_t)(lo & 0xffffffffffffull);
        x.beg = b48 == POS48_NONE ? -1 : b48;
        x.pk = p48 == POS48_NONE ? -1 : p48;
        x.open = (Aw.x >> 3) & 1u;
        x.reset = (Aw.x >> 2) & 1u;
        break;
      }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      Ex y = shfl_ex(x, (lane + o) & 31);
      if (lane + o < 32) x = ex_combine(y, x);
    }
    acc = ex_combine(shfl_ex(x, 0), acc);
    if (acc.reset) return acc;
    base -= 32;
  }
}

// K4 + K5 + K6.  Persistent CTAs, tile = TILE_THREADS chunks (4096 positions at 256 threads), tile ids
// handed out in order by an atomic counter so that a tile only ever waits on tiles that already run.
// Software pipeline per CTA:   local(t1)  local(t2) finish(t1)  local(t3) finish(t2) ...
//   local : packed input -> codes -> table gather -> chunk transform -> block scan -> publish aggregate;
//           everything finish() needs is stashed in shared memory
//   finish: look-back for the state entering the tile (by now the predecessors' aggregates exist),
//           excursion walk, segmented scan of the open-excursion state, emission
// so the look-back latency of one tile hides behind the gather latency of the next.
struct Stash {
  int64_t s[CHUNK][TILE_THREADS];  // fixed-point scores, [j][thread]: conflict-free 8-byte accesses
  fx_t ea[TILE_THREADS];           // exclusive in-tile transform of the thread: a
  fx_t eb[TILE_THREADS];           //                                              b
  int64_t p0[TILE_THREADS];
  uint32_t flags[TILE_THREADS];    // live (16 bits) | head << 16 | excl.kill << 17
  Xf agg;                          // aggregate of the tile
  uint32_t group_last;             // this tile completed its 32-tile group
};

#if defined(KS_EXP_TIMING)
#define KS_T0 long long tk_ = clock64();
#define KS_TICK(slot) { long long n_ = clock64(); if (threadIdx.x == 0 && A.dbg) atomicAdd(A.dbg + (slot), (unsigned long long)(n_ - tk_)); tk_ = n_; }
#else
#define KS_T0
#define KS_TICK(slot)
#endif

template <bool kLut>
__device__ __forceinline__ void scan_local(const LevelArgs &A, int64_t tile, Stash &st, Xf *s_wxf) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KS_T0
  // ---- chunk -> position mapping ----
  const int64_t q = tile * TILE_THREADS + tid;
  int64_t p0 = 16;
  int n_in = 0;
  bool head = true;
  if (q < A.total_chunks) {
    if (A.nseg == 0) {
      p0 = A.dense_start + 16 * q; n_in = 16; head = (q == 0);
    } else {
      int64_t lo = 0, hi = A.nseg;  // largest s in [0, nseg) with seg_chunk0[s] <= q
      while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(&A.seg_chunk0[mid]) <= (uint64_t)q) lo = mid; else hi = mid;
      }
      int64_t c0 = (int64_t)__ldg(&A.seg_chunk0[lo]);
      int64_t sst = __ldg(&A.seg_start[lo]), ln = __ldg(&A.seg_len[lo]);
      p0 = sst + 16 * (q - c0);
      int64_t rem = sst + ln - p0;
      n_in = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
      head = (q == c0);
    }
  }
  // ---- packed window [p0 - 16, p0 + 16) ----
  const int64_t wq = p0 >> 4;
  const int r = (int)(p0 & 15);
  uint32_t hi32 = __ldg(&A.pk[wq - 1]), mid32 = __ldg(&A.pk[wq]);
  uint32_t b0 = __ldg(&A.brk[wq - 1]), b1 = __ldg(&A.brk[wq]);
  uint64_t X;
  uint32_t brk32;
  if (r == 0) {
    X = ((uint64_t)hi32 << 32) | mid32;
    brk32 = b0 | (b1 << 16);
  } else {
    uint32_t lo32 = __ldg(&A.pk[wq + 1]);
    uint32_t b2 = __ldg(&A.brk[wq + 1]);
    X = ((uint64_t)__funnelshift_l(mid32, hi32, 2 * r) << 32) | __funnelshift_l(lo32, mid32, 2 * r);
    uint64_t b48 = (uint64_t)b0 | ((uint64_t)b1 << 16) | ((uint64_t)b2 << 32);
    brk32 = (uint32_t)(b48 >> r);
  }
  // ---- codes + gather ----
  if (X == 0x123456789abcdefull) return;  // (keeps the loads above the tick in timing builds; never true)
  KS_TICK(0)
  uint32_t code[CHUNK], scored;
  decode_scan(X, brk32, A.k, A.kmask, n_in, code, scored);
  int64_t s[CHUNK];
  uint32_t live = 0;
  const uint64_t keep = l2_policy_evict_last();
  if (kLut) {
    uint32_t c[CHUNK];
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) c[j] = (scored & (1u << j)) ? ldg_u32_keep(&A.counts[code[j]], keep) : 0u;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {
      int64_t v = WFX_KILL;
      if (scored & (1u << j)) {
        if (c[j] < A.lut_size) {
          v = __ldg(&A.lut[c[j]]);
        } else {  // rare: very abundant k-mer, look it up in the sorted sparse list
          uint32_t lo = 0, hi = A.sp_n;
          while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&A.sp_count[mid]) <= c[j]) lo = mid; else hi = mid;
          }
          v = __ldg(&A.sp_val[lo]);
        }
      }
      s[j] = v;
    }
  } else {
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) s[j] = (scored & (1u << j)) ? ldg_s64_keep(&A.wfx[code[j]], keep) : WFX_KILL;
  }
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) {
    if (s[j] != WFX_KILL) live |= 1u << j; else s[j] = 0;
  }
  if (A.inscan) {
#pragma unroll
    for (int j = 0; j < CHUNK; ++j)
      if (scored & (1u << j)) atomicAdd(&A.inscan[code[j]], 1);
  }
  // ---- chunk transform + block scan ----
  KS_TICK(1)
  Xf f = chunk_transform(s, live);
  if (head) { fx_t v = xf_apply(f, 0); f.kill = 1; f.a = 0; f.b = v; }
  Xf inc = f;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Xf y = shfl_xf(inc, (lane - o) & 31);
    if (lane >= o) inc = xf_compose(y, inc);
  }
  Xf excl = shfl_xf(inc, (lane - 1) & 31);
  if (lane == 0) excl = xf_identity();
  if (lane == 31) s_wxf[warp] = inc;
  KS_TICK(2)
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) st.s[j][tid] = s[j];
  st.p0[tid] = p0;
  __syncthreads();
  KS_TICK(3)
  if (warp == 0) {  // exclusive scan of the warp totals by one warp: s_wxf[w] <- totals of warps < w
    Xf ti = lane < TILE_WARPS ? s_wxf[lane] : xf_identity();
#pragma unroll
    for (int o = 1; o < TILE_WARPS; o <<= 1) {
      Xf y = shfl_xf(ti, (lane - o) & 31);
      if (lane >= o) ti = xf_compose(y, ti);
    }
    Xf te = shfl_xf(ti, (lane - 1) & 31);
    if (lane == 0) te = xf_identity();
    if (lane < TILE_WARPS) s_wxf[lane] = te;
    if (lane == TILE_WARPS - 1) s_wxf[TILE_WARPS] = ti;  // aggregate of the tile
  }
  __syncthreads();
  excl = xf_compose(s_wxf[warp], excl);
  st.ea[tid] = excl.a;
  st.eb[tid] = excl.b;
  st.flags[tid] = live | (head ? 0x10000u : 0u) | (excl.kill ? 0x20000u : 0u);
  if (tid == 0) {
    Xf agg = s_wxf[TILE_WARPS];
    st.agg = agg;
    publish_xf(A.ts.xfA, A.ts.xfB, tile, A.epoch, agg);
    // the tile that completes a group publishes the group aggregate right away
    const int64_t g = tile / GROUP_TILES;
    int64_t gsize = A.ntiles - g * GROUP_TILES;
    if (gsize > GROUP_TILES) gsize = GROUP_TILES;
    uint32_t done = atomicAdd(&A.ts.gdone[g], 1u);
    st.group_last = (gsize == GROUP_TILES && done == (uint32_t)(GROUP_TILES - 1)) ? 1u : 0u;
  }
  __syncthreads();  // s_wxf may be overwritten by the next local phase; the stash is complete
  KS_TICK(4)
  if (st.group_last && warp == 0) {
    const int64_t g = tile / GROUP_TILES;
    Xf x = poll_xf(A.ts.xfA, A.ts.xfB, g * GROUP_TILES + (GROUP_TILES - 1 - lane), A.epoch);
    Xf grp = warp_fold_xf(x, lane);
    if (lane == 0) {
      uint4 cur = ld_desc(&A.ts.gB[g]);
      if (!((cur.x >> 4) == A.epoch && (cur.x & 3u) == TAG_INC)) publish_xf(A.ts.gA, A.ts.gB, g, A.epoch, grp);
    }
  }
}

__device__ __forceinline__ void scan_finish(const LevelArgs &A, int64_t tile, const Stash &st, Ex *s_wex,
                                            fx_t *s_S, const ScanParams &prm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KS_T0
  if (warp == 0) {
#if defined(KS_EXP_NO_LOOKBACK)
    fx_t S0 = 0;
#else
    fx_t S0 = lookback_xf(A.ts, tile, A.epoch, st.agg, lane);
#endif
    if (lane == 0) *s_S = S0;
  }
  KS_TICK(5)
  __syncthreads();
  const fx_t S_tile = *s_S;
#if defined(KS_EXP_NO_WALK)
  if (S_tile == 12345) A.rec_count[0] = 1;
  return;
#endif
  const uint32_t fl = st.flags[tid];
  const uint32_t live = fl & 0xffffu;
  const bool head = (fl & 0x10000u) != 0;
  const int64_t p0 = st.p0[tid];
  Xf excl;
  excl.a = st.ea[tid]; excl.b = st.eb[tid]; excl.kill = (fl >> 17) & 1u;
  const fx_t S_in = head ? (fx_t)0 : xf_apply(excl, S_tile);
  int64_t s[CHUNK];
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) s[j] = st.s[j][tid];

  // ---- excursions: local walk, segmented scan of the open-excursion state, lazy look-back ----
  DevEmit emit{&A};
  Ex ex;
  fx_t preM;
  int64_t prePk;
  int first_zero;
  KS_TICK(6)
  chunk_walk(s, live, S_in, p0, prm, emit, ex, preM, prePk, first_zero);
  KS_TICK(7)

  Ex einc = ex;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Ex y = shfl_ex(einc, (lane - o) & 31);
    if (lane >= o) einc = ex_combine(y, einc);
  }
  Ex eexcl = shfl_ex(einc, (lane - 1) & 31);
  if (lane == 0) eexcl = ex_identity();
  if (lane == 31) s_wex[warp] = einc;
  __syncthreads();
  if (warp == 0) {
    Ex ti = lane < TILE_WARPS ? s_wex[lane] : ex_identity();
#pragma unroll
    for (int o = 1; o < TILE_WARPS; o <<= 1) {
      Ex y = shfl_ex(ti, (lane - o) & 31);
      if (lane >= o) ti = ex_combine(y, ti);
    }
    Ex te = shfl_ex(ti, (lane - 1) & 31);
    if (lane == 0) te = ex_identity();
    if (lane < TILE_WARPS) s_wex[lane] = te;
    if (lane == TILE_WARPS - 1) publish_ex(A.ts, tile, A.epoch, ti);
  }
  __syncthreads();
  eexcl = ex_combine(s_wex[warp], eexcl);
  KS_TICK(8)
  if (!head && S_in > 0 && first_zero >= 0) {
    if (eexcl.reset) {
      // the entering excursion started inside this tile: everything is known
      chunk_finish_entering(S_in, eexcl, preM, prePk, first_zero, p0, prm, emit);
    } else {
      // it entered the tile from the left (at most one thread per tile gets here): defer
      fx_t M = eexcl.M;
      int64_t pk = eexcl.pk;
      if (preM > M) { M = preM; pk = prePk; }
      ExPending *pe = &A.pending[tile];
      pe->m_lo = fx_lo(M);
      pe->m_hi = (int64_t)fx_hi(M);
      pe->pk = pk;
      pe->c = p0 + first_zero;
      pe->epoch = A.epoch;
    }
  }
  KS_TICK(9)
}

// One warp per tile of the finished launch: resolve the deferred entering excursion (if any) by
// walking back over the per-tile open-excursion aggregates, 32 tiles per step, to the tile that holds
// the excursion's start.  Every aggregate of the launch is final by now, so nothing spins.
__global__ void __launch_bounds__(256) ex_fixup_kernel(const LevelArgs A) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= A.ntiles) return;
  const ExPending pe = A.pending[tile];
  if (pe.epoch != A.epoch) return;
  ScanParams prm;
  prm.min_width = A.prm->min_width;
  prm.min_units = fx_make((uint64_t)A.prm->min_hi, A.prm->min_lo);
  Ex E = lookback_ex(A.ts, tile, A.epoch, lane);
  if (lane == 0) {
    fx_t M = E.M;
    int64_t pk = E.pk;
    const fx_t Mp = fx_make((uint64_t)pe.m_hi, pe.m_lo);
    if (Mp > M) { M = Mp; pk = pe.pk; }
    DevEmit emit{&A};
    if (E.open && qualifies(prm, E.beg, pk, M)) emit(E.beg, pk, pe.c, M);
  }
}

template <bool kLut>
__global__ void __launch_bounds__(TILE_THREADS, KS_SCAN_MINBLOCKS) scan_level_kernel(const LevelArgs A) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  Stash *stash = reinterpret_cast<Stash *>(dyn_smem);  // two buffers
  __shared__ int64_t s_tile[2];
  __shared__ Xf s_wxf[TILE_WARPS + 1];
  __shared__ Ex s_wex[TILE_WARPS + 1];
  __shared__ fx_t s_S;
  const int tid = threadIdx.x;

  ScanParams prm;
  prm.min_width = A.prm->min_width;
  prm.min_units = fx_make((uint64_t)A.prm->min_hi, A.prm->min_lo);

  int slot = 0;
  auto next_tile = [&]() -> int64_t {
    KS_T0
    if (tid == 0) s_tile[slot] = (int64_t)(unsigned int)(atomicAdd(A.tile_counter, 1u) - A.tile_base);
    __syncthreads();
    int64_t t = s_tile[slot];
    slot ^= 1;  // a slot is rewritten only two fetches later, with whole phases (and barriers) in between
    KS_TICK(10)
    if (tid == 0 && A.dbg) atomicAdd(A.dbg + 15, 1ull);
    return t;
  };

  int cur = 0;
  int64_t tA = next_tile();
  if (tA < A.ntiles) scan_local<kLut>(A, tA, stash[0], s_wxf);
  while (tA < A.ntiles) {
    int64_t tB = next_tile();
    if (tB < A.ntiles) scan_local<kLut>(A, tB, stash[cur ^ 1], s_wxf);
    scan_finish(A, tA, stash[cur], s_wex, &s_S, prm);
    tA = tB;
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------
// records [r0, r0 + n) of the finished level -> child segments + their chunk counts
__global__ void __launch_bounds__(256) seg_build_kernel(const int64_t *__restrict__ rec_pk,
                                                        const int64_t *__restrict__ rec_c,
                                                        unsigned long long r0, unsigned long long n,
                                                        uint64_t min_width, int inscan,
                                                        int64_t *__restrict__ seg_start,
                                                        int64_t *__restrict__ seg_len,
                                                        uint64_t *__restrict__ seg_chunks) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t st, ln;
  bool ok = child_segment(rec_pk[r0 + i], rec_c[r0 + i], min_width, inscan != 0, st, ln);
  seg_start[i] = st;
  seg_len[i] = ok ? ln : 0;
  seg_chunks[i] = ok ? (uint64_t)segment_chunks(ln) : 0ull;
}

// ------------------------------------------------------------------------------------------
// sorted record order -> reference layout.  starts = nseq + 1 global offsets (ks_layout.h).
__global__ void __launch_bounds__(256) finalize_kernel(const uint32_t *__restrict__ perm, unsigned long long n,
                                                       const int64_t *__restrict__ rec_beg,
                                                       const int64_t *__restrict__ rec_pk,
                                                       const int64_t *__restrict__ rec_mhi,
                                                       const uint64_t *__restrict__ rec_mlo,
                                                       const int64_t *__restrict__ starts, int nseq,
                                                       const DevScanParams *__restrict__ prm,
                                                       int32_t *__restrict__ pos, double *__restrict__ score) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long r = perm ? perm[i] : i;
  int64_t beg = rec_beg[r], pk = rec_pk[r];
  int lo = 0, hi = nseq;  // largest s with starts[s] <= beg
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (starts[mid] <= beg) lo = mid; else hi = mid;
  }
  pos[3 * i] = lo;
  pos[3 * i + 1] = (int32_t)(beg - starts[lo]);
  pos[3 * i + 2] = (int32_t)(pk - starts[lo]);
  score[2 * i] = fx_to_double(fx_make((uint64_t)rec_mhi[r], rec_mlo[r]), prm->qs);
  score[2 * i + 1] = 0.0;
}

__global__ void __launch_bounds__(256) copy_keys_kernel(const int64_t *__restrict__ rec_beg,
                                                        unsigned long long n, uint64_t *__restrict__ keys) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (uint64_t)rec_beg[i];
}

// ------------------------------------------------------------------------------------------
// K3: score tables
__global__ void __launch_bounds__(256) max_u32_kernel(const uint32_t *__restrict__ v, size_t n,
                                                      uint32_t *__restrict__ out) {
  uint32_t m = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = max(m, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// run heads of the sorted counts -> (count, first sorted position), appended unordered
__global__ void __launch_bounds__(256) rle_heads_kernel(const uint32_t *__restrict__ keys, size_t n,
                                                        uint32_t *__restrict__ gcount,
                                                        uint32_t *__restrict__ gstart, uint32_t *ngroups,
                                                        uint32_t cap) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c = keys[i];
  if (i == 0 || keys[i - 1] != c) {
    uint32_t slot = atomicAdd(ngroups, 1u);
    if (slot < cap) { gcount[slot] = c; gstart[slot] = (uint32_t)i; }
  }
}

// rank of every k-mer from the linear pieces of ks_rankseg.h; p = position in the stable
// (count, index) order, vals[p] = k-mer index.  gstart has ngroups + 1 entries.
__global__ void __launch_bounds__(256) rank_eval_kernel(const uint32_t *__restrict__ vals, size_t n,
                                                        const uint32_t *__restrict__ gstart, uint32_t ngroups,
                                                        const uint32_t *__restrict__ seg_first,
                                                        const unsigned long long *__restrict__ seg_j0,
                                                        const double *__restrict__ seg_x0,
                                                        const double *__restrict__ seg_inc,
                                                        double *__restrict__ ranks) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t lo = 0, hi = ngroups;  // largest g with gstart[g] <= p
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(&gstart[mid]) <= p) lo = mid; else hi = mid;
  }
  unsigned long long j = p - __ldg(&gstart[lo]);
  uint32_t a = __ldg(&seg_first[lo]), b = __ldg(&seg_first[lo + 1]);
  while (b - a > 1) {
    uint32_t mid = (a + b) >> 1;
    if (__ldg(&seg_j0[mid]) <= j) a = mid; else b = mid;
  }
  ranks[vals[p]] = fma((double)(j - __ldg(&seg_j0[a])), __ldg(&seg_inc[a]), __ldg(&seg_x0[a]));
}

// frequency-of-counts histogram h[c] (the "histogram plus prefix sum" of the north star, used by the
// count-function modes, which need no per-k-mer order): block-private shared histogram for c < 4096
// with warp-aggregated updates, global atomics for 4096 <= c < dense, atomic append for c >= dense.
constexpr uint32_t FOC_SMEM_BINS = 4096;
__global__ void __launch_bounds__(256) foc_hist_kernel(const uint32_t *__restrict__ counts, size_t n,
                                                       uint32_t *__restrict__ hist, uint32_t dense,
                                                       uint32_t *__restrict__ big_vals, uint32_t *big_n,
                                                       uint32_t big_cap) {
  __shared__ uint32_t sh[FOC_SMEM_BINS];
  for (int i = threadIdx.x; i < (int)FOC_SMEM_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint32_t lt = (1u << (threadIdx.x & 31)) - 1u;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t nround = (n + stride - 1) / stride;  // uniform trip count: __match_any_sync needs the whole warp
  for (size_t it = 0; it < nround; ++it) {
    size_t i = it * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = i < n;
    uint32_t c = ok ? counts[i] : 0xffffffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, c);
    if (ok && (peers & lt) == 0u) {  // lowest lane of each group of equal counts
      uint32_t m = __popc(peers);
      if (c < FOC_SMEM_BINS) atomicAdd(&sh[c], m);
      else if (c < dense) atomicAdd(&hist[c], m);
      else {
        uint32_t slot = atomicAdd(big_n, m);
        for (uint32_t q = 0; q < m; ++q)
          if (slot + q < big_cap) big_vals[slot + q] = c;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (int)FOC_SMEM_BINS; i += blockDim.x)
    if (sh[i] && (uint32_t)i < dense) atomicAdd(&hist[i], sh[i]);
}

// W[x] = lut[group of counts[x]]  (log2 / +-1 / any pure function of the count)
__global__ void __launch_bounds__(256) lut_apply_kernel(const uint32_t *__restrict__ counts, size_t n,
                                                        const uint32_t *__restrict__ gcount, uint32_t ngroups,
                                                        const double *__restrict__ lut, double *__restrict__ W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c = counts[i];
  uint32_t lo = 0, hi = ngroups;  // largest g with gcount[g] <= c (exists: c is one of them)
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(&gcount[mid]) <= c) lo = mid; else hi = mid;
  }
  W[i] = __ldg(&lut[lo]);
}

// dense count -> score table in exact fixed point: lut[c] = fx(score(c) - thr); entries of counts that
// do not occur are never gathered.  gval = score per distinct count (host libm), gcount ascending.
__global__ void __launch_bounds__(256) lut_build_kernel(const uint32_t *__restrict__ gcount,
                                                        const double *__restrict__ gval, uint32_t ngroups,
                                                        double thr, int qs, uint32_t lut_size,
                                                        int64_t *__restrict__ lut, uint32_t *__restrict__ sp_count,
                                                        int64_t *__restrict__ sp_val, uint32_t sp_first) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  int64_t v = wfx_from_double(gval[g] - thr, qs);
  uint32_t c = gcount[g];
  if (c < lut_size) lut[c] = v;
  else { sp_count[g - sp_first] = c; sp_val[g - sp_first] = v; }
}

__global__ void __launch_bounds__(256) affine_kernel(double *__restrict__ W, size_t n, double sub, double div) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) W[i] = (W[i] - sub) / div;
}

__global__ void __launch_bounds__(256) fill_nan_kernel(double *__restrict__ W, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) W[i] = i == 0 ? 0.0 : __longlong_as_double(0x7ff8000000000000ll);
}

}  // namespace ks
