// ks_window.cuh -- sliding-window occurrence histograms of selected k-mers
// (replaces windowed_kmer_count_distributions, /root/reference/src/kmer_spans.c:398-449).
//
// The reference slides one window per run and keeps a whole 4^k count table up to date.  Restated
// for the GPU: the value of the window starting at s for a selected k-mer x is
//     P_x(s + window - 1) - P_x(s + k - 2),   P_x(p) = #{ e <= p : the k-mer ending at e is x }
// and the window exists iff no run break lies in [s, s + window).  P_x is kept as one 16-bit match
// mask + one 32-bit exclusive prefix per 16 positions (6 B / 16 positions / k-mer), read from the
// 2-bit packed sequence that the pack pass (K1) wrote -- the ASCII is not read again.
//
//   win_match_kernel   packed codes -> match masks + per-chunk match counts      (one thread / chunk)
//   (exclusive scans of the per-chunk counts: ks_sort.cuh)
//   win_hist_kernel    16 window starts per thread -> values -> histogram (shared memory when
//                      kmer_n x (window+1) bins fit, else global reductions) [-> per-position values]
//   win_fix_kernel     takes back the one window of every sequence whose length equals `window`
//                      (the reference leaves those sequences out, :775)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ks_chunk.cuh"

namespace ks {

constexpr int WIN_THREADS = 256;

// match[i * mstride + c] : bit j set iff the k-mer ENDING at position 16c + j is codes[i]
// cnt[i * nch + c]       : popcount of that mask; row kmer_n (first batch only) = breaks per chunk
__global__ void __launch_bounds__(WIN_THREADS) win_match_kernel(const uint32_t *__restrict__ pk,
                                                                const uint16_t *__restrict__ brk, int64_t nch,
                                                                int k, uint32_t kmask,
                                                                const uint32_t *__restrict__ codes, int kmer_n,
                                                                uint16_t *__restrict__ match, int64_t mstride,
                                                                uint8_t *__restrict__ cnt,
                                                                uint8_t *__restrict__ brk_cnt) {
  extern __shared__ uint32_t s_codes[];
  for (int i = threadIdx.x; i < kmer_n; i += blockDim.x) s_codes[i] = codes[i];
  __syncthreads();
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nch; c += (int64_t)gridDim.x * blockDim.x) {
    uint32_t pkp = c ? pk[c - 1] : 0u, pkc = pk[c];
    uint32_t bp = c ? brk[c - 1] : 0xffffu, bc = brk[c];
    const uint64_t X = ((uint64_t)pkp << 32) | pkc;
    const uint32_t valid = run_ending(~(bp | (bc << 16)), k) >> 16;  // all k positions inside one run
    uint32_t code[CHUNK];
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) code[j] = (uint32_t)(X >> (30 - 2 * j)) & kmask;
    for (int i = 0; i < kmer_n; ++i) {
      const uint32_t x = s_codes[i];
      uint32_t m = 0;
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) m |= (code[j] == x ? 1u : 0u) << j;
      m &= valid;
      match[(int64_t)i * mstride + c] = (uint16_t)m;
      cnt[(int64_t)i * nch + c] = (uint8_t)__popc(m);
    }
    if (brk_cnt) brk_cnt[c] = (uint8_t)__popc(bc);
  }
}

struct WinArgs {
  const uint16_t *match;   // kmer_n x mstride
  const uint32_t *pre;     // kmer_n x pstride : exclusive prefix of cnt
  const uint16_t *brk;     // break masks of the set
  const uint32_t *brk_pre; // exclusive prefix of breaks per chunk
  int64_t mstride, pstride, nch;
  int k, window, kmer_n;
  int32_t *hist;           // kmer_n x (window + 1), global
  int32_t *pos;            // NULL or kmer_n x pos_stride : value of the window starting at each position
  int64_t pos_stride;
  int use_smem;
};

// value of P(base + j), j = 0..15, from the two chunks the 16 positions touch
__device__ __forceinline__ uint32_t win_prefix_at(uint32_t pre, uint32_t m32, int off) {
  return pre + __popc(m32 & ((2u << off) - 1u));  // off <= 30
}

// gridDim.y = batches of WIN_KB selected k-mers: a CTA keeps only its batch's bins in shared memory, so
// several CTAs fit an SM
constexpr int WIN_KB = 4;
template <bool kSmem>  // compile-time address space of the bins: shared-memory reductions are native then
__global__ void __launch_bounds__(WIN_THREADS) win_hist_kernel(WinArgs A) {
  extern __shared__ int32_t s_hist[];
  const int bins = A.window + 1;
  const int i0 = blockIdx.y * WIN_KB;
  const int i1 = min(A.kmer_n, i0 + WIN_KB);
  if (kSmem) {
    for (int i = threadIdx.x; i < bins * (i1 - i0); i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
  }
  const int64_t npos = A.nch * 16;
  for (int64_t c = 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < A.nch;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s0 = c * 16;
    if (s0 + A.window > npos) continue;  // every window starting here runs off the buffer
    // breaks in [s, s + window): B(s + window - 1) - B(s - 1)
    const int64_t bh = s0 + A.window - 1, bl = s0 - 1;
    const int64_t ch = bh >> 4, cl = bl >> 4;
    const int oh = (int)(bh & 15), ol = (int)(bl & 15);
    const bool two_h = ch + 1 < A.nch;  // the second chunk of the upper end exists
    const uint32_t mh = A.brk[ch] | (two_h ? (uint32_t)A.brk[ch + 1] << 16 : 0xffff0000u);
    const uint32_t ml = A.brk[cl] | ((uint32_t)A.brk[cl + 1] << 16);
    uint32_t ok = 0;
    {
      // breaks inside the window of start j, updated incrementally: one bit enters, one leaves
      int32_t nb = (int32_t)(win_prefix_at(A.brk_pre[ch], mh, oh) - win_prefix_at(A.brk_pre[cl], ml, ol));
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) {
        if (j) nb += (int32_t)((mh >> (oh + j)) & 1u) - (int32_t)((ml >> (ol + j)) & 1u);
        ok |= ((s0 + j + A.window <= npos && nb == 0) ? 1u : 0u) << j;
      }
    }
    if (!ok) continue;
    const int64_t vl = s0 + A.k - 2;  // matches ending at or before s + k - 2 are outside the window
    const int64_t cvl = vl >> 4;
    const int ovl = (int)(vl & 15);
    for (int i = i0; i < i1; ++i) {
      const uint16_t *m = A.match + (int64_t)i * A.mstride;
      const uint32_t *p = A.pre + (int64_t)i * A.pstride;
      const uint32_t m_h = m[ch] | ((uint32_t)m[ch + 1] << 16);   // match rows carry two spare zero columns
      const uint32_t m_l = m[cvl] | ((uint32_t)m[cvl + 1] << 16);
      int32_t *h = kSmem ? s_hist + (i - i0) * bins : A.hist + (int64_t)i * bins;
      int32_t *po = A.pos ? A.pos + (int64_t)i * A.pos_stride + s0 : nullptr;
      // neighbouring windows mostly hold the same value: add runs of equal values at once
      int32_t v = (int32_t)(win_prefix_at(p[ch], m_h, oh) - win_prefix_at(p[cvl], m_l, ovl));
      if (ok == 0xffffu && !po) {
        // all 16 windows exist: the value changes only where the entering and the leaving match bit differ,
        // so walk the change points (about two per 16 starts) instead of the starts
        const uint32_t e = (m_h >> (oh + 1)) & 0x7fffu, l = (m_l >> (ovl + 1)) & 0x7fffu;  // bit j-1: start j
        const uint32_t up = e & ~l;
        uint32_t chg = e ^ l;
        int start = 0;
        while (chg) {
          const int b = __ffs(chg) - 1;
          chg &= chg - 1;
          atomicAdd(&h[v], b + 1 - start);
          start = b + 1;
          v += ((up >> b) & 1u) ? 1 : -1;
        }
        atomicAdd(&h[v], CHUNK - start);
        continue;
      }
      int32_t run_v = -1;
      int run_n = 0;
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) {
        if (j) v += (int32_t)((m_h >> (oh + j)) & 1u) - (int32_t)((m_l >> (ovl + j)) & 1u);
        if (!(ok & (1u << j))) continue;
        if (po) po[j] = v;
        if (v != run_v) {
          if (run_n) atomicAdd(&h[run_v], run_n);
          run_v = v;
          run_n = 0;
        }
        ++run_n;
      }
      if (run_n) atomicAdd(&h[run_v], run_n);
    }
  }
  if (kSmem) {
    __syncthreads();
    for (int i = threadIdx.x; i < bins * (i1 - i0); i += blockDim.x) {
      int32_t v = s_hist[i];
      if (v) atomicAdd(&A.hist[(int64_t)i0 * bins + i], v);
    }
  }
}

// one thread per (sequence whose length == window, selected k-mer): its only window [start, start + window)
// was added by win_hist_kernel iff it holds no break; take it back
__global__ void win_fix_kernel(WinArgs A, const int64_t *__restrict__ starts, int nfix) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nfix * A.kmer_n) return;
  const int i = t / nfix;
  const int64_t s = starts[t % nfix];
  const int64_t bh = s + A.window - 1, bl = s - 1;
  uint32_t nb = win_prefix_at(A.brk_pre[bh >> 4], A.brk[bh >> 4], (int)(bh & 15)) -
                win_prefix_at(A.brk_pre[bl >> 4], A.brk[bl >> 4], (int)(bl & 15));
  if (nb) return;
  const uint16_t *m = A.match + (int64_t)i * A.mstride;
  const uint32_t *p = A.pre + (int64_t)i * A.pstride;
  const int64_t vl = s + A.k - 2;
  uint32_t v = win_prefix_at(p[bh >> 4], m[bh >> 4], (int)(bh & 15)) -
               win_prefix_at(p[vl >> 4], m[vl >> 4], (int)(vl & 15));
  atomicSub(&A.hist[(int64_t)i * (A.window + 1) + v], 1);
  if (A.pos) A.pos[(int64_t)i * A.pos_stride + s] = 0;
}

}  // namespace ks
