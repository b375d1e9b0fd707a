// ks_sort.cuh -- device-wide exclusive scan and a stable LSD radix sort (8-bit digits), hand
// written for sm_100a.  Used for
//   * the stable (count, index) order of the 4^k count table that replaces qsort_r at
//     /root/reference/src/kmer_spans.c:191-197 (keys = counts, values = k-mer indices)
//   * ordering the emitted spans by start position (keys = global start, values = record id)
//   * prefix sums of per-segment chunk counts and of per-block digit histograms
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

// ------------------------------------------------------------------------------------------
// exclusive scan: out[i] = sum_{j<i} in[j], out[n] = total.  Three small kernels, no spinning.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T &total, T *smem /* >= 33 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0) {
    T w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : (T)0;
    T s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    smem[lane] = s - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = s;
  }
  __syncthreads();
  T res = x - v + smem[warp];
  total = smem[32];
  __syncthreads();
  return res;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const TIn *__restrict__ in, size_t n,
                                                                 TOut *__restrict__ sums) {
  __shared__ TOut sm[33];
  size_t base = (size_t)blockIdx.x * SCAN_TILE;
  TOut acc = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) acc += (TOut)in[idx];
  }
  TOut total;
  block_exclusive_scan<TOut>(acc, total, sm);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one block: in-place exclusive scan of sums[0..m), sums[m] = grand total
template <typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_inplace(TOut *sums, size_t m) {
  __shared__ TOut sm[33];
  TOut carry = 0;
  for (size_t base = 0; base < m; base += SCAN_THREADS) {
    size_t idx = base + threadIdx.x;
    TOut v = idx < m ? sums[idx] : (TOut)0;
    TOut total;
    TOut ex = block_exclusive_scan<TOut>(v, total, sm);
    if (idx < m) sums[idx] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) sums[m] = carry;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const TIn *in, size_t n,
                                                            const TOut *__restrict__ sums, size_t nblocks,
                                                            TOut *out /* may alias in */) {
  __shared__ TOut sm[33];
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  TOut v[SCAN_ITEMS];
  TOut acc = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    size_t idx = base + i;
    v[i] = idx < n ? (TOut)in[idx] : (TOut)0;
    acc += v[i];
  }
  TOut total;
  TOut ex = block_exclusive_scan<TOut>(acc, total, sm) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    size_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
  if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) out[n] = sums[nblocks];
}

// scratch: (nblocks + 1) TOut.  out has n + 1 entries.  Returns launches issued.
template <typename TIn, typename TOut>
static inline int exclusive_scan(const TIn *d_in, size_t n, TOut *d_out, TOut *d_scratch,
                                 cudaStream_t st) {
  if (n == 0) {
    cudaMemsetAsync(d_out, 0, sizeof(TOut), st);
    return 0;
  }
  size_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_block_sums<TIn, TOut><<<(unsigned)nblocks, SCAN_THREADS, 0, st>>>(d_in, n, d_scratch);
  scan_sums_inplace<TOut><<<1, SCAN_THREADS, 0, st>>>(d_scratch, nblocks);
  scan_apply<TIn, TOut><<<(unsigned)nblocks, SCAN_THREADS, 0, st>>>(d_in, n, d_scratch, nblocks, d_out);
  return 3;
}
static inline size_t exclusive_scan_scratch_elems(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 2; }

// the same scan over `rows` independent arrays at once (row r: in + r * in_stride -> out + r * out_stride,
// n + 1 outputs each): three launches in total instead of three per row.  scratch: rows * scratch_elems(n).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_rows(const TIn *__restrict__ in, size_t n,
                                                                      size_t in_stride, TOut *__restrict__ sums,
                                                                      size_t sums_stride) {
  __shared__ TOut sm[33];
  in += (size_t)blockIdx.y * in_stride;
  sums += (size_t)blockIdx.y * sums_stride;
  size_t base = (size_t)blockIdx.x * SCAN_TILE;
  TOut acc = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) acc += (TOut)in[idx];
  }
  TOut total;
  block_exclusive_scan<TOut>(acc, total, sm);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
template <typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_inplace_rows(TOut *sums, size_t m, size_t sums_stride) {
  __shared__ TOut sm[33];
  sums += (size_t)blockIdx.x * sums_stride;
  TOut carry = 0;
  for (size_t base = 0; base < m; base += SCAN_THREADS) {
    size_t idx = base + threadIdx.x;
    TOut v = idx < m ? sums[idx] : (TOut)0;
    TOut total;
    TOut ex = block_exclusive_scan<TOut>(v, total, sm);
    if (idx < m) sums[idx] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) sums[m] = carry;
}
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_rows(const TIn *in, size_t n, size_t in_stride,
                                                                 const TOut *__restrict__ sums, size_t sums_stride,
                                                                 size_t nblocks, TOut *out, size_t out_stride) {
  __shared__ TOut sm[33];
  in += (size_t)blockIdx.y * in_stride;
  sums += (size_t)blockIdx.y * sums_stride;
  out += (size_t)blockIdx.y * out_stride;
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  TOut v[SCAN_ITEMS];
  TOut acc = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    size_t idx = base + i;
    v[i] = idx < n ? (TOut)in[idx] : (TOut)0;
    acc += v[i];
  }
  TOut total;
  TOut ex = block_exclusive_scan<TOut>(acc, total, sm) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    size_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
  if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) out[n] = sums[nblocks];
}
template <typename TIn, typename TOut>
static inline int exclusive_scan_rows(const TIn *d_in, size_t n, int rows, size_t in_stride, TOut *d_out,
                                      size_t out_stride, TOut *d_scratch, cudaStream_t st) {
  if (n == 0 || rows <= 0) return 0;
  const size_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
  const size_t ss = exclusive_scan_scratch_elems(n);
  dim3 grid((unsigned)nblocks, (unsigned)rows);
  scan_block_sums_rows<TIn, TOut><<<grid, SCAN_THREADS, 0, st>>>(d_in, n, in_stride, d_scratch, ss);
  scan_sums_inplace_rows<TOut><<<(unsigned)rows, SCAN_THREADS, 0, st>>>(d_scratch, nblocks, ss);
  scan_apply_rows<TIn, TOut><<<grid, SCAN_THREADS, 0, st>>>(d_in, n, in_stride, d_scratch, ss, nblocks, d_out,
                                                             out_stride);
  return 3;
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort, one 8-bit digit per pass.
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 8;                       // items per thread
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;    // 2048 items per block

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) radix_hist(const K *__restrict__ keys, size_t n, int shift,
                                                          uint32_t *__restrict__ hist, uint32_t nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    size_t idx = base + (size_t)r * RS_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// warp w owns items [base + w*256, base + (w+1)*256) of the block's tile, visited 32 at a time in
// order, so ranks are stable.
template <typename K, bool kIotaVals>
__global__ void __launch_bounds__(RS_THREADS) radix_scatter(const K *__restrict__ keys_in,
                                                             const uint32_t *__restrict__ vals_in,
                                                             K *__restrict__ keys_out,
                                                             uint32_t *__restrict__ vals_out, size_t n,
                                                             int shift, const uint32_t *__restrict__ offs,
                                                             uint32_t nblocks) {
  __shared__ uint32_t wc[RS_WARPS][256];  // per-warp digit counts, then exclusive bases
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
  __syncthreads();
  size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)warp * (32 * RS_ROUNDS);
  K key[RS_ROUNDS];
  uint32_t rank[RS_ROUNDS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    size_t idx = wbase + (size_t)r * 32 + lane;
    bool ok = idx < n;
    key[r] = ok ? keys_in[idx] : (K)0;
    uint32_t d = ok ? ((uint32_t)(key[r] >> shift) & 255u) : 256u;  // 256 = "no item" class
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t prev = ok ? wc[warp][d] : 0u;
    __syncwarp();
    rank[r] = prev + __popc(peers & lt);
    if (ok && (peers & lt) == 0u) wc[warp][d] = prev + __popc(peers);  // lowest peer updates
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: exclusive scan of the 8 warp counts of digit d, plus the global offset
    uint32_t d = threadIdx.x;
    uint32_t run = offs[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c = wc[w][d];
      wc[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    size_t idx = wbase + (size_t)r * 32 + lane;
    if (idx < n) {
      uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
      uint32_t dst = wc[warp][d] + rank[r];
      keys_out[dst] = key[r];
      vals_out[dst] = kIotaVals ? (uint32_t)idx : vals_in[idx];
    }
  }
}

struct RadixScratch {
  uint32_t *hist;     // 256 * nblocks + 1
  uint32_t *scan_tmp; // exclusive_scan_scratch_elems(256 * nblocks)
};
static inline size_t radix_nblocks(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// Sorts by bits [0, nbits) of the key, stable.  keys_a/vals_a hold the input (vals_a ignored in the
// first pass when iota_vals: values are then the original indices); the result ends in
// (*keys_res, *vals_res), one of the two buffer pairs.  n < 2^32.  Returns launches issued.
template <typename K>
static inline int radix_sort_pairs(K *keys_a, uint32_t *vals_a, K *keys_b, uint32_t *vals_b, size_t n,
                                   int nbits, bool iota_vals, const RadixScratch &sc, cudaStream_t st,
                                   K **keys_res, uint32_t **vals_res) {
  int launches = 0;
  uint32_t nb = (uint32_t)radix_nblocks(n);
  K *kin = keys_a, *kout = keys_b;
  uint32_t *vin = vals_a, *vout = vals_b;
  bool first = true;
  if (nbits <= 0 || n == 0) nbits = (iota_vals && n) ? 1 : 0;  // iota still has to be materialised
  for (int shift = 0; shift < nbits; shift += 8) {
    radix_hist<K><<<nb, RS_THREADS, 0, st>>>(kin, n, shift, sc.hist, nb);
    launches += 1 + exclusive_scan<uint32_t, uint32_t>(sc.hist, (size_t)256 * nb, sc.hist, sc.scan_tmp, st);
    if (first && iota_vals)
      radix_scatter<K, true><<<nb, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, sc.hist, nb);
    else
      radix_scatter<K, false><<<nb, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, sc.hist, nb);
    ++launches;
    first = false;
    K *tk = kin; kin = kout; kout = tk;
    uint32_t *tv = vin; vin = vout; vout = tv;
  }
  *keys_res = kin;
  *vals_res = vin;
  return launches;
}

}  // namespace ks
