// unit_peaks3.cu -- random 8-byte gathers from a 32 MiB table: registers (ld.global.nc) against cp.async into shared
// memory (8-byte .ca copies, 16-byte .cg copies), one batch of 16 per thread or double-buffered.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o unit_peaks3 tools/unit_peaks3.cu && ./unit_peaks3
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
constexpr int PER = 16, THREADS = 128;

__global__ void __launch_bounds__(THREADS) ldg_kernel(const uint2 *__restrict__ table, uint32_t mask, uint64_t n,
                                                      unsigned long long *sink) {
  unsigned long long acc = 0;
  for (uint64_t t = (uint64_t)blockIdx.x * THREADS + threadIdx.x; t < n; t += (uint64_t)gridDim.x * THREADS) {
    uint2 v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = __ldg(&table[mix(t * PER + j) & mask]);
#pragma unroll
    for (int j = 0; j < PER; ++j) acc += v[j].x ^ v[j].y;
  }
  if (acc == 0x1234567ull) *sink = acc;
}

// MODE 0: cp.async.ca 8 B; MODE 1: cp.async.cg 16 B (the aligned pair of records, the wanted half picked later)
template <int MODE, int STAGES>
__global__ void __launch_bounds__(THREADS) cpasync_kernel(const uint2 *__restrict__ table, uint32_t mask, uint64_t n,
                                                          unsigned long long *sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int ESZ = MODE ? 16 : 8;
  unsigned long long acc = 0;
  auto slot = [&](int st, int j) { return smem + ((size_t)(st * PER + j) * THREADS + threadIdx.x) * ESZ; };
  auto issue = [&](int st, uint64_t t) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const uint32_t idx = mix(t * PER + j) & mask;
      const uint32_t s = (uint32_t)__cvta_generic_to_shared(slot(st, j));
      if (MODE == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(&table[idx]) : "memory");
      else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(&table[idx & ~1u]) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const uint64_t stride = (uint64_t)gridDim.x * THREADS;
  uint64_t t = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
  if (STAGES == 2 && t < n) issue(0, t);
  for (int it = 0; t < n; t += stride, ++it) {
    int st = 0;
    if (STAGES == 2) {
      st = it & 1;
      if (t + stride < n) issue(st ^ 1, t + stride);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      issue(0, t);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (MODE == 0) { const uint2 v = *reinterpret_cast<const uint2 *>(slot(st, j)); acc += v.x ^ v.y; }
      else { const uint4 v = *reinterpret_cast<const uint4 *>(slot(st, j)); acc += v.x ^ v.y ^ v.z ^ v.w; }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (acc == 0x1234567ull) *sink = acc;
}

template <typename F>
static double time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t foot = 32u << 20;
  void *tab; unsigned long long *sink;
  cudaMalloc(&tab, foot); cudaMemset(tab, 1, foot); cudaMalloc(&sink, 8);
  const uint32_t mask = (uint32_t)(foot / 8 - 1);
  const uint64_t N = 1ull << 27, nthreads = N / PER;  // 134 M gathers
  cudaFuncSetAttribute(cpasync_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(cpasync_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(cpasync_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(cpasync_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  printf("{\n");
  for (int per_sm : {2, 4, 6, 8, 12, 16}) {
    const int grid = 148 * per_sm;
    double a = time_ms([&] { ldg_kernel<<<grid, THREADS>>>((const uint2 *)tab, mask, nthreads, sink); }, 5);
    double b1 = time_ms([&] { cpasync_kernel<0, 1><<<grid, THREADS, PER * THREADS * 8>>>((const uint2 *)tab, mask, nthreads, sink); }, 5);
    double b2 = time_ms([&] { cpasync_kernel<0, 2><<<grid, THREADS, 2 * PER * THREADS * 8>>>((const uint2 *)tab, mask, nthreads, sink); }, 5);
    double c1 = time_ms([&] { cpasync_kernel<1, 1><<<grid, THREADS, PER * THREADS * 16>>>((const uint2 *)tab, mask, nthreads, sink); }, 5);
    double c2 = per_sm * 2 * PER * THREADS * 16 <= 220 * 1024
                    ? time_ms([&] { cpasync_kernel<1, 2><<<grid, THREADS, 2 * PER * THREADS * 16>>>((const uint2 *)tab, mask, nthreads, sink); }, 5)
                    : 0;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
    printf(" \"ctas_per_sm_%d\": {\"ldg_G_per_s\": %.4g, \"cpasync_ca8\": %.4g, \"cpasync_ca8_2stage\": %.4g, \"cpasync_cg16\": %.4g, \"cpasync_cg16_2stage\": %.4g},\n",
           per_sm, N / (a * 1e6), N / (b1 * 1e6), N / (b2 * 1e6), N / (c1 * 1e6), c2 > 0 ? N / (c2 * 1e6) : 0.0);
    fflush(stdout);
  }
  printf(" \"unit\": \"1e9 gathers per second, 32 MiB table of 8-byte records, 16 per thread and batch, 128 threads per CTA\"\n}\n");
  return 0;
}
