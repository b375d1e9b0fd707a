// unit_peaks.cu -- what the B200 sustains for the two random-access primitives this path is made of:
// 32-bit reductions (red.global.add) into, and 32-bit / 64-bit gathers from, a table of 4^k entries,
// with addresses as random as k-mer codes of a random genome.  These are the denominators of the
// "unit_rates" in bench.py's roofline object.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o unit_peaks tools/unit_peaks.cu && ./unit_peaks
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}

template <int PER>
__global__ void __launch_bounds__(256) red_kernel(uint32_t *table, uint32_t mask, uint64_t n_threads) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < PER; ++j) atomicAdd(&table[mix(t * PER + j) & mask], 1u);
  }
}

template <typename T, int PER>
__global__ void __launch_bounds__(256) gather_kernel(const T *__restrict__ table, uint32_t mask, uint64_t n_threads,
                                                     unsigned long long *sink) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
    T v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = __ldg(&table[mix(t * PER + j) & mask]);
#pragma unroll
    for (int j = 0; j < PER; ++j) acc += (unsigned long long)v[j];
  }
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
  const uint64_t N = 256ull << 20;  // operations per launch
  const int PER = 16;
  unsigned long long *sink; cudaMalloc(&sink, 8);
  printf("{\n");
  for (int k = 10; k <= 13; ++k) {
    size_t entries = (size_t)1 << (2 * k);
    uint32_t *t32; uint64_t *t64;
    cudaMalloc(&t32, entries * 4); cudaMalloc(&t64, entries * 8);
    cudaMemset(t32, 0, entries * 4); cudaMemset(t64, 0, entries * 8);
    uint32_t mask = (uint32_t)(entries - 1);
    int grid = 148 * 8;
    double r = time_ms([&] { red_kernel<PER><<<grid, 256>>>(t32, mask, N / PER); }, 5);
    double g4 = time_ms([&] { gather_kernel<uint32_t, PER><<<grid, 256>>>(t32, mask, N / PER, sink); }, 5);
    double g8 = time_ms([&] { gather_kernel<uint64_t, PER><<<grid, 256>>>(t64, mask, N / PER, sink); }, 5);
    printf(" \"k%d\": {\"table_entries\": %zu, \"red_u32_per_s\": %.4g, \"gather_u32_per_s\": %.4g, \"gather_u64_per_s\": %.4g}%s\n",
           k, entries, N / (r * 1e-3), N / (g4 * 1e-3), N / (g8 * 1e-3), k < 13 ? "," : "");
    cudaFree(t32); cudaFree(t64);
  }
  printf("}\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
