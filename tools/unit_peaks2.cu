// unit_peaks2.cu -- round-2 unit rates: which hardware path carries a random table lookup fastest on a B200?
//   * ld.global.nc gathers of 2 / 4 / 8 / 16 bytes per lane (L1TEX t-stage: one wavefront per lane)
//   * texture fetches (tex1Dfetch) of the same element sizes, and a kernel that splits its lookups between both
//   * red.global.add.u32 / .u64, red.global.add.v4.f32, shared-memory atomics over a 64 KiB table
//   * shared-memory random reads (piece tables)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o unit_peaks2 tools/unit_peaks2.cu && ./unit_peaks2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ unsigned long long fold(unsigned short v) { return v; }
__device__ __forceinline__ unsigned long long fold(uint32_t v) { return v; }
__device__ __forceinline__ unsigned long long fold(uint2 v) { return v.x ^ v.y; }
__device__ __forceinline__ unsigned long long fold(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

template <typename T, int PER>
__global__ void __launch_bounds__(256) ldg_kernel(const T *__restrict__ table, uint32_t mask, uint64_t n_threads,
                                                  unsigned long long *sink) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
    T v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = __ldg(&table[mix(t * PER + j) & mask]);
#pragma unroll
    for (int j = 0; j < PER; ++j) acc += fold(v[j]);
  }
  if (acc == 0x1234567ull) *sink = acc;
}

template <typename T, int PER>
__global__ void __launch_bounds__(256) tex_kernel(cudaTextureObject_t tex, uint32_t mask, uint64_t n_threads,
                                                  unsigned long long *sink) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
    T v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = tex1Dfetch<T>(tex, (int)(mix(t * PER + j) & mask));
#pragma unroll
    for (int j = 0; j < PER; ++j) acc += fold(v[j]);
  }
  if (acc == 0x1234567ull) *sink = acc;
}

// TEXN of every PER lookups through the texture unit, the rest through ld.global
template <typename T, int PER, int TEXN>
__global__ void __launch_bounds__(256) mixed_kernel(const T *__restrict__ table, cudaTextureObject_t tex, uint32_t mask,
                                                    uint64_t n_threads, unsigned long long *sink) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
    T v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const uint32_t idx = mix(t * PER + j) & mask;
      if (j < TEXN) v[j] = tex1Dfetch<T>(tex, (int)idx);
      else v[j] = __ldg(&table[idx]);
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) acc += fold(v[j]);
  }
  if (acc == 0x1234567ull) *sink = acc;
}

template <int PER>
__global__ void __launch_bounds__(256) red32_kernel(uint32_t *table, uint32_t mask, uint64_t n_threads) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < PER; ++j) atomicAdd(&table[mix(t * PER + j) & mask], 1u);
  }
}
template <int PER>
__global__ void __launch_bounds__(256) red64_kernel(unsigned long long *table, uint32_t mask, uint64_t n_threads) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const uint32_t h = mix(t * PER + j);
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(&table[h & mask]),
                   "l"(1ull << (8 * (h >> 29))) : "memory");
    }
  }
}
template <int PER>
__global__ void __launch_bounds__(256) redv4_kernel(float4 *table, uint32_t mask, uint64_t n_threads) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const uint32_t h = mix(t * PER + j);
      const int a = h >> 30;
      asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(&table[h & mask]),
                   "f"(a == 0 ? 1.f : 0.f), "f"(a == 1 ? 1.f : 0.f), "f"(a == 2 ? 1.f : 0.f), "f"(a == 3 ? 1.f : 0.f)
                   : "memory");
    }
  }
}

// shared-memory histogram: ENTRIES u32 counters per CTA, random addresses, flushed at the end
template <int PER, int ENTRIES>
__global__ void __launch_bounds__(256) atoms_kernel(uint32_t *out, uint64_t n_threads) {
  extern __shared__ uint32_t sh[];
  for (int i = threadIdx.x; i < ENTRIES; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < PER; ++j) atomicAdd(&sh[mix(t * PER + j) & (ENTRIES - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ENTRIES; i += blockDim.x)
    if (sh[i]) atomicAdd(&out[i], sh[i]);
}

template <int PER, int ENTRIES>
__global__ void __launch_bounds__(256) lds_kernel(const uint32_t *src, uint64_t n_threads, unsigned long long *sink) {
  extern __shared__ uint32_t sh[];
  for (int i = threadIdx.x; i < ENTRIES; i += blockDim.x) sh[i] = src[i];
  __syncthreads();
  unsigned long long acc = 0;
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < PER; ++j) acc += sh[mix(t * PER + j) & (ENTRIES - 1)];
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

template <typename T>
static cudaTextureObject_t make_tex(const void *p, size_t bytes, cudaChannelFormatDesc d) {
  cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
  rd.resType = cudaResourceTypeLinear;
  rd.res.linear.devPtr = const_cast<void *>(p);
  rd.res.linear.desc = d;
  rd.res.linear.sizeInBytes = bytes;
  cudaTextureDesc td; memset(&td, 0, sizeof td);
  td.readMode = cudaReadModeElementType;
  td.filterMode = cudaFilterModePoint;
  td.addressMode[0] = cudaAddressModeClamp;
  cudaTextureObject_t t = 0;
  cudaError_t e = cudaCreateTextureObject(&t, &rd, &td, nullptr);
  if (e != cudaSuccess) { fprintf(stderr, "texture: %s\n", cudaGetErrorString(e)); return 0; }
  return t;
}

#define CHECK() do { cudaError_t e_ = cudaDeviceSynchronize(); if (e_ != cudaSuccess) { fprintf(stderr, "line %d: %s\n", __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
  const uint64_t N = 256ull << 20;  // lookups per launch
  const int PER = 16;
  const int grid = 148 * 8;
  unsigned long long *sink; cudaMalloc(&sink, 8);
  const size_t bytes = 64ull << 20;
  void *tab; cudaMalloc(&tab, bytes); cudaMemset(tab, 1, bytes);
  size_t maxw = 0;
  cudaChannelFormatDesc d16 = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
  cudaChannelFormatDesc d32 = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindUnsigned);
  cudaChannelFormatDesc d64 = cudaCreateChannelDesc(32, 32, 0, 0, cudaChannelFormatKindUnsigned);
  cudaChannelFormatDesc d128 = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindUnsigned);
  cudaDeviceGetTexture1DLinearMaxWidth(&maxw, &d16, 0);
  printf("{\n \"max_tex1d_linear_width_u16\": %zu,\n", maxw);
  // footprint 32 MiB for every element size (k = 12 class table / core table), 64 MiB for the 4-byte count table
  struct Case { const char *name; size_t foot; } cases[] = {{"32MiB", 32ull << 20}, {"64MiB", 64ull << 20}};
  for (auto &c : cases) {
    const uint32_t m16 = (uint32_t)(c.foot / 2 - 1), m32 = (uint32_t)(c.foot / 4 - 1), m64 = (uint32_t)(c.foot / 8 - 1),
                   m128 = (uint32_t)(c.foot / 16 - 1);
    cudaTextureObject_t t16 = make_tex<unsigned short>(tab, c.foot, d16), t32 = make_tex<uint32_t>(tab, c.foot, d32),
                        t64 = make_tex<uint2>(tab, c.foot, d64), t128 = make_tex<uint4>(tab, c.foot, d128);
    double a16 = time_ms([&] { ldg_kernel<unsigned short, PER><<<grid, 256>>>((const unsigned short *)tab, m16, N / PER, sink); }, 5);
    double a32 = time_ms([&] { ldg_kernel<uint32_t, PER><<<grid, 256>>>((const uint32_t *)tab, m32, N / PER, sink); }, 5);
    double a64 = time_ms([&] { ldg_kernel<uint2, PER><<<grid, 256>>>((const uint2 *)tab, m64, N / PER, sink); }, 5);
    double a128 = time_ms([&] { ldg_kernel<uint4, PER><<<grid, 256>>>((const uint4 *)tab, m128, N / PER, sink); }, 5);
    CHECK();
    double b16 = time_ms([&] { tex_kernel<unsigned short, PER><<<grid, 256>>>(t16, m16, N / PER, sink); }, 5);
    double b32 = time_ms([&] { tex_kernel<uint32_t, PER><<<grid, 256>>>(t32, m32, N / PER, sink); }, 5);
    double b64 = time_ms([&] { tex_kernel<uint2, PER><<<grid, 256>>>(t64, m64, N / PER, sink); }, 5);
    double b128 = time_ms([&] { tex_kernel<uint4, PER><<<grid, 256>>>(t128, m128, N / PER, sink); }, 5);
    CHECK();
    double c4 = time_ms([&] { mixed_kernel<uint2, PER, 4><<<grid, 256>>>((const uint2 *)tab, t64, m64, N / PER, sink); }, 5);
    double c8 = time_ms([&] { mixed_kernel<uint2, PER, 8><<<grid, 256>>>((const uint2 *)tab, t64, m64, N / PER, sink); }, 5);
    double c12 = time_ms([&] { mixed_kernel<uint2, PER, 12><<<grid, 256>>>((const uint2 *)tab, t64, m64, N / PER, sink); }, 5);
    double e8 = time_ms([&] { mixed_kernel<unsigned short, PER, 8><<<grid, 256>>>((const unsigned short *)tab, t16, m16, N / PER, sink); }, 5);
    CHECK();
    printf(" \"%s\": {\"ldg_u16\": %.4g, \"ldg_u32\": %.4g, \"ldg_u64\": %.4g, \"ldg_u128\": %.4g,\n"
           "   \"tex_u16\": %.4g, \"tex_u32\": %.4g, \"tex_u64\": %.4g, \"tex_u128\": %.4g,\n"
           "   \"mixed_u64_tex4of16\": %.4g, \"mixed_u64_tex8of16\": %.4g, \"mixed_u64_tex12of16\": %.4g, \"mixed_u16_tex8of16\": %.4g},\n",
           c.name, N / (a16 * 1e-3), N / (a32 * 1e-3), N / (a64 * 1e-3), N / (a128 * 1e-3), N / (b16 * 1e-3),
           N / (b32 * 1e-3), N / (b64 * 1e-3), N / (b128 * 1e-3), N / (c4 * 1e-3), N / (c8 * 1e-3), N / (c12 * 1e-3),
           N / (e8 * 1e-3));
    fflush(stdout);
  }
  // reductions into a 64 MiB table
  {
    double r32 = time_ms([&] { red32_kernel<PER><<<grid, 256>>>((uint32_t *)tab, (uint32_t)(bytes / 4 - 1), N / PER); }, 5);
    double r64 = time_ms([&] { red64_kernel<PER><<<grid, 256>>>((unsigned long long *)tab, (uint32_t)(bytes / 8 - 1), N / PER); }, 5);
    cudaMemset(tab, 0, bytes);
    double rv4 = time_ms([&] { redv4_kernel<PER><<<grid, 256>>>((float4 *)tab, (uint32_t)(bytes / 16 - 1), N / PER); }, 5);
    CHECK();
    printf(" \"red_64MiB\": {\"u32\": %.4g, \"u64\": %.4g, \"v4_f32\": %.4g},\n", N / (r32 * 1e-3), N / (r64 * 1e-3),
           N / (rv4 * 1e-3));
  }
  // shared-memory atomics and reads
  {
    uint32_t *out; cudaMalloc(&out, 1 << 20); cudaMemset(out, 0, 1 << 20);
    cudaFuncSetAttribute(atoms_kernel<PER, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(lds_kernel<PER, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    double s16k = time_ms([&] { atoms_kernel<PER, 16384><<<148 * 3, 256, 65536>>>(out, N / PER); }, 5);
    double s1k = time_ms([&] { atoms_kernel<PER, 1024><<<148 * 8, 256, 4096>>>(out, N / PER); }, 5);
    double l16k = time_ms([&] { lds_kernel<PER, 16384><<<148 * 3, 256, 65536>>>((const uint32_t *)tab, N / PER, sink); }, 5);
    double l1k = time_ms([&] { lds_kernel<PER, 1024><<<148 * 8, 256, 4096>>>((const uint32_t *)tab, N / PER, sink); }, 5);
    CHECK();
    printf(" \"smem\": {\"atoms_64KiB_table\": %.4g, \"atoms_4KiB_table\": %.4g, \"lds_64KiB_table\": %.4g, \"lds_4KiB_table\": %.4g}\n",
           N / (s16k * 1e-3), N / (s1k * 1e-3), N / (l16k * 1e-3), N / (l1k * 1e-3));
  }
  printf("}\n");
  CHECK();
  return 0;
}
