// gather_variants.cu — which load flavour makes L2 fetch 32 B instead of a whole 128 B line
// on a random 8-byte probe?  (ncu on k_search showed 4 DRAM sectors per probe.)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
template <int V>
__device__ __forceinline__ uint64_t ld(const uint64_t *p) {
  uint64_t v;
  if (V == 0) asm volatile("ld.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 1) asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 2) asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 3) asm volatile("ld.global.cg.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 4) asm volatile("ld.global.cs.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 5) asm volatile("ld.global.cv.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 6) asm volatile("ld.global.L1::evict_last.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 7) asm volatile("ld.global.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 8) asm volatile("ld.volatile.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 9) asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
template <int V>
__global__ void k_gather(const uint64_t *__restrict__ table, uint64_t slots, uint64_t n, uint64_t *out, uint64_t seed) {
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint64_t i = tid; i < n; i += stride * 4) {
    uint64_t v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint64_t j = i + (uint64_t)u * stride;
      uint64_t idx = mix(j + seed) % slots;
      v[u] = j < n ? ld<V>(table + idx) : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u];
  }
  if (acc == 0x1234567) out[0] = acc;
}
template <int V>
void run(const char *name, const uint64_t *table, uint64_t slots, uint64_t n, uint64_t *out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k_gather<V><<<148 * 8, 256>>>(table, slots, n, out, rep * 7919 + 1);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("%-44s %.3f ms  %.2f G probes/s  err=%s\n", name, best, n / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char **argv) {
  int gran = argc > 1 ? atoi(argv[1]) : 0;
  if (gran) printf("set limit %d before context use: %s\n", gran, cudaGetErrorString(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)));
  size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity = %zu\n", g);
  uint64_t slots = 1813366968ull, n = 1ull << 26;
  uint64_t *table, *out;
  if (cudaMalloc(&table, slots * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&out, 8);
  cudaMemset(table, 1, slots * 8);
  run<0>("ld.global", table, slots, n, out);
  run<1>("ld.global.nc", table, slots, n, out);
  run<2>("ld.global.nc.L1::no_allocate", table, slots, n, out);
  run<3>("ld.global.cg", table, slots, n, out);
  run<4>("ld.global.cs", table, slots, n, out);
  run<5>("ld.global.cv", table, slots, n, out);
  run<6>("ld.global.L1::evict_last", table, slots, n, out);
  run<7>("ld.global.L1::no_allocate.L2::64B", table, slots, n, out);
  run<8>("ld.volatile.global", table, slots, n, out);
  run<9>("ld.relaxed.gpu.global", table, slots, n, out);
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
