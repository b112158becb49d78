// gather_bench.cu — microbenchmark: how many random 8-byte probes per second does a B200
// sustain over a 14.5 GB direct-address table?  This is the physical ceiling of the lookup
// stage (one 32 B sector per probe); DESIGN.md quotes its output next to the roofline.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
template <int U>
__global__ void k_gather(const uint64_t *__restrict__ table, uint64_t slots, uint64_t n, uint64_t *out, uint64_t seed) {
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint64_t i = tid; i < n; i += stride * U) {
    uint64_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint64_t j = i + (uint64_t)u * stride;
      uint64_t idx = mix(j + seed) % slots;
      v[u] = 0;
      if (j < n) asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v[u]) : "l"(table + idx));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  if (acc == 0x1234567) out[0] = acc;
}
int main(int argc, char **argv) {
  uint64_t slots = 1813366968ull;
  uint64_t n = 1ull << 27;
  uint64_t *table, *out;
  if (cudaMalloc(&table, slots * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&out, 8);
  cudaMemset(table, 1, slots * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int gran = 0; gran < 2; ++gran) {
    if (gran == 1) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32); printf("set L2 fetch granularity 32: %s\n", cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity = %zu\n", g);
    for (int cfg = 0; cfg < 6; ++cfg) {
      int threads = 256, blocks_per_sm = cfg < 3 ? 8 : 4;
      int U = (cfg % 3 == 0) ? 1 : (cfg % 3 == 1 ? 4 : 8);
      int grid = 148 * blocks_per_sm;
      float best = 1e9;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        if (U == 1) k_gather<1><<<grid, threads>>>(table, slots, n, out, rep * 7919);
        else if (U == 4) k_gather<4><<<grid, threads>>>(table, slots, n, out, rep * 7919);
        else k_gather<8><<<grid, threads>>>(table, slots, n, out, rep * 7919);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
      }
      printf("blocks/SM=%d U=%d : %.3f ms  %.2f G probes/s  (%.1f GB/s of 32B sectors)\n", blocks_per_sm, U, best,
             n / best / 1e6, n * 32.0 / best / 1e6);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
