// l2_filter_bench.cu — does an L2-resident folded presence bitmap in front of the 14.5 GB
// direct-address table raise the lookup rate?  Every lookup reads one bitmap word (L2 hit when the
// bitmap stays resident) and probes the table in HBM only when the bit is set.
//   variants: bitmap size (32/64/128 MB), fraction of set bits, L2 eviction hints on/off
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__global__ void k_fill(uint32_t *bits, uint64_t words, uint32_t thresh256) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= words) return;
  uint32_t w = 0;
  for (int b = 0; b < 32; ++b) w |= ((mix(i * 32 + b + 12345) & 255u) < thresh256 ? 1u : 0u) << b;
  bits[i] = w;
}
// HINT 0: plain loads; 1: bitmap evict_last + table evict_first (createpolicy cache hints)
template <int HINT, bool FILTER>
__global__ void k_lookup(const uint64_t *__restrict__ table, uint64_t slots, const uint32_t *__restrict__ bits,
                         uint32_t bit_mask, uint64_t n, uint64_t *out, uint64_t seed, unsigned long long *probes) {
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t pol_last = 0, pol_first = 0;
  if (HINT) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  }
  uint64_t acc = 0;
  unsigned long long np = 0;
  for (uint64_t i = tid; i < n; i += stride * 4) {
    uint32_t d[4], w[4];
    uint64_t v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint64_t j = i + (uint64_t)u * stride;
      d[u] = (uint32_t)(mix(j + seed) % slots);
      w[u] = 0xFFFFFFFFu;
      if (FILTER && j < n) {
        const uint32_t *p = bits + ((d[u] & bit_mask) >> 5);
        if (HINT) asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(w[u]) : "l"(p), "l"(pol_last));
        else asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(w[u]) : "l"(p));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint64_t j = i + (uint64_t)u * stride;
      v[u] = 0;
      if (j < n && ((w[u] >> (d[u] & 31u)) & 1u)) {
        const uint64_t *p = table + d[u];
        if (HINT) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.b64 %0, [%1], %2;" : "=l"(v[u]) : "l"(p), "l"(pol_first));
        else asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v[u]) : "l"(p));
        ++np;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u];
  }
  if (acc == 0x1234567) out[0] = acc;
  atomicAdd(probes, np);
}
template <int HINT, bool FILTER>
void run(const char *name, const uint64_t *table, uint64_t slots, const uint32_t *bits, uint32_t bit_mask, uint64_t n,
         uint64_t *out, unsigned long long *d_probes) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9; unsigned long long probes = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemset(d_probes, 0, 8);
    cudaEventRecord(a);
    k_lookup<HINT, FILTER><<<148 * 8, 256>>>(table, slots, bits, bit_mask, n, out, rep * 7919 + 1, d_probes);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
    cudaMemcpy(&probes, d_probes, 8, cudaMemcpyDeviceToHost);
  }
  printf("  %-46s %7.3f ms  %7.2f G lookups/s  (%.1f %% of the lookups probe the table: %6.2f G probes/s)  %s\n", name, best,
         n / best / 1e6, 100.0 * probes / n, probes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  fflush(stdout);
}
int main() {
  const uint64_t slots = 1813366968ull, n = 1ull << 26;
  uint64_t *table, *out; unsigned long long *d_probes;
  CK(cudaMalloc(&table, slots * 8)); CK(cudaMemset(table, 0, slots * 8)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&d_probes, 8));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("L2 %d MB, persistingL2CacheMaxSize %d MB\n", pr.l2CacheSize >> 20, pr.persistingL2CacheMaxSize >> 20);
  run<0, false>("no filter (every lookup probes)", table, slots, nullptr, 0, n, out, d_probes);
  for (int mb : {32, 64, 128}) {
    const uint64_t nbits = (uint64_t)mb << 23, words = nbits / 32;
    uint32_t *bits; CK(cudaMalloc(&bits, words * 4));
    for (int thresh : {77, 175}) {  // 30 % and 68.5 % of the bits set
      k_fill<<<(unsigned)((words + 255) / 256), 256>>>(bits, words, thresh); CK(cudaDeviceSynchronize());
      char name[128];
      snprintf(name, sizeof name, "%3d MB bitmap, %2d %% set, plain loads", mb, thresh * 100 / 256);
      run<0, true>(name, table, slots, bits, (uint32_t)(nbits - 1), n, out, d_probes);
      snprintf(name, sizeof name, "%3d MB bitmap, %2d %% set, evict_last/first", mb, thresh * 100 / 256);
      run<1, true>(name, table, slots, bits, (uint32_t)(nbits - 1), n, out, d_probes);
    }
    cudaFree(bits);
  }
  printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
