// locality_bench.cu — the "locality path" experiment of VERDICT r1 item 6, measured instead of argued:
// does sorting a batch's (dense code, query) pairs before probing the 14.5 GB table buy more than it costs?
//
//   A  random order (what the search kernels do today): 8-byte probes at codes in query order
//   B  cub radix sort of the (code, query) pairs by the TOP `bits` bits of the code (partial sort: probes of
//      one partition land in one table slice), then the probes in sorted order
//   C  full sort by code, probes in sorted order (several probes per DRAM row where the batch is dense enough)
//   D  the way back: the (query, subject) pairs the probes produce must be regrouped by query before they
//      can be counted — a second sort of ~1.1 pairs per lookup by the query id
//
// Batch sizes: one C3 batch (34.5 M lookups) and ten (345 M).  Prints ms per stage; profiles/ keeps the log.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__global__ void k_make(uint32_t *codes, uint32_t *qid, uint64_t n, uint64_t slots, uint32_t per_query) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  codes[i] = (uint32_t)(mix(i * 0x9E3779B97F4A7C15ull + 1) % slots);
  qid[i] = (uint32_t)(i / per_query);
}
template <int U>
__global__ void k_probe(const uint64_t *__restrict__ table, const uint32_t *__restrict__ codes, uint64_t n, uint64_t *out) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint64_t i = tid; i < n; i += stride * U) {
    uint64_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = i + (uint64_t)u * stride;
      v[u] = 0;
      if (j < n) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v[u]) : "l"(table + codes[j]));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  if (acc == 0x1234567) out[0] = acc;
}
// sorted codes: consecutive threads take consecutive (sorted) codes, so a warp's probes share DRAM rows
__global__ void k_probe_seq(const uint64_t *__restrict__ table, const uint32_t *__restrict__ codes, uint64_t n, uint64_t *out) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint64_t i = tid; i < n; i += stride) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(table + codes[i]));
    acc += v;
  }
  if (acc == 0x1234567) out[0] = acc;
}
static float timeit(cudaEvent_t a, cudaEvent_t b) {
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}
int main() {
  const uint64_t slots = 1813368648ull;
  uint64_t *table, *out;
  if (cudaMalloc(&table, slots * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&out, 8);
  cudaMemset(table, 1, slots * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int big = 0; big < 2; ++big) {
    const uint64_t n = big ? 345000000ull : 34500000ull;
    uint32_t *codes, *qid, *codes2, *qid2;
    cudaMalloc(&codes, n * 4); cudaMalloc(&qid, n * 4); cudaMalloc(&codes2, n * 4); cudaMalloc(&qid2, n * 4);
    k_make<<<(unsigned)((n + 255) / 256), 256>>>(codes, qid, n, slots, 345);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes, codes2, qid, qid2, (int64_t)n, 0, 31);
    void *tmp;
    cudaMalloc(&tmp, tmp_bytes);
    printf("---- %llu lookups (%s) ----\n", (unsigned long long)n, big ? "ten C3 batches" : "one C3 batch");
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      k_probe<4><<<148 * 8, 256>>>(table, codes, n, out);
      cudaEventRecord(b);
      const float ms = timeit(a, b);
      if (ms < best) best = ms;
    }
    printf("A  probes in batch order                 : %8.3f ms  %6.1f G probes/s\n", best, n / best / 1e6);
    for (int bits = 8; bits <= 31; bits += (bits == 8 ? 8 : 15)) {
      float s_best = 1e9, p_best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, codes, codes2, qid, qid2, (int64_t)n, 31 - bits, 31);
        cudaEventRecord(b);
        float ms = timeit(a, b);
        if (ms < s_best) s_best = ms;
        cudaEventRecord(a);
        k_probe_seq<<<148 * 16, 256>>>(table, codes2, n, out);
        cudaEventRecord(b);
        ms = timeit(a, b);
        if (ms < p_best) p_best = ms;
      }
      printf("%c  sort by the top %2d bits %8.3f ms + probes in sorted order %8.3f ms (%6.1f G probes/s) = %8.3f ms\n",
             bits == 31 ? 'C' : 'B', bits, s_best, p_best, n / p_best / 1e6, s_best + p_best);
    }
    {
      // the way back: ~1.13 (query, subject) pairs per lookup sorted by the 17 / 20 query bits
      const uint64_t np = (uint64_t)(n * 1.13);
      const int qbits = big ? 20 : 17;
      float r_best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, qid, qid2, codes, codes2, (int64_t)(np < n ? np : n), 0, qbits);
        cudaEventRecord(b);
        const float ms = timeit(a, b);
        if (ms < r_best) r_best = ms;
      }
      printf("D  regroup %llu (query, subject) pairs by query (%d bits, n capped at the lookups) : %8.3f ms\n",
             (unsigned long long)np, qbits, r_best);
    }
    cudaFree(codes); cudaFree(qid); cudaFree(codes2); cudaFree(qid2); cudaFree(tmp);
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
