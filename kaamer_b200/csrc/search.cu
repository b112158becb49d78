// search.cu — the protein search hot path on the device.
//
// One kernel family replaces, per query, the reference's key-producer loop
// (pkg/search/search_protein.go:94-98), KmerSearch (pkg/search/search.go:414-440),
// sortMapByValue (:132-152) and FilterResults (:189-220):
//
//   k_classify      SizeInKmer per query (search.go:290-293) + size class
//   k_search<S|M>   one CTA per query: residues -> codes in smem -> dense 7-mer code ->
//                   ONE 8-byte table probe per k-mer (posting inlined when the list is a
//                   singleton) -> shared-memory open-addressing histogram keyed by subject
//                   id -> threshold (MinKMatch / MinKRatio in fp64) -> top-MaxResults by
//                   (Kmatch desc, id asc) -> hits appended to a pool
//   k_search_g      same with the histogram in global memory (L2) for queries whose
//                   subject set outgrows shared memory
//   k_scan/k_gather CSR compaction of the pool in query order (host API only)
//
// HBM traffic per k-mer: 1 B residue + 8 B entry (a 32 B sector) [+ 4 B per posting when
// the list is not a singleton].  Counts never touch HBM for classes S and M.
#include "internal.cuh"

namespace kaamer {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
enum { CNT_POOL = 0, CNT_LOOKUPS = 1, CNT_INCR = 2, CNT_STATUS = 3, CNT_CLS_LOOKUPS = 4, CNT_CLS_INCR = 8, CNT_N = 16 };
enum { ST_POOL_OVERFLOW = 1, ST_GHASH_OVERFLOW = 2 };

// size classes by SizeInKmer
constexpr int S_THREADS = 128, S_H = 1024, S_MAXK = 512;
constexpr int M_THREADS = 256, M_H = 4096, M_MAXK = 2048;
constexpr int G_THREADS = 512;
constexpr int FAST_C = 64;  // candidates ranked by counting below this, bitonic sort above

struct SearchArgs {
  const uint64_t *table;
  uint64_t d_lo, d_hi;
  const uint32_t *postings;
  const uint8_t *res;
  const uint64_t *off;
  uint32_t nq;
  long long min_kmatch;
  double min_kratio;
  int max_results;
  uint32_t *n_hits, *hit_base;
  int32_t *size_in_kmer;
  uint64_t *pool;
  uint64_t pool_cap;
  unsigned long long *counters;
  uint32_t *lists;       // [3][nq]
  uint32_t *list_count;  // [4]
  uint32_t *ghash;       // class G scratch: per CTA [keys HG][cnt HG][cand HG]
  uint32_t ghash_slots;  // HG (power of two)
};

__device__ __forceinline__ uint64_t ldg_entry(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// smallest Kmatch that survives FilterResults (search.go:195): the hit is dropped when
// float64(Kmatch)/float64(SizeInKmer) < MinKRatio || Kmatch < MinKMatch; both tests are
// monotone in Kmatch, so the kept set is {Kmatch >= kmin}.
__device__ uint32_t filter_kmin(long long min_kmatch, double ratio, int32_t size) {
  long long k = min_kmatch > 1 ? min_kmatch : 1;
  double ds = (double)size;
  if (ratio != ratio) {
    // NaN: `x < NaN` is false, the ratio test never drops a hit
  } else if (ratio > 0.0) {
    double g = ceil(ratio * ds);
    if (!(g < 4.0e9)) return 0xFFFFFFFFu;  // nothing can pass (Kmatch <= SizeInKmer < 2^31)
    long long kr = (long long)g;
    while (kr > 0 && !((double)(kr - 1) / ds < ratio)) --kr;
    while ((double)kr / ds < ratio) ++kr;
    if (kr > k) k = kr;
  }
  return k > 0xFFFFFFFEll ? 0xFFFFFFFFu : (uint32_t)k;
}

__global__ void k_classify(SearchArgs a) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.nq) return;
  uint64_t b = a.off[q], e = a.off[q + 1];
  long long len = (long long)(e - b);
  long long K = len - KAAMER_KMER_SIZE + 1;          // search.go:290
  if (len > 0 && a.res[e - 1] == '*') K--;           // search.go:291-293
  a.size_in_kmer[q] = (int32_t)K;
  a.n_hits[q] = 0;
  a.hit_base[q] = 0;
  if (K < 7) return;  // search_protein.go:74-76 (documented deviation: skipped, not fatal)
  int cls = K <= S_MAXK ? 0 : (K <= M_MAXK ? 1 : 2);
  uint32_t slot = atomicAdd(&a.list_count[cls], 1u);
  a.lists[(size_t)cls * a.nq + slot] = q;
}

struct HashView {
  uint32_t *keys, *cnt;
  uint32_t mask;
  int shift;  // 32 - log2(slots)
};

struct SelectScratch {
  uint32_t hist[256];
  uint32_t ncand, distinct, overflow, kmin, nout;
  unsigned long long base;
  unsigned long long prefix;
  uint32_t remaining;
};

// histogram[id]++ ; flags ss.overflow when the table holds more than max_distinct subjects.
// Callers test ss.overflow before every call, so at most one insertion per thread can land
// after the flag is raised: max_distinct + THREADS < slots keeps the probe loop finite.
__device__ __forceinline__ void hash_add(const HashView &hv, uint32_t id, SelectScratch &ss,
                                         uint32_t max_distinct) {
  uint32_t slot = (id * 2654435761u) >> hv.shift;
  for (uint32_t probe = 0; probe <= hv.mask; ++probe) {
    uint32_t cur = *(volatile uint32_t *)(hv.keys + slot);
    if (cur == EMPTY) {
      cur = atomicCAS(hv.keys + slot, EMPTY, id);
      if (cur == EMPTY) {
        if (atomicAdd(&ss.distinct, 1u) >= max_distinct) ss.overflow = 1;
        cur = id;
      }
    }
    if (cur == id) {
      atomicAdd(hv.cnt + slot, 1u);
      return;
    }
    slot = (slot + 1) & hv.mask;
  }
  ss.overflow = 1;
}

// composite sort key: ascending order == (Kmatch desc, subject id asc)
__device__ __forceinline__ uint64_t composite(uint32_t id, uint32_t cnt) {
  return ((uint64_t)(0xFFFFFFFFu - cnt) << 32) | id;
}
__device__ __forceinline__ uint64_t decomposite(uint64_t c) {
  uint32_t cnt = 0xFFFFFFFFu - (uint32_t)(c >> 32);
  return ((uint64_t)cnt << 32) | (uint32_t)c;  // pool format: subject | kmatch << 32
}

// Candidate c_i lives in hash slot cand[i] (CandT = u16 for smem classes, u32 for class G).
// Emits the top-N candidates into the pool in rank order and records n_hits / hit_base.
template <int THREADS, class CandT>
__device__ void select_and_emit(const SearchArgs &a, uint32_t q, const HashView &hv, const CandT *cand,
                                SelectScratch &ss) {
  const int tid = threadIdx.x;
  const uint32_t c = ss.ncand;
  const uint32_t N = a.max_results > 0 ? (uint32_t)a.max_results : 0u;
  const uint32_t nout = c < N ? c : N;
  if (nout == 0) {
    if (tid == 0) {
      a.n_hits[q] = 0;
      a.hit_base[q] = 0;
    }
    return;
  }
  if (c <= FAST_C) {
    // fast path (the common case: a handful of family hits): rank by counting
    if (tid == 0) ss.base = atomicAdd(&a.counters[CNT_POOL], (unsigned long long)nout);
    __syncthreads();
    const unsigned long long base = ss.base;
    const bool fits = base + nout <= a.pool_cap;
    if (tid < (int)c) {
      uint32_t s = cand[tid];
      uint64_t me = composite(hv.keys[s], hv.cnt[s]);
      uint32_t rank = 0;
      for (uint32_t j = 0; j < c; ++j) {
        uint32_t sj = cand[j];
        rank += composite(hv.keys[sj], hv.cnt[sj]) < me ? 1u : 0u;
      }
      if (rank < nout && fits) a.pool[base + rank] = decomposite(me);
    }
    if (tid == 0) {
      if (fits) {
        a.n_hits[q] = nout;
        a.hit_base[q] = (uint32_t)base;
      } else {
        atomicOr(&a.counters[CNT_STATUS], (unsigned long long)ST_POOL_OVERFLOW);
      }
    }
    return;
  }
  // slow path: (1) radix-select the N-th smallest composite when c > N, (2) write the kept
  // composites to a power-of-two pool segment padded with +inf, (3) bitonic sort in place.
  unsigned long long thresh = ~0ull;
  if (c > N) {
    if (tid == 0) {
      ss.prefix = 0;
      ss.remaining = N;  // N >= 1 here
    }
    for (int byte = 7; byte >= 0; --byte) {
      for (int i = tid; i < 256; i += THREADS) ss.hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = ss.prefix;
      const unsigned long long himask = byte == 7 ? 0ull : (~0ull << (8 * (byte + 1)));
      for (uint32_t i = tid; i < c; i += THREADS) {
        uint32_t s = cand[i];
        uint64_t k = composite(hv.keys[s], hv.cnt[s]);
        if ((k & himask) == prefix) atomicAdd(&ss.hist[(k >> (8 * byte)) & 0xFF], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t rem = ss.remaining, acc = 0;
        int b = 0;
        for (; b < 256; ++b) {
          if (acc + ss.hist[b] >= rem) break;
          acc += ss.hist[b];
        }
        ss.remaining = rem - acc;
        ss.prefix = prefix | ((unsigned long long)b << (8 * byte));
      }
      __syncthreads();
    }
    thresh = ss.prefix;  // exactly N composites are <= thresh (keys are unique)
  }
  uint32_t P = 1;
  while (P < nout) P <<= 1;
  if (tid == 0) {
    ss.base = atomicAdd(&a.counters[CNT_POOL], (unsigned long long)P);
    ss.nout = 0;
  }
  __syncthreads();
  const unsigned long long base = ss.base;
  const bool fits = base + P <= a.pool_cap;
  if (!fits) {
    if (tid == 0) atomicOr(&a.counters[CNT_STATUS], (unsigned long long)ST_POOL_OVERFLOW);
    return;
  }
  uint64_t *seg = a.pool + base;
  for (uint32_t i = tid; i < c; i += THREADS) {
    uint32_t s = cand[i];
    uint64_t k = composite(hv.keys[s], hv.cnt[s]);
    if (k <= thresh) seg[atomicAdd(&ss.nout, 1u)] = k;
  }
  for (uint32_t i = nout + tid; i < P; i += THREADS) seg[i] = ~0ull;
  __syncthreads();
  for (uint32_t k = 2; k <= P; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = tid; i < P; i += THREADS) {
        uint32_t l = i ^ j;
        if (l > i) {
          uint64_t x = seg[i], y = seg[l];
          bool up = (i & k) == 0;
          if ((x > y) == up) {
            seg[i] = y;
            seg[l] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = tid; i < nout; i += THREADS) seg[i] = decomposite(seg[i]);
  if (tid == 0) {
    a.n_hits[q] = nout;
    a.hit_base[q] = (uint32_t)base;
  }
}

// Lookup + count for positions [0,K) of one query; codes come from `code_at(pos)`.
template <int THREADS, class CodeAt>
__device__ __forceinline__ void lookup_and_count(const SearchArgs &a, int K, CodeAt code_at, const HashView &hv,
                                                 uint32_t max_distinct, SelectScratch &ss,
                                                 unsigned long long &q_incr) {
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;
  constexpr int U = 4;  // independent table probes in flight per thread
  for (int base = 0; base < K; base += U * THREADS) {
    uint64_t ent[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int pos = base + u * THREADS + tid;
      ent[u] = 0;
      if (pos < K) {
        uint32_t d = dense_from_codes(code_at(pos), code_at(pos + 1), code_at(pos + 2), code_at(pos + 3),
                                      code_at(pos + 4), code_at(pos + 5), code_at(pos + 6));
        if (d >= a.d_lo && d < a.d_hi) ent[u] = ldg_entry(a.table + (d - a.d_lo));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t cnt = (uint32_t)(ent[u] >> ENTRY_VALUE_BITS);
      const uint64_t val = ent[u] & ENTRY_VALUE_MASK;
      q_incr += cnt;
      if (cnt == 1) {
        if (!*(volatile uint32_t *)&ss.overflow) hash_add(hv, (uint32_t)val, ss, max_distinct);
      } else if (cnt > 1 && cnt < 32) {
        for (uint32_t i = 0; i < cnt; ++i) {
          if (*(volatile uint32_t *)&ss.overflow) break;
          hash_add(hv, __ldg(a.postings + val + i), ss, max_distinct);
        }
      }
      // long posting lists: the whole warp walks them together (coalesced)
      unsigned big = __ballot_sync(0xFFFFFFFFu, cnt >= 32);
      while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t bc = __shfl_sync(0xFFFFFFFFu, cnt, src);
        const uint64_t bv = __shfl_sync(0xFFFFFFFFu, val, src);
        for (uint32_t i = lane; i < bc; i += 32) {
          if (*(volatile uint32_t *)&ss.overflow) break;
          hash_add(hv, __ldg(a.postings + bv + i), ss, max_distinct);
        }
      }
    }
    // warp-uniform early exit (the ballots above need converged warps)
    if (__any_sync(0xFFFFFFFFu, *(volatile uint32_t *)&ss.overflow != 0)) break;
  }
}

template <int THREADS, int H, int MAXK, int CLS>
__global__ void __launch_bounds__(THREADS) k_search(SearchArgs a) {
  static_assert((H & (H - 1)) == 0, "H must be a power of two");
  __shared__ uint32_t hkeys[H];
  __shared__ uint32_t hcnt[H];
  __shared__ __align__(16) uint8_t codes[MAXK + 16];
  __shared__ uint16_t cand[H];
  __shared__ SelectScratch ss;
  constexpr uint32_t MAXD = (H - THREADS - 1) < (3 * H / 4) ? (H - THREADS - 1) : (3 * H / 4);
  int log2h = 0;
  while ((1 << log2h) < H) ++log2h;
  const HashView hv{hkeys, hcnt, (uint32_t)H - 1u, 32 - log2h};
  const int tid = threadIdx.x;
  const uint32_t count = a.list_count[CLS];
  unsigned long long my_incr = 0, my_lookups = 0;
  for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
    const uint32_t q = a.lists[(size_t)CLS * a.nq + it];
    const uint64_t b = a.off[q];
    const int K = a.size_in_kmer[q];
    const int ncodes = K + KAAMER_KMER_SIZE - 1;
    for (int i = tid; i < H; i += THREADS) {
      hkeys[i] = EMPTY;
      hcnt[i] = 0;
    }
    for (int i = tid; i < ncodes; i += THREADS) codes[i] = (uint8_t)aa_code(a.res[b + i]);
    if (tid == 0) {
      ss.ncand = 0;
      ss.distinct = 0;
      ss.overflow = 0;
      ss.kmin = filter_kmin(a.min_kmatch, a.min_kratio, K);
    }
    __syncthreads();
    unsigned long long q_incr = 0;
    lookup_and_count<THREADS>(a, K, [&](int p) -> uint32_t { return codes[p]; }, hv, MAXD, ss, q_incr);
    __syncthreads();
    if (ss.overflow) {
      // subject set outgrew this class: hand the query to the next one (stream order
      // guarantees that kernel has not started yet)
      if (tid == 0) {
        uint32_t slot = atomicAdd(&a.list_count[CLS + 1], 1u);
        a.lists[(size_t)(CLS + 1) * a.nq + slot] = q;
      }
      __syncthreads();
      continue;
    }
    my_incr += q_incr;
    if (tid == 0) my_lookups += (unsigned long long)K;
    const uint32_t kmin = ss.kmin;
    for (int i = tid; i < H; i += THREADS)
      if (hkeys[i] != EMPTY && hcnt[i] >= kmin) cand[atomicAdd(&ss.ncand, 1u)] = (uint16_t)i;
    __syncthreads();
    select_and_emit<THREADS, uint16_t>(a, q, hv, cand, ss);
    __syncthreads();
  }
  // work counters: one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    my_incr += __shfl_down_sync(0xFFFFFFFFu, my_incr, o);
    my_lookups += __shfl_down_sync(0xFFFFFFFFu, my_lookups, o);
  }
  if ((tid & 31) == 0) {
    if (my_incr) {
      atomicAdd(&a.counters[CNT_INCR], my_incr);
      atomicAdd(&a.counters[CNT_CLS_INCR + CLS], my_incr);
    }
    if (my_lookups) {
      atomicAdd(&a.counters[CNT_LOOKUPS], my_lookups);
      atomicAdd(&a.counters[CNT_CLS_LOOKUPS + CLS], my_lookups);
    }
  }
}

// class G: histogram in global memory (per-CTA scratch, stays in L2), codes read on the fly
__global__ void __launch_bounds__(G_THREADS) k_search_g(SearchArgs a) {
  __shared__ SelectScratch ss;
  __shared__ unsigned long long s_total;
  constexpr int THREADS = G_THREADS;
  const int tid = threadIdx.x;
  const uint32_t HG = a.ghash_slots;
  uint32_t *gkeys = a.ghash + (size_t)blockIdx.x * 3 * HG;
  uint32_t *gcnt = gkeys + HG;
  uint32_t *gcand = gcnt + HG;
  const uint32_t count = a.list_count[2];
  unsigned long long my_incr = 0, my_lookups = 0;
  for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
    const uint32_t q = a.lists[(size_t)2 * a.nq + it];
    const uint64_t b = a.off[q];
    const int K = a.size_in_kmer[q];
    const uint8_t *s = a.res + b;
    auto code_at = [&](int p) -> uint32_t { return aa_code(s[p]); };
    // pass 1: total postings of the query bounds the number of distinct subjects
    if (tid == 0) s_total = 0;
    __syncthreads();
    unsigned long long tot = 0;
    for (int pos = tid; pos < K; pos += THREADS) {
      uint32_t d = dense_from_codes(code_at(pos), code_at(pos + 1), code_at(pos + 2), code_at(pos + 3),
                                    code_at(pos + 4), code_at(pos + 5), code_at(pos + 6));
      if (d >= a.d_lo && d < a.d_hi) tot += ldg_entry(a.table + (d - a.d_lo)) >> ENTRY_VALUE_BITS;
    }
    atomicAdd(&s_total, tot);
    __syncthreads();
    unsigned long long T = s_total;
    uint32_t Hq = 1024;
    while (Hq < HG && (unsigned long long)Hq < 2 * T) Hq <<= 1;
    int log2h = 0;
    while ((1u << log2h) < Hq) ++log2h;
    const HashView hv{gkeys, gcnt, Hq - 1u, 32 - log2h};
    for (uint32_t i = tid; i < Hq; i += THREADS) {
      gkeys[i] = EMPTY;
      gcnt[i] = 0;
    }
    if (tid == 0) {
      ss.ncand = 0;
      ss.distinct = 0;
      ss.overflow = 0;
      ss.kmin = filter_kmin(a.min_kmatch, a.min_kratio, K);
    }
    __syncthreads();
    unsigned long long q_incr = 0;
    const uint32_t maxd = Hq - THREADS - 1 < (Hq / 4) * 3 ? Hq - THREADS - 1 : (Hq / 4) * 3;
    lookup_and_count<THREADS>(a, K, code_at, hv, maxd, ss, q_incr);
    __syncthreads();
    if (ss.overflow) {
      if (tid == 0) atomicOr(&a.counters[CNT_STATUS], (unsigned long long)ST_GHASH_OVERFLOW);
      __syncthreads();
      continue;
    }
    my_incr += q_incr;
    if (tid == 0) my_lookups += (unsigned long long)K;
    const uint32_t kmin = ss.kmin;
    for (uint32_t i = tid; i < Hq; i += THREADS)
      if (gkeys[i] != EMPTY && gcnt[i] >= kmin) gcand[atomicAdd(&ss.ncand, 1u)] = i;
    __syncthreads();
    select_and_emit<THREADS, uint32_t>(a, q, hv, gcand, ss);
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_incr += __shfl_down_sync(0xFFFFFFFFu, my_incr, o);
    my_lookups += __shfl_down_sync(0xFFFFFFFFu, my_lookups, o);
  }
  if ((tid & 31) == 0) {
    if (my_incr) {
      atomicAdd(&a.counters[CNT_INCR], my_incr);
      atomicAdd(&a.counters[CNT_CLS_INCR + 2], my_incr);
    }
    if (my_lookups) {
      atomicAdd(&a.counters[CNT_LOOKUPS], my_lookups);
      atomicAdd(&a.counters[CNT_CLS_LOOKUPS + 2], my_lookups);
    }
  }
}

// ---- CSR compaction (host API) ----------------------------------------------------------
// single-CTA exclusive scan of n_hits -> hit_off[nq+1]
__global__ void __launch_bounds__(1024) k_scan_hits(const uint32_t *__restrict__ n_hits, uint32_t nq,
                                                    uint64_t *hit_off) {
  __shared__ uint64_t warp_sum[32];
  __shared__ uint64_t carry;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nq; base += 1024) {
    uint32_t i = base + tid;
    uint64_t v = i < nq ? n_hits[i] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[w] = x;
    __syncthreads();
    if (w == 0) {
      uint64_t s = warp_sum[lane], t = s;
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xFFFFFFFFu, t, o);
        if (lane >= o) t += y;
      }
      warp_sum[lane] = t - s;  // exclusive
    }
    __syncthreads();
    uint64_t excl = carry + warp_sum[w] + x - v;
    if (i < nq) hit_off[i] = excl;
    __syncthreads();
    if (tid == 1023) carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) hit_off[nq] = carry;
}

__global__ void k_gather_hits(const uint32_t *__restrict__ n_hits, const uint32_t *__restrict__ hit_base,
                              const uint64_t *__restrict__ hit_off, const uint64_t *__restrict__ pool,
                              uint32_t nq, uint64_t *out) {
  // one warp per query
  uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  uint32_t n = n_hits[q];
  uint64_t src = hit_base[q], dst = hit_off[q];
  for (uint32_t i = threadIdx.x & 31; i < n; i += 32) out[dst + i] = pool[src + i];
}

// ---- launch ---------------------------------------------------------------------------
void profile_begin(kaamer_gpu *h, cudaStream_t st, int cls) {
  if (!h->profile) return;
  ProfSpan sp;
  sp.cls = cls;
  cudaEventCreate(&sp.a);
  cudaEventCreate(&sp.b);
  h->prof_pending.push_back(sp);
  cudaEventRecord(sp.a, st);
}
void profile_end(kaamer_gpu *h, cudaStream_t st) {
  if (!h->profile || h->prof_pending.empty()) return;
  cudaEventRecord(h->prof_pending.back().b, st);
}

static uint32_t ghash_slots_for(kaamer_gpu *h) {
  (void)h;
  return 1u << 20;
}

int search_proteins_device(kaamer_gpu *h, const uint8_t *d_res, const uint64_t *d_off, uint32_t nq,
                           const kaamer_opts *o, const kaamer_dev_result *out, cudaStream_t st) {
  if (!h->idx.table) {
    set_error("no index resident");
    return KAAMER_ERR_ARG;
  }
  if (nq == 0) return KAAMER_OK;
  SearchWorkspace &ws = h->ws;
  KCHECK(ws.lists.ensure((size_t)3 * nq + 8));
  uint32_t *list_count = ws.lists.p + (size_t)3 * nq;
  SearchArgs a{};
  a.table = h->idx.table;
  a.d_lo = h->idx.d_lo;
  a.d_hi = h->idx.d_hi;
  a.postings = h->idx.postings;
  a.res = d_res;
  a.off = d_off;
  a.nq = nq;
  a.min_kmatch = o->min_kmatch;
  a.min_kratio = o->min_kratio;
  a.max_results = o->max_results;
  a.n_hits = out->n_hits;
  a.hit_base = out->hit_base;
  a.size_in_kmer = out->size_in_kmer;
  a.pool = out->pool;
  a.pool_cap = out->pool_cap;
  a.counters = (unsigned long long *)out->counters;
  a.lists = ws.lists.p;
  a.list_count = list_count;
  a.ghash_slots = ghash_slots_for(h);
  const int g_ctas = h->sm_count;
  KCHECK(ws.ghash.ensure((size_t)g_ctas * 3 * a.ghash_slots));
  a.ghash = ws.ghash.p;
  KCUDA(cudaMemsetAsync(list_count, 0, 4 * sizeof(uint32_t), st));
  KCUDA(cudaMemsetAsync(out->counters, 0, CNT_N * sizeof(uint64_t), st));
  k_classify<<<(nq + 255) / 256, 256, 0, st>>>(a);
  // persistent grids: a multiple of the SM count, CTAs loop over their class list
  const unsigned s_grid = (unsigned)h->sm_count * 14u;
  const unsigned m_grid = (unsigned)h->sm_count * 3u;
  profile_begin(h, st, 0);
  k_search<S_THREADS, S_H, S_MAXK, 0><<<s_grid < nq ? s_grid : nq, S_THREADS, 0, st>>>(a);
  profile_end(h, st);
  profile_begin(h, st, 1);
  k_search<M_THREADS, M_H, M_MAXK, 1><<<m_grid < nq ? m_grid : nq, M_THREADS, 0, st>>>(a);
  profile_end(h, st);
  profile_begin(h, st, 2);
  k_search_g<<<g_ctas, G_THREADS, 0, st>>>(a);
  profile_end(h, st);
  h->prof_all_launches += 4;
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

// host-buffer entry point: H2D, search, CSR compaction on the device, D2H
int search_proteins_host(kaamer_gpu *h, const uint8_t *res, const uint64_t *off, uint32_t nq,
                         const kaamer_opts *o, kaamer_hits **out_hits) {
  SearchWorkspace &ws = h->ws;
  cudaStream_t st = h->stream;
  const uint64_t n_res = nq ? off[nq] - off[0] : 0;
  auto *hits = new kaamer_hits();
  memset(hits, 0, sizeof *hits);
  auto *owner = new HitsOwner();
  hits->_owner = owner;
  auto fail = [&](int rc) {
    delete owner;
    delete hits;
    return rc;
  };
#define HCHECK(x)                       \
  do {                                  \
    int _r = (x);                       \
    if (_r != KAAMER_OK) return fail(_r); \
  } while (0)
#define HCUDA(call)                                                                     \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));  \
      return fail(KAAMER_ERR_CUDA);                                                     \
    }                                                                                   \
  } while (0)
  hits->n_rows = nq;
  HCHECK(owner->alloc(&hits->hit_off, (size_t)nq + 1));
  HCHECK(owner->alloc(&hits->size_in_kmer, (size_t)nq));
  hits->hit_off[0] = 0;
  if (nq == 0) {
    HCHECK(owner->alloc(&hits->subject_id, 1));
    HCHECK(owner->alloc(&hits->kmatch, 1));
    *out_hits = hits;
    return KAAMER_OK;
  }
  if (off[0] != 0) {
    set_error("seq_off[0] must be 0");
    return fail(KAAMER_ERR_ARG);
  }
  HCHECK(ws.residues.ensure((size_t)n_res + 16));
  HCHECK(ws.seq_off.ensure((size_t)nq + 1));
  HCHECK(ws.n_hits.ensure(nq));
  HCHECK(ws.hit_base.ensure(nq));
  HCHECK(ws.size_in_kmer.ensure(nq));
  HCHECK(ws.hit_off.ensure((size_t)nq + 1));
  HCHECK(ws.counters.ensure(CNT_N));
  HCHECK(ws.h_counters.ensure(CNT_N + 2));
  HCUDA(cudaMemcpyAsync(ws.residues.p, res, (size_t)n_res, cudaMemcpyHostToDevice, st));
  HCUDA(cudaMemcpyAsync(ws.seq_off.p, off, ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, st));
  uint64_t per_q = o->max_results > 0 ? (uint64_t)(o->max_results < 16 ? o->max_results : 16) : 1;
  uint64_t pool_cap = (uint64_t)nq * per_q + 4096;
  for (int attempt = 0;; ++attempt) {
    HCHECK(ws.pool.ensure((size_t)pool_cap));
    kaamer_dev_result dr{};
    dr.n_hits = ws.n_hits.p;
    dr.hit_base = ws.hit_base.p;
    dr.size_in_kmer = ws.size_in_kmer.p;
    dr.pool = ws.pool.p;
    dr.pool_cap = pool_cap;
    dr.counters = ws.counters.p;
    HCHECK(search_proteins_device(h, ws.residues.p, ws.seq_off.p, nq, o, &dr, st));
    k_scan_hits<<<1, 1024, 0, st>>>(ws.n_hits.p, nq, ws.hit_off.p);
    h->prof_all_launches += 1;
    HCUDA(cudaMemcpyAsync(ws.h_counters.p, ws.counters.p, CNT_N * 8, cudaMemcpyDeviceToHost, st));
    HCUDA(cudaMemcpyAsync(ws.h_counters.p + CNT_N, ws.hit_off.p + nq, 8, cudaMemcpyDeviceToHost, st));
    HCUDA(cudaStreamSynchronize(st));
    uint64_t status = ws.h_counters.p[CNT_STATUS];
    if (status & ST_GHASH_OVERFLOW) {
      set_error("a query matched more distinct subjects than the class-G histogram holds (%u slots)",
                ghash_slots_for(h));
      return fail(KAAMER_ERR_LIMIT);
    }
    if (status & ST_POOL_OVERFLOW) {
      if (attempt >= 3) {
        set_error("hit pool overflow after %d attempts", attempt + 1);
        return fail(KAAMER_ERR_LIMIT);
      }
      pool_cap = ws.h_counters.p[CNT_POOL] + 4096;  // exact demand of the failed pass
      continue;
    }
    break;
  }
  const uint64_t n_hits = ws.h_counters.p[CNT_N];
  hits->n_hits = n_hits;
  hits->n_lookups = ws.h_counters.p[CNT_LOOKUPS];
  hits->n_increments = ws.h_counters.p[CNT_INCR];
  HCHECK(owner->alloc(&hits->subject_id, (size_t)n_hits));
  HCHECK(owner->alloc(&hits->kmatch, (size_t)n_hits));
  HCHECK(ws.out_hits.ensure((size_t)n_hits + 1));
  PinBuf<uint64_t> packed;
  HCHECK(packed.ensure((size_t)n_hits + 1));
  if (n_hits) {
    unsigned grid = (unsigned)(((uint64_t)nq * 32 + 255) / 256);
    k_gather_hits<<<grid, 256, 0, st>>>(ws.n_hits.p, ws.hit_base.p, ws.hit_off.p, ws.pool.p, nq, ws.out_hits.p);
    h->prof_all_launches += 1;
    HCUDA(cudaMemcpyAsync(packed.p, ws.out_hits.p, (size_t)n_hits * 8, cudaMemcpyDeviceToHost, st));
  }
  HCUDA(cudaMemcpyAsync(hits->hit_off, ws.hit_off.p, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, st));
  HCUDA(cudaMemcpyAsync(hits->size_in_kmer, ws.size_in_kmer.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  HCUDA(cudaStreamSynchronize(st));
  for (uint64_t i = 0; i < n_hits; ++i) {
    hits->subject_id[i] = (uint32_t)packed.p[i];
    hits->kmatch[i] = (uint32_t)(packed.p[i] >> 32);
  }
  packed.release();
  *out_hits = hits;
#undef HCHECK
#undef HCUDA
  return KAAMER_OK;
}

}  // namespace kaamer
