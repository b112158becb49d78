// shard.cu — key-range sharded search (mode S of DESIGN.md §7).  New with respect to
// the reference, which is single-process: `Kmatch[q,s]` is a plain sum over the query's k-mers
// (pkg/search/search.go:431-436), so it decomposes over any partition of the key space.
//
//   home rank     k_route<false/true>  dense code of every query k-mer, bucketed by owner shard
//                                      (counts[s][q], then codes[] in (shard, query) order)
//                 -- all-to-all #1: codes + per-query counts to the owners (host: NCCL) --
//   owner shard   k_shard_count[_g]    per segment (= one query's k-mers on this shard): probe the
//                                      resident table range, count subjects in a shared-memory
//                                      histogram, emit ALL (subject, partial count) pairs: a
//                                      partial count cannot be thresholded (10 may be 4 + 6)
//                 k_shard_gather       pool -> segment order
//                 -- all-to-all #2: partial lists back to the query's home rank --
//   home rank     k_shard_merge[_g]    per query: sum the partial counts of all shards, then the
//                                      same FilterResults threshold and top-N as search.cu
#include "search_common.cuh"

namespace kaamer {

constexpr int MAX_SHARDS = 16;
constexpr int SH_THREADS = 256;
constexpr int SH_H = 8192;            // shared-memory histogram slots
constexpr uint32_t SEG_MAX_U16 = 60000;  // counts are u16 in shared memory

struct RouteArgs {
  const uint8_t *res;
  const uint64_t *off;
  uint32_t nq;
  int n_shards;
  uint32_t fences[MAX_SHARDS + 1];  // shard s owns dense codes [fences[s], fences[s+1])
  uint32_t *counts;                 // [n_shards][nq]
  const uint64_t *offsets;          // exclusive scan of counts (fill pass)
  uint32_t *codes;
  int32_t *size_in_kmer;
};

template <bool FILL>
__global__ void __launch_bounds__(256) k_route(RouteArgs a) {
  __shared__ uint32_t run[8][MAX_SHARDS];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x * 8 + w;
  if (q >= a.nq) return;
  const uint64_t b = a.off[q], e = a.off[q + 1];
  const long long len = (long long)(e - b);
  long long K = len - KAAMER_KMER_SIZE + 1;  // search.go:290
  if (len > 0 && a.res[e - 1] == '*') K--;   // search.go:291-293
  if (!FILL && lane == 0) a.size_in_kmer[q] = (int32_t)K;
  const int Keff = K >= 7 ? (int)K : 0;  // search_protein.go:74-76
  if (lane < (unsigned)a.n_shards) run[w][lane] = 0;
  __syncwarp();
  const uint8_t *s = a.res + b;
  for (int base = 0; base < Keff; base += 32) {
    const int pos = base + (int)lane;
    const bool valid = pos < Keff;
    uint32_t d = 0;
    int sh = -1;
    if (valid) {
      d = dense_from_codes(aa_code(s[pos]), aa_code(s[pos + 1]), aa_code(s[pos + 2]), aa_code(s[pos + 3]),
                           aa_code(s[pos + 4]), aa_code(s[pos + 5]), aa_code(s[pos + 6]));
      sh = 0;
      for (int t = 1; t < a.n_shards; ++t) sh += d >= a.fences[t] ? 1 : 0;
    }
    for (int t = 0; t < a.n_shards; ++t) {
      const unsigned mask = __ballot_sync(0xFFFFFFFFu, sh == t);
      if (mask == 0) continue;
      if (FILL && sh == t)
        a.codes[a.offsets[(size_t)t * a.nq + q] + run[w][t] + __popc(mask & ((1u << lane) - 1u))] = d;
      __syncwarp();
      if (lane == 0) run[w][t] += __popc(mask);
      __syncwarp();
    }
  }
  if (!FILL && lane < (unsigned)a.n_shards) a.counts[(size_t)lane * a.nq + q] = run[w][lane];
}

// ---- partial counting on the owner shard -----------------------------------------------------
struct ShardArgs {
  SearchArgs sa;  // table, postings, counters, pool (reused field meanings)
  const uint32_t *codes;
  const uint64_t *seg_off;
  uint32_t nseg;
  uint32_t *part_n;
  uint64_t *part_base;
  uint32_t *work;  // [0] CTA-tier cursor, [1] global-tier count, [3] warp-tier cursor, [4] CTA-tier count
  uint32_t *ovf;   // [nseg] segments that need the global-memory histogram
  uint32_t *mid;   // [nseg] segments the warp tier passed on to the CTA tier
};

// ---- warp tier: one warp per segment, warp-private 512-slot histogram ------------------------
// With 8 shards a 350-residue query leaves ~43 k-mers per segment: a CTA per segment spends its
// time in barriers and table sweeps.  Here the candidate mechanism of the search kernels is
// reused with kmin = 1: a slot enters the list when its subject is first seen, so the list IS the
// set of distinct subjects and no sweep is needed.
constexpr int SWP_H = 512, SWP_WARPS = 8, SWP_MAXN = 256;
struct __align__(16) ShardWarpSmem {
  uint32_t hkeys[SWP_H];
  uint16_t hcnt[SWP_H];
  uint16_t cand[SWP_H];
  uint32_t ncand, flags, pad0, pad1;
};

__device__ __forceinline__ void swp_clear(ShardWarpSmem &s, unsigned lane) {
  uint4 *hk = reinterpret_cast<uint4 *>(s.hkeys);
  uint4 *hc = reinterpret_cast<uint4 *>(s.hcnt);
  const uint4 E = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY), Z = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < SWP_H / 4 / 32; ++i) hk[i * 32 + lane] = E;
#pragma unroll
  for (int i = 0; i < SWP_H / 8 / 32; ++i) hc[i * 32 + lane] = Z;
  if (lane == 0) {
    s.ncand = 0;
    s.flags = 0;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(SWP_WARPS * 32, 6) k_shard_count_w(ShardArgs a) {
  __shared__ ShardWarpSmem sm[SWP_WARPS];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  ShardWarpSmem &s = sm[w];
  const WarpHashT<SWP_H> hv{s.hkeys, s.hcnt};
  const CandList cl{&s.ncand, &s.flags, s.cand, nullptr, (uint32_t)SWP_H};
  unsigned long long my_incr = 0, my_lookups = 0;
  uint32_t it_next = 0;
  if (lane == 0) it_next = atomicAdd(&a.work[3], 1u);
  for (;;) {
    const uint32_t seg = __shfl_sync(0xFFFFFFFFu, it_next, 0);
    if (seg >= a.nseg) break;
    if (lane == 0) it_next = atomicAdd(&a.work[3], 1u);
    const uint64_t b = a.seg_off[seg];
    const uint64_t n64 = a.seg_off[seg + 1] - b;
    if (n64 == 0) {
      if (lane == 0) {
        a.part_n[seg] = 0;
        a.part_base[seg] = 0;
      }
      continue;
    }
    if (n64 > SWP_MAXN) {
      if (lane == 0) a.mid[atomicAdd(&a.work[4], 1u)] = seg;
      continue;
    }
    const int n = (int)n64;
    swp_clear(s, lane);
    unsigned long long incr = 0;
    constexpr int U = 4;
    for (int base = 0; base < n; base += U * 32) {
      uint64_t ent[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pos = base + u * 32 + (int)lane;
        ent[u] = 0;
        if (pos < n) {
          const uint32_t d = a.codes[b + pos];
          if (d >= a.sa.d_lo && d < a.sa.d_hi) ent[u] = ldg_entry(a.sa.table + (d - a.sa.d_lo));
        }
      }
      warp_consume<U>(a.sa, ent, hv, 1u, cl, incr);
    }
    __syncwarp();
    if (*(volatile uint32_t *)&s.flags) {
      if (lane == 0) a.mid[atomicAdd(&a.work[4], 1u)] = seg;
      __syncwarp();
      continue;
    }
    const uint32_t c = *(volatile uint32_t *)&s.ncand;
    unsigned long long pbase = 0;
    if (lane == 0 && c) pbase = atomicAdd(&a.sa.counters[CNT_POOL], (unsigned long long)c);
    pbase = __shfl_sync(0xFFFFFFFFu, pbase, 0);
    const bool fits = pbase + c <= a.sa.pool_cap;
    if (fits)
      for (uint32_t i = lane; i < c; i += 32) {
        const uint32_t sl = s.cand[i];
        a.sa.pool[pbase + i] = (uint64_t)hv.key_at(sl) | ((uint64_t)hv.count_at(sl) << 32);
      }
    if (lane == 0) {
      a.part_n[seg] = fits ? c : 0;
      a.part_base[seg] = fits ? pbase : 0;
      if (!fits) atomicOr(&a.sa.counters[CNT_STATUS], (unsigned long long)ST_POOL_OVERFLOW);
    }
    my_incr += incr;
    if (lane == 0) my_lookups += (unsigned long long)n;
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_incr += __shfl_down_sync(0xFFFFFFFFu, my_incr, o);
    my_lookups += __shfl_down_sync(0xFFFFFFFFu, my_lookups, o);
  }
  if (lane == 0) {
    if (my_incr) atomicAdd(&a.sa.counters[CNT_INCR], my_incr);
    if (my_lookups) atomicAdd(&a.sa.counters[CNT_LOOKUPS], my_lookups);
  }
}

template <class Hash>
__device__ __forceinline__ void hash_add(const Hash &hv, uint32_t id, uint32_t c, uint32_t *flags) {
  uint32_t slot = hv.home(id);
#pragma unroll 1
  for (int probe = 0; probe < Hash::kMaxProbe; ++probe) {
    const uint32_t cur = hv.cas(slot, id);
    if (cur == EMPTY || cur == id) {
      hv.add(slot, c);
      return;
    }
    slot = (slot + 1) & hv.mask;
  }
  atomicOr(flags, 1u);
}

// probe + count all codes of one segment into hv; returns through ss.flags whether it overflowed
template <int THREADS, class Hash>
__device__ __forceinline__ void count_segment(const ShardArgs &a, const Hash &hv, uint64_t b, uint32_t n,
                                              SelectScratch &ss, unsigned long long &incr) {
  const int tid = threadIdx.x;
  const CandList cl{&ss.ncand, &ss.flags, nullptr, nullptr, 0u};
  constexpr int U = 4;
  for (uint32_t base = 0; base < n; base += U * THREADS) {
    uint64_t ent[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t pos = base + u * THREADS + tid;
      ent[u] = 0;
      if (pos < n) {
        const uint32_t d = a.codes[b + pos];
        if (d >= a.sa.d_lo && d < a.sa.d_hi) ent[u] = ldg_entry(a.sa.table + (d - a.sa.d_lo));
      }
    }
    warp_consume<U>(a.sa, ent, hv, 0xFFFFFFFFu, cl, incr);  // kmin never reached: no candidates
  }
}

// write every (subject, count) of the histogram to the pool
template <int THREADS, class Hash>
__device__ __forceinline__ void emit_segment(const ShardArgs &a, const Hash &hv, uint32_t slots, uint32_t seg,
                                             SelectScratch &ss) {
  const int tid = threadIdx.x;
  if (tid == 0) ss.nout = 0;
  __syncthreads();
  uint32_t mine = 0;
  for (uint32_t i = tid; i < slots; i += THREADS) mine += hv.key_at(i) != EMPTY ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
  if ((tid & 31) == 0 && mine) atomicAdd(&ss.nout, mine);
  __syncthreads();
  const uint32_t total = ss.nout;
  if (tid == 0) {
    ss.base = total ? atomicAdd(&a.sa.counters[CNT_POOL], (unsigned long long)total) : 0ull;
    ss.ncand = 0;
  }
  __syncthreads();
  const unsigned long long base = ss.base;
  const bool fits = base + total <= a.sa.pool_cap;
  if (fits) {
    for (uint32_t i = tid; i < slots; i += THREADS) {
      const uint32_t k = hv.key_at(i);
      if (k != EMPTY) a.sa.pool[base + atomicAdd(&ss.ncand, 1u)] = (uint64_t)k | ((uint64_t)hv.count_at(i) << 32);
    }
  }
  if (tid == 0) {
    if (fits) {
      a.part_n[seg] = total;
      a.part_base[seg] = base;
    } else {
      a.part_n[seg] = 0;
      a.part_base[seg] = 0;
      atomicOr(&a.sa.counters[CNT_STATUS], (unsigned long long)ST_POOL_OVERFLOW);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int ilog2_pow2(uint32_t x) { return 31 - __clz(x); }

__global__ void __launch_bounds__(SH_THREADS) k_shard_count(ShardArgs a) {
  extern __shared__ __align__(16) uint32_t dyn[];
  uint32_t *hkeys = dyn;            // [SH_H]
  uint32_t *hcnt2 = dyn + SH_H;     // [SH_H/2]
  __shared__ SelectScratch ss;
  __shared__ uint32_t s_seg;
  constexpr int THREADS = SH_THREADS;
  const int tid = threadIdx.x;
  unsigned long long my_incr = 0, my_lookups = 0;
  for (;;) {
    if (tid == 0) s_seg = atomicAdd(&a.work[0], 1u);
    __syncthreads();
    const uint32_t it = s_seg;
    __syncthreads();
    if (it >= a.work[4]) break;
    const uint32_t seg = a.mid[it];
    const uint64_t b = a.seg_off[seg];
    const uint64_t n64 = a.seg_off[seg + 1] - b;
    if (n64 == 0) {
      if (tid == 0) {
        a.part_n[seg] = 0;
        a.part_base[seg] = 0;
      }
      continue;
    }
    if (n64 > SEG_MAX_U16) {
      if (tid == 0) a.ovf[atomicAdd(&a.work[1], 1u)] = seg;
      continue;
    }
    const uint32_t n = (uint32_t)n64;
    uint32_t Hq = 256;
    while (Hq < 8 * n && Hq < (uint32_t)SH_H) Hq <<= 1;
    for (;;) {
      const uint4 E = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY), Z = make_uint4(0, 0, 0, 0);
      for (uint32_t i = tid; i < Hq / 4; i += THREADS) reinterpret_cast<uint4 *>(hkeys)[i] = E;
      for (uint32_t i = tid; i < Hq / 8; i += THREADS) reinterpret_cast<uint4 *>(hcnt2)[i] = Z;
      if (tid == 0) {
        ss.ncand = 0;
        ss.flags = 0;
      }
      __syncthreads();
      const SmemHash hv{hkeys, hcnt2, Hq - 1u, 32 - ilog2_pow2(Hq)};
      unsigned long long incr = 0;
      count_segment<THREADS>(a, hv, b, n, ss, incr);
      __syncthreads();
      if (ss.flags == 0) {
        my_incr += incr;
        if (tid == 0) my_lookups += n;
        emit_segment<THREADS>(a, hv, Hq, seg, ss);
        break;
      }
      __syncthreads();
      if (Hq < (uint32_t)SH_H) {
        Hq = SH_H;  // retry once with the full table
        continue;
      }
      if (tid == 0) a.ovf[atomicAdd(&a.work[1], 1u)] = seg;
      break;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_incr += __shfl_down_sync(0xFFFFFFFFu, my_incr, o);
    my_lookups += __shfl_down_sync(0xFFFFFFFFu, my_lookups, o);
  }
  if ((tid & 31) == 0) {
    if (my_incr) atomicAdd(&a.sa.counters[CNT_INCR], my_incr);
    if (my_lookups) atomicAdd(&a.sa.counters[CNT_LOOKUPS], my_lookups);
  }
}

// segments whose subject set outgrew shared memory: histogram in per-CTA global scratch (L2)
__global__ void __launch_bounds__(SH_THREADS) k_shard_count_g(ShardArgs a) {
  __shared__ SelectScratch ss;
  __shared__ unsigned long long s_total;
  constexpr int THREADS = SH_THREADS;
  const int tid = threadIdx.x;
  const uint32_t HG = a.sa.ghash_slots;
  uint32_t *gkeys = a.sa.ghash + (size_t)blockIdx.x * 3 * HG;
  uint32_t *gcnt = gkeys + HG;
  const uint32_t count = a.work[1];
  unsigned long long my_incr = 0, my_lookups = 0;
  for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
    const uint32_t seg = a.ovf[it];
    const uint64_t b = a.seg_off[seg];
    const uint64_t n64 = a.seg_off[seg + 1] - b;
    const uint32_t n = (uint32_t)n64;
    if (tid == 0) s_total = 0;
    __syncthreads();
    unsigned long long tot = 0;
    for (uint32_t pos = tid; pos < n; pos += THREADS) {
      const uint32_t d = a.codes[b + pos];
      if (d >= a.sa.d_lo && d < a.sa.d_hi) tot += ldg_entry(a.sa.table + (d - a.sa.d_lo)) >> ENTRY_VALUE_BITS;
    }
    atomicAdd(&s_total, tot);
    __syncthreads();
    const unsigned long long T = s_total;
    uint32_t Hq = 1024;
    while (Hq < HG && (unsigned long long)Hq < 2 * T) Hq <<= 1;
    const GmemHash hv{gkeys, gcnt, Hq - 1u, 32 - ilog2_pow2(Hq)};
    for (uint32_t i = tid; i < Hq; i += THREADS) {
      gkeys[i] = EMPTY;
      gcnt[i] = 0;
    }
    if (tid == 0) {
      ss.ncand = 0;
      ss.flags = 0;
    }
    __syncthreads();
    unsigned long long incr = 0;
    count_segment<THREADS>(a, hv, b, n, ss, incr);
    __syncthreads();
    if (ss.flags) {
      if (tid == 0) {
        a.part_n[seg] = 0;
        a.part_base[seg] = 0;
        atomicOr(&a.sa.counters[CNT_STATUS], (unsigned long long)ST_GHASH_OVERFLOW);
      }
      __syncthreads();
      continue;
    }
    my_incr += incr;
    if (tid == 0) my_lookups += n;
    emit_segment<THREADS>(a, hv, Hq, seg, ss);
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_incr += __shfl_down_sync(0xFFFFFFFFu, my_incr, o);
    my_lookups += __shfl_down_sync(0xFFFFFFFFu, my_lookups, o);
  }
  if ((tid & 31) == 0) {
    if (my_incr) atomicAdd(&a.sa.counters[CNT_INCR], my_incr);
    if (my_lookups) atomicAdd(&a.sa.counters[CNT_LOOKUPS], my_lookups);
  }
}

__global__ void k_shard_gather(const uint32_t *__restrict__ part_n, const uint64_t *__restrict__ part_base,
                               const uint64_t *__restrict__ part_off, const uint64_t *__restrict__ pool,
                               uint32_t nseg, uint64_t *out) {
  const uint32_t seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (seg >= nseg) return;
  const uint32_t n = part_n[seg];
  const uint64_t src = part_base[seg], dst = part_off[seg];
  for (uint32_t i = threadIdx.x & 31; i < n; i += 32) out[dst + i] = pool[src + i];
}

// ---- merge of the partial lists on the query's home rank ---------------------------------------
struct MergeArgs {
  SearchArgs sa;             // options, n_hits / hit_base / pool / counters, size_in_kmer (input)
  const uint64_t *part;      // (subject | count << 32)
  const uint64_t *part_off;  // [n_shards][nq+1] absolute offsets into part
  int n_shards;
  uint32_t *work;            // [0] CTA-tier cursor, [1] global-tier count, [3] warp-tier cursor, [4] CTA-tier count
  uint32_t *ovf;
  uint32_t *mid;
};

// add `c` to the count of `id`; duplicates of an id inside the batch of 32 are summed first
template <int H>
__device__ __forceinline__ void warp_add(const WarpHashT<H> &hv, bool valid, uint32_t id, uint32_t c, uint32_t kmin,
                                         const CandList &cl) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned act = __ballot_sync(0xFFFFFFFFu, valid);
  if (act == 0) return;
  unsigned peers = 0;
  uint32_t mult = 0;
  if (valid) {
    peers = __match_any_sync(act, id);
    mult = __reduce_add_sync(peers, c);
  }
  if (valid && lane == (unsigned)(__ffs(peers) - 1)) {
    uint32_t slot = hv.home(id);
    int probe = 0;
#pragma unroll 1
    for (; probe < MAX_PROBE; ++probe) {
      uint32_t key = *(volatile uint32_t *)(hv.keys + slot);
      if (key == EMPTY) {
        key = atomicCAS(hv.keys + slot, EMPTY, id);
        if (key == EMPTY) key = id;
      }
      if (key == id) {
        const uint32_t old = hv.cnt[slot];
        hv.cnt[slot] = (uint16_t)(old + mult);
        if (old < kmin && old + mult >= kmin) push_candidate(cl, slot);
        break;
      }
      slot = (slot + 1) & (H - 1);
    }
    if (probe == MAX_PROBE) atomicOr(cl.flags, 1u);
  }
  __syncwarp();
}

// warp tier of the merge: queries whose partial lists hold <= SWP_MAXN entries in total
__global__ void __launch_bounds__(SWP_WARPS * 32, 6) k_shard_merge_w(MergeArgs a) {
  __shared__ ShardWarpSmem sm[SWP_WARPS];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  ShardWarpSmem &s = sm[w];
  const WarpHashT<SWP_H> hv{s.hkeys, s.hcnt};
  const CandList cl{&s.ncand, &s.flags, s.cand, nullptr, (uint32_t)SWP_H};
  const uint32_t N = a.sa.max_results > 0 ? (uint32_t)a.sa.max_results : 0u;
  uint32_t it_next = 0;
  if (lane == 0) it_next = atomicAdd(&a.work[3], 1u);
  for (;;) {
    const uint32_t q = __shfl_sync(0xFFFFFFFFu, it_next, 0);
    if (q >= a.sa.nq) break;
    if (lane == 0) it_next = atomicAdd(&a.work[3], 1u);
    const int K = a.sa.size_in_kmer[q];
    if (lane == 0) {
      a.sa.n_hits[q] = 0;
      a.sa.hit_base[q] = 0;
    }
    if (K < 7) continue;  // search_protein.go:74-76
    uint64_t total = 0;
    for (int sh = 0; sh < a.n_shards; ++sh) {
      const uint64_t *po = a.part_off + (size_t)sh * (a.sa.nq + 1) + q;
      total += po[1] - po[0];
    }
    if (total == 0) continue;
    if (total > SWP_MAXN || (uint32_t)K > SEG_MAX_U16) {
      if (lane == 0) a.mid[atomicAdd(&a.work[4], 1u)] = q;
      continue;
    }
    const uint32_t kmin = filter_kmin(a.sa.min_kmatch, a.sa.min_kratio, K);
    swp_clear(s, lane);
    for (int sh = 0; sh < a.n_shards; ++sh) {
      const uint64_t *po = a.part_off + (size_t)sh * (a.sa.nq + 1) + q;
      const uint64_t b = po[0], e = po[1];
      for (uint64_t i0 = b; i0 < e; i0 += 32) {
        const uint64_t i = i0 + lane;
        const uint64_t v = i < e ? a.part[i] : 0ull;
        warp_add(hv, i < e, (uint32_t)v, (uint32_t)(v >> 32), kmin, cl);
      }
    }
    __syncwarp();
    const uint32_t c = *(volatile uint32_t *)&s.ncand;
    if (*(volatile uint32_t *)&s.flags || c > (uint32_t)W_CAND) {  // the CTA tier ranks large candidate sets
      if (lane == 0) a.mid[atomicAdd(&a.work[4], 1u)] = q;
      __syncwarp();
      continue;
    }
    const uint32_t nout = c < N ? c : N;
    if (nout) {
      // c <= 64: rank by counting, two candidates per lane (same as the search kernels)
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&a.sa.counters[CNT_POOL], (unsigned long long)nout);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      const bool fits = base + nout <= a.sa.pool_cap;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t i = lane + 32 * hh;
        if (i < c) {
          const uint32_t sl = s.cand[i];
          const uint64_t me = composite(hv.key_at(sl), hv.count_at(sl));
          uint32_t rank = 0;
          for (uint32_t j = 0; j < c; ++j) {
            const uint32_t sj = s.cand[j];
            rank += composite(hv.key_at(sj), hv.count_at(sj)) < me ? 1u : 0u;
          }
          if (rank < nout && fits) a.sa.pool[base + rank] = decomposite(me);
        }
      }
      if (lane == 0) {
        if (fits) {
          a.sa.n_hits[q] = nout;
          a.sa.hit_base[q] = (uint32_t)base;
        } else {
          atomicOr(&a.sa.counters[CNT_STATUS], (unsigned long long)ST_POOL_OVERFLOW);
        }
      }
    }
    __syncwarp();
  }
}

template <int THREADS, class Hash, class CandPush>
__device__ __forceinline__ void merge_query(const MergeArgs &a, uint32_t q, const Hash &hv, uint32_t slots,
                                            uint32_t kmin, SelectScratch &ss, CandPush push) {
  const int tid = threadIdx.x;
  for (int s = 0; s < a.n_shards; ++s) {
    const uint64_t *po = a.part_off + (size_t)s * (a.sa.nq + 1) + q;
    const uint64_t b = po[0], e = po[1];
    for (uint64_t i = b + tid; i < e; i += THREADS) {
      const uint64_t v = a.part[i];
      hash_add(hv, (uint32_t)v, (uint32_t)(v >> 32), &ss.flags);
    }
  }
  __syncthreads();
  if (ss.flags) return;
  for (uint32_t i = tid; i < slots; i += THREADS)
    if (hv.key_at(i) != EMPTY && hv.count_at(i) >= kmin) push(atomicAdd(&ss.ncand, 1u), i);
  __syncthreads();
}

__global__ void __launch_bounds__(SH_THREADS) k_shard_merge(MergeArgs a) {
  extern __shared__ __align__(16) uint32_t dyn[];
  uint32_t *hkeys = dyn;                                            // [SH_H]
  uint32_t *hcnt2 = dyn + SH_H;                                     // [SH_H/2]
  uint16_t *cand = reinterpret_cast<uint16_t *>(dyn + SH_H + SH_H / 2);  // [SH_H]
  __shared__ SelectScratch ss;
  __shared__ uint32_t s_q;
  constexpr int THREADS = SH_THREADS;
  const int tid = threadIdx.x;
  for (;;) {
    if (tid == 0) s_q = atomicAdd(&a.work[0], 1u);
    __syncthreads();
    const uint32_t it = s_q;
    __syncthreads();
    if (it >= a.work[4]) break;
    const uint32_t q = a.mid[it];
    const int K = a.sa.size_in_kmer[q];
    uint64_t total = 0;
    for (int s = 0; s < a.n_shards; ++s) {
      const uint64_t *po = a.part_off + (size_t)s * (a.sa.nq + 1) + q;
      total += po[1] - po[0];
    }
    if (total > SH_H / 2 || (uint32_t)K > SEG_MAX_U16) {
      if (tid == 0) a.ovf[atomicAdd(&a.work[1], 1u)] = q;
      continue;
    }
    const uint32_t kmin = filter_kmin(a.sa.min_kmatch, a.sa.min_kratio, K);
    uint32_t Hq = 256;
    while (Hq < 2 * (uint32_t)total) Hq <<= 1;
    const uint4 E = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY), Z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < Hq / 4; i += THREADS) reinterpret_cast<uint4 *>(hkeys)[i] = E;
    for (uint32_t i = tid; i < Hq / 8; i += THREADS) reinterpret_cast<uint4 *>(hcnt2)[i] = Z;
    if (tid == 0) {
      ss.ncand = 0;
      ss.flags = 0;
    }
    __syncthreads();
    const SmemHash hv{hkeys, hcnt2, Hq - 1u, 32 - ilog2_pow2(Hq)};
    merge_query<THREADS>(a, q, hv, Hq, kmin, ss, [&](uint32_t i, uint32_t slot) { cand[i] = (uint16_t)slot; });
    if (ss.flags) {  // probe budget exhausted (pathological clustering): global pass
      if (tid == 0) a.ovf[atomicAdd(&a.work[1], 1u)] = q;
      __syncthreads();
      continue;
    }
    const uint32_t c = ss.ncand;
    select_and_emit<THREADS>(a.sa, q, hv, [&](uint32_t i) -> uint32_t { return cand[i]; }, c, ss);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SH_THREADS) k_shard_merge_g(MergeArgs a) {
  __shared__ SelectScratch ss;
  constexpr int THREADS = SH_THREADS;
  const int tid = threadIdx.x;
  const uint32_t HG = a.sa.ghash_slots;
  uint32_t *gkeys = a.sa.ghash + (size_t)blockIdx.x * 3 * HG;
  uint32_t *gcnt = gkeys + HG;
  uint32_t *gcand = gcnt + HG;
  const uint32_t count = a.work[1];
  for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
    const uint32_t q = a.ovf[it];
    const int K = a.sa.size_in_kmer[q];
    uint64_t total = 0;
    for (int s = 0; s < a.n_shards; ++s) {
      const uint64_t *po = a.part_off + (size_t)s * (a.sa.nq + 1) + q;
      total += po[1] - po[0];
    }
    const uint32_t kmin = filter_kmin(a.sa.min_kmatch, a.sa.min_kratio, K);
    uint32_t Hq = 1024;
    while (Hq < HG && (uint64_t)Hq < 2 * total) Hq <<= 1;
    for (uint32_t i = tid; i < Hq; i += THREADS) {
      gkeys[i] = EMPTY;
      gcnt[i] = 0;
    }
    if (tid == 0) {
      ss.ncand = 0;
      ss.flags = 0;
    }
    __syncthreads();
    const GmemHash hv{gkeys, gcnt, Hq - 1u, 32 - ilog2_pow2(Hq)};
    merge_query<THREADS>(a, q, hv, Hq, kmin, ss, [&](uint32_t i, uint32_t slot) { gcand[i] = slot; });
    if (ss.flags) {
      if (tid == 0) atomicOr(&a.sa.counters[CNT_STATUS], (unsigned long long)ST_GHASH_OVERFLOW);
      __syncthreads();
      continue;
    }
    const uint32_t c = ss.ncand;
    select_and_emit<THREADS>(a.sa, q, hv, [&](uint32_t i) -> uint32_t { return gcand[i]; }, c, ss);
    __syncthreads();
  }
}

constexpr size_t COUNT_SMEM = (SH_H + SH_H / 2) * 4;
constexpr size_t MERGE_SMEM = (SH_H + SH_H / 2) * 4 + SH_H * 2;

static int ensure_attrs(kaamer_gpu *h) {  // function attributes belong to the device context
  if (h->shard_attrs_ready) return KAAMER_OK;
  KCUDA(cudaFuncSetAttribute(k_shard_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COUNT_SMEM));
  KCUDA(cudaFuncSetAttribute(k_shard_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MERGE_SMEM));
  h->shard_attrs_ready = true;
  return KAAMER_OK;
}

}  // namespace kaamer

using namespace kaamer;

static int check_fences(const uint64_t *fences, int n_shards) {
  if (!fences || n_shards < 1 || n_shards > MAX_SHARDS) {
    set_error("n_shards must be in 1..%d", MAX_SHARDS);
    return KAAMER_ERR_ARG;
  }
  if (fences[0] != 0 || fences[n_shards] != DENSE_SPACE) {
    set_error("fences must start at 0 and end at the dense code space size (%llu)", (unsigned long long)DENSE_SPACE);
    return KAAMER_ERR_ARG;
  }
  for (int s = 0; s < n_shards; ++s)
    if (fences[s] > fences[s + 1]) {
      set_error("fences must be non-decreasing");
      return KAAMER_ERR_ARG;
    }
  return KAAMER_OK;
}

extern "C" {

uint64_t kaamer_gpu_dense_space(void) { return DENSE_SPACE; }

int kaamer_gpu_shard_route(kaamer_gpu_t *h, const uint8_t *d_residues, const uint64_t *d_seq_off, uint32_t nq,
                           const uint64_t *fences, int n_shards, uint32_t *d_counts, const uint64_t *d_offsets,
                           uint32_t *d_codes, int32_t *d_size_in_kmer, void *stream) {
  if (!h || (nq && (!d_residues || !d_seq_off)) || !d_counts) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  KCHECK(check_fences(fences, n_shards));
  if (nq == 0) return KAAMER_OK;
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  RouteArgs a{};
  a.res = d_residues;
  a.off = d_seq_off;
  a.nq = nq;
  a.n_shards = n_shards;
  for (int s = 0; s <= n_shards; ++s) a.fences[s] = (uint32_t)fences[s];  // DENSE_SPACE < 2^32
  a.counts = d_counts;
  a.offsets = d_offsets;
  a.codes = d_codes;
  a.size_in_kmer = d_size_in_kmer;
  const unsigned grid = (nq + 7) / 8;
  if (!d_codes) {
    if (!d_size_in_kmer) {
      set_error("count pass needs d_size_in_kmer");
      return KAAMER_ERR_ARG;
    }
    k_route<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  } else {
    if (!d_offsets) {
      set_error("fill pass needs d_offsets");
      return KAAMER_ERR_ARG;
    }
    k_route<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  }
  h->prof_all_launches += 1;
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

int kaamer_gpu_shard_count(kaamer_gpu_t *h, const uint32_t *d_codes, const uint64_t *d_seg_off, uint32_t n_segments,
                           uint32_t *d_part_n, uint64_t *d_part_base, uint64_t *d_pool, uint64_t pool_cap,
                           uint64_t *d_counters, void *stream) {
  if (!h || !d_seg_off || !d_part_n || !d_part_base || !d_counters || (pool_cap && !d_pool)) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  if (!h->idx.table) {
    set_error("no index resident");
    return KAAMER_ERR_ARG;
  }
  KCHECK(ensure_attrs(h));
  cudaStream_t st = (cudaStream_t)stream;
  KCUDA(cudaMemsetAsync(d_counters, 0, CNT_N * sizeof(uint64_t), st));
  if (n_segments == 0) return KAAMER_OK;
  SearchWorkspace &ws = h->ws;
  KCHECK(ws.lists.ensure((size_t)2 * n_segments + 16));
  const uint32_t HG = 1u << 20;
  const int g_ctas = h->sm_count;
  KCHECK(ws.ghash.ensure((size_t)g_ctas * 3 * HG));
  ShardArgs a{};
  a.sa.table = h->idx.table;
  a.sa.d_lo = h->idx.d_lo;
  a.sa.d_hi = h->idx.d_hi;
  a.sa.postings = h->idx.postings;
  a.sa.pool = d_pool;
  a.sa.pool_cap = pool_cap;
  a.sa.counters = (unsigned long long *)d_counters;
  a.sa.ghash = ws.ghash.p;
  a.sa.ghash_slots = HG;
  a.codes = d_codes;
  a.seg_off = d_seg_off;
  a.nseg = n_segments;
  a.part_n = d_part_n;
  a.part_base = d_part_base;
  a.ovf = ws.lists.p;
  a.mid = ws.lists.p + n_segments;
  a.work = ws.lists.p + (size_t)2 * n_segments;
  KCUDA(cudaMemsetAsync(a.work, 0, 8 * sizeof(uint32_t), st));
  unsigned grid = (unsigned)h->sm_count * 4u;
  unsigned wgrid = (unsigned)h->sm_count * 6u;
  if (wgrid > (n_segments + SWP_WARPS - 1) / SWP_WARPS) wgrid = (n_segments + SWP_WARPS - 1) / SWP_WARPS;
  profile_begin(h, st, 0);
  k_shard_count_w<<<wgrid, SWP_WARPS * 32, 0, st>>>(a);
  profile_end(h, st);
  profile_begin(h, st, 1);
  k_shard_count<<<grid, SH_THREADS, COUNT_SMEM, st>>>(a);
  profile_end(h, st);
  k_shard_count_g<<<g_ctas, SH_THREADS, 0, st>>>(a);
  h->prof_all_launches += 3;
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

int kaamer_gpu_shard_gather(kaamer_gpu_t *h, const uint32_t *d_part_n, const uint64_t *d_part_base,
                            const uint64_t *d_part_off, const uint64_t *d_pool, uint32_t n_segments, uint64_t *d_out,
                            void *stream) {
  if (!h || (n_segments && (!d_part_n || !d_part_base || !d_part_off))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  if (n_segments == 0) return KAAMER_OK;
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)(((uint64_t)n_segments * 32 + 255) / 256);
  k_shard_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(d_part_n, d_part_base, d_part_off, d_pool, n_segments, d_out);
  h->prof_all_launches += 1;
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

int kaamer_gpu_shard_merge(kaamer_gpu_t *h, const uint64_t *d_part, const uint64_t *d_part_off, int n_shards,
                           uint32_t nq, const int32_t *d_size_in_kmer, const kaamer_opts *opts,
                           const kaamer_dev_result *d_out, void *stream) {
  if (!h || !opts || !d_out || (nq && (!d_part_off || !d_size_in_kmer))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  if (n_shards < 1 || n_shards > MAX_SHARDS) {
    set_error("n_shards must be in 1..%d", MAX_SHARDS);
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  KCHECK(ensure_attrs(h));
  cudaStream_t st = (cudaStream_t)stream;
  KCUDA(cudaMemsetAsync(d_out->counters, 0, CNT_N * sizeof(uint64_t), st));
  if (nq == 0) return KAAMER_OK;
  SearchWorkspace &ws = h->ws;
  KCHECK(ws.kmin.ensure((size_t)2 * nq + 16));  // tier lists + work counters (ws.lists may be in use by a shard pass)
  const uint32_t HG = 1u << 20;
  const int g_ctas = h->sm_count;
  KCHECK(ws.ghash.ensure((size_t)g_ctas * 3 * HG));
  MergeArgs a{};
  a.sa.nq = nq;
  a.sa.min_kmatch = opts->min_kmatch;
  a.sa.min_kratio = opts->min_kratio;
  a.sa.max_results = opts->max_results;
  a.sa.n_hits = d_out->n_hits;
  a.sa.hit_base = d_out->hit_base;
  a.sa.size_in_kmer = const_cast<int32_t *>(d_size_in_kmer);
  a.sa.pool = d_out->pool;
  a.sa.pool_cap = d_out->pool_cap;
  a.sa.counters = (unsigned long long *)d_out->counters;
  a.sa.ghash = ws.ghash.p;
  a.sa.ghash_slots = HG;
  a.part = d_part;
  a.part_off = d_part_off;
  a.n_shards = n_shards;
  a.ovf = ws.kmin.p;
  a.mid = ws.kmin.p + nq;
  a.work = ws.kmin.p + (size_t)2 * nq;
  KCUDA(cudaMemsetAsync(a.work, 0, 8 * sizeof(uint32_t), st));
  unsigned grid = (unsigned)h->sm_count * 3u;
  unsigned wgrid = (unsigned)h->sm_count * 6u;
  if (wgrid > (nq + SWP_WARPS - 1) / SWP_WARPS) wgrid = (nq + SWP_WARPS - 1) / SWP_WARPS;
  k_shard_merge_w<<<wgrid, SWP_WARPS * 32, 0, st>>>(a);
  k_shard_merge<<<grid, SH_THREADS, MERGE_SMEM, st>>>(a);
  k_shard_merge_g<<<g_ctas, SH_THREADS, 0, st>>>(a);
  h->prof_all_launches += 3;
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

}  // extern "C"
