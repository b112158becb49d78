// search.cu — the protein search hot path on the device.
//
// One kernel family replaces, per query, the reference's key-producer loop
// (pkg/search/search_protein.go:94-98), KmerSearch (pkg/search/search.go:414-440),
// sortMapByValue (:132-152) and FilterResults (:189-220):
//
//   k_classify    SizeInKmer per query (search.go:290-293), the smallest surviving Kmatch
//                 (FilterResults thresholds, fp64) and the size class
//   k_search_w    class W (SizeInKmer <= 512): ONE WARP PER QUERY, no block barriers.
//                 residues -> packed pair/single codes in smem -> dense 7-mer code -> one
//                 8-byte table probe per k-mer (singleton posting inlined in the entry) ->
//                 per-warp shared-memory open-addressing histogram keyed by subject id ->
//                 candidates collected the moment their count reaches the threshold ->
//                 top-MaxResults by (Kmatch desc, id asc) -> pool
//   k_search_m    class M (<= 2048): same with one CTA per query and a 4096-slot histogram
//   k_search_g    class G: histogram in global memory (L2-resident per-CTA scratch)
//   k_scan/k_gather  CSR compaction of the pool in query order (host API only)
//
// HBM traffic per k-mer: 1 B residue + one 8 B entry (one 64 B HBM access) [+ 4 B per
// posting when the list is not a singleton].  Counts never touch HBM for classes W and M.
// Measured ceiling of the probe stage on B200: 36.5 G random probes/s (profiles/).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include <cub/cub.cuh>

#include "internal.cuh"

#include "search_common.cuh"
#include "search_dense.cuh"
#include "search_dense2.cuh"
#include "search_dense3.cuh"

namespace kaamer {

__global__ void k_classify(SearchArgs a) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  int cls = -1;
  if (q < a.nq) {
    uint64_t b = a.off[q], e = a.off[q + 1];
    long long len = (long long)(e - b);
    long long K = len - KAAMER_KMER_SIZE + 1;  // search.go:290
    if (len > 0 && a.res[e - 1] == '*') K--;   // search.go:291-293
    a.size_in_kmer[q] = (int32_t)K;
    a.n_hits[q] = 0;
    a.hit_base[q] = 0;
    bool go;
    if (a.nt_mode) {
      a.any0[q] = 0;
      go = K >= 1;  // always true for ORFs (>= 21 residues, dna.go:26)
      const long long k = a.min_kmatch > 1 ? a.min_kmatch : 1;
      if (go) a.kmin[q] = k > 0xFFFFFFFEll ? 0xFFFFFFFFu : (uint32_t)k;
    } else {
      go = K >= 7;  // search_protein.go:74-76 (documented deviation: skipped, not fatal)
      if (go) a.kmin[q] = filter_kmin(a.min_kmatch, a.min_kratio, (int32_t)K);
    }
    // Class W2 (warp-per-query with a 1024-slot table for 512 < K <= 2048) was measured and is not
    // used: the warp-per-query kernels are bound by the per-warp counting work, W2's larger
    // per-warp state halves the resident warps and it ran at 10 G lookups/s against 24 G/s for
    // the CTA-per-query class M (profiles/r1_notes.md).
    if (go) {
      if (a.dense) cls = (a.kmin[q] >= 3u && K <= D_MAXK) ? ((a.dense >= 2 && K > a.e_kcap) ? (K > a.e_kcap_l ? 6 : 5) : 4) : 2;
      else cls = K <= a.w_maxk ? 0 : (K <= a.m_maxk ? 1 : 2);
    }
  }
  // warp-aggregated append to the class lists (one atomic per warp and class)
#pragma unroll
  for (int c = 0; c < N_LISTS; ++c) {
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, cls == c);
    if (mask == 0) continue;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&a.list_count[c], (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (cls == c) a.lists[(size_t)c * a.nq + base + __popc(mask & ((1u << lane) - 1u))] = q;
  }
}

// ---- class W: one warp per query ------------------------------------------------------------
template <int H, int MAXK>
struct __align__(16) WarpSmemT {
  uint32_t hkeys[H];
  uint16_t hcnt[H];
  uint16_t pp[MAXK + 8];
  uint16_t cand[W_CAND];
  uint32_t ncand, flags, pad0, pad1;
};

// CLS: index of the class list / counters (0 = W)
template <int H, int MAXK, int WARPS, int MINB, int CLS, bool PEER>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_search_wt(SearchArgs a) {
  __shared__ WarpSmemT<H, MAXK> sm[WARPS];
  __shared__ uint8_t lut[256];
  const PeerView *pv = nullptr;
  if constexpr (PEER) {
    __shared__ PeerView s_peer;
    load_peer_view(&s_peer, a.peer, threadIdx.x, WARPS * 32);
    pv = &s_peer;
  }
  for (int i = threadIdx.x; i < 256; i += WARPS * 32) lut[i] = (uint8_t)aa_code(i);
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  WarpSmemT<H, MAXK> &s = sm[w];
  const WarpHashT<H, false> hv{s.hkeys, s.hcnt};  // fire-and-forget counts, candidates collected by a final sweep
  const CandList cl{&s.ncand, &s.flags, s.cand, nullptr, (uint32_t)W_CAND};
  const uint32_t count = a.list_count[CLS];
  const uint8_t *res_end = a.res + a.off[a.nq];
  const uint32_t nwarps = gridDim.x * WARPS;
  const uint32_t N = a.max_results > 0 ? (uint32_t)a.max_results : 0u;
  unsigned long long my_incr = 0, my_lookups = 0;
  // dynamic scheduling: warps pull the next query from a global cursor (query lengths vary by
  // 10x, a static split left the SMs idle for ~20 % of the kernel); the next index is
  // fetched one query ahead so its latency is hidden
  uint32_t it_next = 0;
  if (lane == 0) it_next = atomicAdd(&a.list_count[N_LISTS + CLS], 1u);
  (void)nwarps;
  for (;;) {
    const uint32_t it = __shfl_sync(0xFFFFFFFFu, it_next, 0);
    if (it >= count) break;
    if (lane == 0) it_next = atomicAdd(&a.list_count[N_LISTS + CLS], 1u);
    const uint32_t q = a.lists[(size_t)CLS * a.nq + it];
    const uint64_t b = a.off[q];
    const int len = (int)(a.off[q + 1] - b);
    const int K = a.size_in_kmer[q];
    const uint32_t kmin = a.kmin[q];
    // residues -> packed codes; the raw bytes are staged in the (not yet cleared) histogram
    {
      uint8_t *raw = reinterpret_cast<uint8_t *>(s.hkeys);
      const int head = stage_bytes<32>(raw, a.res + b, len, res_end, (int)lane);
      __syncwarp();
      const uint8_t *r = raw + head;
      const int ncodes = K + KAAMER_KMER_SIZE - 1;
      for (int i = lane; i < ncodes; i += 32) {
        uint32_t c0 = lut[r[i]];
        uint32_t c1 = (i + 1 < len) ? (uint32_t)lut[r[i + 1]] : CODE_UNKNOWN;
        s.pp[i] = (uint16_t)packed_code(c0, c1);
      }
      __syncwarp();
    }
    unsigned long long q_incr = 0;
    constexpr int U = 4;
    // software pipeline: the probes of round r+1 are in flight while round r is counted
    auto load_round = [&](int base, uint64_t(&e)[U]) {
      uint32_t d[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pos = base + u * 32 + (int)lane;
        ok[u] = pos < K;
        d[u] = ok[u] ? dense_from_packed(s.pp[pos], s.pp[pos + 2], s.pp[pos + 4], s.pp[pos + 6]) : 0u;
      }
      probe_entries<PEER, U>(a, pv, d, ok, e);
    };
    uint64_t nxt[U];
    load_round(0, nxt);
    // (the first probes are in flight while the histogram is cleared)
    {
      uint4 *hk = reinterpret_cast<uint4 *>(s.hkeys);
      uint4 *hc = reinterpret_cast<uint4 *>(s.hcnt);
      const uint4 E = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY), Z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < H / 4 / 32; ++i) hk[i * 32 + lane] = E;
#pragma unroll
      for (int i = 0; i < H / 8 / 32; ++i) hc[i * 32 + lane] = Z;
      if (lane == 0) {
        s.ncand = 0;
        s.flags = 0;
      }
    }
    __syncwarp();
    for (int base = 0; base < K; base += U * 32) {
      uint64_t ent[U];
#pragma unroll
      for (int u = 0; u < U; ++u) ent[u] = nxt[u];
      if (base + U * 32 < K) load_round(base + U * 32, nxt);
      warp_consume<U, PEER>(a, ent, hv, kmin, cl, q_incr, pv);
    }
    __syncwarp();
    uint32_t flags = *(volatile uint32_t *)&s.flags;
    uint32_t c = 0;
    if (!flags) {
      c = warp_collect_candidates(hv, kmin, cl);
      if (c > (uint32_t)W_CAND) flags = 2u;
    }
    if (flags) {
      // histogram or candidate list outgrew the warp's shared memory: hand the query to
      // class M (that kernel starts after this one in stream order)
      if (lane == 0) {
        uint32_t slot = atomicAdd(&a.list_count[1], 1u);
        a.lists[(size_t)a.nq + slot] = q;
      }
      __syncwarp();
      continue;
    }
    my_incr += q_incr;
    if (lane == 0) my_lookups += (unsigned long long)K;
    const uint32_t nout = c < N ? c : N;
    if (nout) {
      // c <= W_CAND = 64: rank by counting, two candidates per lane
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&a.counters[CNT_POOL], (unsigned long long)nout);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      const bool fits = base + nout <= a.pool_cap;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t i = lane + 32 * h;
        if (i < c) {
          const uint32_t sl = s.cand[i];
          const uint64_t me = composite(hv.key_at(sl), hv.count_at(sl));
          uint32_t rank = 0;
          for (uint32_t j = 0; j < c; ++j) {
            const uint32_t sj = s.cand[j];
            rank += composite(hv.key_at(sj), hv.count_at(sj)) < me ? 1u : 0u;
          }
          if (rank < nout && fits) a.pool[base + rank] = decomposite(me);
        }
      }
      if (lane == 0) {
        if (fits) {
          a.n_hits[q] = nout;
          a.hit_base[q] = (uint32_t)base;
        } else {
          atomicOr(&a.counters[CNT_STATUS], (unsigned long long)ST_POOL_OVERFLOW);
        }
      }
      if (a.nt_mode && fits) {
        __syncwarp();
        const uint64_t top = a.pool[base];  // rank 0, written by this warp
        const bool any = warp_any0<PEER>(a, pv, hv, dense_from_packed(s.pp[0], s.pp[2], s.pp[4], s.pp[6]),
                                         (uint32_t)top, (uint32_t)(top >> 32));
        if (lane == 0) a.any0[q] = any ? 1 : 0;
      }
    }
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_incr += __shfl_down_sync(0xFFFFFFFFu, my_incr, o);
    my_lookups += __shfl_down_sync(0xFFFFFFFFu, my_lookups, o);
  }
  if (lane == 0) {
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

template <bool PEER>
__global__ void __launch_bounds__(M_THREADS, M_CTAS) k_search_m(SearchArgs a) {
  __shared__ __align__(16) uint32_t hkeys[M_H];
  __shared__ __align__(16) uint32_t hcnt2[M_H / 2];
  __shared__ uint16_t pp[M_MAXK + 8];
  __shared__ __align__(16) uint16_t cand[M_H];  // one entry per slot: can never overflow
  __shared__ uint8_t lut[256];
  __shared__ SelectScratch ss;
  constexpr int THREADS = M_THREADS;
  const int tid = threadIdx.x;
  const PeerView *pv = nullptr;
  if constexpr (PEER) {
    __shared__ PeerView s_peer;
    load_peer_view(&s_peer, a.peer, tid, THREADS);
    pv = &s_peer;
  }
  for (int i = tid; i < 256; i += THREADS) lut[i] = (uint8_t)aa_code(i);
  __syncthreads();
  // fire-and-forget counts; the candidates are collected by one sweep over the slots per query
  const SmemHashT<false> hv{hkeys, hcnt2, (uint32_t)M_H - 1u, 32 - ilog2_c(M_H)};
  const CandList cl{&ss.ncand, &ss.flags, cand, nullptr, (uint32_t)M_H};
  const uint32_t count = a.list_count[1];
  const uint8_t *res_end = a.res + a.off[a.nq];
  unsigned long long my_incr = 0, my_lookups = 0;
  __shared__ uint32_t s_it;
  for (;;) {
    if (tid == 0) s_it = atomicAdd(&a.list_count[N_LISTS + 1], 1u);
    __syncthreads();
    const uint32_t it = s_it;
    if (it >= count) break;
    const uint32_t q = a.lists[(size_t)a.nq + it];
    const uint64_t b = a.off[q];
    const int len = (int)(a.off[q + 1] - b);
    const int K = a.size_in_kmer[q];
    const uint32_t kmin = a.kmin[q];
    // raw residues staged in the (idle) candidate list with aligned 16-byte loads
    uint8_t *raw = reinterpret_cast<uint8_t *>(cand);
    const int head = stage_bytes<THREADS>(raw, a.res + b, len, res_end, tid);
    __syncthreads();
    const uint8_t *r = raw + head;
    const int ncodes = K + KAAMER_KMER_SIZE - 1;
    for (int i = tid; i < ncodes; i += THREADS) {
      uint32_t c0 = lut[r[i]];
      uint32_t c1 = (i + 1 < len) ? (uint32_t)lut[r[i + 1]] : CODE_UNKNOWN;
      pp[i] = (uint16_t)packed_code(c0, c1);
    }
    __syncthreads();
    unsigned long long q_incr = 0;
    constexpr int U = 4;
    auto load_round = [&](int base, uint64_t(&e)[U]) {
      uint32_t d[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pos = base + u * THREADS + tid;
        ok[u] = pos < K;
        d[u] = ok[u] ? dense_from_packed(pp[pos], pp[pos + 2], pp[pos + 4], pp[pos + 6]) : 0u;
      }
      probe_entries<PEER, U>(a, pv, d, ok, e);
    };
    uint64_t nxt[U];
    load_round(0, nxt);
    // (the first probes are in flight while the histogram is cleared)
    {
      uint4 *hk = reinterpret_cast<uint4 *>(hkeys);
      uint4 *hc = reinterpret_cast<uint4 *>(hcnt2);
      const uint4 E = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY), Z = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < M_H / 4; i += THREADS) hk[i] = E;
      for (int i = tid; i < M_H / 8; i += THREADS) hc[i] = Z;
      if (tid == 0) {
        ss.ncand = 0;
        ss.flags = 0;
      }
    }
    __syncthreads();
    for (int base = 0; base < K; base += U * THREADS) {
      uint64_t ent[U];
#pragma unroll
      for (int u = 0; u < U; ++u) ent[u] = nxt[u];
      if (base + U * THREADS < K) load_round(base + U * THREADS, nxt);
      warp_consume<U, PEER>(a, ent, hv, kmin, cl, q_incr, pv);
    }
    __syncthreads();
    if (!ss.flags) {
      for (uint32_t i = tid; i < (uint32_t)M_H; i += THREADS) {
        const uint32_t cnt = hv.count_at(i);
        if (cnt >= kmin && cnt != 0u) cand[atomicAdd(&ss.ncand, 1u)] = (uint16_t)i;
      }
      __syncthreads();
    }
    if (ss.flags) {
      // histogram full: class G
      if (tid == 0) {
        uint32_t slot = atomicAdd(&a.list_count[3], 1u);  // list 3: hand-offs to class G (second G launch)
        a.lists[(size_t)3 * a.nq + slot] = q;
      }
      __syncthreads();
      continue;
    }
    my_incr += q_incr;
    if (tid == 0) my_lookups += (unsigned long long)K;
    const uint32_t c = ss.ncand;
    select_and_emit<THREADS>(a, q, hv, [&](uint32_t i) -> uint32_t { return cand[i]; }, c, ss);
    __syncthreads();
    if (a.nt_mode) {
      if (tid < 32 && a.n_hits[q]) {
        const uint64_t top = a.pool[a.hit_base[q]];
        const bool any = warp_any0<PEER>(a, pv, hv, dense_from_packed(pp[0], pp[2], pp[4], pp[6]), (uint32_t)top,
                                         (uint32_t)(top >> 32));
        if (tid == 0) a.any0[q] = any ? 1 : 0;
      }
      __syncthreads();
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_incr += __shfl_down_sync(0xFFFFFFFFu, my_incr, o);
    my_lookups += __shfl_down_sync(0xFFFFFFFFu, my_lookups, o);
  }
  if ((tid & 31) == 0) {
    if (my_incr) {
      atomicAdd(&a.counters[CNT_INCR], my_incr);
      atomicAdd(&a.counters[CNT_CLS_INCR + 1], my_incr);
    }
    if (my_lookups) {
      atomicAdd(&a.counters[CNT_LOOKUPS], my_lookups);
      atomicAdd(&a.counters[CNT_CLS_LOOKUPS + 1], my_lookups);
    }
  }
}

// class G: histogram in global memory (per-CTA scratch, stays in L2), codes computed on the fly
constexpr int G_STAGE = 40 * 1024;
template <bool PEER, int GT>
__global__ void __launch_bounds__(GT) k_search_g(SearchArgs a) {
  __shared__ __align__(16) uint8_t s_res[G_STAGE];
  __shared__ SelectScratch ss;
  __shared__ unsigned long long s_total;
  constexpr int THREADS = GT;
  const int tid = threadIdx.x;
  const PeerView *pv = nullptr;
  if constexpr (PEER) {
    __shared__ PeerView s_peer;
    load_peer_view(&s_peer, a.peer, tid, THREADS);
    pv = &s_peer;  // (the first use is behind the __syncthreads() of the query loop)
  }
  const uint32_t HG = a.ghash_slots;
  uint32_t *gkeys = a.ghash + (size_t)blockIdx.x * 3 * HG;
  uint32_t *gcnt = gkeys + HG;
  uint32_t *gcand = gcnt + HG;
  const uint32_t count = a.list_count[a.g_list];
  const uint8_t *res_end = a.res + a.off[a.nq];
  unsigned long long my_incr = 0, my_lookups = 0;
  for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
    const uint32_t q = a.lists[(size_t)a.g_list * a.nq + it];
    const uint64_t b = a.off[q];
    const int K = a.size_in_kmer[q];
    const uint32_t kmin = a.kmin[q];
    // stage the query in shared memory (it is read 14 times per position below, and may live in
    // pinned host memory when the caller's buffer is used without a copy)
    const int len = (int)(a.off[q + 1] - b);
    const bool staged = len + 31 <= G_STAGE;
    __syncthreads();
    int head = 0;
    if (staged) head = stage_bytes<THREADS>(s_res, a.res + b, len, res_end, tid);
    __syncthreads();
    const uint8_t *s = staged ? s_res + head : a.res + b;
    auto dense_at = [&](int pos) -> uint32_t {
      return dense_from_codes(aa_code(s[pos]), aa_code(s[pos + 1]), aa_code(s[pos + 2]), aa_code(s[pos + 3]),
                              aa_code(s[pos + 4]), aa_code(s[pos + 5]), aa_code(s[pos + 6]));
    };
    // pass 1: total postings of the query bounds the number of distinct subjects
    if (tid == 0) s_total = 0;
    __syncthreads();
    unsigned long long tot = 0;
    for (int pos = tid; pos < K; pos += THREADS) {
      tot += probe_entry<PEER>(a, pv, dense_at(pos)) >> ENTRY_VALUE_BITS;
    }
    atomicAdd(&s_total, tot);
    __syncthreads();
    const unsigned long long T = s_total;
    if (tid == 0 && 2 * T > (unsigned long long)HG) atomicMax(&a.counters[CNT_GNEED], 2 * T);
    uint32_t Hq = 1024;
    while (Hq < HG && (unsigned long long)Hq < 2 * T) Hq <<= 1;
    int log2h = 0;
    while ((1u << log2h) < Hq) ++log2h;
    const GmemHash hv{gkeys, gcnt, Hq - 1u, 32 - log2h};
    const CandList cl{&ss.ncand, &ss.flags, nullptr, gcand, HG};
    for (uint32_t i = tid; i < Hq; i += THREADS) {
      gkeys[i] = EMPTY;
      gcnt[i] = 0;
    }
    if (tid == 0) {
      ss.ncand = 0;
      ss.flags = 0;
    }
    __syncthreads();
    unsigned long long q_incr = 0;
    constexpr int U = 4;
    for (int base = 0; base < K; base += U * THREADS) {
      uint64_t ent[U];
      uint32_t d[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pos = base + u * THREADS + tid;
        ok[u] = pos < K;
        d[u] = ok[u] ? dense_at(pos) : 0u;
      }
      probe_entries<PEER, U>(a, pv, d, ok, ent);
      warp_consume<U, PEER>(a, ent, hv, kmin, cl, q_incr, pv);
    }
    __syncthreads();
    if (ss.flags) {
      if (tid == 0) atomicOr(&a.counters[CNT_STATUS], (unsigned long long)ST_GHASH_OVERFLOW);
      __syncthreads();
      continue;
    }
    my_incr += q_incr;
    if (tid == 0) my_lookups += (unsigned long long)K;
    const uint32_t c = ss.ncand;
    select_and_emit<THREADS>(a, q, hv, [&](uint32_t i) -> uint32_t { return gcand[i]; }, c, ss);
    __syncthreads();
    if (a.nt_mode) {
      if (tid < 32 && a.n_hits[q]) {
        const uint64_t top = a.pool[a.hit_base[q]];
        const bool any = warp_any0<PEER>(a, pv, hv, dense_at(0), (uint32_t)top, (uint32_t)(top >> 32));
        if (tid == 0) a.any0[q] = any ? 1 : 0;
      }
      __syncthreads();
    }
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

// ---- presence filter of the attached shards (mode P) ------------------------------------------
// One bit per dense code: does the owner shard hold postings for it?  Every lane reads the entry of
// one code (coalesced 256-byte rows, remote shards streamed through NVLink), a ballot makes the word.
__global__ void __launch_bounds__(256) k_presence(PeerView pv, uint32_t *__restrict__ bits, uint64_t n_words) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t w = warp; w < n_words; w += nwarps) {
    const uint64_t d = w * 32 + lane;
    bool present = false;
    if (d < DENSE_SPACE) {
      uint32_t s = 0;
#pragma unroll
      for (int i = 1; i < MAX_PEER_SHARDS; ++i) s += (uint32_t)d >= pv.fence[i] ? 1u : 0u;
      present = (pv.table[s][d - pv.fence[s]] >> ENTRY_VALUE_BITS) != 0ull;
    }
    const unsigned word = __ballot_sync(0xFFFFFFFFu, present);
    if (lane == 0) bits[w] = word;
  }
}

// Copy of every shard's table range into one local table over the whole key space; the shard of a
// multi-posting entry goes into the top bits of its value (probe_entries / post_ptr).  14.5 GB, read
// from the owners at NVLink streaming speed, once per attach.
__global__ void __launch_bounds__(256) k_replicate_table(PeerView pv, uint64_t *__restrict__ full) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; d < DENSE_SPACE; d += stride) {
    uint32_t s = 0;
#pragma unroll
    for (int i = 1; i < MAX_PEER_SHARDS; ++i) s += (uint32_t)d >= pv.fence[i] ? 1u : 0u;
    uint64_t e = pv.table[s][d - pv.fence[s]];
    if ((e >> ENTRY_VALUE_BITS) >= 2ull) e |= (uint64_t)s << PEER_SHARD_SHIFT;
    full[d] = e;
  }
}

int replicate_table(kaamer_gpu *h, const PeerView &pv, uint64_t *d_full, cudaStream_t st) {
  k_replicate_table<<<h->sm_count * 16, 256, 0, st>>>(pv, d_full);
  KCUDA(cudaGetLastError());
  KCUDA(cudaStreamSynchronize(st));
  return KAAMER_OK;
}

// Built sharded, searched replicated: every shard's postings were copied into one local array (shard s at
// base[s]), so the shard tag of a multi-posting entry becomes a plain offset: value = base[shard] + local.  The
// PEER kernels still decode such an entry (shard 0, whose pointer is the start of the array), and the protein
// search can use its non-PEER kernels on (full table, replicated postings).
struct FlatBases {
  uint64_t base[MAX_PEER_SHARDS];
};
__global__ void __launch_bounds__(256) k_flatten_table(uint64_t *__restrict__ full, FlatBases fb) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; d < DENSE_SPACE; d += stride) {
    const uint64_t e = full[d];
    if ((e >> ENTRY_VALUE_BITS) >= 2ull) {
      const uint64_t val = e & ENTRY_VALUE_MASK;
      const uint64_t flat = fb.base[val >> PEER_SHARD_SHIFT] + (val & PEER_LOCAL_MASK);
      full[d] = (e & ~ENTRY_VALUE_MASK) | flat;
    }
  }
}
int flatten_table(kaamer_gpu *h, uint64_t *d_full, const uint64_t *base, int n_shards, cudaStream_t st) {
  FlatBases fb{};
  for (int i = 0; i < n_shards && i < MAX_PEER_SHARDS; ++i) fb.base[i] = base[i];
  k_flatten_table<<<h->sm_count * 16, 256, 0, st>>>(d_full, fb);
  KCUDA(cudaGetLastError());
  KCUDA(cudaStreamSynchronize(st));
  return KAAMER_OK;
}

int build_presence(kaamer_gpu *h, const PeerView &pv, uint32_t *d_bits, cudaStream_t st) {
  const uint64_t n_words = (DENSE_SPACE + 31) / 32;
  k_presence<<<h->sm_count * 16, 256, 0, st>>>(pv, d_bits, n_words);
  KCUDA(cudaGetLastError());
  KCUDA(cudaStreamSynchronize(st));
  return KAAMER_OK;
}

// ---- CSR compaction (host API) ----------------------------------------------------------
struct WidenU32 {
  __host__ __device__ uint64_t operator()(uint32_t x) const { return x; }
};
__global__ void k_gather_hits(const uint32_t *__restrict__ n_hits, const uint32_t *__restrict__ hit_base,
                              const uint64_t *__restrict__ hit_off, const uint64_t *__restrict__ pool,
                              uint32_t nq, uint32_t *subject, uint32_t *kmatch) {
  // one warp per query
  uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  uint32_t n = n_hits[q];
  uint64_t src = hit_base[q], dst = hit_off[q];
  for (uint32_t i = threadIdx.x & 31; i < n; i += 32) {
    const uint64_t v = pool[src + i];
    subject[dst + i] = (uint32_t)v;
    kmatch[dst + i] = (uint32_t)(v >> 32);
  }
}

// compact device arrays -> the caller-visible pinned result arrays, coalesced 16-byte stores over PCIe
// (the per-query fragments of k_gather_hits would cross the bus as 40-byte writes); n = hit_off[nq] is read
// on the device, so the host needs no round trip to size the copy
__global__ void __launch_bounds__(256) k_hits_to_host(const uint32_t *__restrict__ subj, const uint32_t *__restrict__ km,
                                                      const uint64_t *__restrict__ total, uint64_t cap,
                                                      uint32_t *h_subj, uint32_t *h_km) {
  uint64_t n = *total;
  if (n > cap) n = cap;
  const uint64_t n4 = (n + 3) / 4;  // (the arrays are allocated in multiples of 16 bytes)
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    reinterpret_cast<uint4 *>(h_subj)[i] = reinterpret_cast<const uint4 *>(subj)[i];
    reinterpret_cast<uint4 *>(h_km)[i] = reinterpret_cast<const uint4 *>(km)[i];
  }
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

// Size-class limits for this database.  A query of K k-mers touches about K * p distinct subjects, p =
// background postings per query k-mer = (k-mer occurrences of the database) x P(two 7-mers drawn from the
// amino-acid composition are equal) = n_kmers x 2.7e-9 for Swiss-Prot composition, plus its
// true hits.  The class histograms work up to ~55 % load; a query that overflows its class is handed to
// the next one and searched again, so on dense databases the limits must come down (measured on a
// 2 M-protein database with fixed limits: class W spent 2.9 ms on 7.8 M lookups, most of them repeated
// in class M).  Swiss-Prot scale (p = 0.53) keeps the full 512 / 2048.
static void class_limits(const kaamer_gpu *h, int *w_maxk, int *m_maxk, int *dense) {
  double p = (double)h->idx.n_kmers * 2.7e-9;
  // Dense database: above ~3.5 background postings per query k-mer (measured: 4 M proteins, p = 3.7, 5.08 ms with
  // class D against 5.77 ms; 2 M proteins, p = 1.9, 4.97 ms against 2.62 ms; profiles/r2_dense_probe.jsonl)
  // the shared-memory histograms of classes W
  // and M overflow for ordinary queries, and class D (search_dense.cuh), whose cost per posting is a byte
  // load and a byte store instead of an atomic, takes every query.  KAAMER_DENSE=0/1 forces the choice
  // (A/B measurements, parity tests).
  *dense = p >= 3.5 ? 2 : 0;
  if (const char *env = getenv("KAAMER_DENSE")) {
    const int v = atoi(env);
    // 1: class D; 12 / 11: its second / first design (A/B measurements); 0: off
    *dense = v == 1 ? 2 : (v == 12 ? 3 : (v == 11 ? 1 : 0));
  }
  if (p < 0.5) p = 0.5;
  double w = 0.55 * W_H / p, m = 0.55 * M_H / p;
  *w_maxk = w >= W_MAXK ? W_MAXK : (w < 32 ? 32 : (int)w);
  *m_maxk = m >= M_MAXK ? M_MAXK : (m < *w_maxk ? *w_maxk : (int)m);
  // test hook: KAAMER_CLASS_LIMITS="w,m" forces the limits (the parity tests push ordinary queries
  // through classes M and G with it)
  if (const char *env = getenv("KAAMER_CLASS_LIMITS")) {
    int ew = 0, em = 0;
    if (sscanf(env, "%d,%d", &ew, &em) == 2 && ew >= 0 && ew <= W_MAXK && em >= ew && em <= M_MAXK) {
      *w_maxk = ew;
      *m_maxk = em;
    }
  }
}

// bytes per byte map of class D (test / tuning hook KAAMER_D_MAPKB)
static uint32_t dense_mapb() {
  uint32_t b = D_MAPB_DEFAULT;
  if (const char *env = getenv("KAAMER_D_MAPKB")) {
    const int kb = atoi(env);
    if (kb >= 1 && kb <= 96) b = (uint32_t)kb * 1024u;
  }
  return b;
}

int search_proteins_device(kaamer_gpu *h, const uint8_t *d_res, const uint64_t *d_off, uint32_t nq,
                           const kaamer_opts *o, const kaamer_dev_result *out, cudaStream_t st, int nt_mode,
                           uint8_t *d_any0, const uint64_t *d_prev_counters) {
  if (!h->idx.table) {
    set_error("no index resident");
    return KAAMER_ERR_ARG;
  }
  const bool flat = h->idx.flat_view;  // built sharded, everything replicated: a plain whole index
  const bool peer = h->idx.peer.n > 0 && !flat;
  if (!peer && !flat && (h->idx.d_lo != 0 || h->idx.d_hi != DENSE_SPACE)) {
    set_error("this handle holds the key-range shard [%llu, %llu) only: attach the other shards "
              "(kaamer_gpu_attach_shards) or use the kaamer_gpu_shard_* steps",
              (unsigned long long)h->idx.d_lo, (unsigned long long)h->idx.d_hi);
    return KAAMER_ERR_ARG;
  }
  if (nq == 0) return KAAMER_OK;
  SearchWorkspace &ws = h->ws;
  KCHECK(ws.lists.ensure((size_t)N_LISTS * nq + 16));
  KCHECK(ws.kmin.ensure(nq));
  uint32_t *list_count = ws.lists.p + (size_t)N_LISTS * nq;
  SearchArgs a{};
  a.table = flat ? h->idx.full_table : h->idx.table;
  a.d_lo = flat ? 0 : h->idx.d_lo;
  a.d_hi = flat ? DENSE_SPACE : h->idx.d_hi;
  a.postings = flat ? h->idx.repl_postings : h->idx.postings;
  a.res = d_res;
  a.off = d_off;
  a.nq = nq;
  a.min_kmatch = o->min_kmatch;
  a.min_kratio = o->min_kratio;
  a.max_results = o->max_results;
  a.n_hits = out->n_hits;
  a.hit_base = out->hit_base;
  a.size_in_kmer = out->size_in_kmer;
  a.kmin = ws.kmin.p;
  a.pool = out->pool;
  a.pool_cap = out->pool_cap;
  a.counters = (unsigned long long *)out->counters;
  a.lists = ws.lists.p;
  a.list_count = list_count;
  a.ghash_slots = h->ghash_slots;
  a.nt_mode = nt_mode;
  a.any0 = d_any0;
  a.peer = h->idx.d_peer;
  a.filter = flat ? nullptr : h->idx.filter;  // (the presence bits cover the handle's own key range only)
  class_limits(h, &a.w_maxk, &a.m_maxk, &a.dense);
  a.d_mapb = dense_mapb();
  const size_t d_smem = ((sizeof(DenseSmem) + 15) & ~(size_t)15) + 2 * (size_t)a.d_mapb;
  int d_pf = 8;
  if (const char *env = getenv("KAAMER_D_PF")) d_pf = atoi(env);
  auto d_kernel = [&](bool pr) -> void (*)(SearchArgs) {
    if (d_pf >= 32) return pr ? k_search_d<true, 32> : k_search_d<false, 32>;
    if (d_pf >= 16) return pr ? k_search_d<true, 16> : k_search_d<false, 16>;
    return pr ? k_search_d<true, 8> : k_search_d<false, 8>;
  };
  if (a.dense == 1) {
    KCUDA(cudaFuncSetAttribute(d_kernel(false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d_smem));
    KCUDA(cudaFuncSetAttribute(d_kernel(true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d_smem));
  }
  // class D (search_dense2.cuh): two launches, queries of up to E_KCAP_S k-mers and the longer ones
  constexpr int E_KCAP_S = 512, E_KCAP_L = 1024;
  a.e_kcap = E_KCAP_S;
  a.e_kcap_l = E_KCAP_L;
  a.e_mapw_small = 512;
  a.e_mapw_large = 1024;
  a.e_mapw_xl = 4096;
  a.lists_sorted = h->idx.lists_sorted ? 1 : 0;
  if (getenv("KAAMER_D_NO_BSEARCH")) a.lists_sorted = 0;  // A/B hook: stream the lists again instead
  if (const char *env = getenv("KAAMER_E_MAPW")) {  // tuning hook: "small,large" words per map per warp (powers of two)
    unsigned ms = 0, ml = 0;
    if (sscanf(env, "%u,%u", &ms, &ml) == 2 && ms >= 64 && ms <= 4096 && ml >= 64 && ml <= 4096 &&
        (ms & (ms - 1)) == 0 && (ml & (ml - 1)) == 0) {
      a.e_mapw_small = ms;
      a.e_mapw_large = ml;
      a.e_mapw_xl = ml > 4096 ? ml : (ml < 1024 ? ml : 4096);
    }
  }
  constexpr int EH_S = 512, EH_L = 512, EH_XL = 2048;
  const size_t e_smem_s = ((sizeof(Dense2Smem<E_KCAP_S, EH_S>) + 15) & ~(size_t)15) + (size_t)E_WARPS * 2 * a.e_mapw_small * 4;
  const size_t e_smem_l = ((sizeof(Dense2Smem<E_KCAP_L, EH_L>) + 15) & ~(size_t)15) + (size_t)E_WARPS * 2 * a.e_mapw_large * 4;
  const size_t e_smem_xl = ((sizeof(Dense2Smem<E_KCAP_L, EH_XL>) + 15) & ~(size_t)15) + (size_t)E_WARPS * 2 * a.e_mapw_xl * 4;
  auto e_small = peer ? k_search_e<true, E_KCAP_S, EH_S, 4> : k_search_e<false, E_KCAP_S, EH_S, 4>;
  auto e_large = peer ? k_search_e<true, E_KCAP_L, EH_L, 5> : k_search_e<false, E_KCAP_L, EH_L, 5>;
  auto e_xl = peer ? k_search_e<true, E_KCAP_L, EH_XL, 6> : k_search_e<false, E_KCAP_L, EH_XL, 6>;
  if (a.dense == 3) {
    KCUDA(cudaFuncSetAttribute(e_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e_smem_s));
    KCUDA(cudaFuncSetAttribute(e_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e_smem_l));
    KCUDA(cudaFuncSetAttribute(e_xl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e_smem_xl));
  }
  // class D, third design (search_dense3.cuh): 4 / 8 / 16 warps per CTA for the three lengths; a warp's map is
  // f_mapw 32-bit words (16 + 16 bits each)
  constexpr int F_NW_S = 4, F_NW_L = 8, F_NW_XL = 16;
  uint32_t f_mapw_s = 1024, f_mapw_l = 1024, f_mapw_xl = 2048;
  if (const char *env = getenv("KAAMER_F_MAPW")) {  // tuning hook: "small,large,xl" words per warp map (powers of two)
    unsigned ms = 0, ml = 0, mx = 0;
    if (sscanf(env, "%u,%u,%u", &ms, &ml, &mx) == 3 && ms >= 64 && ms <= 8192 && ml >= 64 && ml <= 8192 && mx >= 64 &&
        mx <= 2048 && (ms & (ms - 1)) == 0 && (ml & (ml - 1)) == 0 && (mx & (mx - 1)) == 0) {
      f_mapw_s = ms;
      f_mapw_l = ml;
      f_mapw_xl = mx;
    }
  }
  const int f_ph = peer ? F_PH_PEER : F_PH_LOCAL;
  auto f_smem = [&](size_t s3, size_t s7, int nw, uint32_t mapw) {
    return (((f_ph == F_PH_PEER ? s7 : s3) + 15) & ~(size_t)15) + (size_t)nw * mapw * 4;
  };
  const size_t f_smem_s = f_smem(sizeof(Dense3Smem<E_KCAP_S, EH_S, F_NW_S, F_PH_LOCAL>), sizeof(Dense3Smem<E_KCAP_S, EH_S, F_NW_S, F_PH_PEER>), F_NW_S, f_mapw_s);
  const size_t f_smem_l = f_smem(sizeof(Dense3Smem<E_KCAP_L, EH_L, F_NW_L, F_PH_LOCAL>), sizeof(Dense3Smem<E_KCAP_L, EH_L, F_NW_L, F_PH_PEER>), F_NW_L, f_mapw_l);
  const size_t f_smem_xl = f_smem(sizeof(Dense3Smem<E_KCAP_L, EH_XL, F_NW_XL, F_PH_LOCAL>), sizeof(Dense3Smem<E_KCAP_L, EH_XL, F_NW_XL, F_PH_PEER>), F_NW_XL, f_mapw_xl);
  auto f_small = peer ? k_search_f<true, E_KCAP_S, EH_S, 4, F_NW_S, 5> : k_search_f<false, E_KCAP_S, EH_S, 4, F_NW_S, 7>;
  auto f_large = peer ? k_search_f<true, E_KCAP_L, EH_L, 5, F_NW_L, 3> : k_search_f<false, E_KCAP_L, EH_L, 5, F_NW_L, 3>;
  auto f_xl = peer ? k_search_f<true, E_KCAP_L, EH_XL, 6, F_NW_XL, 1> : k_search_f<false, E_KCAP_L, EH_XL, 6, F_NW_XL, 1>;
  auto f_xl2 = peer ? k_search_f<true, E_KCAP_L, EH_XL, 7, F_NW_XL, 1> : k_search_f<false, E_KCAP_L, EH_XL, 7, F_NW_XL, 1>;
  if (a.dense == 2) {
    a.e_mapw_small = f_mapw_s;
    a.e_mapw_large = f_mapw_l;
    a.e_mapw_xl = f_mapw_xl;
    KCUDA(cudaFuncSetAttribute(f_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f_smem_s));
    KCUDA(cudaFuncSetAttribute(f_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f_smem_l));
    KCUDA(cudaFuncSetAttribute(f_xl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f_smem_xl));
    KCUDA(cudaFuncSetAttribute(f_xl2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f_smem_xl));
  }
  // Class G: one CTA per query, histogram in a per-CTA global scratch.  At Swiss-Prot density it holds a
  // handful of very long queries and gets one CTA per SM (it runs underneath W and M and must leave them
  // their shared memory); on a dense database, where the lowered class limits send many queries here,
  // four CTAs of 256 threads per SM (measured with every C3 query forced into class G: 14.2 ms at one
  // 512-thread CTA per SM, 7.5 ms at two).
  const bool dense = a.dense || a.w_maxk < W_MAXK || a.m_maxk < M_MAXK;
  int g_ctas = h->sm_count * (dense ? 4 : 1);
  // a grown histogram (retry after ST_GHASH_OVERFLOW) gets fewer CTAs: at most 8 GB of scratch
  while (g_ctas > 1 && (size_t)2 * g_ctas * 3 * a.ghash_slots * 4 > ((size_t)8 << 30)) g_ctas = (g_ctas + 1) / 2;
  auto launch_g = [&](cudaStream_t s_) {
    if (dense) {
      if (peer) k_search_g<true, 256><<<g_ctas, 256, 0, s_>>>(a);
      else k_search_g<false, 256><<<g_ctas, 256, 0, s_>>>(a);
    } else {
      if (peer) k_search_g<true, G_THREADS><<<g_ctas, G_THREADS, 0, s_>>>(a);
      else k_search_g<false, G_THREADS><<<g_ctas, G_THREADS, 0, s_>>>(a);
    }
  };
  KCHECK(ws.ghash.ensure((size_t)2 * g_ctas * 3 * a.ghash_slots));
  a.ghash = ws.ghash.p;
  KCUDA(cudaMemsetAsync(list_count, 0, 2 * N_LISTS * sizeof(uint32_t), st));
  KCUDA(cudaMemsetAsync(out->counters, 0, CNT_N * sizeof(uint64_t), st));
  if (d_prev_counters)  // chunked host call: the pool cursor continues where the previous chunk stopped
    KCUDA(cudaMemcpyAsync(out->counters + CNT_POOL, d_prev_counters + CNT_POOL, 8, cudaMemcpyDeviceToDevice, st));
  k_classify<<<(nq + 255) / 256, 256, 0, st>>>(a);
  // persistent grids: a multiple of the SM count, warps / CTAs loop over their class list
  const unsigned w_grid = (unsigned)h->sm_count * 5u;
  const unsigned m_grid = (unsigned)h->sm_count * (unsigned)M_CTAS;
  // Class G holds a handful of very long queries, one CTA each: its kernel is a long tail on a few
  // SMs.  It is launched first, on the side stream, so that it runs underneath W and M; a second
  // (normally empty) G launch after M takes the queries whose histograms outgrew class M.
  cudaStream_t side = h->side_stream;
  KCUDA(cudaEventRecord(h->chunk_ev[6], st));
  KCUDA(cudaStreamWaitEvent(side, h->chunk_ev[6], 0));
  profile_begin(h, side, 2);
  a.g_list = 2;
  launch_g(side);
  profile_end(h, side);
  // Posting lists read through NVLink (mode P without replicated postings): fewer CTAs per SM — more outstanding
  // small remote reads make the peer path slower, not faster (measured at 2 GPUs: 20.1 ms per step with four CTAs
  // per SM, 24.2 ms with six; profiles/r2_classD_notes.md)
  const bool lists_remote = peer && h->idx.repl_postings == nullptr;
  if (a.dense == 2) {
    // the long queries of class D run on the side stream underneath the short ones: the (few) longest first
    int per_sm = 1;
    f_xl<<<(unsigned)h->sm_count, F_NW_XL * 32, f_smem_xl, side>>>(a);
    KCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f_large, F_NW_L * 32, f_smem_l));
    if (lists_remote && per_sm > 2) per_sm = 2;
    profile_begin(h, side, 1);
    f_large<<<(unsigned)h->sm_count * (unsigned)(per_sm < 1 ? 1 : per_sm), F_NW_L * 32, f_smem_l, side>>>(a);
    profile_end(h, side);
  } else if (a.dense == 3) {
    int per_sm = 1;
    e_xl<<<(unsigned)h->sm_count, E_THREADS, e_smem_xl, side>>>(a);
    KCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, e_large, E_THREADS, e_smem_l));
    profile_begin(h, side, 1);
    e_large<<<(unsigned)h->sm_count * (unsigned)(per_sm < 1 ? 1 : per_sm), E_THREADS, e_smem_l, side>>>(a);
    profile_end(h, side);
  }
  KCUDA(cudaEventRecord(h->chunk_ev[7], side));
  if (a.dense == 2) {
    int per_sm = 1;
    KCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f_small, F_NW_S * 32, f_smem_s));
    if (lists_remote && per_sm > 4) per_sm = 4;
    profile_begin(h, st, 6);
    f_small<<<(unsigned)h->sm_count * (unsigned)(per_sm < 1 ? 1 : per_sm), F_NW_S * 32, f_smem_s, st>>>(a);
    profile_end(h, st);
    KCUDA(cudaStreamWaitEvent(st, h->chunk_ev[7], 0));  // the hand-offs of both launches are complete
    // queries whose repeated subjects overflowed H (large families): the 2048-slot launch takes them
    f_xl2<<<(unsigned)h->sm_count, F_NW_XL * 32, f_smem_xl, st>>>(a);
  } else if (a.dense == 3) {
    int per_sm = 1;
    KCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, e_small, E_THREADS, e_smem_s));
    profile_begin(h, st, 6);
    e_small<<<(unsigned)h->sm_count * (unsigned)(per_sm < 1 ? 1 : per_sm), E_THREADS, e_smem_s, st>>>(a);
    profile_end(h, st);
    KCUDA(cudaStreamWaitEvent(st, h->chunk_ev[7], 0));  // the hand-offs of both launches are complete
  } else if (a.dense) {
    // first design of class D (A/B measurements): one launch takes every query
    int d_per_sm = 1;
    KCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d_per_sm, d_kernel(peer), D_THREADS, d_smem));
    if (d_per_sm < 1) d_per_sm = 1;
    const unsigned d_grid = (unsigned)h->sm_count * (unsigned)d_per_sm;
    profile_begin(h, st, 6);
    d_kernel(peer)<<<d_grid, D_THREADS, d_smem, st>>>(a);
    profile_end(h, st);
  } else {
    profile_begin(h, st, 0);
    if (peer) k_search_wt<W_H, W_MAXK, W_WARPS, 5, 0, true><<<w_grid, W_WARPS * 32, 0, st>>>(a);
    else k_search_wt<W_H, W_MAXK, W_WARPS, 5, 0, false><<<w_grid, W_WARPS * 32, 0, st>>>(a);
    profile_end(h, st);
    profile_begin(h, st, 1);
    if (peer) k_search_m<true><<<m_grid < nq ? m_grid : nq, M_THREADS, 0, st>>>(a);
    else k_search_m<false><<<m_grid < nq ? m_grid : nq, M_THREADS, 0, st>>>(a);
    profile_end(h, st);
  }
  a.g_list = 3;
  a.ghash = ws.ghash.p + (size_t)g_ctas * 3 * a.ghash_slots;  // own scratch: the first G launch may still run
  launch_g(st);
  KCUDA(cudaStreamWaitEvent(st, h->chunk_ev[7], 0));
  h->prof_all_launches += a.dense == 1 ? 4 : (a.dense == 2 ? 7 : (a.dense == 3 ? 6 : 5));
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

// A class-G query touched more distinct subjects than the global-memory histogram holds: grow the
// histogram to what the batch asked for (counters[CNT_GNEED] = 2 x the largest posting total) and let the
// caller run the batch again.  The reference returns results for such queries; so do we.
int grow_ghash(kaamer_gpu *h, uint64_t need) {
  uint64_t slots = h->ghash_slots;
  while (slots < need || slots <= h->ghash_slots) slots <<= 1;
  if (slots > (1ull << 30)) {
    set_error("a query matched more distinct subjects (%llu) than a 2^30-slot histogram holds",
              (unsigned long long)(need / 2));
    return KAAMER_ERR_LIMIT;
  }
  h->ghash_slots = (uint32_t)slots;
  return KAAMER_OK;
}

// Count pass over queries that are already on the device, with the hit-pool retry loop.
// Leaves n_hits / hit_base / size_in_kmer / pool in the workspace and the counters in
// ws.h_counters (host).  Synchronises the stream.
int search_counted(kaamer_gpu *h, const uint8_t *d_res, const uint64_t *d_off, uint32_t nq, const kaamer_opts *o,
                   int nt_mode, uint8_t *d_any0, cudaStream_t st) {
  SearchWorkspace &ws = h->ws;
  KCHECK(ws.n_hits.ensure(nq));
  KCHECK(ws.hit_base.ensure(nq));
  KCHECK(ws.size_in_kmer.ensure(nq));
  KCHECK(ws.counters.ensure(CNT_N));
  KCHECK(ws.h_counters.ensure(CNT_N + 2));
  uint64_t per_q = o->max_results > 0 ? (uint64_t)(o->max_results < 16 ? o->max_results : 16) : 1;
  uint64_t pool_cap = (uint64_t)nq * per_q + 4096;
  for (int attempt = 0;; ++attempt) {
    KCHECK(ws.pool.ensure((size_t)pool_cap));
    kaamer_dev_result dr{};
    dr.n_hits = ws.n_hits.p;
    dr.hit_base = ws.hit_base.p;
    dr.size_in_kmer = ws.size_in_kmer.p;
    dr.pool = ws.pool.p;
    dr.pool_cap = pool_cap;
    dr.counters = ws.counters.p;
    KCHECK(search_proteins_device(h, d_res, d_off, nq, o, &dr, st, nt_mode, d_any0));
    KCUDA(cudaMemcpyAsync(ws.h_counters.p, ws.counters.p, CNT_N * 8, cudaMemcpyDeviceToHost, st));
    KCUDA(cudaStreamSynchronize(st));
    uint64_t status = ws.h_counters.p[CNT_STATUS];
    if (status & ST_GHASH_OVERFLOW) {
      if (attempt >= 6) {
        set_error("class-G histogram overflow after %d attempts", attempt + 1);
        return KAAMER_ERR_LIMIT;
      }
      KCHECK(grow_ghash(h, ws.h_counters.p[CNT_GNEED]));
      continue;
    }
    if (status & ST_POOL_OVERFLOW) {
      if (attempt >= 6) {
        set_error("hit pool overflow after %d attempts", attempt + 1);
        return KAAMER_ERR_LIMIT;
      }
      pool_cap = ws.h_counters.p[CNT_POOL] + 4096;  // exact demand of the failed pass
      continue;
    }
    break;
  }
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
  if (nq == 0) {
    HCHECK(owner->alloc(&hits->hit_off, 1));
    HCHECK(owner->alloc(&hits->size_in_kmer, 1));
    hits->hit_off[0] = 0;
    HCHECK(owner->alloc(&hits->subject_id, 1));
    HCHECK(owner->alloc(&hits->kmatch, 1));
    *out_hits = hits;
    return KAAMER_OK;
  }
  if (off[0] != 0) {
    set_error("seq_off[0] must be 0");
    return fail(KAAMER_ERR_ARG);
  }
  HCHECK(ws.seq_off.ensure((size_t)nq + 1));
  // Residues: a pinned (page-locked, device-mapped) caller buffer — kaamer_gpu_pinned_alloc, or
  // any cudaHostAlloc/cudaHostRegister memory — is read by the kernels IN PLACE over PCIe: every
  // residue is needed exactly once, so the transfer overlaps the table probes instead of
  // preceding them.  Pageable memory is staged with one H2D copy.
  const uint8_t *d_res = nullptr;
  {
    cudaPointerAttributes at;
    if (n_res && cudaPointerGetAttributes(&at, res) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
      d_res = (const uint8_t *)at.devicePointer;
    else
      cudaGetLastError();
  }
  const int n_chunks = 1;
  uint32_t cq[9];
  cq[0] = 0;
  cq[n_chunks] = nq;
  profile_begin(h, h->copy_stream, 4);
  HCUDA(cudaMemcpyAsync(ws.seq_off.p, off, ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, h->copy_stream));
  if (!d_res) {
    HCHECK(ws.residues.ensure((size_t)n_res + 16));
    if (n_res) HCUDA(cudaMemcpyAsync(ws.residues.p, res, (size_t)n_res, cudaMemcpyHostToDevice, h->copy_stream));
    d_res = ws.residues.p;
  }
  HCUDA(cudaEventRecord(h->chunk_ev[0], h->copy_stream));
  profile_end(h, h->copy_stream);
  if (o->want_positions) {
    // PositionHits wanted (search.go:416,442-452): rows, hits and positions assembled by finish.cu
    h->arena.reset();
    HCUDA(cudaStreamWaitEvent(st, h->chunk_ev[0], 0));
    HCHECK(search_counted(h, d_res, ws.seq_off.p, nq, o, 0, nullptr, st));
    hits->n_lookups = ws.h_counters.p[CNT_LOOKUPS];
    hits->n_increments = ws.h_counters.p[CNT_INCR];
    HCHECK(finish_rows(h, d_res, ws.seq_off.p, nq, o, 0, nullptr, nullptr, hits, owner, st));
    *out_hits = hits;
    return KAAMER_OK;
  }
  // count pass + CSR offsets in one stream pass, one host sync for (status, counters, total)
  HCHECK(ws.n_hits.ensure((size_t)nq + 1));
  HCHECK(ws.hit_base.ensure(nq));
  HCHECK(ws.size_in_kmer.ensure(nq));
  size_t scan_tmp = 0;
  {
    cub::TransformInputIterator<uint64_t, WidenU32, const uint32_t *> in(ws.n_hits.p, WidenU32());
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, in, ws.hit_off.p, (int64_t)nq + 1, st);
    HCHECK(ws.f_tmp.ensure(scan_tmp + 16));
  }
  HCHECK(ws.counters.ensure(CNT_N * 8));
  HCHECK(ws.h_counters.ensure(CNT_N * 8 + 2));
  HCHECK(ws.hit_off.ensure((size_t)nq + 1));
  HCHECK(owner->alloc(&hits->hit_off, (size_t)nq + 1));
  HCHECK(owner->alloc(&hits->size_in_kmer, (size_t)nq));
  uint64_t per_q = o->max_results > 0 ? (uint64_t)(o->max_results < 16 ? o->max_results : 16) : 1;
  uint64_t pool_cap = (uint64_t)nq * per_q + 4096;
  if (ws.pool.n > pool_cap) pool_cap = ws.pool.n;  // a previous batch already grew the pool
  for (int attempt = 0;; ++attempt) {
    HCHECK(ws.pool.ensure((size_t)pool_cap));
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t qb = cq[c], nqc = cq[c + 1] - cq[c];
      if (attempt == 0) HCUDA(cudaStreamWaitEvent(st, h->chunk_ev[c], 0));
      kaamer_dev_result dr{};
      dr.n_hits = ws.n_hits.p + qb;
      dr.hit_base = ws.hit_base.p + qb;
      dr.size_in_kmer = ws.size_in_kmer.p + qb;
      dr.pool = ws.pool.p;
      dr.pool_cap = pool_cap;
      dr.counters = ws.counters.p + (size_t)c * CNT_N;
      if (nqc == 0) {
        HCUDA(cudaMemsetAsync(dr.counters, 0, CNT_N * 8, st));
        if (c) HCUDA(cudaMemcpyAsync(dr.counters + CNT_POOL, dr.counters - CNT_N + CNT_POOL, 8, cudaMemcpyDeviceToDevice, st));
        continue;
      }
      HCHECK(search_proteins_device(h, d_res, ws.seq_off.p + qb, nqc, o, &dr, st, 0, nullptr,
                                    c ? ws.counters.p + (size_t)(c - 1) * CNT_N : nullptr));
    }
    profile_begin(h, st, 5);
    {
      // hit_off[0..nq] = exclusive sum of n_hits[0..nq) (+ one zero): total lands in hit_off[nq]
      HCUDA(cudaMemsetAsync(ws.n_hits.p + nq, 0, 4, st));
      cub::TransformInputIterator<uint64_t, WidenU32, const uint32_t *> in(ws.n_hits.p, WidenU32());
      size_t tb = scan_tmp;
      HCUDA(cub::DeviceScan::ExclusiveSum(ws.f_tmp.p, tb, in, ws.hit_off.p, (int64_t)nq + 1, st));
      h->prof_all_launches += 2;
    }
    HCUDA(cudaMemcpyAsync(ws.h_counters.p, ws.counters.p, (size_t)n_chunks * CNT_N * 8, cudaMemcpyDeviceToHost, st));
    HCUDA(cudaMemcpyAsync(ws.h_counters.p + CNT_N * 8, ws.hit_off.p + nq, 8, cudaMemcpyDeviceToHost, st));
    HCUDA(cudaMemcpyAsync(hits->hit_off, ws.hit_off.p, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, st));
    HCUDA(cudaMemcpyAsync(hits->size_in_kmer, ws.size_in_kmer.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    profile_end(h, st);
    HCUDA(cudaStreamSynchronize(st));
    uint64_t status = 0, lookups = 0, incr = 0, gneed = 0;
    for (int c = 0; c < n_chunks; ++c) {
      status |= ws.h_counters.p[(size_t)c * CNT_N + CNT_STATUS];
      gneed = std::max<uint64_t>(gneed, ws.h_counters.p[(size_t)c * CNT_N + CNT_GNEED]);
      lookups += ws.h_counters.p[(size_t)c * CNT_N + CNT_LOOKUPS];
      incr += ws.h_counters.p[(size_t)c * CNT_N + CNT_INCR];
    }
    if (status & ST_GHASH_OVERFLOW) {
      if (attempt >= 6) {
        set_error("class-G histogram overflow after %d attempts", attempt + 1);
        return fail(KAAMER_ERR_LIMIT);
      }
      HCHECK(grow_ghash(h, gneed));
      continue;
    }
    if (status & ST_POOL_OVERFLOW) {
      if (attempt >= 6) {
        set_error("hit pool overflow after %d attempts", attempt + 1);
        return fail(KAAMER_ERR_LIMIT);
      }
      pool_cap = ws.h_counters.p[(size_t)(n_chunks - 1) * CNT_N + CNT_POOL] + 4096;  // exact demand of the failed pass
      continue;
    }
    hits->n_lookups = lookups;
    hits->n_increments = incr;
    break;
  }
  const uint64_t n_hits = ws.h_counters.p[CNT_N * 8];
  hits->n_hits = n_hits;
  HCHECK(owner->alloc(&hits->subject_id, (size_t)n_hits));
  HCHECK(owner->alloc(&hits->kmatch, (size_t)n_hits));
  if (n_hits) {
    HCHECK(ws.out_hits.ensure((size_t)n_hits + 1));  // subject u32[n_hits] | kmatch u32[n_hits]
    uint32_t *d_subj = reinterpret_cast<uint32_t *>(ws.out_hits.p), *d_km = d_subj + n_hits;
    unsigned grid = (unsigned)(((uint64_t)nq * 32 + 255) / 256);
    profile_begin(h, st, 5);
    k_gather_hits<<<grid, 256, 0, st>>>(ws.n_hits.p, ws.hit_base.p, ws.hit_off.p, ws.pool.p, nq, d_subj, d_km);
    h->prof_all_launches += 1;
    HCUDA(cudaMemcpyAsync(hits->subject_id, d_subj, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, st));
    HCUDA(cudaMemcpyAsync(hits->kmatch, d_km, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, st));
    profile_end(h, st);
    HCUDA(cudaStreamSynchronize(st));
  }
  *out_hits = hits;
#undef HCHECK
#undef HCUDA
  return KAAMER_OK;
}

// ---- pipelined host-buffer path: submit / wait ------------------------------------------------
// A blocking kaamer_gpu_search_proteins call costs, besides its kernels, two stream synchronisations, the
// H2D of the offsets and the D2H of the hits — with 8 ranks on one host that tax grew to +60 % of a call
// (SCALE_r01: e2e efficiency 0.64 at 8 GPUs).  The submit / wait pair keeps the device busy across calls:
// submit enqueues EVERYTHING of a batch (offsets H2D, search kernels, CSR scan, and a gather kernel that
// writes the hits straight into the caller-visible pinned result arrays over PCIe), wait is ONE event
// synchronisation.  Two slots: batch k+1 is submitted before batch k is waited for, its kernels queue up
// behind batch k's on the same stream (the scratch of the kernels is reused in stream order; only what the
// host reads lives in the slot).  Result capacity is bounded before the launch: a query keeps at most
// MaxResults hits (FilterResults, search.go:208-210).
struct PendingSlot {
  bool busy = false;
  uint32_t nq = 0;
  kaamer_opts opts{};
  const uint8_t *res = nullptr;
  const uint64_t *off = nullptr;
  DevBuf<uint64_t> seq_off, hit_off, pool, counters;
  DevBuf<uint32_t> n_hits, hit_base;
  DevBuf<int32_t> size_in_kmer;
  DevBuf<uint8_t> residues;  // pageable caller memory is staged here
  DevBuf<uint32_t> d_subj, d_km;  // compact hits on the device
  kaamer_hits *hits = nullptr;
  HitsOwner *owner = nullptr;
  uint64_t *h_tail = nullptr;  // pinned: counters[CNT_N] + total hits
  uint64_t cap_hits = 0, pool_cap = 0;
  cudaEvent_t done = nullptr, copied = nullptr;
  void drop_result() {
    delete owner;
    delete hits;
    owner = nullptr;
    hits = nullptr;
  }
};

static PendingSlot *slots_of(kaamer_gpu *h) {
  if (!h->pending) {
    auto *sl = new PendingSlot[2];
    for (int i = 0; i < 2; ++i) {
      cudaEventCreateWithFlags(&sl[i].done, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&sl[i].copied, cudaEventDisableTiming);
    }
    h->pending = sl;
  }
  return static_cast<PendingSlot *>(h->pending);
}

void release_pending(kaamer_gpu *h) {
  if (!h->pending) return;
  auto *sl = static_cast<PendingSlot *>(h->pending);
  for (int i = 0; i < 2; ++i) {
    if (sl[i].busy) cudaEventSynchronize(sl[i].done);
    sl[i].drop_result();
    sl[i].seq_off.release(); sl[i].hit_off.release(); sl[i].pool.release(); sl[i].counters.release();
    sl[i].n_hits.release(); sl[i].hit_base.release(); sl[i].size_in_kmer.release(); sl[i].residues.release(); sl[i].d_subj.release(); sl[i].d_km.release();
    if (sl[i].done) cudaEventDestroy(sl[i].done);
    if (sl[i].copied) cudaEventDestroy(sl[i].copied);
  }
  delete[] sl;
  h->pending = nullptr;
}

constexpr uint64_t SUBMIT_MAX_HITS = 32ull << 20;  // larger bounds (MaxResults in the thousands) take the blocking path

// returns the slot (0 / 1), or -1: this batch cannot be pipelined (positions wanted, huge MaxResults)
int search_proteins_submit(kaamer_gpu *h, const uint8_t *res, const uint64_t *off, uint32_t nq, const kaamer_opts *o,
                           int *slot_out) {
  *slot_out = -1;
  const uint64_t per_q = o->max_results > 0 ? (uint64_t)o->max_results : 0;
  if (o->want_positions || nq == 0 || (uint64_t)nq * per_q > SUBMIT_MAX_HITS) return KAAMER_OK;
  if (off[0] != 0) {
    set_error("seq_off[0] must be 0");
    return KAAMER_ERR_ARG;
  }
  PendingSlot *sl = slots_of(h);
  int s = -1;
  for (int i = 0; i < 2; ++i)
    if (!sl[i].busy) {
      s = i;
      break;
    }
  if (s < 0) {
    set_error("two batches are already in flight on this handle: wait for one of them first");
    return KAAMER_ERR_ARG;
  }
  PendingSlot &p = sl[s];
  cudaStream_t st = h->stream;
  const uint64_t n_res = off[nq];
  p.nq = nq;
  p.opts = *o;
  p.res = res;
  p.off = off;
  p.cap_hits = (uint64_t)nq * per_q;
  const uint64_t pq = per_q < 16 ? (per_q ? per_q : 1) : 16;
  p.pool_cap = (uint64_t)nq * pq + 4096;
  if (p.pool.n > p.pool_cap) p.pool_cap = p.pool.n;
  KCHECK(p.seq_off.ensure((size_t)nq + 1));
  KCHECK(p.hit_off.ensure((size_t)nq + 1));
  KCHECK(p.n_hits.ensure((size_t)nq + 1));
  KCHECK(p.hit_base.ensure(nq));
  KCHECK(p.size_in_kmer.ensure(nq));
  KCHECK(p.counters.ensure(CNT_N));
  KCHECK(p.pool.ensure((size_t)p.pool_cap));
  p.drop_result();
  p.hits = new kaamer_hits();
  memset(p.hits, 0, sizeof *p.hits);
  p.owner = new HitsOwner();
  p.hits->_owner = p.owner;
  p.hits->n_rows = nq;
  int rc = p.owner->alloc(&p.hits->hit_off, (size_t)nq + 1);
  if (rc == KAAMER_OK) rc = p.owner->alloc(&p.hits->size_in_kmer, (size_t)nq);
  const size_t cap4 = ((size_t)p.cap_hits + 3) & ~(size_t)3;  // whole 16-byte groups
  if (rc == KAAMER_OK) rc = p.owner->alloc(&p.hits->subject_id, cap4 ? cap4 : 4);
  if (rc == KAAMER_OK) rc = p.owner->alloc(&p.hits->kmatch, cap4 ? cap4 : 4);
  if (rc == KAAMER_OK) rc = p.d_subj.ensure(cap4 + 4);
  if (rc == KAAMER_OK) rc = p.d_km.ensure(cap4 + 4);
  if (rc == KAAMER_OK) rc = p.owner->alloc(&p.h_tail, (size_t)CNT_N + 2);
  if (rc != KAAMER_OK) {
    p.drop_result();
    return rc;
  }
  auto fail = [&](const char *what, cudaError_t e) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    p.drop_result();
    return KAAMER_ERR_CUDA;
  };
  // Residues.  A lone call reads a page-locked caller buffer IN PLACE (the PCIe transfer overlaps the call's own
  // table probes; the kernels then run at ~26 GB/s of zero-copy reads).  When another batch is in flight the
  // copy engine stages them instead: the DMA (~2x the zero-copy rate) runs underneath the other batch's
  // kernels and this batch's kernels run at device speed.  Pageable memory is always staged.
  const uint8_t *d_res = nullptr;
  const bool other_in_flight = sl[1 - s].busy;
  if (!other_in_flight) {
    cudaPointerAttributes at;
    if (n_res && cudaPointerGetAttributes(&at, res) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
      d_res = (const uint8_t *)at.devicePointer;
    else
      cudaGetLastError();
  }
  cudaError_t e = cudaMemcpyAsync(p.seq_off.p, off, ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, h->copy_stream);
  if (e != cudaSuccess) return fail("offsets H2D", e);
  if (!d_res) {
    rc = p.residues.ensure((size_t)n_res + 16);
    if (rc != KAAMER_OK) {
      p.drop_result();
      return rc;
    }
    if (n_res) {
      e = cudaMemcpyAsync(p.residues.p, res, (size_t)n_res, cudaMemcpyHostToDevice, h->copy_stream);
      if (e != cudaSuccess) return fail("residues H2D", e);
    }
    d_res = p.residues.p;
  }
  e = cudaEventRecord(p.copied, h->copy_stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(st, p.copied, 0);
  if (e != cudaSuccess) return fail("copy event", e);
  kaamer_dev_result dr{};
  dr.n_hits = p.n_hits.p;
  dr.hit_base = p.hit_base.p;
  dr.size_in_kmer = p.size_in_kmer.p;
  dr.pool = p.pool.p;
  dr.pool_cap = p.pool_cap;
  dr.counters = p.counters.p;
  rc = search_proteins_device(h, d_res, p.seq_off.p, nq, o, &dr, st, 0, nullptr, nullptr);
  if (rc != KAAMER_OK) {
    p.drop_result();
    return rc;
  }
  // CSR offsets, then the hits straight into the pinned result arrays (PCIe writes from the gather kernel)
  size_t scan_tmp = 0;
  cub::TransformInputIterator<uint64_t, WidenU32, const uint32_t *> in(p.n_hits.p, WidenU32());
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, in, p.hit_off.p, (int64_t)nq + 1, st);
  rc = h->ws.f_tmp.ensure(scan_tmp + 16);
  if (rc != KAAMER_OK) {
    p.drop_result();
    return rc;
  }
  e = cudaMemsetAsync(p.n_hits.p + nq, 0, 4, st);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(h->ws.f_tmp.p, scan_tmp, in, p.hit_off.p, (int64_t)nq + 1, st);
  if (e != cudaSuccess) return fail("scan", e);
  uint32_t *out_subj = nullptr, *out_km = nullptr;
  e = cudaHostGetDevicePointer((void **)&out_subj, p.hits->subject_id, 0);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&out_km, p.hits->kmatch, 0);
  if (e != cudaSuccess) return fail("cudaHostGetDevicePointer(result arrays)", e);
  if (p.cap_hits) {
    const unsigned grid = (unsigned)(((uint64_t)nq * 32 + 255) / 256);
    k_gather_hits<<<grid, 256, 0, st>>>(p.n_hits.p, p.hit_base.p, p.hit_off.p, p.pool.p, nq, p.d_subj.p, p.d_km.p);
    k_hits_to_host<<<h->sm_count * 2, 256, 0, st>>>(p.d_subj.p, p.d_km.p, p.hit_off.p + nq, p.cap_hits, out_subj, out_km);
  }
  h->prof_all_launches += 4;
  e = cudaMemcpyAsync(p.h_tail, p.counters.p, CNT_N * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p.h_tail + CNT_N, p.hit_off.p + nq, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p.hits->hit_off, p.hit_off.p, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p.hits->size_in_kmer, p.size_in_kmer.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaEventRecord(p.done, st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail("result D2H", e);
  p.busy = true;
  *slot_out = s;
  return KAAMER_OK;
}

int search_proteins_wait(kaamer_gpu *h, int slot, kaamer_hits **out_hits) {
  PendingSlot *sl = slots_of(h);
  if (slot < 0 || slot > 1 || !sl[slot].busy) {
    set_error("no batch in flight in slot %d", slot);
    return KAAMER_ERR_ARG;
  }
  PendingSlot &p = sl[slot];
  cudaError_t e = cudaEventSynchronize(p.done);
  p.busy = false;
  if (e != cudaSuccess) {
    set_error("search batch: %s", cudaGetErrorString(e));
    p.drop_result();
    return KAAMER_ERR_CUDA;
  }
  const uint64_t status = p.h_tail[CNT_STATUS];
  if (status & (ST_POOL_OVERFLOW | ST_GHASH_OVERFLOW)) {
    // rare: more candidates than the hit pool / the class-G histogram was sized for.  The blocking path has the
    // retry loops; the caller's buffers are still valid (they are until wait returns)
    if (status & ST_GHASH_OVERFLOW) KCHECK(grow_ghash(h, p.h_tail[CNT_GNEED]));
    p.drop_result();
    return search_proteins_host(h, p.res, p.off, p.nq, &p.opts, out_hits);
  }
  p.hits->n_hits = p.h_tail[CNT_N];
  p.hits->n_lookups = p.h_tail[CNT_LOOKUPS];
  p.hits->n_increments = p.h_tail[CNT_INCR];
  *out_hits = p.hits;
  p.hits = nullptr;  // handed over
  p.owner = nullptr;
  return KAAMER_OK;
}

// nucleotide contigs on the host: H2D -> ORFs -> count pass (nt mode) -> finish
int search_nucleotide_host(kaamer_gpu *h, const uint8_t *nt, const uint64_t *coff, uint32_t nc,
                           const kaamer_opts *o, kaamer_hits **out_hits) {
  SearchWorkspace &ws = h->ws;
  cudaStream_t st = h->stream;
  const uint64_t zero = 0;
  if (nc == 0) coff = &zero;
  if (coff[0] != 0) {
    set_error("contig_off[0] must be 0");
    return KAAMER_ERR_ARG;
  }
  const uint64_t total = coff[nc];
  h->arena.reset();
  HostPhase ph_orf(h, 0);
  KCHECK(ws.residues.ensure((size_t)total + 16));
  if (total) KCUDA(cudaMemcpyAsync(ws.residues.p, nt, (size_t)total, cudaMemcpyHostToDevice, st));
  OrfSet os;
  KCHECK(orfs_device(h, ws.residues.p, coff, nc, &os, st));
  ph_orf.stop();
  auto *hits = new kaamer_hits();
  memset(hits, 0, sizeof *hits);
  auto *owner = new HitsOwner();
  hits->_owner = owner;
  int rc = KAAMER_OK;
  if (os.n > 0xFFFFFFFFull) {
    set_error("too many ORFs in one batch");
    rc = KAAMER_ERR_LIMIT;
  }
  const uint32_t nq = (uint32_t)os.n;
  if (rc == KAAMER_OK) rc = ws.any0.ensure((size_t)nq + 1);
  HostPhase ph_count(h, 1);
  if (rc == KAAMER_OK && nq) rc = search_counted(h, os.seq, os.seq_off, nq, o, 1, ws.any0.p, st);
  ph_count.stop();
  HostPhase ph_finish(h, 2);
  if (rc == KAAMER_OK && nq) {
    hits->n_lookups = ws.h_counters.p[CNT_LOOKUPS];
    hits->n_increments = ws.h_counters.p[CNT_INCR];
  }
  if (rc == KAAMER_OK) rc = finish_rows(h, os.seq, os.seq_off, nq, o, 1, ws.any0.p, &os, hits, owner, st);
  ph_finish.stop();
  orfset_release(&os);
  if (rc != KAAMER_OK) {
    delete owner;
    delete hits;
    return rc;
  }
  *out_hits = hits;
  return KAAMER_OK;
}

}  // namespace kaamer
