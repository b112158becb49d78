// search_dense.cuh — class D: the search kernel of the saturated-key-space regime (C4: ~41 postings per
// query k-mer, ~14 000 (query, subject) increments per 350-aa query, almost all of them subjects seen once).
//
// Replaces, like the other classes, KmerSearch + sortMapByValue + FilterResults
// (pkg/search/search.go:414-440, 132-152, 189-220) for one query per CTA.  What changes is the counting:
// a shared-memory atomic costs ~2 cycles per lane on the SM's atomic unit (0.5 increments/clk/SM =
// 145 G increments/s per GPU), an order of magnitude below what streaming 164-byte posting lists from HBM
// delivers.  So the ~99 % of the postings that belong to subjects seen ONCE are filtered out with plain
// byte loads/stores, and only the repeated subjects ever reach an atomic:
//
//   pass 1  every posting id tests-and-sets two byte maps in shared memory, M1[h1(id)] then M2[h2(id)]
//           (idempotent stores of the constant 1: no read-modify-write, hence no race to lose); an id
//           that finds both bytes already set is PUSHED into a small exact hash H (CAS + atomic count);
//   sweep   subjects pushed >= kmin - 2*W times become FINAL candidates (bloom bits in registers);
//   pass 2  the lists are streamed again (L2 hits) and every occurrence of a final candidate is counted
//           exactly; then the same threshold / top-N epilogue as the other classes.
//
// No false negatives, deterministically: W warps stream disjoint lists of the query, each warp handles
// ONE posting list per step (ids of a list are unique) and synchronises (__syncwarp) between steps, so a
// warp's later steps see its own earlier stores.  In one warp a subject is left unpushed at most twice
// (the step that sets M1, the step that sets M2), so a subject with Kmatch >= kmin is pushed at least
// kmin - 2W times; W = min(4, (kmin-1)/2) warps stream in pass 1 (kmin < 3: class G).  False positives
// (hash collisions) only cost work: pass 2 counts exactly.
#pragma once
#include "search_common.cuh"

namespace kaamer {

constexpr int D_WARPS = 4, D_THREADS = D_WARPS * 32;
constexpr int D_KCH = 512;     // query k-mers per chunk (table entries staged in shared memory)
constexpr int D_H = 1024;      // slots of the exact hash of pushed subjects
constexpr int D_MAXK = 60000;  // 16-bit counts
constexpr uint32_t D_MAPB_DEFAULT = 20 * 1024;

struct __align__(16) DenseSmem {
  uint64_t ent[D_KCH];
  uint32_t hkeys[D_H];
  uint32_t hcnt2[D_H / 2];
  uint32_t fin[D_H / 32];
  uint16_t cand[D_H];
  uint16_t pp[D_KCH + 8];
  uint8_t raw[D_KCH + 64];
  uint8_t lut[256];
  SelectScratch ss;
  uint32_t bloom[4];
  uint32_t nfinal, it;
  unsigned long long total;
};

__device__ __forceinline__ uint32_t dense_h1(uint32_t id, uint32_t S) { return __umulhi(id * 0x9E3779B1u, S); }
__device__ __forceinline__ uint32_t dense_h2(uint32_t id, uint32_t S) {
  uint32_t x = id * 0x85EBCA77u;
  x ^= x >> 13;
  x *= 0xC2B2AE3Du;
  return __umulhi(x, S);
}

// One chunk of the query: residues -> packed codes -> table entries in s.ent[0, kn).  Returns kn.
// Every thread returns in `tot` the posting total of the entries it probed.
template <bool PEER>
__device__ __forceinline__ int dense_load_chunk(const SearchArgs &a, const PeerView *pv, DenseSmem &s, uint64_t b,
                                                int len, int K, int c, const uint8_t *res_end,
                                                unsigned long long &tot) {
  const int tid = threadIdx.x;
  const int kbeg = c * D_KCH;
  const int kn = K - kbeg < D_KCH ? K - kbeg : D_KCH;
  const int nres = len - kbeg < kn + 7 ? len - kbeg : kn + 7;
  __syncthreads();  // the previous users of raw / pp / ent are done
  const int head = stage_bytes<D_THREADS>(s.raw, a.res + b + kbeg, nres, res_end, tid);
  __syncthreads();
  const uint8_t *r = s.raw + head;
  const int ncodes = kn + KAAMER_KMER_SIZE - 1;
  for (int i = tid; i < ncodes; i += D_THREADS) {
    const uint32_t c0 = s.lut[r[i]];
    const uint32_t c1 = (i + 1 < nres) ? (uint32_t)s.lut[r[i + 1]] : CODE_UNKNOWN;
    s.pp[i] = (uint16_t)packed_code(c0, c1);
  }
  __syncthreads();
  constexpr int U = D_KCH / D_THREADS;
  uint32_t d[U];
  bool ok[U];
  uint64_t e[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int pos = u * D_THREADS + tid;
    ok[u] = pos < kn;
    d[u] = ok[u] ? dense_from_packed(s.pp[pos], s.pp[pos + 2], s.pp[pos + 4], s.pp[pos + 6]) : 0u;
  }
  probe_entries<PEER, U>(a, pv, d, ok, e);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int pos = u * D_THREADS + tid;
    if (ok[u]) {
      s.ent[pos] = e[u];
      tot += e[u] >> ENTRY_VALUE_BITS;
    }
  }
  __syncthreads();
  return kn;
}

// Warp `wi` of `nw` streams the posting lists of entries wi, wi + nw, ... of the staged chunk, one list per
// step, D_PF steps in flight.  PASS 1: byte-map test-and-set + pushes.  PASS 2: exact counts of the final
// candidates (bloom bits b0..b3).
template <int PASS, bool PEER, int D_PF>
__device__ __forceinline__ void dense_stream(const SearchArgs &a, const PeerView *pv, DenseSmem &s, uint8_t *m1,
                                             uint8_t *m2, uint32_t S, int kn, int wi, int nw,
                                             const SmemHashT<false> &hv, const CandList &cl, uint32_t b0, uint32_t b1,
                                             uint32_t b2, uint32_t b3) {
  const unsigned lane = threadIdx.x & 31;
  int k = wi - nw;
  uint32_t off = 0, cnt = 0, single = 0;
  const uint32_t *ptr = nullptr;
  // warp-uniform iterator over (list, 32-id window) steps
  auto next_step = [&](uint32_t &id, uint32_t &nv) -> bool {
    while (off >= cnt) {
      k += nw;
      if (k >= kn) return false;
      const uint64_t e = s.ent[k];
      cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);
      off = 0;
      const uint64_t val = e & ENTRY_VALUE_MASK;
      single = (uint32_t)val;
      if (cnt >= 2) ptr = post_ptr<PEER>(a, pv, val);
    }
    nv = cnt - off < 32u ? cnt - off : 32u;
    if (cnt == 1) id = single;
    else id = lane < nv ? __ldg(ptr + off + lane) : 0u;
    off += 32;
    return true;
  };
  auto process = [&](uint32_t id, bool valid) {
    if constexpr (PASS == 1) {
      if (valid) {
        const uint32_t h1 = dense_h1(id, S);
        if (m1[h1] == 0) {
          m1[h1] = 1;
        } else {
          const uint32_t h2 = dense_h2(id, S);
          if (m2[h2] == 0) m2[h2] = 1;
          else count_subject(hv, id, 0xFFFFFFFFu, cl);
        }
      }
      __syncwarp();  // this step's stores are visible to the warp's next step
    } else {
      if (valid) {
        const uint32_t hb = (id * 0x9E3779B1u) >> 25;
        const uint32_t w = hb < 64u ? (hb < 32u ? b0 : b1) : (hb < 96u ? b2 : b3);
        if ((w >> (hb & 31u)) & 1u) {
          uint32_t slot = hv.home(id);
#pragma unroll 1
          for (int probe = 0; probe < SmemHashT<false>::kMaxProbe; ++probe) {
            const uint32_t key = s.hkeys[slot];
            if (key == id) {
              if ((s.fin[slot >> 5] >> (slot & 31u)) & 1u) hv.add(slot, 1u);
              break;
            }
            if (key == EMPTY) break;
            slot = (slot + 1) & hv.mask;
          }
        }
      }
    }
  };
  uint32_t ids[D_PF], nvs[D_PF];
  bool ok[D_PF];
#pragma unroll
  for (int u = 0; u < D_PF; ++u) ok[u] = next_step(ids[u], nvs[u]);
  while (ok[0]) {
#pragma unroll
    for (int u = 0; u < D_PF; ++u) {
      if (ok[u]) {
        process(ids[u], lane < nvs[u]);
        ok[u] = next_step(ids[u], nvs[u]);
      }
    }
  }
}

template <bool PEER, int D_PF>
__global__ void __launch_bounds__(D_THREADS) k_search_d(SearchArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  DenseSmem &s = *reinterpret_cast<DenseSmem *>(dsm);
  const uint32_t mapb = a.d_mapb;
  uint8_t *m1 = dsm + ((sizeof(DenseSmem) + 15) & ~(size_t)15);
  uint8_t *m2 = m1 + mapb;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;
  const int w = tid >> 5;
  const PeerView *pv = nullptr;
  if constexpr (PEER) {
    __shared__ PeerView s_peer;
    load_peer_view(&s_peer, a.peer, tid, D_THREADS);
    pv = &s_peer;
  }
  for (int i = tid; i < 256; i += D_THREADS) s.lut[i] = (uint8_t)aa_code(i);
  __syncthreads();
  const SmemHashT<false> hv{s.hkeys, s.hcnt2, (uint32_t)D_H - 1u, 32 - ilog2_c(D_H)};
  const CandList cl{&s.ss.ncand, &s.ss.flags, s.cand, nullptr, (uint32_t)D_H};
  const uint32_t count = a.list_count[4];
  const uint8_t *res_end = a.res + a.off[a.nq];
  unsigned long long my_incr = 0, my_lookups = 0;
  for (;;) {
    __syncthreads();
    if (tid == 0) s.it = atomicAdd(&a.list_count[N_LISTS + 4], 1u);
    __syncthreads();
    const uint32_t it = s.it;
    if (it >= count) break;
    const uint32_t q = a.lists[(size_t)4 * a.nq + it];
    const uint64_t b = a.off[q];
    const int len = (int)(a.off[q + 1] - b);
    const int K = a.size_in_kmer[q];
    const uint32_t kmin = a.kmin[q];  // >= 3 (k_classify)
    const int w_act = (int)((kmin - 1u) / 2u) < D_WARPS ? (int)((kmin - 1u) / 2u) : D_WARPS;
    const uint32_t thr = kmin - 2u * (uint32_t)w_act;  // >= 1
    const int nchunks = (K + D_KCH - 1) / D_KCH;
    {
      uint4 *hk = reinterpret_cast<uint4 *>(s.hkeys);
      uint4 *hc = reinterpret_cast<uint4 *>(s.hcnt2);
      const uint4 E = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY), Z = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < D_H / 4; i += D_THREADS) hk[i] = E;
      for (int i = tid; i < D_H / 8; i += D_THREADS) hc[i] = Z;
      if (tid < 4) s.bloom[tid] = 0;
      if (tid == 0) {
        s.ss.ncand = 0;
        s.ss.flags = 0;
        s.total = 0;
        s.nfinal = 0;
      }
    }
    uint32_t S = mapb;
    unsigned long long q_incr = 0;
    // ---- pass 1 ----
    for (int c = 0; c < nchunks; ++c) {
      unsigned long long tot = 0;
      const int kn = dense_load_chunk<PEER>(a, pv, s, b, len, K, c, res_end, tot);
      q_incr += tot;
      if (c == 0) {
        if (nchunks == 1) {
          // the byte maps only need to be as large as the query's posting total asks for
          for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
          if (lane == 0 && tot) atomicAdd(&s.total, tot);
          __syncthreads();
          const unsigned long long want = 4ull * s.total;
          if (want < (unsigned long long)mapb) S = want < 1024ull ? 1024u : (((uint32_t)want + 15u) & ~15u);
        }
        uint4 *v1 = reinterpret_cast<uint4 *>(m1), *v2 = reinterpret_cast<uint4 *>(m2);
        const uint4 Z = make_uint4(0, 0, 0, 0);
        for (uint32_t i = tid; i < S / 16; i += D_THREADS) {
          v1[i] = Z;
          v2[i] = Z;
        }
        __syncthreads();
      }
      if (w < w_act) dense_stream<1, PEER, D_PF>(a, pv, s, m1, m2, S, kn, w, w_act, hv, cl, 0, 0, 0, 0);
    }
    __syncthreads();
    if (s.ss.flags) {
      // more repeated subjects than H holds: class G counts this query exactly in global memory
      if (tid == 0) {
        const uint32_t slot = atomicAdd(&a.list_count[3], 1u);
        a.lists[(size_t)3 * a.nq + slot] = q;
      }
      continue;
    }
    // ---- sweep: final candidates ----
    for (int base = 0; base < D_H; base += D_THREADS) {
      const uint32_t slot = base + tid;
      const uint32_t key = s.hkeys[slot];
      const bool isfin = key != EMPTY && hv.count_at(slot) >= thr;
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, isfin);
      if (lane == 0) s.fin[slot >> 5] = bal;
      if (isfin) {
        const uint32_t hb = (key * 0x9E3779B1u) >> 25;
        atomicOr(&s.bloom[hb >> 5], 1u << (hb & 31u));
      }
      if (lane == 0 && bal) atomicAdd(&s.nfinal, (uint32_t)__popc(bal));
    }
    __syncthreads();
    {
      uint4 *hc = reinterpret_cast<uint4 *>(s.hcnt2);
      const uint4 Z = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < D_H / 8; i += D_THREADS) hc[i] = Z;
    }
    __syncthreads();
    my_incr += q_incr;
    if (tid == 0) my_lookups += (unsigned long long)K;
    if (s.nfinal == 0) continue;  // nothing can reach kmin: no hits (n_hits[q] was zeroed by k_classify)
    // ---- pass 2: exact counts of the final candidates ----
    const uint32_t b0 = s.bloom[0], b1 = s.bloom[1], b2 = s.bloom[2], b3 = s.bloom[3];
    for (int c = 0; c < nchunks; ++c) {
      int kn = K < D_KCH ? K : D_KCH;
      if (nchunks > 1) {
        unsigned long long tot = 0;
        kn = dense_load_chunk<PEER>(a, pv, s, b, len, K, c, res_end, tot);
      }
      dense_stream<2, PEER, D_PF>(a, pv, s, m1, m2, S, kn, w, D_WARPS, hv, cl, b0, b1, b2, b3);
    }
    __syncthreads();
    for (int base = 0; base < D_H; base += D_THREADS) {
      const uint32_t slot = base + tid;
      if (((s.fin[slot >> 5] >> (slot & 31u)) & 1u) && hv.count_at(slot) >= kmin)
        s.cand[atomicAdd(&s.ss.ncand, 1u)] = (uint16_t)slot;
    }
    __syncthreads();
    const uint32_t c = s.ss.ncand;
    select_and_emit<D_THREADS>(a, q, hv, [&](uint32_t i) -> uint32_t { return s.cand[i]; }, c, s.ss);
    __syncthreads();
    if (a.nt_mode) {
      if (tid < 32 && a.n_hits[q]) {
        const uint64_t top = a.pool[a.hit_base[q]];
        const uint8_t *r = a.res + b;
        const uint32_t d0 = dense_from_codes(aa_code(r[0]), aa_code(r[1]), aa_code(r[2]), aa_code(r[3]),
                                             aa_code(r[4]), aa_code(r[5]), aa_code(r[6]));
        const bool any = warp_any0<PEER>(a, pv, hv, d0, (uint32_t)top, (uint32_t)(top >> 32));
        if (tid == 0) a.any0[q] = any ? 1 : 0;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_incr += __shfl_down_sync(0xFFFFFFFFu, my_incr, o);
    my_lookups += __shfl_down_sync(0xFFFFFFFFu, my_lookups, o);
  }
  if (lane == 0) {
    if (my_incr) {
      atomicAdd(&a.counters[CNT_INCR], my_incr);
      atomicAdd(&a.counters[CNT_CLS_INCR + 3], my_incr);
    }
    if (my_lookups) {
      atomicAdd(&a.counters[CNT_LOOKUPS], my_lookups);
      atomicAdd(&a.counters[CNT_CLS_LOOKUPS + 3], my_lookups);
    }
  }
}

}  // namespace kaamer
