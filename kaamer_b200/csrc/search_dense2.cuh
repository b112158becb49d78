// search_dense2.cuh — class D, second design: the search kernel of the saturated-key-space regime
// (C4: ~47 postings per query k-mer, ~16 000 (query, subject) increments per 350-aa query, almost all of
// them subjects seen once).
//
// Replaces, like the other classes, KmerSearch + sortMapByValue + FilterResults
// (pkg/search/search.go:414-440, 132-152, 189-220) for one query per CTA of four warps.
//
// Why not a histogram: a shared-memory atomic costs ~2 cycles per lane on the SM's atomic unit (0.5
// increments/clk/SM, 145 G/s per GPU), an order of magnitude below what streaming 190-byte posting lists
// out of HBM delivers.  So the ~99 % of the postings that belong to subjects seen ONCE are filtered out
// with plain shared-memory loads and stores, and only repeated subjects ever reach an atomic:
//
//   pass 1  every posting id tests-and-sets one bit in M1 and, if that was already set, one in M2 (two
//           Bloom-style bit maps with independent hashes); an id that finds both set is PUSHED into a
//           small exact hash H (CAS + atomic count, shared by the CTA);
//   sweep   subjects pushed >= kmin - 2C times are FINAL candidates (C = concurrent chains, below);
//   pass 2  the lists are streamed again (L2 hits) and every occurrence of a final candidate is counted
//           exactly (a 64-bit Bloom word in registers rejects the rest without touching shared memory);
//           then the same threshold / top-N epilogue as the other classes.
//
// The bit maps are PRIVATE to a warp (each warp streams its own quarter of the query's lists), and a warp
// works on R posting lists at a time: all their loads, then all their stores, __syncwarp, then a VERIFY
// read — two lanes of the same group may have written different bits of one word, the loser repairs its bit
// with an atomic OR (rare).  No false negatives, deterministically: ids of one list are unique, so a chain
// (warp x R) leaves a subject unpushed at most twice (the group that sets its M1 bit, the group that sets its
// M2 bit); a subject with Kmatch >= kmin is therefore pushed at least kmin - 2C times with C = W x R chains,
// W = min(4, (kmin-1)/2) streaming warps, R = min(2, (kmin-1)/(2W)).  kmin < 3 goes to class G.  False
// positives (hash collisions) only cost work: pass 2 counts exactly.
#pragma once
#include "search_common.cuh"

namespace kaamer {

constexpr int E_WARPS = 4, E_THREADS = E_WARPS * 32;
constexpr int E_NF = 16;       // final candidates verified by binary search (more: the lists are streamed again)
constexpr int E_MAXK = 60000;  // 16-bit counts

// KCAP: query k-mers staged at once; EH: slots of the exact hash of pushed subjects
template <int KCAP, int EH>
struct __align__(16) Dense2Smem {
  uint64_t ent[KCAP];
  uint32_t hkeys[EH];
  uint32_t hcnt2[EH / 2];
  uint32_t fin[EH / 32];
  uint16_t cand[EH];
  uint16_t pp[KCAP + 8];
  uint8_t raw[KCAP + 64];
  uint8_t lut[256];
  SelectScratch ss;
  unsigned long long bloom;
  uint32_t nfinal, it;
  uint32_t fid[E_NF], fslot[E_NF], fcnt[E_NF];  // final candidates: subject id, slot in H, exact count
  // window descriptors of the staged chunk: entry index | window number << 10 (one 64-id window of one list);
  // lists of more than 64 windows, and what does not fit, go to the long-list queue lq (entry indices)
  uint32_t nd, nlq;
  uint16_t desc[(5 * KCAP) / 2];
  uint16_t lq[KCAP];
};

__device__ __forceinline__ uint32_t e_hash1(uint32_t id, uint32_t nbits) { return __umulhi(id * 0x9E3779B1u, nbits); }
__device__ __forceinline__ uint32_t e_hash2(uint32_t id, uint32_t nbits) {
  uint32_t x = id * 0x85EBCA77u;
  x ^= x >> 13;
  x *= 0xC2B2AE3Du;
  return __umulhi(x, nbits);
}

// One chunk of the query: residues -> packed codes -> table entries in s.ent[0, kn).  Returns kn; every
// thread returns in `tot` the posting total of the entries it probed.
template <bool PEER, int KCAP, int EH>
__device__ __forceinline__ int dense2_load_chunk(const SearchArgs &a, const PeerView *pv, Dense2Smem<KCAP, EH> &s,
                                                 uint64_t b, int len, int K, int c, const uint8_t *res_end,
                                                 unsigned long long &tot) {
  const int tid = threadIdx.x;
  const int kbeg = c * KCAP;
  const int kn = K - kbeg < KCAP ? K - kbeg : KCAP;
  const int nres = len - kbeg < kn + 7 ? len - kbeg : kn + 7;
  __syncthreads();  // the previous users of raw / pp / ent are done
  if (tid == 0) {
    s.nd = 0;
    s.nlq = 0;
  }
  const int head = stage_bytes<E_THREADS>(s.raw, a.res + b + kbeg, nres, res_end, tid);
  __syncthreads();
  const uint8_t *r = s.raw + head;
  const int ncodes = kn + KAAMER_KMER_SIZE - 1;
  for (int i = tid; i < ncodes; i += E_THREADS) {
    const uint32_t c0 = s.lut[r[i]];
    const uint32_t c1 = (i + 1 < nres) ? (uint32_t)s.lut[r[i + 1]] : CODE_UNKNOWN;
    s.pp[i] = (uint16_t)packed_code(c0, c1);
  }
  __syncthreads();
  constexpr int U = 4;
#pragma unroll 1
  for (int base = 0; base < kn; base += U * E_THREADS) {
    uint32_t d[U];
    bool ok[U];
    uint64_t e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pos = base + u * E_THREADS + tid;
      ok[u] = pos < kn;
      d[u] = ok[u] ? dense_from_packed(s.pp[pos], s.pp[pos + 2], s.pp[pos + 4], s.pp[pos + 6]) : 0u;
    }
    probe_entries<PEER, U>(a, pv, d, ok, e);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pos = base + u * E_THREADS + tid;
      uint32_t nwin = 0;
      if (ok[u]) {
        s.ent[pos] = e[u];
        const uint32_t cnt = (uint32_t)(e[u] >> ENTRY_VALUE_BITS);
        tot += cnt;
        nwin = (cnt + 63u) >> 6;
      }
      // window descriptors, reserved with one shared-memory atomic per warp
      const bool is_long = nwin > 64u;
      const uint32_t want = is_long ? 0u : nwin;
      uint32_t incl = want;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
      }
      const uint32_t wtot = __shfl_sync(0xFFFFFFFFu, incl, 31);
      uint32_t wbase = 0;
      if ((threadIdx.x & 31) == 31 && wtot) wbase = atomicAdd(&s.nd, wtot);
      wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
      const uint32_t first = wbase + incl - want;
      constexpr uint32_t DCAP = (5 * KCAP) / 2;
      if (want && first + want <= DCAP) {
        for (uint32_t wn = 0; wn < want; ++wn) s.desc[first + wn] = (uint16_t)((uint32_t)pos | (wn << 10));
      } else if (is_long) {
        s.lq[atomicAdd(&s.nlq, 1u)] = (uint16_t)pos;  // (rare: a list of more than 4096 ids)
      }
    }
  }
  __syncthreads();
  if (s.nd > (uint32_t)((5 * KCAP) / 2)) {
    // more windows than descriptors (a database ~3x denser than C4): every list of the chunk is walked
    // through the long-list queue instead — slower per window, same result
    __syncthreads();
    if (tid == 0) s.nd = 0;
    for (int pos = tid; pos < kn; pos += E_THREADS) {
      const uint32_t cnt = (uint32_t)(s.ent[pos] >> ENTRY_VALUE_BITS);
      if (cnt > 0 && ((cnt + 63u) >> 6) <= 64u) s.lq[atomicAdd(&s.nlq, 1u)] = (uint16_t)pos;
    }
    __syncthreads();
  }
  return kn;
}

// Warp `wi` of `nw` streams the posting lists of entries wi, wi + nw, ... of the staged chunk.  The unit is one
// 64-id window of ONE list (two ids per lane: positions lane and lane + 32 of the window); R windows of R
// different lists form a group, and the loads of the next group are in flight while a group is processed.
// The body is flat — predicated loads and stores, no nested branches — and exists once per instantiation:
// the kernel is bound by issued instructions (and, before this layout, by instruction-cache misses), not by HBM.
//
// M1 and M2 are interleaved, mm[word] = (M1 word, M2 word): an id names ONE word index and two bit positions
// (from two multiplicative hashes), so the test is one 8-byte load and the verify another.
// LONGQ = false: the windows are the descriptors desc[wi], desc[wi + nw], ... (one shared-memory read names the
// entry and the window: no list-walking state).  LONGQ = true: the lists of the long-list queue, window by window.
template <int PASS, bool PEER, int R, bool LONGQ, int KCAP, int EH>
__device__ __forceinline__ void dense2_stream(const SearchArgs &a, const PeerView *pv, Dense2Smem<KCAP, EH> &s, uint2 *mm,
                                              int lg, int wi, int nw, const SmemHashT<false> &hv, const CandList &cl,
                                              unsigned long long bloom) {
  const unsigned lane = threadIdx.x & 31;
  const int kn = LONGQ ? (int)s.nlq : (int)s.nd;
  int k = LONGQ ? wi - nw : wi;
  uint32_t off = 0, cnt = 0;
  const uint32_t *ptr = nullptr;
  // warp-uniform iterator over 64-id windows; nv = 0 once exhausted
  auto fetch = [&](uint32_t &ia, uint32_t &ib, uint32_t &nv) {
    nv = 0;
    ia = ib = 0;
    if constexpr (!LONGQ) {
      if (k >= kn) return;
      const uint32_t d = s.desc[k];
      k += nw;
      const uint64_t e = s.ent[d & 1023u];
      const uint32_t c = (uint32_t)(e >> ENTRY_VALUE_BITS), o = (d >> 10) << 6;
      nv = c - o < 64u ? c - o : 64u;
      if (c == 1) {
        ia = (uint32_t)e;  // the posting inlined in the entry (only lane 0 is valid: nv = 1)
      } else {
        const uint32_t *p = post_ptr<PEER>(a, pv, e & ENTRY_VALUE_MASK) + o + lane;
        if (lane < nv) ia = __ldg(p);
        if (lane + 32u < nv) ib = __ldg(p + 32);
      }
    } else {
      if (off >= cnt) {
        k += nw;
        if (k >= kn) {
          k = kn;  // stay exhausted
          cnt = off = 0;
          return;
        }
        const uint64_t e = s.ent[s.lq[k]];
        cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);  // >= 2: a long list, or one that did not fit the descriptors
        off = 0;
        ptr = post_ptr<PEER>(a, pv, e & ENTRY_VALUE_MASK);
        if (cnt == 1) {  // (an inlined posting that did not fit the descriptor array)
          ia = (uint32_t)e;
          nv = 1;
          off = 64;
          return;
        }
      }
      const uint32_t rem = cnt - off;
      nv = rem < 64u ? rem : 64u;
      const uint32_t *p = ptr + off + lane;
      if (lane < nv) ia = __ldg(p);
      if (lane + 32u < nv) ib = __ldg(p + 32);
      off += 64;
    }
  };
  constexpr int J = 2 * R;
  const int sh_w = 37 - lg, sh_b = 32 - lg;  // word index = x1 >> (32 - lg + 5), bit index = (x1 >> (32 - lg)) & 31
  // AHEAD groups in flight beyond the one being processed.  Two on local HBM: ncu showed 22 % of the stall
  // samples on the first use of a window fetched only one group earlier (a posting list is a dependent HBM
  // access behind its table entry).  Six when the lists may live on another GPU (mode P): an NVLink round trip
  // is several times an HBM one.
  constexpr int AHEAD = PEER ? 6 : 2;
  uint32_t qa[AHEAD][R], qb[AHEAD][R], qn[AHEAD][R];
#pragma unroll
  for (int d = 0; d < AHEAD; ++d)
#pragma unroll
    for (int r = 0; r < R; ++r) fetch(qa[d][r], qb[d][r], qn[d][r]);
#pragma unroll 1
  while (qn[0][0] != 0) {
    uint32_t id[J];
    bool v[J];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      id[2 * r] = qa[0][r];
      id[2 * r + 1] = qb[0][r];
      v[2 * r] = lane < qn[0][r];
      v[2 * r + 1] = lane + 32u < qn[0][r];
    }
#pragma unroll
    for (int d = 0; d + 1 < AHEAD; ++d)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        qa[d][r] = qa[d + 1][r];
        qb[d][r] = qb[d + 1][r];
        qn[d][r] = qn[d + 1][r];
      }
#pragma unroll
    for (int r = 0; r < R; ++r) fetch(qa[AHEAD - 1][r], qb[AHEAD - 1][r], qn[AHEAD - 1][r]);
    if constexpr (PASS == 1) {
      uint32_t wd[J], b1[J], b2[J];
      uint2 w[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t x1 = id[j] * 0x9E3779B1u, x2 = id[j] * 0x85EBCA77u;
        wd[j] = x1 >> sh_w;
        b1[j] = 1u << ((x1 >> sh_b) & 31u);
        b2[j] = 1u << (x2 >> 27);
        w[j] = mm[wd[j]];  // (id 0 of an invalid lane still names a valid word)
      }
      bool set1[J], set2[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const bool seen1 = (w[j].x & b1[j]) != 0u, seen2 = (w[j].y & b2[j]) != 0u;
        set1[j] = v[j] && !seen1;
        set2[j] = v[j] && seen1 && !seen2;
        if (set1[j]) mm[wd[j]].x = w[j].x | b1[j];
        if (set2[j]) mm[wd[j]].y = w[j].y | b2[j];
        if (v[j] && seen1 && seen2) count_subject(hv, id[j], 0xFFFFFFFFu, cl);
      }
      __syncwarp();
      // verify: a store of this group may have been overwritten by another lane's store to the same word
#pragma unroll
      for (int j = 0; j < J; ++j) {
        uint2 c = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
        if (set1[j] || set2[j]) c = mm[wd[j]];
        if (set1[j] && (c.x & b1[j]) == 0u) atomicOr(&mm[wd[j]].x, b1[j]);
        if (set2[j] && (c.y & b2[j]) == 0u) atomicOr(&mm[wd[j]].y, b2[j]);
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t hb = (id[j] * 0x9E3779B1u) >> 26;
        if (v[j] && ((bloom >> hb) & 1ull)) {
          uint32_t slot = hv.home(id[j]);
#pragma unroll 1
          for (int probe = 0; probe < SmemHashT<false>::kMaxProbe; ++probe) {
            const uint32_t key = s.hkeys[slot];
            if (key == id[j]) {
              if ((s.fin[slot >> 5] >> (slot & 31u)) & 1u) hv.add(slot, 1u);
              break;
            }
            if (key == EMPTY) break;
            slot = (slot + 1) & hv.mask;
          }
        }
      }
    }
  }
}

// Exact counts of the (few) final candidates without streaming the lists again: a posting list is sorted
// (ids strictly descending, pkg/kvstore/kv_store.go:284-305), so "does list k hold subject X" is a binary
// search — ~6 dependent loads against ~47 ids scanned.  One (list, candidate) pair per thread.
template <bool PEER, int KCAP, int EH>
__device__ __forceinline__ void dense2_verify(const SearchArgs &a, const PeerView *pv, Dense2Smem<KCAP, EH> &s, int kn,
                                              int nf) {
  const int total = kn * nf;
  for (int p = threadIdx.x; p < total; p += E_THREADS) {
    const int k = p / nf, f = p - k * nf;
    const uint64_t e = s.ent[k];
    const uint32_t cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);
    if (cnt == 0) continue;
    const uint32_t x = s.fid[f];
    bool hit = false;
    if (cnt == 1) {
      hit = (uint32_t)e == x;
    } else {
      const uint32_t *pl = post_ptr<PEER>(a, pv, e & ENTRY_VALUE_MASK);
      uint32_t lo = 0, hi = cnt;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t v = __ldg(pl + mid);
        if (v == x) {
          hit = true;
          break;
        }
        if (v > x) lo = mid + 1;  // descending order
        else hi = mid;
      }
    }
    if (hit) atomicAdd(&s.fcnt[f], 1u);
  }
}

// CLS: class list (4: queries up to KCAP k-mers staged at once; 5: longer ones; 6: the longest, chunked)
template <bool PEER, int KCAP, int EH, int CLS>
__global__ void __launch_bounds__(E_THREADS, CLS == 4 ? (PEER ? 4 : 6) : (CLS == 5 ? (PEER ? 3 : 4) : 1)) k_search_e(SearchArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  using Smem = Dense2Smem<KCAP, EH>;
  constexpr int E_H = EH;
  Smem &s = *reinterpret_cast<Smem *>(dsm);
  const uint32_t mapw = CLS == 4 ? a.e_mapw_small : (CLS == 5 ? a.e_mapw_large : a.e_mapw_xl);  // words per map per warp (2^n)
  const int lg = 31 - __clz(mapw) + 5;                                  // log2 of the bits per map
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;
  const int w = tid >> 5;
  uint2 *mm = reinterpret_cast<uint2 *>(dsm + ((sizeof(Smem) + 15) & ~(size_t)15)) + (size_t)w * mapw;
  const PeerView *pv = nullptr;
  if constexpr (PEER) {
    __shared__ PeerView s_peer;
    load_peer_view(&s_peer, a.peer, tid, E_THREADS);
    pv = &s_peer;
  }
  for (int i = tid; i < 256; i += E_THREADS) s.lut[i] = (uint8_t)aa_code(i);
  __syncthreads();
  const SmemHashT<false> hv{s.hkeys, s.hcnt2, (uint32_t)E_H - 1u, 32 - ilog2_c(E_H)};
  const CandList cl{&s.ss.ncand, &s.ss.flags, s.cand, nullptr, (uint32_t)E_H};
  const uint32_t count = a.list_count[CLS];
  const uint8_t *res_end = a.res + a.off[a.nq];
  unsigned long long my_incr = 0, my_lookups = 0;
  for (;;) {
    __syncthreads();
    if (tid == 0) s.it = atomicAdd(&a.list_count[N_LISTS + CLS], 1u);
    __syncthreads();
    const uint32_t it = s.it;
    if (it >= count) break;
    const uint32_t q = a.lists[(size_t)CLS * a.nq + it];
    const uint64_t b = a.off[q];
    const int len = (int)(a.off[q + 1] - b);
    const int K = a.size_in_kmer[q];
    const uint32_t kmin = a.kmin[q];  // >= 3 (k_classify)
    const int w_act = (int)((kmin - 1u) / 2u) < E_WARPS ? (int)((kmin - 1u) / 2u) : E_WARPS;
    int R = (int)((kmin - 1u) / (2u * (uint32_t)w_act));
    R = R >= 2 ? 2 : 1;  // (four lists at a time cost 40 more registers than they are worth)
    const uint32_t thr = kmin - 2u * (uint32_t)(w_act * R);  // >= 1
    const int nchunks = (K + KCAP - 1) / KCAP;
    {
      uint4 *hk = reinterpret_cast<uint4 *>(s.hkeys);
      uint4 *hc = reinterpret_cast<uint4 *>(s.hcnt2);
      const uint4 E = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY), Z = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < E_H / 4; i += E_THREADS) hk[i] = E;
      for (int i = tid; i < E_H / 8; i += E_THREADS) hc[i] = Z;
      if (tid == 0) {
        s.bloom = 0ull;
        s.ss.ncand = 0;
        s.ss.flags = 0;
        s.nfinal = 0;
      }
      // the warp's own maps (the first probes of the chunk are issued right after)
      uint4 *mv = reinterpret_cast<uint4 *>(mm);
      for (uint32_t i = lane; i < mapw / 2; i += 32) mv[i] = Z;  // mapw word pairs = mapw / 2 uint4
    }
    unsigned long long q_incr = 0;
    // ---- pass 1 ----
    for (int c = 0; c < nchunks; ++c) {
      unsigned long long tot = 0;
      dense2_load_chunk<PEER, KCAP, EH>(a, pv, s, b, len, K, c, res_end, tot);
      q_incr += tot;
      if (w < w_act) {
        if (R == 2) dense2_stream<1, PEER, 2, false, KCAP, EH>(a, pv, s, mm, lg, w, w_act, hv, cl, 0ull);
        else dense2_stream<1, PEER, 1, false, KCAP, EH>(a, pv, s, mm, lg, w, w_act, hv, cl, 0ull);
        if (s.nlq) dense2_stream<1, PEER, 1, true, KCAP, EH>(a, pv, s, mm, lg, w, w_act, hv, cl, 0ull);
      }
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
    for (int base = 0; base < E_H; base += E_THREADS) {
      const uint32_t slot = base + tid;
      const uint32_t key = s.hkeys[slot];
      const bool isfin = key != EMPTY && hv.count_at(slot) >= thr;
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, isfin);
      if (lane == 0) s.fin[slot >> 5] = bal;
      if (isfin) {
        atomicOr(&s.bloom, 1ull << ((key * 0x9E3779B1u) >> 26));
        const uint32_t fi = atomicAdd(&s.nfinal, 1u);
        if (fi < (uint32_t)E_NF) {
          s.fid[fi] = key;
          s.fslot[fi] = slot;
          s.fcnt[fi] = 0;
        }
      }
    }
    __syncthreads();
    {
      uint4 *hc = reinterpret_cast<uint4 *>(s.hcnt2);
      const uint4 Z = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < E_H / 8; i += E_THREADS) hc[i] = Z;
    }
    __syncthreads();
    my_incr += q_incr;
    if (tid == 0) my_lookups += (unsigned long long)K;
    if (s.nfinal == 0) continue;  // nothing can reach kmin: no hits (n_hits[q] was zeroed by k_classify)
    // ---- pass 2: exact counts of the final candidates ----
    const unsigned long long bloom = s.bloom;
    const int nf = (int)s.nfinal;
    const bool by_search = nf <= E_NF && a.lists_sorted;
    for (int c = 0; c < nchunks; ++c) {
      int kn = K < KCAP ? K : KCAP;
      if (nchunks > 1) {
        unsigned long long tot = 0;
        kn = dense2_load_chunk<PEER, KCAP, EH>(a, pv, s, b, len, K, c, res_end, tot);
      }
      if (by_search) {
        dense2_verify<PEER, KCAP, EH>(a, pv, s, kn, nf);
      } else {
        dense2_stream<2, PEER, 1, false, KCAP, EH>(a, pv, s, mm, lg, w, E_WARPS, hv, cl, bloom);
        if (s.nlq) dense2_stream<2, PEER, 1, true, KCAP, EH>(a, pv, s, mm, lg, w, E_WARPS, hv, cl, bloom);
      }
    }
    __syncthreads();
    if (by_search) {
      if (tid < nf) hv.add(s.fslot[tid], s.fcnt[tid]);  // (the counts of H were zeroed after the sweep)
      __syncthreads();
    }
    for (int base = 0; base < E_H; base += E_THREADS) {
      const uint32_t slot = base + tid;
      if (((s.fin[slot >> 5] >> (slot & 31u)) & 1u) && hv.count_at(slot) >= kmin)
        s.cand[atomicAdd(&s.ss.ncand, 1u)] = (uint16_t)slot;
    }
    __syncthreads();
    const uint32_t c = s.ss.ncand;
    select_and_emit<E_THREADS>(a, q, hv, [&](uint32_t i) -> uint32_t { return s.cand[i]; }, c, s.ss);
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
