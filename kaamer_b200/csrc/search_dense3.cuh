// search_dense3.cuh — class D, third design: the search kernel of the saturated-key-space regime
// (C4: ~47 postings per query k-mer, ~16 000 (query, subject) increments per 350-aa query, almost all of
// them subjects seen once).
//
// Replaces, like the other classes, KmerSearch + sortMapByValue + FilterResults
// (pkg/search/search.go:414-440, 132-152, 189-220) for one query per CTA of NW warps.
//
// The filter is the one of the second design (search_dense2.cuh): every posting id tests-and-sets one bit
// in a warp-private map M1 and, if that was already set, one in M2; an id that finds both set is PUSHED
// into a small exact hash H; subjects pushed >= kmin - 2W times (W streaming warps, one list window per
// warp at a time) are the FINAL candidates, counted exactly by binary search in the sorted posting lists.
// What the ncu source view of the second design showed (profiles/r2p_*): the kernel is bound by ISSUED
// INSTRUCTIONS — 60 per 32 streamed ids, 56 % of the lanes holding an id — not by HBM.  Hence:
//
//   * 32-id windows, one id per lane (lists have a median of 32 ids: 64-id windows ran 44 % empty); the
//     first window of a list is described by its table entry itself, further windows by 8-byte descriptors
//     of the same layout, both arrays padded with null windows: a fetch is one LDS.64, a min and an LEA;
//   * M1 and M2 share one 32-bit word (16 + 16 bits, one multiplicative hash names the word and both
//     bits): one LDS.32 to test, one STS.32 to set, one LDS.32 to verify;
//   * HOT subjects: a subject pushed three times by a warp is kept in a register and from then on counted
//     there (one compare per id) instead of going through the map and the hash — the true hits of a query
//     recur in every other window and their pushes were a quarter of all instructions;
//   * the software pipeline of windows (two ahead, six in mode P) is unrolled by its depth: no register
//     rotation;
//   * the longest queries get sixteen warps per CTA instead of four (they ran at 6 % occupancy).
//
// No false negatives, deterministically (the argument of search_dense2.cuh with R = 1): ids of one list are
// unique, a warp handles one window at a time with __syncwarp + verify between its stores and its next
// loads, so a warp leaves a subject unpushed at most twice; hot counting only adds exact occurrences.
#pragma once
#include "search_common.cuh"

namespace kaamer {

constexpr int F_NF = 16;       // final candidates verified by binary search (more: the lists are streamed again)
constexpr int F_MAXK = 60000;  // 16-bit counts
constexpr uint32_t F_NONE = 0xFFFFFFFFu;
#ifndef KAAMER_F_PH
#define KAAMER_F_PH 4
#endif
constexpr int F_PH_LOCAL = KAAMER_F_PH, F_PH_PEER = 8;

// KCAP: query k-mers staged at once; EH: slots of the exact hash; NW warps per CTA; PH windows in flight per warp.
//
// Windows of 64 ids (two per lane: i = lane and lane + 32; the second half is skipped when the first is not full).
// The FIRST window of the list of k-mer `pos` is described by its table entry ent[pos]
// itself (count | offset of the list); every further window gets a descriptor of the same layout behind the kn
// entries of the chunk: (ids in the window) << 37 | offset of its first id.  The array ends in NW * PH null
// descriptors so that the pipelined fetch of a warp never checks a bound.  Lists of more than 64 windows, and
// windows that do not fit, are walked through the long-list queue lq (entry indices) from their second window on.
// (Lists have a median of 32 ids: half of them need no second half, a quarter one, and the per-window work —
// descriptor, address, cp.async, loop — is paid once per 64 ids where it is needed most.)
template <int KCAP, int EH, int NW, int PH>
struct __align__(16) Dense3Smem {
  static constexpr int PAD = NW * PH;
  static constexpr int XCAP = KCAP / 2;
  uint64_t ent[KCAP + XCAP + PAD];  // [0, kn): table entries = first windows; [kn, kn + nx): further windows; nulls
  uint32_t hkeys[EH];
  uint32_t hcnt2[EH / 2];
  uint32_t fin[EH / 32];
  uint16_t lq[KCAP];
  uint8_t lut[256];
  // The posting windows in flight (pass 1) share their space with what only the other phases of a query use:
  // the staged residues and codes (load_chunk) and the candidate list and selection scratch (epilogue).
  struct Phases {
    uint16_t cand[EH];
    uint16_t pp[KCAP + 8];
    uint8_t raw[KCAP + 64];
    SelectScratch ss;
  };
  union {
    Phases e;
    uint32_t ring[NW][PH][64];  // ring[w][slot][i]: id i of window `slot` of warp w (cp.async destination)
  } u;
  unsigned long long bloom;
  uint32_t nfinal, it, pflags;                  // pflags: bit0 = H is full (pass 1)
  uint32_t fid[F_NF], fslot[F_NF], fcnt[F_NF];  // final candidates: subject id, slot in H, exact count
  uint32_t nd, nx_end, nlq;                     // xd slots reserved, first slot of a failed reservation, lq entries
};

// One chunk of the query: residues -> packed codes -> table entries in s.ent[0, kn) + the descriptors of the
// further windows.  Returns kn; every thread returns in `tot` the posting total of the entries it probed.
template <bool PEER, int NT, class Smem>
__device__ __forceinline__ int dense3_load_chunk(const SearchArgs &a, const PeerView *pv, Smem &s, int KCAP, uint64_t b,
                                                 int len, int K, int c, const uint8_t *res_end,
                                                 unsigned long long &tot) {
  constexpr uint32_t XCAP = Smem::XCAP;
  const int tid = threadIdx.x;
  const int kbeg = c * KCAP;
  const int kn = K - kbeg < KCAP ? K - kbeg : KCAP;
  const int nres = len - kbeg < kn + 7 ? len - kbeg : kn + 7;
  __syncthreads();  // the previous users of raw / pp / ent / xd are done
  if (tid == 0) {
    s.nd = 0;
    s.nx_end = XCAP;
    s.nlq = 0;
  }
  const int head = stage_bytes<NT>(s.u.e.raw, a.res + b + kbeg, nres, res_end, tid);
  __syncthreads();
  const uint8_t *r = s.u.e.raw + head;
  const int ncodes = kn + KAAMER_KMER_SIZE - 1;
  for (int i = tid; i < ncodes; i += NT) {
    const uint32_t c0 = s.lut[r[i]];
    const uint32_t c1 = (i + 1 < nres) ? (uint32_t)s.lut[r[i + 1]] : CODE_UNKNOWN;
    s.u.e.pp[i] = (uint16_t)packed_code(c0, c1);
  }
  __syncthreads();
  constexpr int U = 4;
#pragma unroll 1
  for (int base = 0; base < kn; base += U * NT) {
    uint32_t d[U];
    bool ok[U];
    uint64_t e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pos = base + u * NT + tid;
      ok[u] = pos < kn;
      d[u] = ok[u] ? dense_from_packed(s.u.e.pp[pos], s.u.e.pp[pos + 2], s.u.e.pp[pos + 4], s.u.e.pp[pos + 6]) : 0u;
    }
    probe_entries<PEER, U>(a, pv, d, ok, e);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pos = base + u * NT + tid;
      uint32_t cnt = 0;
      if (ok[u]) {
        s.ent[pos] = e[u];
        cnt = (uint32_t)(e[u] >> ENTRY_VALUE_BITS);
        tot += cnt;
      }
      // descriptors of windows 1, 2, ... of the list, reserved with one shared-memory atomic per warp
      const uint32_t extra = cnt > 64u ? (cnt - 1u) >> 6 : 0u;
      const bool is_long = extra > 63u;
      const uint32_t want = is_long ? 0u : extra;
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
      if (want && first + want <= XCAP) {
        const uint64_t val = e[u] & ENTRY_VALUE_MASK;
        for (uint32_t wn = 1; wn <= want; ++wn) {
          const uint32_t o = wn << 6, rem = cnt - o;
          s.ent[kn + first + wn - 1u] = ((uint64_t)(rem < 64u ? rem : 64u) << ENTRY_VALUE_BITS) | (val + o);
        }
      } else if (want || is_long) {
        if (want) atomicMin(&s.nx_end, first);      // (a database ~3x denser than C4: the descriptors are full)
        s.lq[atomicAdd(&s.nlq, 1u)] = (uint16_t)pos;  // walked list by list from its second window on
      }
    }
  }
  __syncthreads();
  {
    const uint32_t nx = s.nd < s.nx_end ? s.nd : s.nx_end;
    if (tid < Smem::PAD) s.ent[kn + nx + tid] = 0ull;
  }
  __syncthreads();
  return kn;
}

// Per-warp state of pass 1: the hot subjects (warp-uniform ids, per-lane occurrence counts)
struct HotState {
  uint32_t id0, id1;
  uint32_t n0, n1;
};

// One step of pass 1: J ids per lane (of ONE list: distinct), F_NONE where there is none.  mm: the warp's map,
// word = M1 bits 0-15 | M2 bits 16-31; the word index is the top lgw bits of the multiplicative hash, the two bit
// numbers the 4 + 4 bits below.  All loads, then all stores, one __syncwarp, then the verify reads.
template <int J>
__device__ __forceinline__ void dense3_window(const uint32_t (&id)[J], uint32_t *mm, int sh_w, int sh_1, int sh_2,
                                              const SmemHashT<true> &hv, uint32_t *flags, HotState &hot) {
  uint32_t *wp[J];
  uint32_t m1[J], m2[J], w[J], setbit[J];
  bool push[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const uint32_t x = id[j] * 0x9E3779B1u;
    wp[j] = mm + (x >> sh_w);
    // (opaque shifts: the compiler would turn `w & (1 << f)` into shift-and-mask sequences)
    asm("shl.b32 %0, 1, %1;" : "=r"(m1[j]) : "r"((x >> sh_1) & 15u));
    asm("shl.b32 %0, 0x10000, %1;" : "=r"(m2[j]) : "r"((x >> sh_2) & 15u));
    w[j] = *wp[j];  // (the id of an invalid lane still names a valid word)
  }
  bool any_push = false;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const bool v = id[j] != F_NONE;
    const bool seen1 = (w[j] & m1[j]) != 0u, seen2 = (w[j] & m2[j]) != 0u;
    push[j] = v && seen1 && seen2;
    const bool need = v && !(seen1 && seen2);
    setbit[j] = need ? (seen1 ? m2[j] : m1[j]) : 0u;  // the bit this lane sets for this id, if any
    if (need) *wp[j] = w[j] | setbit[j];
    any_push = any_push || push[j];
  }
  if (__any_sync(0xFFFFFFFFu, any_push)) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      // a hot subject is counted in a register (its bits are set: it always arrives here)
      const bool h0 = push[j] && id[j] == hot.id0, h1 = push[j] && id[j] == hot.id1;
      if (h0) ++hot.n0;
      if (h1) ++hot.n1;
      const bool cold = push[j] && !h0 && !h1;
      if (__any_sync(0xFFFFFFFFu, cold)) {
        uint32_t after = 0;
        if (cold) {
          uint32_t slot = hv.home(id[j]);
          bool done = false;
#pragma unroll 1
          for (int probe = 0; probe < SmemHashT<true>::kMaxProbe; ++probe) {
            const uint32_t cur = hv.cas(slot, id[j]);
            if (cur == EMPTY || cur == id[j]) {
              after = hv.inc(slot) + 1u;
              done = true;
              break;
            }
            slot = (slot + 1) & hv.mask;
          }
          if (!done) atomicOr(flags, 1u);  // more repeated subjects than H holds
        }
        // a subject this warp has pushed three times becomes hot (a false positive of the maps rarely is)
        const unsigned hm = __ballot_sync(0xFFFFFFFFu, after >= 3u);
        if (hm && hot.id1 == F_NONE) {
          const uint32_t cand = __shfl_sync(0xFFFFFFFFu, id[j], __ffs(hm) - 1);
          if (hot.id0 == F_NONE) hot.id0 = cand;
          else hot.id1 = cand;
        }
      }
    }
  }
  __syncwarp();
  // verify: a store may have been overwritten by another store to the same word (another lane's, or the lane's own
  // second id)
#pragma unroll
  for (int j = 0; j < J; ++j)
    if (~(*wp[j]) & setbit[j]) atomicOr(wp[j], setbit[j]);
  __syncwarp();
}

// Pass 1 of warp `wi` of `nw` over the windows win[wi], win[wi + nw], ... (n of them in all, the first kn being
// table entries).  The ids of PH - 1 windows are in flight while one is processed: each lane copies its id of a
// window into the warp's ring with cp.async (LDGSTS) — lanes without an id store F_NONE — and the consumer waits
// with cp.async.wait_group.  Loads into registers cannot be pipelined this deep: ptxas tracks all of them with
// one of the six scoreboards of a warp, so waiting for the oldest waits for the youngest (profiles/r2t_*).  One
// copy of the loop body (slot numbers are run-time values): unrolled by PH it missed the instruction cache.
template <bool PEER, int PH>
__device__ __forceinline__ void dense3_pass1(const SearchArgs &a, const PeerView *pv, const uint64_t *win, int kn, int n,
                                             uint32_t *mm, int lgw, int wi, int nw, uint32_t *ring,
                                             const SmemHashT<true> &hv, uint32_t *flags, HotState &hot) {
  const unsigned lane = threadIdx.x & 31;
  const int sh_w = 32 - lgw, sh_1 = 28 - lgw, sh_2 = 24 - lgw;
  int left = wi < n ? (n - wi + nw - 1) / nw : 0;
  const uint64_t *p = win + wi, *first_end = win + kn;
  const uint32_t *plane = PEER ? nullptr : a.postings + lane;
  asm volatile("" : "+l"(plane));  // (kept in registers: the compiler would rebuild it from the constant bank per window)
  const uint32_t rs = (uint32_t)__cvta_generic_to_shared(ring + lane);  // the lane's word of slot 0
  constexpr uint32_t RING_BYTES = PH * 256u;
  uint32_t wslot = 0;  // byte offset of the slot to fill; the slot behind it is the oldest one in flight
  auto fetch = [&]() {
    const uint64_t e = *p;
    const bool is_first = p < first_end;
    p += nw;
    const uint32_t c = (uint32_t)(e >> ENTRY_VALUE_BITS);
    uint32_t nl = c < 64u ? c : 64u;  // ids of the window
    uint32_t other = F_NONE;          // what the lanes without an id store
    if (is_first && c == 1u) {        // the posting inlined in the entry (lane 0)
      nl = 0;
      if (lane == 0) other = (uint32_t)e;
    }
    uint64_t ptr;
    if constexpr (PEER) {
      ptr = (uint64_t)(post_ptr<PEER>(a, pv, e & ENTRY_VALUE_MASK) + lane);
    } else {
      // plane + 4 * value: one wide multiply-add for the low word, the 5 high bits of the value go to the high
      // word (in PTX: the compiler expands the 64-bit form into six instructions)
      asm("{\n .reg .b32 l, h, t;\n mad.wide.u32 %0, %1, 4, %3;\n and.b32 t, %2, 31;\n mov.b64 {l, h}, %0;\n"
          " mad.lo.u32 h, t, 4, h;\n mov.b64 %0, {l, h};\n}"
          : "=l"(ptr)
          : "r"((uint32_t)e), "r"((uint32_t)(e >> 32)), "l"(plane));
    }
    asm volatile(
        "{\n .reg .pred p;\n setp.lt.u32 p, %2, %3;\n @p cp.async.ca.shared.global [%0], [%1], 4;\n"
        "@!p st.shared.u32 [%0], %4;\n}"
        :
        : "r"(rs + wslot), "l"(ptr), "r"(lane), "r"(nl), "r"(other));
    if (nl >= 32u)  // (warp-uniform) the second half is read only behind a full first half
      asm volatile(
          "{\n .reg .pred p;\n setp.lt.u32 p, %2, %3;\n @p cp.async.ca.shared.global [%0+128], [%1+128], 4;\n"
          "@!p st.shared.u32 [%0+128], %4;\n}"
          :
          : "r"(rs + wslot), "l"(ptr), "r"(lane + 32u), "r"(nl), "r"(F_NONE));
    asm volatile("cp.async.commit_group;");
    wslot = wslot + 256u == RING_BYTES ? 0u : wslot + 256u;
  };
#pragma unroll
  for (int d = 0; d + 1 < PH; ++d) fetch();
#pragma unroll 1
  for (; left > 0; --left) {
    fetch();
    asm volatile("cp.async.wait_group %0;" ::"n"(PH - 1) : "memory");
    uint32_t ida, idb;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ida) : "r"(rs + wslot));
    asm volatile("ld.shared.u32 %0, [%1+128];" : "=r"(idb) : "r"(rs + wslot));  // (stale unless the first half is full)
    if (__all_sync(0xFFFFFFFFu, ida != F_NONE)) {
      const uint32_t id[2] = {ida, idb};
      dense3_window<2>(id, mm, sh_w, sh_1, sh_2, hv, flags, hot);
    } else {
      const uint32_t id[1] = {ida};  // (an empty window — a k-mer the database lacks — is harmless)
      dense3_window<1>(id, mm, sh_w, sh_1, sh_2, hv, flags, hot);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // (the null windows behind the last one: empty groups)
}

// The lists of the long-list queue, window by window from their second window on (rare: no pipelining)
template <bool PEER, class Smem>
__device__ __forceinline__ void dense3_pass1_long(const SearchArgs &a, const PeerView *pv, Smem &s, uint32_t *mm, int lgw,
                                                  int wi, int nw, const SmemHashT<true> &hv, uint32_t *flags,
                                                  HotState &hot) {
  const unsigned lane = threadIdx.x & 31;
  const int sh_w = 32 - lgw, sh_1 = 28 - lgw, sh_2 = 24 - lgw;
  const int n = (int)s.nlq;
#pragma unroll 1
  for (int k = wi; k < n; k += nw) {
    const uint64_t e = s.ent[s.lq[k]];
    const uint32_t cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);
    const uint32_t *ptr = post_ptr<PEER>(a, pv, e & ENTRY_VALUE_MASK);
#pragma unroll 1
    for (uint32_t off = 64; off < cnt; off += 32) {
      const uint32_t nv = cnt - off < 32u ? cnt - off : 32u;
      uint32_t id[1] = {F_NONE};
      if (lane < nv) id[0] = __ldg(ptr + off + lane);
      dense3_window<1>(id, mm, sh_w, sh_1, sh_2, hv, flags, hot);
    }
  }
}

// Pass 2 by streaming (more final candidates than F_NF, or lists of unknown order): every occurrence of a
// final candidate is counted exactly; a 64-bit Bloom word in registers rejects the rest.  Warp w of NW takes the
// lists w, w + NW, ... of the staged chunk.
template <bool PEER, int NW, class Smem>
__device__ __forceinline__ void dense3_pass2(const SearchArgs &a, const PeerView *pv, Smem &s, int kn, int w,
                                             const SmemHashT<true> &hv, unsigned long long bloom) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll 1
  for (int k = w; k < kn; k += NW) {
    const uint64_t e = s.ent[k];
    const uint32_t cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);
    const uint32_t *ptr = cnt >= 2u ? post_ptr<PEER>(a, pv, e & ENTRY_VALUE_MASK) : nullptr;
#pragma unroll 1
    for (uint32_t off = 0; off < cnt; off += 32) {
      uint32_t id = (uint32_t)e;
      bool v = lane == 0;
      if (cnt >= 2u) {
        v = off + lane < cnt;
        if (v) id = __ldg(ptr + off + lane);
      }
      const uint32_t hb = (id * 0x9E3779B1u) >> 26;
      if (v && ((bloom >> hb) & 1ull)) {
        uint32_t slot = hv.home(id);
#pragma unroll 1
        for (int probe = 0; probe < SmemHashT<true>::kMaxProbe; ++probe) {
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
}

// Exact counts of the (few) final candidates without streaming the lists again: a posting list is sorted
// (ids strictly descending, pkg/kvstore/kv_store.go:284-305), so "does list k hold subject X" is a binary
// search — ~6 dependent loads against ~47 ids scanned.  One (list, candidate) pair per thread.
template <bool PEER, int NT, class Smem>
__device__ __forceinline__ void dense3_verify(const SearchArgs &a, const PeerView *pv, Smem &s, int kn, int nf) {
  const int total = kn * nf;
  for (int p = threadIdx.x; p < total; p += NT) {
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

// CLS: class list (4: queries up to KCAP k-mers staged at once; 5: longer ones; 6: the longest, chunked; 7: the
// queries whose repeated subjects overflowed H in launches 4 and 5);
// NW warps per CTA, MINB CTAs per SM
template <bool PEER, int KCAP, int EH, int CLS, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) k_search_f(SearchArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  constexpr int PH = PEER ? F_PH_PEER : F_PH_LOCAL;  // windows in flight per warp: one being processed + PH - 1 ahead
  using Smem = Dense3Smem<KCAP, EH, NW, PH>;
  constexpr int NT = NW * 32;
  constexpr int E_H = EH;
  Smem &s = *reinterpret_cast<Smem *>(dsm);
  const uint32_t mapw = CLS == 4 ? a.e_mapw_small : (CLS == 5 ? a.e_mapw_large : a.e_mapw_xl);  // 32-bit words per warp map (2^n)
  const int lgw = 31 - __clz(mapw);
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;
  const int w = tid >> 5;
  uint32_t *mm = reinterpret_cast<uint32_t *>(dsm + ((sizeof(Smem) + 15) & ~(size_t)15)) + (size_t)w * mapw;
  const PeerView *pv = nullptr;
  if constexpr (PEER) {
    __shared__ PeerView s_peer;
    load_peer_view(&s_peer, a.peer, tid, NT);
    pv = &s_peer;
  }
  for (int i = tid; i < 256; i += NT) s.lut[i] = (uint8_t)aa_code(i);
  __syncthreads();
  const SmemHashT<true> hv{s.hkeys, s.hcnt2, (uint32_t)E_H - 1u, 32 - ilog2_c(E_H)};
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
    const int w_act = (int)((kmin - 1u) / 2u) < NW ? (int)((kmin - 1u) / 2u) : NW;
    const uint32_t thr = kmin - 2u * (uint32_t)w_act;  // >= 1
    const int nchunks = (K + KCAP - 1) / KCAP;
    {
      uint4 *hk = reinterpret_cast<uint4 *>(s.hkeys);
      uint4 *hc = reinterpret_cast<uint4 *>(s.hcnt2);
      const uint4 E = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY), Z = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < E_H / 4; i += NT) hk[i] = E;
      for (int i = tid; i < E_H / 8; i += NT) hc[i] = Z;
      if (tid == 0) {
        s.bloom = 0ull;
        s.pflags = 0;
        s.nfinal = 0;
      }
      // the warp's own map (the first probes of the chunk are issued right after)
      uint4 *mv = reinterpret_cast<uint4 *>(mm);
      for (uint32_t i = lane; i < mapw / 4; i += 32) mv[i] = Z;
    }
    unsigned long long q_incr = 0;
    // ---- pass 1 ----
    HotState hot{F_NONE, F_NONE, 0u, 0u};
    for (int c = 0; c < nchunks; ++c) {
      unsigned long long tot = 0;
      const int kn = dense3_load_chunk<PEER, NT>(a, pv, s, KCAP, b, len, K, c, res_end, tot);
      q_incr += tot;
      if (w < w_act) {
        const int nx = (int)(s.nd < s.nx_end ? s.nd : s.nx_end);
        dense3_pass1<PEER, PH>(a, pv, s.ent, kn, kn + nx, mm, lgw, w, w_act, &s.u.ring[w][0][0], hv, &s.pflags, hot);
        if (s.nlq) dense3_pass1_long<PEER>(a, pv, s, mm, lgw, w, w_act, hv, &s.pflags, hot);
      }
    }
    // the occurrences counted in registers join the pushes of their subject (it is in H: it was pushed)
    if (hot.id0 != F_NONE) {
      uint32_t n0 = hot.n0, n1 = hot.n1;
      for (int o = 16; o > 0; o >>= 1) {
        n0 += __shfl_down_sync(0xFFFFFFFFu, n0, o);
        n1 += __shfl_down_sync(0xFFFFFFFFu, n1, o);
      }
      n1 = __shfl_sync(0xFFFFFFFFu, n1, 0);
      if (lane < 2) {
        const uint32_t id = lane == 0 ? hot.id0 : hot.id1;
        const uint32_t n = lane == 0 ? n0 : n1;
        if (id != F_NONE && n) {
          uint32_t slot = hv.home(id);
#pragma unroll 1
          for (int probe = 0; probe < E_H; ++probe) {
            if (s.hkeys[slot] == id) {
              hv.add(slot, n);
              break;
            }
            slot = (slot + 1) & hv.mask;
          }
        }
      }
    }
    __syncthreads();
    if (s.pflags) {
      // more repeated subjects than H holds (a query from a large family): the launch with the largest H takes
      // it (list 7); if that one overflows too, class G counts the query exactly in global memory (list 3)
      if (tid == 0) {
        constexpr int TO = CLS <= 5 ? 7 : 3;
        const uint32_t slot = atomicAdd(&a.list_count[TO], 1u);
        a.lists[(size_t)TO * a.nq + slot] = q;
      }
      continue;
    }
    // ---- sweep: final candidates ----
    if (tid == 0) {
      s.u.e.ss.ncand = 0;  // (the selection scratch shares its space with the windows of pass 1)
      s.u.e.ss.flags = 0;
    }
    for (int base = 0; base < E_H; base += NT) {
      const uint32_t slot = base + tid;
      const uint32_t key = s.hkeys[slot];
      const bool isfin = key != EMPTY && hv.count_at(slot) >= thr;
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, isfin);
      if (lane == 0) s.fin[slot >> 5] = bal;
      if (isfin) {
        atomicOr(&s.bloom, 1ull << ((key * 0x9E3779B1u) >> 26));
        const uint32_t fi = atomicAdd(&s.nfinal, 1u);
        if (fi < (uint32_t)F_NF) {
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
      for (int i = tid; i < E_H / 8; i += NT) hc[i] = Z;
    }
    __syncthreads();
    my_incr += q_incr;
    if (tid == 0) my_lookups += (unsigned long long)K;
    if (s.nfinal == 0) continue;  // nothing can reach kmin: no hits (n_hits[q] was zeroed by k_classify)
    // ---- pass 2: exact counts of the final candidates ----
    const unsigned long long bloom = s.bloom;
    const int nf = (int)s.nfinal;
    const bool by_search = nf <= F_NF && a.lists_sorted;
    for (int c = 0; c < nchunks; ++c) {
      int kn = K < KCAP ? K : KCAP;
      if (nchunks > 1) {
        unsigned long long tot = 0;
        kn = dense3_load_chunk<PEER, NT>(a, pv, s, KCAP, b, len, K, c, res_end, tot);
      }
      if (by_search) dense3_verify<PEER, NT>(a, pv, s, kn, nf);
      else dense3_pass2<PEER, NW>(a, pv, s, kn, w, hv, bloom);
    }
    __syncthreads();
    if (by_search) {
      if (tid < nf) hv.add(s.fslot[tid], s.fcnt[tid]);  // (the counts of H were zeroed after the sweep)
      __syncthreads();
    }
    for (int base = 0; base < E_H; base += NT) {
      const uint32_t slot = base + tid;
      if (((s.fin[slot >> 5] >> (slot & 31u)) & 1u) && hv.count_at(slot) >= kmin)
        s.u.e.cand[atomicAdd(&s.u.e.ss.ncand, 1u)] = (uint16_t)slot;
    }
    __syncthreads();
    const uint32_t c = s.u.e.ss.ncand;
    select_and_emit<NT>(a, q, hv, [&](uint32_t i) -> uint32_t { return s.u.e.cand[i]; }, c, s.u.e.ss);
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
