// search_common.cuh — device code shared by the search kernels (search.cu) and the key-range
// shard kernels (shard.cu): histograms, the per-round consume logic, top-N selection.
#pragma once
#include <type_traits>

#include "internal.cuh"

namespace kaamer {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
enum { CNT_POOL = 0, CNT_LOOKUPS = 1, CNT_INCR = 2, CNT_STATUS = 3, CNT_CLS_LOOKUPS = 4, CNT_CLS_INCR = 8, CNT_GNEED = 12, CNT_N = 16 };
enum { ST_POOL_OVERFLOW = 1, ST_GHASH_OVERFLOW = 2 };

// size classes by SizeInKmer
constexpr int W_H = 512, W_MAXK = 512, W_WARPS = 8, W_CAND = 64;
constexpr int M_THREADS = 256, M_H = 4096, M_MAXK = 2048, M_CTAS = 5;
constexpr int G_THREADS = 512;
constexpr int FAST_C = 64;     // candidates ranked by counting below this, bitonic sort above
constexpr int BIG_LIST = 32;   // posting lists at least this long are walked by the whole warp
constexpr int MAX_PROBE = 96;  // linear-probe budget before a histogram is declared full
constexpr int N_LISTS = 8;     // class lists: W, M, G, hand-offs to G, D short / long / longest queries, D hand-offs

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
  uint32_t *kmin;
  uint64_t *pool;
  uint64_t pool_cap;
  unsigned long long *counters;
  uint32_t *lists;       // [N_LISTS][nq]: W, M, G, hand-offs to G (from M or D), D short, D long
  uint32_t *list_count;  // [2 * N_LISTS]: list sizes, then work cursors (dynamic scheduling)
  uint32_t *ghash;       // class G scratch: per CTA [keys HG][cnt HG][cand HG]
  uint32_t ghash_slots;  // HG (power of two)
  // nucleotide / reads mode (search_nucleotide.go:76-124): queries are ORFs, the candidate
  // threshold is the gate `Hits[0].Kmatch >= MinKMatch` (:116) — FilterResults runs later with
  // the SizeInKmer that SetBestStartCodon leaves (finish.cu) — and any0[q] records whether a
  // hit tied with the best one, other than the first of them, matches at query position 0
  // (the only thing SetBestStartCodon reads from later tied hits, dna.go:224-237)
  int nt_mode;
  uint8_t *any0;
  int g_list;  // class list processed by k_search_g (2: classified, 3: hand-offs from class M)
  const PeerView *peer;  // mode P (kernels instantiated with PEER = true): shards of all ranks
  const uint32_t *filter;  // folded presence bits of the handle's own keys (L2-resident), or nullptr
  // size-class limits of this database (search.cu class_limits): queries up to w_maxk k-mers go to class
  // W, up to m_maxk to class M — at most W_MAXK / M_MAXK, less when the database is so dense that a
  // query of that size would overflow the class's histogram and be searched twice
  int w_maxk, m_maxk;
  // dense database (search.cu class_limits): every query goes to class D (search_dense.cuh), or to class G
  // when its threshold is too small for D's filter; d_mapb = bytes per byte map of class D
  int dense;             // 0: classes W / M / G; 1: class D first design (A/B only); 2: class D (search_dense2.cuh)
  uint32_t d_mapb;
  int e_kcap, e_kcap_l;  // class D launches by query length: up to e_kcap k-mers, up to e_kcap_l, longer
  uint32_t e_mapw_small, e_mapw_large, e_mapw_xl;  // words per bit map per warp of the three launches
  int lists_sorted;      // every posting list is strictly descending (builders, checked views): class D may
                         // verify its final candidates by binary search
};

// Copy src[0, len) into shared memory with aligned 16-byte loads (one request per 16 residues:
// coalesced into full lines from HBM and into large read requests when the source is pinned
// host memory read in place over PCIe).  dst must be 16-byte aligned with room for len + 31
// bytes; returns `head`: the sequence starts at dst + head.  Bytes outside [src, src_end) are
// never dereferenced past src_end (the chunk that would cross it is read bytewise).
template <int NT>
__device__ __forceinline__ int stage_bytes(uint8_t *dst, const uint8_t *src, int len, const uint8_t *src_end,
                                           int tid) {
  const uintptr_t base = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15;
  const int head = (int)(reinterpret_cast<uintptr_t>(src) - base);
  const int nchunks = (head + len + 15) >> 4;
  const uint4 *g = reinterpret_cast<const uint4 *>(base);
  for (int c = tid; c < nchunks; c += NT) {
    if (reinterpret_cast<const uint8_t *>(g + c + 1) <= src_end) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "l"(g + c));
      reinterpret_cast<uint4 *>(dst)[c] = v;
    } else {
      const uint8_t *p = reinterpret_cast<const uint8_t *>(g + c);
      for (int k = 0; k < 16; ++k)
        if (p + k < src_end && p + k >= src) dst[c * 16 + k] = p[k];
    }
  }
  return head;
}

__device__ __forceinline__ uint64_t ldg_entry(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// ---- table probe / posting access -----------------------------------------------------------
// PEER = false: the handle's own table over [d_lo, d_hi).  PEER = true (mode P): the dense code
// space is tiled by the shards of pv (a shared-memory copy of SearchArgs::peer); the owner is found
// with MAX_PEER_SHARDS-1 compares and the entry is read from its HBM — local or through NVLink.
// A multi-posting entry gets its shard folded into the value so that post_ptr finds the list.
//
// Every probe is preceded by a presence-filter lookup: the folded 64 MB bitmap of the handle's own
// keys (L2-resident, internal.cuh) for the local table, the exact 227 MB replica for remote shards.
// U probes are issued in two phases — U filter words, then the surviving table entries — so that the
// loads of each phase are in flight together.
template <bool PEER, int U>
__device__ __forceinline__ void probe_entries(const SearchArgs &a, const PeerView *pv, const uint32_t (&d)[U],
                                              const bool (&ok)[U], uint64_t (&e)[U]) {
  uint32_t w[U];
  const uint64_t *p[U];
  uint32_t sh[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    bool valid = ok[u];
    const uint32_t *f;
    uint32_t fi;
    if constexpr (!PEER) {
      (void)pv;
      valid = valid && d[u] >= a.d_lo && d[u] < a.d_hi;
      p[u] = a.table + (d[u] - a.d_lo);
      f = a.filter;
      fi = d[u] & FILTER_MASK;
      sh[u] = 0;
    } else if (pv->full_table != nullptr) {
      // replicated table: local probe, the entry already carries the shard of its posting list
      p[u] = pv->full_table + d[u];
      f = nullptr;
      fi = 0;
      sh[u] = 0;
    } else {
      uint32_t s = 0;
#pragma unroll
      for (int i = 1; i < MAX_PEER_SHARDS; ++i) s += d[u] >= pv->fence[i] ? 1u : 0u;
      p[u] = pv->table[s] + (d[u] - pv->fence[s]);
      const bool mine = (int32_t)s == pv->self;
      f = mine ? a.filter : pv->presence;
      fi = mine ? (d[u] & FILTER_MASK) : d[u];
      sh[u] = s;
    }
    w[u] = valid ? 0xFFFFFFFFu : 0u;
    if (valid && f != nullptr) w[u] = __ldg(f + (fi >> 5));
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    e[u] = 0ull;
    if ((w[u] >> (d[u] & 31u)) & 1u) {
      e[u] = ldg_entry(p[u]);
      if constexpr (PEER) {
        if ((e[u] >> ENTRY_VALUE_BITS) >= 2ull) e[u] |= (uint64_t)sh[u] << PEER_SHARD_SHIFT;
      }
    }
  }
}
template <bool PEER>
__device__ __forceinline__ uint64_t probe_entry(const SearchArgs &a, const PeerView *pv, uint32_t d) {
  const uint32_t dd[1] = {d};
  const bool ok[1] = {true};
  uint64_t e[1];
  probe_entries<PEER, 1>(a, pv, dd, ok, e);
  return e[0];
}
template <bool PEER>
__device__ __forceinline__ const uint32_t *post_ptr(const SearchArgs &a, const PeerView *pv, uint64_t val) {
  if constexpr (!PEER) {
    (void)pv;
    return a.postings + val;
  } else {
    return pv->postings[val >> PEER_SHARD_SHIFT] + (val & PEER_LOCAL_MASK);
  }
}
// CTA-wide copy of the peer view into shared memory (callers barrier afterwards)
__device__ __forceinline__ void load_peer_view(PeerView *dst, const PeerView *src, int tid, int nthreads) {
  const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
  uint32_t *d = reinterpret_cast<uint32_t *>(dst);
  for (int i = tid; i < (int)(sizeof(PeerView) / 4); i += nthreads) d[i] = s[i];
}

// ---- histograms -------------------------------------------------------------------------
// add(id) returns the count BEFORE the increment, or 0xFFFFFFFF when the table is full.
// Shared-memory flavour: keys u32, counts u16 packed two per word (counts <= SizeInKmer <= 2048).
template <bool EVENTS>
struct SmemHashT {
  static constexpr bool kWarp = false;
  static constexpr bool kEvents = EVENTS;  // false: counts are fire-and-forget, candidates swept at the end
  uint32_t *keys;
  uint32_t *cnt2;  // [slots/2]
  uint32_t mask;
  int shift;  // 32 - log2(slots)
  __device__ __forceinline__ uint32_t home(uint32_t id) const { return (id * 2654435761u) >> shift; }
  // claim-or-match: returns the previous key of the slot (EMPTY or id => the slot is ours).
  // Plain load first: shared-memory atomics are serialised per lane on the SM's atomic unit
  // (~2 cycles/lane for ADD, twice that for CAS), a repeated subject must not pay the CAS.
  __device__ __forceinline__ uint32_t cas(uint32_t slot, uint32_t id) const {
    uint32_t cur = *(volatile uint32_t *)(keys + slot);
    if (cur == EMPTY) cur = atomicCAS(keys + slot, EMPTY, id);
    return cur;
  }
  __device__ __forceinline__ uint32_t inc(uint32_t slot) const {  // returns the count before
    uint32_t sh = (slot & 1u) * 16u;
    uint32_t old = atomicAdd(cnt2 + (slot >> 1), 1u << sh);
    return (old >> sh) & 0xFFFFu;
  }
  __device__ __forceinline__ void add(uint32_t slot, uint32_t c) const { atomicAdd(cnt2 + (slot >> 1), c << ((slot & 1u) * 16u)); }
  __device__ __forceinline__ uint32_t key_at(uint32_t slot) const { return keys[slot]; }
  __device__ __forceinline__ uint32_t count_at(uint32_t slot) const {
    return (cnt2[slot >> 1] >> ((slot & 1u) * 16u)) & 0xFFFFu;
  }
  static constexpr int kMaxProbe = MAX_PROBE;
};
using SmemHash = SmemHashT<true>;
// Global-memory flavour (class G): keys u32, counts u32.
struct GmemHash {
  static constexpr bool kWarp = false;
  static constexpr bool kEvents = true;
  uint32_t *keys;
  uint32_t *cnt;
  uint32_t mask;
  int shift;
  __device__ __forceinline__ uint32_t home(uint32_t id) const { return (id * 2654435761u) >> shift; }
  __device__ __forceinline__ uint32_t cas(uint32_t slot, uint32_t id) const { return atomicCAS(keys + slot, EMPTY, id); }
  __device__ __forceinline__ uint32_t inc(uint32_t slot) const { return atomicAdd(cnt + slot, 1u); }
  __device__ __forceinline__ void add(uint32_t slot, uint32_t c) const { atomicAdd(cnt + slot, c); }
  __device__ __forceinline__ uint32_t key_at(uint32_t slot) const { return keys[slot]; }
  __device__ __forceinline__ uint32_t count_at(uint32_t slot) const { return cnt[slot]; }
  static constexpr int kMaxProbe = 4 * MAX_PROBE;
};

// candidate list shared by the lanes/threads that work on one query
struct CandList {
  uint32_t *ncand;  // smem counter
  uint32_t *flags;  // smem: bit0 histogram full, bit1 candidate list full
  uint16_t *slots16;
  uint32_t *slots32;
  uint32_t cap;
};

__device__ __forceinline__ void push_candidate(const CandList &cl, uint32_t slot) {
  uint32_t i = atomicAdd(cl.ncand, 1u);
  if (i < cl.cap) {
    if (cl.slots16) cl.slots16[i] = (uint16_t)slot;
    else cl.slots32[i] = slot;
  } else {
    atomicOr(cl.flags, 2u);
  }
}

// One increment + candidate bookkeeping (general path: linear probing from `slot`).  The
// subject becomes a candidate exactly when its count reaches kmin (counts only grow), so no
// scan of the histogram is needed afterwards.
template <class Hash>
__device__ __noinline__ void count_subject_from(const Hash hv, uint32_t id, uint32_t slot, uint32_t kmin,
                                                const CandList cl) {
#pragma unroll 1
  for (int probe = 0; probe < Hash::kMaxProbe; ++probe) {
    uint32_t cur = hv.cas(slot, id);
    if (cur == EMPTY || cur == id) {
      if constexpr (Hash::kEvents) {
        if (hv.inc(slot) + 1 == kmin) push_candidate(cl, slot);
      } else {
        hv.add(slot, 1u);
      }
      return;
    }
    slot = (slot + 1) & hv.mask;
  }
  atomicOr(cl.flags, 1u);  // histogram full
}
template <class Hash>
__device__ __forceinline__ void count_subject(const Hash &hv, uint32_t id, uint32_t kmin, const CandList &cl) {
  const uint32_t slot = hv.home(id);
  const uint32_t cur = hv.cas(slot, id);
  if (cur == EMPTY || cur == id) {
    if constexpr (Hash::kEvents) {
      if (hv.inc(slot) + 1 == kmin) push_candidate(cl, slot);
    } else {
      hv.add(slot, 1u);
    }
  } else {
    count_subject_from(hv, id, (slot + 1) & hv.mask, kmin, cl);
  }
}

// ---- warp-private histogram (class W) ----------------------------------------------------
// The table belongs to one warp: no block barriers, 16-bit counts packed two per word.
constexpr int ilog2_c(int x) { return x <= 1 ? 0 : 1 + ilog2_c(x / 2); }
// EVENTS = true: a subject is pushed on the candidate list the moment its count reaches kmin (the
// increment has to return the old count).  EVENTS = false: increments are fire-and-forget reductions
// (no wait on the shared-memory atomic unit) and the caller sweeps the H slots once at the end of the
// query (warp_collect_candidates) — cheaper for the 512-slot table of class W.
template <int H, bool EVENTS = true>
struct WarpHashT {
  uint32_t *keys;   // [H]
  uint16_t *cnt;    // [H]
  static constexpr bool kWarp = true;
  static constexpr bool kEvents = EVENTS;
  static constexpr uint32_t mask = H - 1;
  static constexpr int kMaxProbe = MAX_PROBE;
  __device__ __forceinline__ uint32_t home(uint32_t id) const { return (id * 2654435761u) >> (32 - ilog2_c(H)); }
  __device__ __forceinline__ uint32_t key_at(uint32_t slot) const { return keys[slot]; }
  __device__ __forceinline__ uint32_t count_at(uint32_t slot) const { return cnt[slot]; }
};

// count of a subject after the histogram is complete (0 if absent)
template <class Hash>
__device__ __forceinline__ uint32_t hist_count(const Hash &hv, uint32_t id) {
  uint32_t slot = hv.home(id);
#pragma unroll 1
  for (int probe = 0; probe < Hash::kMaxProbe; ++probe) {
    const uint32_t k = hv.key_at(slot);
    if (k == id) return hv.count_at(slot);
    if (k == EMPTY) return 0;
    slot = (slot + 1) & hv.mask;
  }
  return 0;
}

// nucleotide mode, one warp: does a subject tied with the best hit (count T), other than the
// best hit itself, hold the query's first k-mer (dense code d0)?
template <bool PEER = false, class Hash>
__device__ __forceinline__ bool warp_any0(const SearchArgs &a, const PeerView *pv, const Hash &hv, uint32_t d0,
                                       uint32_t best_id, uint32_t T) {
  const unsigned lane = threadIdx.x & 31;
  bool any = false;
  const uint64_t e = probe_entry<PEER>(a, pv, d0);
  const uint32_t cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);
  const uint64_t val = e & ENTRY_VALUE_MASK;
  if (cnt == 1) {
    const uint32_t id = (uint32_t)val;
    any = id != best_id && hist_count(hv, id) == T;
  } else if (cnt >= 2) {
    const uint32_t *pl = post_ptr<PEER>(a, pv, val);
    for (uint32_t i = lane; i < cnt; i += 32) {
      const uint32_t id = __ldg(pl + i);
      if (id != best_id && hist_count(hv, id) == T) any = true;
    }
  }
  return __any_sync(0xFFFFFFFFu, any);
}

// all 32 lanes call this together.  Every valid lane walks the probe sequence of its id on its own:
// a slot is claimed by CAS only when it is seen EMPTY (once per distinct subject of the query) and
// the count is one shared-memory atomic add on the 16-bit half of its word.  (An earlier version
// merged duplicate ids of the 32 lanes with __match_any_sync and let one leader do a plain
// read-modify-write: MATCH.ANY alone was 19 % of the kernel's stall samples.)
template <int H, bool EVENTS>
__device__ __forceinline__ void warp_count(const WarpHashT<H, EVENTS> &hv, bool valid, uint32_t id, uint32_t kmin,
                                           const CandList &cl) {
  if (valid) {
    uint32_t slot = hv.home(id);
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>(hv.cnt);
    int probe = 0;
#pragma unroll 1
    for (; probe < MAX_PROBE; ++probe) {
      uint32_t key = *(volatile uint32_t *)(hv.keys + slot);
      if (key == EMPTY) {
        key = atomicCAS(hv.keys + slot, EMPTY, id);
        if (key == EMPTY) key = id;
      }
      if (key == id) {
        const uint32_t sh = (slot & 1u) * 16u;
        if constexpr (EVENTS) {
          const uint32_t old = (atomicAdd(cnt32 + (slot >> 1), 1u << sh) >> sh) & 0xFFFFu;
          if (old + 1 == kmin) push_candidate(cl, slot);
        } else {
          atomicAdd(cnt32 + (slot >> 1), 1u << sh);  // result unused: RED.ADD, nothing to wait for
        }
        break;
      }
      slot = (slot + 1) & (H - 1);
    }
    if (probe == MAX_PROBE) atomicOr(cl.flags, 1u);
  }
  __syncwarp();
}

// EVENTS = false: one sweep over the H slots at the end of a query; slots whose count reached kmin
// are appended to the candidate list in slot order (warp-aggregated).  Returns the candidate count
// (it may exceed cl.cap: the caller hands the query to the next class).
template <int H>
__device__ __forceinline__ uint32_t warp_collect_candidates(const WarpHashT<H, false> &hv, uint32_t kmin,
                                                            const CandList &cl) {
  const unsigned lane = threadIdx.x & 31;
  uint32_t c = 0;
#pragma unroll 1
  for (int base = 0; base < H; base += 32 * 8) {
    // 8 consecutive slots per lane: one 16-byte load of counts
    const int s0 = base + (int)lane * 8;
    const uint4 cv = *reinterpret_cast<const uint4 *>(hv.cnt + s0);
    const uint32_t w[4] = {cv.x, cv.y, cv.z, cv.w};
    unsigned hit = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t cnt = (w[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
      if (cnt >= kmin && cnt != 0u) hit |= 1u << i;
    }
    // exclusive prefix of popc(hit) over the lanes
    const uint32_t n = __popc(hit);
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (unsigned)o) incl += t;
    }
    uint32_t pos = c + incl - n;
    while (hit) {
      const int i = __ffs(hit) - 1;
      hit &= hit - 1;
      if (pos < cl.cap) cl.slots16[pos] = (uint16_t)(s0 + i);
      ++pos;
    }
    c += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  __syncwarp();
  return c;
}

// ---- one warp-round of lookups ------------------------------------------------------------
// Every lane holds up to U table entries.  Singletons are counted directly; short posting
// lists (2..BIG_LIST-1) of the whole warp are flattened (warp scan + search by shuffles) so
// that all 32 lanes fetch and count postings together; long lists are walked cooperatively.
template <int U, bool PEER = false, class Hash>
__device__ __forceinline__ void warp_consume(const SearchArgs &a, const uint64_t (&ent)[U], const Hash &hv,
                                             uint32_t kmin, const CandList &cl, unsigned long long &q_incr,
                                             const PeerView *pv = nullptr) {
  static_assert(U <= 4, "vhi packs one byte per entry");
  const unsigned lane = threadIdx.x & 31;
  uint32_t m[U];  // postings of this lane's short multi lists
  uint32_t vlo[U];
  uint32_t vhi = 0;  // the 5 high bits of each value, one byte each (U <= 4)
  uint32_t mt = 0;
  bool any_big = false;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t cnt = (uint32_t)(ent[u] >> ENTRY_VALUE_BITS);
    const uint64_t val = ent[u] & ENTRY_VALUE_MASK;
    q_incr += cnt;
    vlo[u] = (uint32_t)val;
    vhi |= (uint32_t)(val >> 32) << (8 * u);
    m[u] = (cnt >= 2 && cnt < BIG_LIST) ? cnt : 0u;
    mt += m[u];
    any_big |= cnt >= BIG_LIST;
  }
  // (2a) short lists, flattened across the warp: posting j of the warp's lists belongs to the lane
  // `owner` whose inclusive prefix first exceeds j.  The FIRST 32 postings are fetched now and counted
  // after the singletons: the table is probed at the HBM random-access ceiling, so a dependent load
  // waits in the same queue as the probes (~5 us) — ncu showed 22 % of the class-W stall samples on
  // the first use of these ids when they were fetched right before being counted.
  uint32_t incl = mt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (unsigned)o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  const uint32_t excl = incl - mt;
  auto fetch_posting = [&](uint32_t j) -> uint32_t {  // all lanes call it together
    uint32_t owner = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, (owner + step - 1) & 31);
      if (v <= j) owner += step;
    }
    owner &= 31;
    uint32_t r = j - __shfl_sync(0xFFFFFFFFu, excl, owner);
    const uint32_t ohi = __shfl_sync(0xFFFFFFFFu, vhi, owner);
    uint32_t sel_lo = 0, sel_hi = 0;
    bool found = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t mu = __shfl_sync(0xFFFFFFFFu, m[u], owner);
      uint32_t lo = __shfl_sync(0xFFFFFFFFu, vlo[u], owner);
      if (!found) {
        if (r < mu) {
          sel_lo = lo;
          sel_hi = (ohi >> (8 * u)) & 0xFFu;
          found = true;
        } else {
          r -= mu;
        }
      }
    }
    uint32_t pid = 0;
    if (j < total) pid = __ldg(post_ptr<PEER>(a, pv, ((uint64_t)sel_hi << 32) | sel_lo) + r);
    return pid;
  };
  uint32_t pid0 = 0;
  if (total) pid0 = fetch_posting(lane);  // (total is warp-uniform)
  // (1) singletons
  if constexpr (Hash::kWarp) {
#pragma unroll
    for (int u = 0; u < U; ++u) warp_count(hv, (uint32_t)(ent[u] >> ENTRY_VALUE_BITS) == 1u, vlo[u], kmin, cl);
  } else {
    // the first probe of all U entries is issued back to back (independent shared-memory
    // operations in flight), collisions fall back to the probing loop
    uint32_t slot[U], cur[U];
    bool act[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      act[u] = (uint32_t)(ent[u] >> ENTRY_VALUE_BITS) == 1u;
      slot[u] = hv.home(vlo[u]);
      cur[u] = act[u] ? hv.cas(slot[u], vlo[u]) : 0u;
    }
    uint32_t old[U];
    bool got[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      got[u] = act[u] && (cur[u] == EMPTY || cur[u] == vlo[u]);
      old[u] = 0u;
      if constexpr (Hash::kEvents) {
        if (got[u]) old[u] = hv.inc(slot[u]);
      } else {
        if (got[u]) hv.add(slot[u], 1u);
      }
      act[u] = act[u] && !got[u];  // still pending
    }
    if constexpr (Hash::kEvents) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (got[u] && old[u] + 1 == kmin) push_candidate(cl, slot[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (act[u]) count_subject_from(hv, vlo[u], (slot[u] + 1) & hv.mask, kmin, cl);
  }
  // (2b) count the short-list postings: the prefetched batch, then (rarely) the rest
#pragma unroll 1
  for (uint32_t jb = 0; jb < total; jb += 32) {
    const uint32_t j = jb + lane;
    const uint32_t pid = jb == 0 ? pid0 : fetch_posting(j);
    if constexpr (Hash::kWarp) {
      warp_count(hv, j < total, pid, kmin, cl);
    } else {
      if (j < total) count_subject(hv, pid, kmin, cl);
    }
  }
  // (3) long lists: the whole warp walks each of them (coalesced)
  if (__any_sync(0xFFFFFFFFu, any_big)) {
#pragma unroll 1
    for (int u = 0; u < U; ++u) {
      uint64_t e = ent[0];
#pragma unroll
      for (int t = 1; t < U; ++t)
        if (u == t) e = ent[t];
      const uint32_t cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);
      unsigned big = __ballot_sync(0xFFFFFFFFu, cnt >= BIG_LIST);
      while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t bc = __shfl_sync(0xFFFFFFFFu, cnt, src);
        const uint64_t bv = __shfl_sync(0xFFFFFFFFu, e & ENTRY_VALUE_MASK, src);
        const uint32_t *pl = post_ptr<PEER>(a, pv, bv);
        if constexpr (Hash::kWarp) {
          for (uint32_t ib = 0; ib < bc; ib += 32) {
            const uint32_t i = ib + lane;
            warp_count(hv, i < bc, i < bc ? __ldg(pl + i) : 0u, kmin, cl);
          }
        } else {
          for (uint32_t i = lane; i < bc; i += 32) {
            if (*(volatile uint32_t *)cl.flags & 1u) break;
            count_subject(hv, __ldg(pl + i), kmin, cl);
          }
        }
      }
    }
  }
}

// packed per-position code: bits 0..8 pair code p'(c_i, c_i+1), bits 9..13 single code s(c_i)
__device__ __forceinline__ uint32_t packed_code(uint32_t c0, uint32_t c1) {
  return pair_dense(c0, c1) | (single_dense(c0) << 9);
}
__device__ __forceinline__ uint32_t dense_from_packed(uint32_t w0, uint32_t w2, uint32_t w4, uint32_t w6) {
  return (((w0 & 511u) * PAIR_RADIX + (w2 & 511u)) * PAIR_RADIX + (w4 & 511u)) * 21u + (w6 >> 9);
}

// composite sort key: ascending order == (Kmatch desc, subject id asc)
__device__ __forceinline__ uint64_t composite(uint32_t id, uint32_t cnt) {
  return ((uint64_t)(0xFFFFFFFFu - cnt) << 32) | id;
}
__device__ __forceinline__ uint64_t decomposite(uint64_t c) {
  uint32_t cnt = 0xFFFFFFFFu - (uint32_t)(c >> 32);
  return ((uint64_t)cnt << 32) | (uint32_t)c;  // pool format: subject | kmatch << 32
}

// ---- classes M and G: one CTA per query -----------------------------------------------------
struct SelectScratch {
  uint32_t hist[256];
  uint32_t ncand, flags, nout, remaining;
  unsigned long long base;
  unsigned long long prefix;
};

// Candidate i lives in histogram slot cand(i).  Emits the top-N candidates into the pool in
// rank order and records n_hits / hit_base.
template <int THREADS, class Hash, class CandAt>
__device__ void select_and_emit(const SearchArgs &a, uint32_t q, const Hash &hv, CandAt cand, uint32_t c,
                                SelectScratch &ss) {
  const int tid = threadIdx.x;
  const uint32_t N = a.max_results > 0 ? (uint32_t)a.max_results : 0u;
  const uint32_t nout = c < N ? c : N;
  if (nout == 0) return;  // n_hits / hit_base were zeroed by k_classify
  if (c <= FAST_C) {
    if (tid == 0) ss.base = atomicAdd(&a.counters[CNT_POOL], (unsigned long long)nout);
    __syncthreads();
    const unsigned long long base = ss.base;
    const bool fits = base + nout <= a.pool_cap;
    if (tid < (int)c) {
      uint32_t s = cand(tid);
      uint64_t me = composite(hv.key_at(s), hv.count_at(s));
      uint32_t rank = 0;
      for (uint32_t j = 0; j < c; ++j) {
        uint32_t sj = cand(j);
        rank += composite(hv.key_at(sj), hv.count_at(sj)) < me ? 1u : 0u;
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
        uint32_t s = cand(i);
        uint64_t k = composite(hv.key_at(s), hv.count_at(s));
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
    uint32_t s = cand(i);
    uint64_t k = composite(hv.key_at(s), hv.count_at(s));
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


}  // namespace kaamer
