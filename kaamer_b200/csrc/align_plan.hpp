// align_plan.hpp — host-side schedule of one kaamer_gpu_align call (pure C++: no device code, no CUDA calls).
//
// The (query, subject) pairs of a call become three kinds of work, each list ordered longest first (short tail,
// similar pairs share a CTA):
//   long pairs    one CTA each (k_sw_affine_cta), n x m >= BIG_CELLS;
//   packed jobs   two pairs of similar geometry per warp in int16x2 lanes (k_sw_affine_pk, align_packed.cuh);
//   single pairs  one warp each in 32-bit lanes (k_sw_affine): what does not fit a job.
// The items are walked longest first and cut into chunks whose traceback state fits the memory budget.
// tests/csrc/align_plan_host.cu checks the invariants of the schedule (every pair once, regions disjoint and
// inside the budget, job geometry covers both pairs) and times it; align.cu only launches what this returns.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/kaamer_gpu.h"
#include "align_packed.cuh"

namespace kaamer {

struct AlnPair {
  uint32_t out_index;  // position in the caller's pair list
  uint32_t q;          // query index
  uint32_t s;          // subject protein id
  uint32_t cw;         // columns per lane (4, 8, 12 or 16; packed jobs: 4 or 8, see pk_geo)
  uint64_t scratch;    // byte offset of the pair's traceback region
};

constexpr uint64_t BIG_CELLS = 2ull << 20;  // pairs at least this large get a whole CTA

// columns per lane of the warp-per-pair kernel: the cost of a pair is
// blocks x (rows + 31) x (per-step overhead + CW x per-cell work), ~40 and ~22.5 instructions
inline int choose_cw(uint64_t m) {
  int best = 4;
  double best_cost = 1e300;
  for (int cw = 4; cw <= 16; cw += 4) {
    const uint64_t bw = 32ull * cw, nblk = (m + bw - 1) / bw;
    const double cost = (double)nblk * (40.0 + 22.5 * cw);
    if (cost < best_cost) {
      best_cost = cost;
      best = cw;
    }
  }
  return best;
}

inline uint64_t pair_scratch_bytes(uint64_t n, uint64_t m, int cw, bool big) {
  const uint64_t bw = 32ull * cw;
  const uint64_t nblk = (m + bw - 1) / bw;
  // traceback bytes + boundary columns (one per block in the pipelined kernel, one in place otherwise)
  uint64_t b = nblk * (n + 31) * 32 * cw + (big ? nblk : 1) * 3 * 4 * n;
  return (b + 255) & ~255ull;
}

// estimated warp instructions of one pair in the one-warp-per-pair kernel: instructions per wavefront step counted
// in the SASS of dp_block<CW> (158 / 245 / 341 / 435 for 4 / 8 / 12 / 16 columns per lane)
inline double single_work(uint64_t n, uint64_t m, int cw) {
  const uint64_t bw = 32ull * cw, nblk = (m + bw - 1) / bw;
  return (double)nblk * (double)(n + 31) * (66.0 + 23.0 * cw);
}

// ---- packed jobs: two pairs per warp (k_sw_affine_pk) ---------------------------------------------
// Instructions per wavefront step of dp_block_packed<CW>, counted in the SASS: ~187 / ~305 for 4 / 8 columns per
// lane, i.e. ~50 + 32 per column for BOTH pairs (two single pairs: 2 x (66 + 23 per column)).  12 and 16 columns
// per lane were built and measured: 187 registers, two CTAs per SM, 50 ms against 36 ms for the C5 batch.
constexpr double PK_OVH = 50.0, PK_CELL = 32.0;
constexpr double PK_ACCEPT = 0.9;  // a job must cost less than this share of its two pairs run one by one

struct PackedConfig {
  bool on;
  int maxcw;          // 4 or 8 columns per lane
  uint64_t max_cells; // pairs below this many cells may be packed
  uint32_t max_min_dim;  // min(n, m) bound that keeps every DP value below PK_MAX_SCORE
};

struct PkJob {
  uint32_t a, b;  // pair indices
  int cw;
  uint32_t N, Mx;  // rows / columns of the job
  double work;
};

// Test / measurement hooks: KAAMER_ALIGN_PACKED=0 keeps every pair on the 32-bit kernels,
// KAAMER_ALIGN_PK_MAXCW=4|8 bounds the columns per lane, KAAMER_ALIGN_PK_CELLS=<cells> lets pairs of up
// to that many cells be packed (default: below the one-CTA-per-pair threshold).
inline PackedConfig packed_config(const kaamer_aln_model &model, bool zero_gap) {
  PackedConfig c{zero_gap, 8, BIG_CELLS, 0};
  if (const char *e = getenv("KAAMER_ALIGN_PACKED")) c.on = c.on && atoi(e) != 0;
  if (const char *e = getenv("KAAMER_ALIGN_PK_MAXCW")) {
    const int v = atoi(e);
    if (v == 4 || v == 8) c.maxcw = v;
  }
  if (const char *e = getenv("KAAMER_ALIGN_PK_CELLS")) {
    const long long v = atoll(e);
    if (v > 0) c.max_cells = (uint64_t)v;
  }
  int max_entry = 1;
  for (int i = 1; i < 26; ++i)
    for (int j = 1; j < 26; ++j) max_entry = model.matrix[i * 26 + j] > max_entry ? model.matrix[i * 26 + j] : max_entry;
  c.max_min_dim = (uint32_t)(PK_MAX_SCORE / max_entry);
  return c;
}

// instructions per row of a job of `cols` columns: its blocks (pk_geo) x (overhead + per-column work)
inline double packed_row_work(uint64_t cols, int cw) {
  const PkGeo g = pk_geo((int)cols, cw);
  return (double)g.nfull * (PK_OVH + PK_CELL * cw) + (g.tail_cw ? PK_OVH + PK_CELL * 4 : 0.0);
}

inline int choose_cw_pk(uint64_t m, int maxcw) {
  if (maxcw < 8) return 4;
  return packed_row_work(m, 8) <= packed_row_work(m, 4) ? 8 : 4;
}

// every block of a job sweeps all rows (+ 31 steps to fill and drain the wavefront)
inline double packed_work(uint64_t N, uint64_t Mx, int cw) {
  return (double)(N + 31) * packed_row_work(Mx, cw);
}

// Per-subject-length tables of the cost model (the planner runs once per call over 10^5 pairs: the divisions
// and the loops of choose_cw / pk_geo were 13 of its 16 ms): columns per lane and instructions per row of a
// single pair, and of a packed job with 8 (+ tail) or 4 columns per lane, padded columns of the job.
struct PlanTables {
  static constexpr uint32_t NS = 65536, NP = PK_MAX_DIM + 1;
  std::vector<uint8_t> cw_single, cw_pk8;
  std::vector<float> row_single, row_pk8, row_pk4;
  std::vector<uint16_t> padded_pk8, padded_pk4;
  PlanTables() : cw_single(NS), cw_pk8(NP), row_single(NS), row_pk8(NP), row_pk4(NP), padded_pk8(NP), padded_pk4(NP) {
    for (uint32_t m = 0; m < NS; ++m) {
      const int cw = choose_cw(m);
      const uint32_t bw = 32u * (uint32_t)cw, nblk = (m + bw - 1) / bw;
      cw_single[m] = (uint8_t)cw;
      row_single[m] = (float)(nblk * (66.0 + 23.0 * cw));
    }
    for (uint32_t m = 0; m < NP; ++m) {
      const int cw = choose_cw_pk(m, 8);
      cw_pk8[m] = (uint8_t)cw;
      row_pk8[m] = (float)packed_row_work(m, cw);
      row_pk4[m] = (float)packed_row_work(m, 4);
      padded_pk8[m] = (uint16_t)pk_padded_cols(m, cw);
      padded_pk4[m] = (uint16_t)pk_padded_cols(m, 4);
    }
  }
};
inline const PlanTables &plan_tables() {
  static const PlanTables t;
  return t;
}

// Pairs of similar geometry become jobs: the eligible pairs are sorted by (padded columns, rows) descending
// (two counting-sort passes) and neighbours are joined when the job is cheaper than the two pairs run singly.
// The jobs come out ordered by estimated work, largest first.
inline void plan_packed_jobs(const PackedConfig &pk, const std::vector<uint32_t> &dim_n, const std::vector<uint32_t> &dim_m,
                             const std::vector<uint64_t> &cost, std::vector<PkJob> &jobs, std::vector<uint8_t> &in_job) {
  const uint32_t n_pairs = (uint32_t)dim_n.size();
  const PlanTables &T = plan_tables();
  const bool wide = pk.maxcw >= 8;
  // (pair, key, subject length) travel through the sort together: the walk below then reads them in order
  std::vector<uint32_t> el, key, mm;
  el.reserve(n_pairs);
  key.reserve(n_pairs);
  mm.reserve(n_pairs);
  for (uint32_t i = 0; i < n_pairs; ++i) {
    const uint32_t n = dim_n[i], m = dim_m[i];
    if (n < 1 || m < 1 || n > (uint32_t)PK_MAX_DIM || m > (uint32_t)PK_MAX_DIM) continue;
    if ((n < m ? n : m) > pk.max_min_dim || cost[i] >= pk.max_cells) continue;
    const uint32_t padded = wide ? T.padded_pk8[m] : T.padded_pk4[m];  // <= 16384 + 255
    const uint32_t k = ((padded / 128u) << 14) | n;                    // 8 + 14 bits
    el.push_back(i);
    key.push_back(0x3FFFFFu - k);  // ascending sort of the complement = descending (padded, n)
    mm.push_back(m);
  }
  const uint32_t ne = (uint32_t)el.size();
  if (ne < 2) return;
  {
    std::vector<uint32_t> el2(ne), key2(ne), mm2(ne);
    for (int pass = 0; pass < 2; ++pass) {
      const int shift = pass * 11;
      uint32_t cnt[2049] = {0};
      for (uint32_t x = 0; x < ne; ++x) cnt[((key[x] >> shift) & 2047u) + 1]++;
      for (int b = 0; b < 2048; ++b) cnt[b + 1] += cnt[b];
      for (uint32_t x = 0; x < ne; ++x) {
        const uint32_t at = cnt[(key[x] >> shift) & 2047u]++;
        el2[at] = el[x];
        key2[at] = key[x];
        mm2[at] = mm[x];
      }
      el.swap(el2);
      key.swap(key2);
      mm.swap(mm2);
    }
  }
  std::vector<PkJob> raw;
  raw.reserve(ne / 2);
  double max_work = 1.0;
  for (uint32_t x = 0; x + 1 < ne;) {
    const uint32_t nA = (0x3FFFFFu - key[x]) & 0x3FFFu, nB = (0x3FFFFFu - key[x + 1]) & 0x3FFFu, mA = mm[x], mB = mm[x + 1];
    const uint32_t N = nA > nB ? nA : nB, Mx = mA > mB ? mA : mB;
    const double w = (double)(N + 31) * (wide ? T.row_pk8[Mx] : T.row_pk4[Mx]);  // = packed_work(N, Mx, cw)
    const double singly = (double)(nA + 31) * T.row_single[mA] + (double)(nB + 31) * T.row_single[mB];
    if (w <= PK_ACCEPT * singly) {
      const uint32_t A = el[x], B = el[x + 1];
      raw.push_back(PkJob{A, B, wide ? (int)T.cw_pk8[Mx] : 4, N, Mx, w});
      in_job[A] = in_job[B] = 1;
      max_work = w > max_work ? w : max_work;
      x += 2;
    } else {
      x += 1;
    }
  }
  // largest first (1024 buckets, as for the single pairs)
  constexpr int NB = 1024;
  std::vector<uint32_t> start(NB + 1, 0);
  std::vector<uint16_t> bucket(raw.size());
  const double scale = (NB - 1) / max_work;
  for (size_t x = 0; x < raw.size(); ++x) {
    int bk = NB - 1 - (int)(raw[x].work * scale);
    bk = bk < 0 ? 0 : (bk > NB - 1 ? NB - 1 : bk);
    bucket[x] = (uint16_t)bk;
    start[bk + 1]++;
  }
  for (int b = 0; b < NB; ++b) start[b + 1] += start[b];
  jobs.resize(raw.size());
  for (size_t x = 0; x < raw.size(); ++x) jobs[start[bucket[x]]++] = raw[x];
}

struct AlnChunk {
  uint32_t big_end, job_end, small_end;  // ends of the chunk's ranges in the three lists
};

struct AlnPlan {
  std::vector<AlnPair> big_pairs, small_pairs, job_pairs;  // device order: [long | single | jobs (two entries each)]
  std::vector<AlnChunk> chunks;
  uint64_t max_used = 0;       // largest traceback footprint of a chunk
  int pk_maxcw_used = 0;       // widest packed job (profile columns of k_sw_affine_pk)
  int64_t too_large_pair = -1; // a pair whose traceback state alone exceeds the budget
  uint64_t too_large_bytes = 0;
  // every scheduled entry carries the query / subject of its pair (test aid)
  bool pair_q_ok(const std::vector<uint32_t> &pq, const std::vector<uint32_t> &ps) const {
    for (const std::vector<AlnPair> *l : {&big_pairs, &small_pairs, &job_pairs})
      for (const AlnPair &p : *l)
        if (p.q != pq[p.out_index] || p.s != ps[p.out_index]) return false;
    return true;
  }
};

// false: plan.too_large_pair does not fit the budget on its own
inline bool build_align_plan(uint32_t n_pairs, const uint32_t *pair_q, const uint32_t *pair_s,
                             const std::vector<uint32_t> &dim_n, const std::vector<uint32_t> &dim_m,
                             const std::vector<uint64_t> &cost, bool zero_gap, const PackedConfig &pk, uint64_t budget,
                             AlnPlan &plan) {
  const PlanTables &T = plan_tables();
  std::vector<PkJob> jobs;
  std::vector<uint8_t> in_job(n_pairs, 0);
  if (pk.on && n_pairs >= 2) plan_packed_jobs(pk, dim_n, dim_m, cost, jobs, in_job);
  std::vector<uint32_t> order;  // the pairs outside the jobs, by descending cost (1024 buckets)
  {
    uint64_t max_cost = 1;
    uint32_t n_rest = 0;
    for (uint32_t i = 0; i < n_pairs; ++i)
      if (!in_job[i]) {
        max_cost = cost[i] > max_cost ? cost[i] : max_cost;
        ++n_rest;
      }
    constexpr int NB = 1024;
    std::vector<uint32_t> bucket_start(NB + 1, 0);
    order.resize(n_rest);
    const double scale = (double)(NB - 1) / (double)max_cost;
    auto bucket_of = [&](uint64_t c) {
      const int b = NB - 1 - (int)((double)c * scale);
      return (uint32_t)(b < 0 ? 0 : (b > NB - 1 ? NB - 1 : b));
    };
    for (uint32_t i = 0; i < n_pairs; ++i)
      if (!in_job[i]) bucket_start[bucket_of(cost[i]) + 1]++;
    for (int b = 0; b < NB; ++b) bucket_start[b + 1] += bucket_start[b];
    std::vector<uint32_t> cur(bucket_start.begin(), bucket_start.end() - 1);
    for (uint32_t i = 0; i < n_pairs; ++i)
      if (!in_job[i]) order[cur[bucket_of(cost[i])]++] = i;
    // the long pairs (one CTA each) first, exactly
    std::stable_partition(order.begin(), order.end(), [&](uint32_t i) { return cost[i] >= BIG_CELLS; });
  }
  std::vector<AlnPair> &big_pairs = plan.big_pairs, &small_pairs = plan.small_pairs, &job_pairs = plan.job_pairs;
  job_pairs.reserve(jobs.size() * 2);
  small_pairs.reserve(order.size());
  uint64_t used = 0;
  bool too_large = false;
  auto cut = [&]() {
    plan.chunks.push_back(AlnChunk{(uint32_t)big_pairs.size(), (uint32_t)(job_pairs.size() / 2), (uint32_t)small_pairs.size()});
  };
  auto place = [&](uint64_t b, uint32_t i) -> uint64_t {
    if (b > budget) {
      plan.too_large_pair = i;
      plan.too_large_bytes = b;
      too_large = true;
      return 0;
    }
    if (used + b > budget) {
      cut();
      used = 0;
    }
    const uint64_t at = used;
    used += b;
    plan.max_used = used > plan.max_used ? used : plan.max_used;
    return at;
  };
  size_t k = 0;
  for (; k < order.size() && cost[order[k]] >= BIG_CELLS && !too_large; ++k) {
    const uint32_t i = order[k];
    const uint64_t at = place(pair_scratch_bytes(dim_n[i], dim_m[i], 8, true), i);
    big_pairs.push_back(AlnPair{i, pair_q[i], pair_s[i], 8u, at});
  }
  // then the jobs and the single pairs, merged by their estimated work
  size_t j = 0;
  while (!too_large && (k < order.size() || j < jobs.size())) {
    bool take_job = j < jobs.size();
    if (take_job && k < order.size()) {
      const uint32_t i = order[k];
      take_job = jobs[j].work >= (zero_gap ? (double)(dim_n[i] + 31) * T.row_single[dim_m[i]] : single_work(dim_n[i], dim_m[i], 8));
    }
    if (take_job) {
      const PkJob &jb = jobs[j++];
      const uint64_t fb = ((uint64_t)(jb.cw == 8 ? T.padded_pk8[jb.Mx] : T.padded_pk4[jb.Mx]) * (jb.N + 31ull) + 255) & ~255ull;  // pk_flags_bytes
      const uint64_t at = place(2 * fb + pk_bnd_bytes(jb.N), jb.a);
      job_pairs.push_back(AlnPair{jb.a, pair_q[jb.a], pair_s[jb.a], (uint32_t)jb.cw, at});
      job_pairs.push_back(AlnPair{jb.b, pair_q[jb.b], pair_s[jb.b], (uint32_t)jb.cw, at + fb});
      plan.pk_maxcw_used = jb.cw > plan.pk_maxcw_used ? jb.cw : plan.pk_maxcw_used;
    } else {
      const uint32_t i = order[k++];
      const int cw = zero_gap ? (int)T.cw_single[dim_m[i]] : 8;
      const uint64_t at = place(pair_scratch_bytes(dim_n[i], dim_m[i], cw, false), i);
      small_pairs.push_back(AlnPair{i, pair_q[i], pair_s[i], (uint32_t)cw, at});
    }
  }
  if (too_large) return false;
  cut();
  return true;
}

}  // namespace kaamer
