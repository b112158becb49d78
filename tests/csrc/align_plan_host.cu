// Host test of the alignment schedule (kaamer_b200/csrc/align_plan.hpp): invariants of build_align_plan on
// random pair sets, with and without packing, under a large and a small traceback-memory budget; prints the
// planning time of a C5-sized call.  Compiled by nvcc as host code, run by tests/test_align_packed_cpu.py.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "../../kaamer_b200/csrc/align_plan.hpp"

using namespace kaamer;

static int fails = 0;
#define CHECK(c, ...)                  \
  do {                                 \
    if (!(c)) {                        \
      if (fails++ < 20) {              \
        printf("FAIL %s: ", #c);       \
        printf(__VA_ARGS__);           \
        printf("\n");                  \
      }                                \
    }                                  \
  } while (0)

struct Interval {
  uint64_t b, e;
};

static void check_plan(const char *name, uint32_t n_pairs, const std::vector<uint32_t> &pq, const std::vector<uint32_t> &ps,
                       const std::vector<uint32_t> &dn, const std::vector<uint32_t> &dm, const std::vector<uint64_t> &cost,
                       bool zero_gap, const PackedConfig &pk, uint64_t budget) {
  AlnPlan plan;
  const bool ok = build_align_plan(n_pairs, pq.data(), ps.data(), dn, dm, cost, zero_gap, pk, budget, plan);
  CHECK(ok, "%s: plan failed on pair %lld", name, (long long)plan.too_large_pair);
  if (!ok) return;
  std::vector<int> seen(n_pairs, 0);
  for (const AlnPair &p : plan.big_pairs) {
    seen[p.out_index]++;
    CHECK(cost[p.out_index] >= BIG_CELLS && p.cw == 8, "%s: long pair %u", name, p.out_index);
  }
  for (const AlnPair &p : plan.small_pairs) {
    seen[p.out_index]++;
    CHECK(cost[p.out_index] < BIG_CELLS, "%s: single pair %u has %llu cells", name, p.out_index, (unsigned long long)cost[p.out_index]);
    CHECK(p.cw == (zero_gap ? (uint32_t)choose_cw(dm[p.out_index]) : 8u), "%s: cw of single pair %u", name, p.out_index);
  }
  CHECK(plan.job_pairs.size() % 2 == 0, "%s: odd job list", name);
  if (!pk.on) CHECK(plan.job_pairs.empty(), "%s: jobs although packing is off", name);
  for (size_t k = 0; k + 1 < plan.job_pairs.size(); k += 2) {
    const AlnPair &a = plan.job_pairs[k], &b = plan.job_pairs[k + 1];
    seen[a.out_index]++;
    seen[b.out_index]++;
    CHECK(a.cw == b.cw && (a.cw == 4 || a.cw == 8) && (int)a.cw <= pk.maxcw, "%s: job cw %u / %u", name, a.cw, b.cw);
    const uint32_t N = std::max(dn[a.out_index], dn[b.out_index]), Mx = std::max(dm[a.out_index], dm[b.out_index]);
    CHECK(b.scratch == a.scratch + pk_flags_bytes(N, Mx, (int)a.cw), "%s: second region of job %zu", name, k / 2);
    for (const AlnPair *p : {&a, &b}) {
      const uint32_t n = dn[p->out_index], m = dm[p->out_index];
      CHECK(n >= 1 && m >= 1 && n <= (uint32_t)PK_MAX_DIM && m <= (uint32_t)PK_MAX_DIM, "%s: job pair %u is %u x %u", name, p->out_index, n, m);
      CHECK(std::min(n, m) <= pk.max_min_dim && cost[p->out_index] < pk.max_cells, "%s: job pair %u breaks the score / cell bound", name, p->out_index);
    }
    // the kernel sizes its blocks from the job geometry: the padded columns must cover both subjects
    CHECK(pk_padded_cols(Mx, (int)a.cw) >= Mx, "%s: padded columns", name);
  }
  for (uint32_t i = 0; i < n_pairs; ++i) CHECK(seen[i] == 1, "%s: pair %u scheduled %d times", name, i, seen[i]);
  CHECK(plan.pair_q_ok(pq, ps), "%s: query / subject ids copied wrongly", name);
  // chunks: monotone ranges that end at the list sizes; regions of a chunk disjoint and inside the budget
  AlnChunk prev{0, 0, 0};
  CHECK(!plan.chunks.empty(), "%s: no chunk", name);
  for (const AlnChunk &c : plan.chunks) {
    CHECK(c.big_end >= prev.big_end && c.job_end >= prev.job_end && c.small_end >= prev.small_end, "%s: chunk order", name);
    std::vector<Interval> iv;
    for (uint32_t k = prev.big_end; k < c.big_end; ++k) {
      const AlnPair &p = plan.big_pairs[k];
      iv.push_back({p.scratch, p.scratch + pair_scratch_bytes(dn[p.out_index], dm[p.out_index], 8, true)});
    }
    for (uint32_t k = prev.small_end; k < c.small_end; ++k) {
      const AlnPair &p = plan.small_pairs[k];
      iv.push_back({p.scratch, p.scratch + pair_scratch_bytes(dn[p.out_index], dm[p.out_index], (int)p.cw, false)});
    }
    for (uint32_t k = prev.job_end; k < c.job_end; ++k) {
      const AlnPair &a = plan.job_pairs[2 * k], &b = plan.job_pairs[2 * k + 1];
      const uint32_t N = std::max(dn[a.out_index], dn[b.out_index]), Mx = std::max(dm[a.out_index], dm[b.out_index]);
      iv.push_back({a.scratch, a.scratch + 2 * pk_flags_bytes(N, Mx, (int)a.cw) + pk_bnd_bytes(N)});
    }
    std::sort(iv.begin(), iv.end(), [](const Interval &x, const Interval &y) { return x.b < y.b; });
    for (size_t k = 0; k < iv.size(); ++k) {
      CHECK(iv[k].e <= budget && iv[k].e <= plan.max_used, "%s: region beyond the budget", name);
      if (k) CHECK(iv[k].b >= iv[k - 1].e, "%s: regions overlap", name);
      CHECK((iv[k].b & 255) == 0, "%s: region not 256-byte aligned", name);
    }
    prev = c;
  }
  CHECK(prev.big_end == plan.big_pairs.size() && 2 * (size_t)prev.job_end == plan.job_pairs.size() &&
            prev.small_end == plan.small_pairs.size(),
        "%s: last chunk", name);
  printf("%-34s %u pairs -> %zu long, %zu single, %zu jobs, %zu chunk(s), %.1f MB peak\n", name, n_pairs, plan.big_pairs.size(),
         plan.small_pairs.size(), plan.job_pairs.size() / 2, plan.chunks.size(), plan.max_used / 1e6);
}

int main() {
  std::mt19937 rng(7);
  kaamer_aln_model model;
  memset(&model, 0, sizeof model);
  for (int i = 1; i < 26; ++i)
    for (int j = 1; j < 26; ++j) model.matrix[i * 26 + j] = (int8_t)(i == j ? 11 : -4);
  model.gap_open = -11;
  unsetenv("KAAMER_ALIGN_PACKED");
  unsetenv("KAAMER_ALIGN_PK_MAXCW");
  unsetenv("KAAMER_ALIGN_PK_CELLS");
  PackedConfig on = packed_config(model, true), off = packed_config(model, false);
  CHECK(on.on && !off.on && on.maxcw == 8 && on.max_min_dim == 32000 / 11, "packed_config defaults");
  auto make = [&](uint32_t n_pairs, int kind, std::vector<uint32_t> &pq, std::vector<uint32_t> &ps, std::vector<uint32_t> &dn,
                  std::vector<uint32_t> &dm, std::vector<uint64_t> &cost) {
    pq.resize(n_pairs), ps.resize(n_pairs), dn.resize(n_pairs), dm.resize(n_pairs), cost.resize(n_pairs);
    std::lognormal_distribution<double> len(5.6, 0.6);
    for (uint32_t i = 0; i < n_pairs; ++i) {
      uint32_t n = (uint32_t)std::min(30000.0, std::max(kind == 2 ? 0.0 : 7.0, len(rng)));
      uint32_t m = rng() % 3 ? (uint32_t)(n * (0.8 + (rng() % 400) / 1000.0)) : (uint32_t)std::min(30000.0, std::max(7.0, len(rng)));
      if (kind == 1 && rng() % 50 == 0) n = 3000 + rng() % 3000, m = 3000 + rng() % 3000;  // long pairs
      if (kind == 2 && rng() % 9 == 0) m = rng() % 3;                                        // empty / tiny
      pq[i] = rng() % 1000, ps[i] = rng() % 5000, dn[i] = n, dm[i] = m, cost[i] = (uint64_t)n * m;
    }
  };
  std::vector<uint32_t> pq, ps, dn, dm;
  std::vector<uint64_t> cost;
  for (uint32_t n_pairs : {0u, 1u, 2u, 3u, 17u, 1000u, 20000u}) {
    for (int kind = 0; kind < 3; ++kind) {
      make(n_pairs, kind, pq, ps, dn, dm, cost);
      char name[64];
      snprintf(name, sizeof name, "n=%u kind=%d packed", n_pairs, kind);
      check_plan(name, n_pairs, pq, ps, dn, dm, cost, true, on, 16ull << 30);
      snprintf(name, sizeof name, "n=%u kind=%d packed, 256 MB", n_pairs, kind);
      check_plan(name, n_pairs, pq, ps, dn, dm, cost, true, on, 256ull << 20);
      PackedConfig cw4 = on;
      cw4.maxcw = 4;
      snprintf(name, sizeof name, "n=%u kind=%d packed, 4 columns", n_pairs, kind);
      check_plan(name, n_pairs, pq, ps, dn, dm, cost, true, cw4, 16ull << 30);
      snprintf(name, sizeof name, "n=%u kind=%d 32-bit only", n_pairs, kind);
      check_plan(name, n_pairs, pq, ps, dn, dm, cost, false, off, 1ull << 30);
    }
  }
  {  // the per-length tables state the cost model functions
    const PlanTables &T = plan_tables();
    for (int k = 0; k < 20000; ++k) {
      const uint32_t n = 1 + rng() % 16000, m = 1 + rng() % PK_MAX_DIM, big_m = 1 + rng() % 65534;
      const int cw8 = choose_cw_pk(m, 8);
      CHECK(T.cw_pk8[m] == cw8 && T.padded_pk8[m] == pk_padded_cols(m, cw8) && T.padded_pk4[m] == pk_padded_cols(m, 4), "packed tables at m = %u", m);
      const double w8 = packed_work(n, m, cw8), w4 = packed_work(n, m, 4), ws = single_work(n, big_m, choose_cw(big_m));
      CHECK(std::abs((double)(n + 31) * T.row_pk8[m] - w8) <= 1e-6 * w8 && std::abs((double)(n + 31) * T.row_pk4[m] - w4) <= 1e-6 * w4, "packed work at %u x %u", n, m);
      CHECK(T.cw_single[big_m] == choose_cw(big_m) && std::abs((double)(n + 31) * T.row_single[big_m] - ws) <= 1e-6 * ws, "single work at %u x %u", n, big_m);
      CHECK(pk_flags_bytes(n, m, cw8) == (((uint64_t)T.padded_pk8[m] * (n + 31ull) + 255) & ~255ull), "region bytes at %u x %u", n, m);
    }
  }
  {  // a pair that cannot fit
    make(100, 1, pq, ps, dn, dm, cost);
    dn[40] = 60000, dm[40] = 60000, cost[40] = 3600000000ull;
    AlnPlan plan;
    const bool ok = build_align_plan(100, pq.data(), ps.data(), dn, dm, cost, true, on, 1ull << 30, plan);
    CHECK(!ok && plan.too_large_pair == 40, "the oversized pair is reported (%lld)", (long long)plan.too_large_pair);
  }
  {  // planning time of a C5-sized call
    make(150000, 1, pq, ps, dn, dm, cost);
    double best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      AlnPlan plan;
      const auto t0 = std::chrono::steady_clock::now();
      build_align_plan(150000, pq.data(), ps.data(), dn, dm, cost, true, on, 16ull << 30, plan);
      best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    printf("planning 150000 pairs: %.2f ms\n", best);
  }
  printf("%d failures\n", fails);
  return fails ? 1 : 0;
}
