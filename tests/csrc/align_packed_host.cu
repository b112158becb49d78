// CPU lock-step emulation of the packed (int16x2) Smith-Waterman forward pass.
//
// Test infrastructure: compiled by nvcc as HOST code (no GPU needed) and run by tests/test_align_packed_cpu.py.
// The per-lane step `pk_step`, the end-cell merge and the profile builder are the very functions the kernel
// k_sw_affine_pk executes (kaamer_b200/csrc/align_packed.cuh, __host__ __device__); the only thing emulated is
// the warp: 32 lane states stepped in lock-step, the warp shuffle replaced by a snapshot of the previous step.
// Every traceback byte of every real cell and the end cell of both pairs are compared with a plain scalar DP
// that states the cell of align.cu `sw_cell` (the definition the 32-bit kernels implement, bit-exact against the
// oracle on the GPU).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "../../kaamer_b200/csrc/align_packed.cuh"

using namespace kaamer;

static int8_t B62[26 * 32];
static int8_t LIDX[256];
static const char ORDER[] = "-ABCDEFGHIJKLMNPQRSTVWXYZ*";

struct Ref {
  std::vector<uint8_t> f;  // [n][m] traceback bytes
  int best_s = 0;
  uint32_t best_pos = 0;
};

static Ref scalar_dp(const std::vector<uint8_t> &q, const std::vector<uint8_t> &s, int open) {
  const int n = (int)q.size(), m = (int)s.size();
  Ref R;
  R.f.assign((size_t)n * m, 0);
  std::vector<int> Mp(m + 1, 0), Up(m + 1, 0), Bp(m + 1, 0), Mc(m + 1), Uc(m + 1), Bc(m + 1), Lc(m + 1);
  for (int i = 1; i <= n; ++i) {
    Mc[0] = 0, Uc[0] = 0, Bc[0] = 0, Lc[0] = 0;
    const int qi = LIDX[pk_fix_u(q[i - 1])];
    for (int j = 1; j <= m; ++j) {
      const int sc = B62[qi * 32 + LIDX[pk_fix_u(s[j - 1])]];
      const int mm = std::max(Bp[j - 1] + sc, 0);
      const int uo = Mp[j] + open, lo = Mc[j - 1] + open;
      const int u = std::max(uo, Up[j]), l = std::max(lo, Lc[j - 1]);
      const int mu = std::max(mm, u), b = std::max(mu, l);
      // PK encoding of the traceback byte: bits 0 / 1 are the raw comparisons, the layer is min(f & 3, 2)
      uint32_t f = (mm >= u ? 0u : 1u) | (mu >= l ? 0u : 2u);
      f |= (mm == 0 ? 4u : 0u) | (uo >= Up[j] ? 8u : 0u) | (lo >= Lc[j - 1] ? 16u : 0u);
      R.f[(size_t)(i - 1) * m + (j - 1)] = (uint8_t)f;
      Mc[j] = mm, Uc[j] = u, Lc[j] = l, Bc[j] = b;
      if (mm > 0 && mm >= R.best_s) {  // last cell in row-major order holding the maximum
        R.best_s = mm;
        R.best_pos = ((uint32_t)i << 16) | (uint32_t)j;
      }
    }
    Mp.swap(Mc), Up.swap(Uc), Bp.swap(Bc);
  }
  return R;
}

struct Rows {
  const uint8_t *qA, *qB;
  int nA, nB;
};

template <int CW>
static void emulate_block(const PkBlockArgs &g, const Rows &q, int col0, int &sA, uint32_t &pA, int &sB, uint32_t &pB) {
  static PkLane<CW> st[32];
  for (int l = 0; l < 32; ++l) st[l].init();
  const int steps = g.N + 31;
  for (int t = 0; t < steps; ++t) {
    uint32_t pm[32], pl[32], pb[32];
    for (int l = 0; l < 32; ++l) pm[l] = st[l].pubM, pl[l] = st[l].pubL, pb[l] = st[l].pubB;
    for (int l = 0; l < 32; ++l) {
      const int src = l ? l - 1 : 0;  // __shfl_up_sync: lane 0 reads itself
      const int r = t - l;  // (the kernel passes the rows round by shuffle; a row outside [0, N) is never used)
      const uint32_t qi2 = r >= 0 ? pk_query_rows(LIDX, q.qA, q.nA, q.qB, q.nB, r) : 0x1A1Au;
      pk_step<CW>(st[l], g, l, t, pm[src], pl[src], pb[src], qi2);
    }
  }
  // the kernel's butterfly reduction: higher score, then larger position
  for (int l = 0; l < 32; ++l) {
    int a = 0, b = 0;
    uint32_t ap = 0, bp = 0;
    pk_block_end<CW>(st[l], col0 + l * CW, a, ap, b, bp);
    if (a > sA || (a == sA && ap > pA)) sA = a, pA = ap;
    if (b > sB || (b == sB && bp > pB)) sB = b, pB = bp;
  }
}

static uint8_t dir_at(const uint8_t *scratch, int nrows, int cw, int tail_from, int i, int j) {  // align.cu dir_at
  int col = j - 1;
  const int r = i - 1;
  size_t base = 0;
  if (col >= tail_from) {
    base = (size_t)(tail_from / (32 * cw)) * ((size_t)(nrows + 31) * 32 * cw);
    col -= tail_from;
    cw = 4;
  }
  const int bw = 32 * cw, blk = col / bw, in = col - blk * bw, lane = in / cw, c = in - lane * cw;
  return scratch[base + (size_t)blk * ((size_t)(nrows + 31) * 32 * cw) + ((size_t)(r + lane) * 32 + lane) * cw + c];
}

static long run_job(const std::vector<uint8_t> &qA, const std::vector<uint8_t> &sA, const std::vector<uint8_t> &qB,
                    const std::vector<uint8_t> &sB, bool badB, int cw, int open) {
  const int nA = (int)qA.size(), mA = (int)sA.size(), nB = (int)qB.size(), mB = (int)sB.size();
  const int N = std::max(nA, nB), Mx = std::max(mA, mB), pcols = 256;
  const PkGeo geo = pk_geo(Mx, cw);
  const int nblk = geo.blocks(), tail_from = pk_tail_from(Mx, cw);
  const size_t fb = (size_t)pk_flags_bytes(N, Mx, cw);
  std::vector<uint8_t> regA(fb + 64, 0xEE), regB(fb + 64, 0xEE);
  std::vector<uint4> bnd((size_t)N + 2, make_uint4(0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0u));
  alignas(16) static int8_t profA[PK_PROF_ROWS * 256], profB[PK_PROF_ROWS * 256];
  int bsA = 0, bsB = 0;
  uint32_t bpA = 0, bpB = 0;
  const int mBd = badB ? 0 : mB, nBd = badB ? 0 : nB;
  const Rows rows{qA.data(), qB.data(), nA, nBd};
  for (int blk = 0; blk < nblk; ++blk) {
    const int bcw = blk < geo.nfull ? cw : geo.tail_cw, bw = 32 * bcw, col0 = blk * 32 * cw;
    memset(profA, 0x55, sizeof profA);
    memset(profB, 0x55, sizeof profB);
    for (int l = 0; l < 32; ++l) {
      pk_build_profile(profA, pcols, B62, LIDX, sA.data(), mA, col0, bw, l);
      pk_build_profile(profB, pcols, B62, LIDX, sB.data(), mBd, col0, bw, l);
    }
    PkBlockArgs g{};
    g.profA = profA, g.profB = profB, g.pcols = pcols, g.N = N;
    g.open2 = ((uint32_t)(uint16_t)(int16_t)open) * 0x00010001u;
    g.zero2 = 0u;
    g.dirsA = regA.data() + (size_t)blk * ((size_t)(N + 31) * 32 * cw);
    g.dirsB = regB.data() + (size_t)blk * ((size_t)(N + 31) * 32 * cw);
    g.bnd_in = blk > 0 ? bnd.data() : nullptr;
    g.bnd_out = blk + 1 < nblk ? bnd.data() : nullptr;
    if (bcw == 8) emulate_block<8>(g, rows, col0, bsA, bpA, bsB, bpB);
    else emulate_block<4>(g, rows, col0, bsA, bpA, bsB, bpB);
  }
  if (regA[fb] != 0xEE || regB[fb] != 0xEE) {
    printf("traceback region overrun (cw %d, %d x %d)\n", cw, N, Mx);
    return 1;
  }
  long bad = 0;
  auto check = [&](const char *name, const std::vector<uint8_t> &q, const std::vector<uint8_t> &s,
                   const std::vector<uint8_t> &reg, int bs, uint32_t bp, bool is_bad) {
    if (is_bad) {
      if (bs != 0) ++bad, printf("%s: bad pair reported score %d\n", name, bs);
      return;
    }
    const Ref R = scalar_dp(q, s, open);
    const int n = (int)q.size(), m = (int)s.size();
    if (R.best_s != bs || (bs > 0 && R.best_pos != bp)) {
      ++bad;
      printf("%s (%d x %d, cw %d, N %d): end cell %d@%08x, expected %d@%08x\n", name, n, m, cw, N, bs, bp, R.best_s, R.best_pos);
    }
    long cells_bad = 0;
    for (int i = 1; i <= n; ++i)
      for (int j = 1; j <= m; ++j)
        if (dir_at(reg.data(), N, cw, tail_from, i, j) != R.f[(size_t)(i - 1) * m + (j - 1)]) {
          if (cells_bad++ < 3)
            printf("%s (%d x %d, cw %d): cell (%d, %d) flags %02x, expected %02x\n", name, n, m, cw, i, j,
                   dir_at(reg.data(), N, cw, tail_from, i, j), R.f[(size_t)(i - 1) * m + (j - 1)]);
        }
    bad += cells_bad;
  };
  check("A", qA, sA, regA, bsA, bpA, false);
  check("B", qB, sB, regB, bsB, bpB, badB);
  return bad;
}

int main(int argc, char **argv) {
  const int jobs = argc > 1 ? atoi(argv[1]) : 60;
  std::mt19937 rng(12345);
  // a BLOSUM-like integer matrix in biogo's alphabet order with a zero gap row / column
  memset(B62, 0, sizeof B62);
  for (int i = 1; i < 26; ++i)
    for (int j = i; j < 26; ++j) {
      const int v = i == j ? 4 + (int)(rng() % 8) : -4 + (int)(rng() % 8);
      B62[i * 32 + j] = B62[j * 32 + i] = (int8_t)v;
    }
  memset(LIDX, -1, sizeof LIDX);
  for (int i = 0; i < 26; ++i) LIDX[(unsigned char)ORDER[i]] = (int8_t)i, LIDX[(unsigned char)tolower(ORDER[i])] = (int8_t)i;
  const char AA[] = "ARNDCQEGHILKMFPSTWYVUu*BZXJ";
  auto rnd = [&](int n) {
    std::vector<uint8_t> v(n);
    for (auto &c : v) c = (uint8_t)AA[rng() % (rng() % 8 ? 20 : 27)];
    return v;
  };
  auto mutate = [&](const std::vector<uint8_t> &s, double rate, double indel) {
    std::vector<uint8_t> o;
    std::uniform_real_distribution<double> U(0, 1);
    for (uint8_t c : s) {
      const double u = U(rng);
      if (u < indel) continue;
      if (u < 2 * indel) {
        auto ins = rnd(1 + (int)(rng() % 6));
        o.insert(o.end(), ins.begin(), ins.end());
      }
      o.push_back(U(rng) > rate ? c : (uint8_t)AA[rng() % 20]);
    }
    if (o.empty()) o.push_back('A');
    return o;
  };
  long bad = 0, cells = 0;
  for (int k = 0; k < jobs; ++k) {
    const int cw = k % 3 == 1 ? 4 : 8;  // 8: blocks of 256 columns + a 128-column tail block when that is enough
    const int lenA = 1 + (int)(rng() % (k % 7 == 0 ? 1300 : 420)), lenB = 1 + (int)(rng() % (k % 5 == 0 ? 900 : 420));
    auto baseA = rnd(lenA), baseB = rnd(lenB);
    auto qA = k % 3 ? mutate(baseA, 0.25, 0.03) : rnd(1 + (int)(rng() % 300));
    auto qB = k % 4 ? mutate(baseB, 0.1, 0.02) : rnd(1 + (int)(rng() % 300));
    if (k % 11 == 3) qA.assign(qA.size(), 'W'), baseA.assign(baseA.size(), 'W');  // saturating scores, many ties
    const bool badB = k % 9 == 4;
    const int open = k % 13 == 5 ? -3 : (k % 17 == 6 ? 0 : -11);
    bad += run_job(qA, baseA, qB, baseB, badB, cw, open);
    cells += (long)qA.size() * baseA.size() + (badB ? 0 : (long)qB.size() * baseB.size());
    if (bad > 20) break;
  }
  // geometry corners: single cells, the longest query the packed path takes, subjects at the block / tail boundaries
  {
    const int ms[] = {1, 127, 128, 129, 255, 256, 257, 383, 384, 385, 511, 512, 513, 640, 641};
    for (int m : ms)
      for (int cw : {4, 8}) {
        auto sA = rnd(m), sB = rnd(std::max(1, m - 1 - (int)(rng() % 5)));
        auto qA = mutate(sA, 0.2, 0.02), qB = mutate(sB, 0.3, 0.05);
        bad += run_job(qA, sA, qB, sB, false, cw, -11);
        cells += (long)qA.size() * sA.size() + (long)qB.size() * sB.size();
      }
    auto one = rnd(1);
    bad += run_job(one, one, one, rnd(3), false, 8, -11);
    auto longq = rnd(PK_MAX_DIM), narrow = rnd(5), shortq = rnd(9), wide = rnd(300);
    bad += run_job(longq, narrow, shortq, wide, false, 8, -11);  // N = 16383 rows, the other pair 9 x 300
    bad += run_job(shortq, wide, longq, narrow, true, 4, -11);
    cells += 2 * ((long)PK_MAX_DIM * 5 + 9 * 300);
  }
  printf("%d jobs, %ld cells, %ld mismatches\n", jobs, cells, bad);
  return bad ? 1 : 0;
}
