// align_packed.cuh — Smith-Waterman forward pass of TWO (query, subject) pairs at once.
//
// Same DP, same traceback bytes and same end-cell rule as `dp_block` / `sw_cell` in align.cu (the restated
// biogo SWAffine of align.Align, pkg/align/align.go:62-67), but every 32-bit register carries two cells:
// pair A in the low 16 bits, pair B in the high 16 bits, updated by the sm_90+/sm_100 DPX instructions
//     VIADDMNMX.S16x2[.RELU]   max(a + b, c[, 0]) per half            (__viaddmax_s16x2[_relu])
//     VIMNMX.S16x2 R, P0, P1   max(a, b) per half + "a >= b" per half  (__vibmax_s16x2)
// so a cell of BOTH pairs costs 10 score instructions + 10 flag instructions + 2 for the end cell instead of
// 22 per pair.  Valid while every DP value fits a signed 16-bit half: the host takes this path only for the zero-gap-row
// model and pairs with (largest matrix entry) x min(n, m) <= PK_MAX_SCORE; everything else keeps the 32-bit
// kernels.  Rows / columns past the end of the shorter pair score PK_PAD_SCORE: they never influence a real
// cell (dependencies run right / down) and never hold the maximum (they are <= maximum - 100).
//
// The per-lane step is `__host__ __device__`: tests/csrc/align_packed_host.cu runs the identical code for 32
// emulated lanes on the CPU (the DPX intrinsics have host definitions in the CUDA headers) and compares every
// traceback byte and the end cell with a plain scalar DP.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PK_HD __host__ __device__ __forceinline__
#else
#define PK_HD inline
#endif

namespace kaamer {

constexpr int PK_PROF_ROWS = 27;  // 26 letters + the row of a query position past the end of its pair
constexpr int PK_PAD_ROW = 26;
constexpr int PK_PAD_SCORE = -100;
constexpr int PK_MAX_DIM = 16383;    // rows / columns per pair: row indices are packed 16-bit values, 0xFFFF = none
constexpr int PK_MAX_SCORE = 32000;  // bound on any DP value of a packed pair

PK_HD uint8_t pk_fix_u(uint8_t c) { return (c == 'u' || c == 'U') ? (uint8_t)'*' : c; }  // align.go:54-55

// prmt.b32, default mode: selector nibble n picks byte (n & 7) of {b:a}; bit 3 replicates that byte's sign
PK_HD uint32_t pk_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
#else
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t n = (sel >> (4 * i)) & 0xFu;
    uint32_t byte = (uint32_t)(v >> (8 * (n & 7u))) & 0xFFu;
    if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
    r |= byte << (8 * i);
  }
  return r;
#endif
}

// per-halfword addition with wrap-around (VIADD.16x2)
PK_HD uint32_t pk_add2(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __vadd2(a, b);
#else
  return ((a & 0xFFFF0000u) + (b & 0xFFFF0000u)) | ((a + b) & 0xFFFFu);
#endif
}

PK_HD uint4 pk_ld_l2(const uint4 *p) {
#ifdef __CUDA_ARCH__
  return __ldcg(p);  // the boundary column is written by this warp's lane 31: read it from L2
#else
  return *p;
#endif
}

// Column blocks of a packed job.  cw = 4: blocks of 128 columns.  cw = 8: blocks of 256 columns, and a last
// block of 128 columns (4 per lane) when at most 128 columns are left over -- a 350-column subject costs
// 256 + 128 columns instead of 512.
struct PkGeo {
  int nfull;     // blocks of 32 * cw columns
  int tail_cw;   // 0: no further block; 4: one more block of 128 columns
  PK_HD int blocks() const { return nfull + (tail_cw ? 1 : 0); }
};
PK_HD PkGeo pk_geo(int cols, int cw) {
  PkGeo g;
  const int bw = 32 * cw;
  g.nfull = cols / bw;
  const int rem = cols - g.nfull * bw;
  g.tail_cw = 0;
  if (rem > 0) {
    if (cw == 8 && rem <= 128) g.tail_cw = 4;
    else g.nfull += 1;
  }
  return g;
}
PK_HD uint64_t pk_padded_cols(uint64_t cols, int cw) {
  const PkGeo g = pk_geo((int)cols, cw);
  return (uint64_t)g.nfull * 32ull * (uint64_t)cw + (g.tail_cw ? 128ull : 0ull);
}
// first column (0-based) of the 4-wide tail block, or "never"
PK_HD int pk_tail_from(int cols, int cw) {
  const PkGeo g = pk_geo(cols, cw);
  return g.tail_cw ? g.nfull * 32 * cw : 0x7FFFFFFF;
}
// bytes of the block-boundary column of a packed job (16 per row: one 128-bit store / load per step)
PK_HD uint64_t pk_bnd_bytes(uint64_t rows) { return (16ull * rows + 255) & ~255ull; }
// bytes of one pair's traceback region of a packed job: per block [rows + 31 steps][32 lanes][block cw bytes]
PK_HD uint64_t pk_flags_bytes(uint64_t rows, uint64_t cols, int cw) {
  return (pk_padded_cols(cols, cw) * (rows + 31) + 255) & ~255ull;
}

template <int V>
struct PkInt {
  static constexpr int value = V;
};
template <class F>
PK_HD void pk_for4(F f) {
  f(PkInt<0>());
  f(PkInt<1>());
  f(PkInt<2>());
  f(PkInt<3>());
}

#ifdef __CUDA_ARCH__
// One packed cell in PTX (device): the compiler turned the C++ form below into select chains and spilled
// predicates into a register (432 instructions per 8-column step against 236 here).  Each max.s16x2 is followed
// by the unpack + setp.eq pattern of the CUDA header's __vibmax_s16x2, which ptxas fuses into ONE
// VIMNMX.S16x2 with two predicate outputs ("the maximum is the first operand": a >= b per half); every
// traceback bit is then one predicated add.  Byte CC of wA / wB receives the cell's traceback byte.
template <int CC>
__device__ __forceinline__ void pk_cell_ptx(uint32_t diag, uint32_t sc, uint32_t Mup, uint32_t Uup, uint32_t left_m,
                                            uint32_t left_l, uint32_t open2, uint32_t zero2, uint32_t rr, uint32_t &m, uint32_t &u,
                                            uint32_t &l, uint32_t &b, uint32_t &wA, uint32_t &wB, uint32_t &bs,
                                            uint32_t &br) {
  asm("{\n\t"
      ".reg .pred pl, ph;\n\t"
      ".reg .b32 dd, uo, lo, mu, nb;\n\t"
      ".reg .s16 x0, x1, y0, y1;\n\t"
      // M = max(0, diag + sc); "0 >= diag + sc" <=> M == 0 (bit 2)
      "add.s16x2 dd, %8, %9;\n\t"
      "max.s16x2 %0, %21, dd;\n\t"
      "mov.b32 {x0, x1}, %0;\n\t"
      "mov.b32 {y0, y1}, %21;\n\t"
      "setp.eq.s16 pl, x0, y0;\n\t"
      "setp.eq.s16 ph, x1, y1;\n\t"
      "@pl add.u32 %4, %4, %18;\n\t"
      "@ph add.u32 %5, %5, %18;\n\t"
      // U = max(Mup + open, Uup); "opened" (bit 3): Mup + open >= Uup
      "add.s16x2 uo, %10, %14;\n\t"
      "max.s16x2 %1, uo, %11;\n\t"
      "mov.b32 {x0, x1}, %1;\n\t"
      "mov.b32 {y0, y1}, uo;\n\t"
      "setp.eq.s16 pl, x0, y0;\n\t"
      "setp.eq.s16 ph, x1, y1;\n\t"
      "@pl add.u32 %4, %4, %19;\n\t"
      "@ph add.u32 %5, %5, %19;\n\t"
      // L = max(left M + open, left L); "opened" (bit 4)
      "add.s16x2 lo, %12, %14;\n\t"
      "max.s16x2 %2, lo, %13;\n\t"
      "mov.b32 {x0, x1}, %2;\n\t"
      "mov.b32 {y0, y1}, lo;\n\t"
      "setp.eq.s16 pl, x0, y0;\n\t"
      "setp.eq.s16 ph, x1, y1;\n\t"
      "@pl add.u32 %4, %4, %20;\n\t"
      "@ph add.u32 %5, %5, %20;\n\t"
      // max(M, U); bit 0 when NOT M >= U
      "max.s16x2 mu, %0, %1;\n\t"
      "mov.b32 {x0, x1}, mu;\n\t"
      "mov.b32 {y0, y1}, %0;\n\t"
      "setp.eq.s16 pl, x0, y0;\n\t"
      "setp.eq.s16 ph, x1, y1;\n\t"
      "@!pl add.u32 %4, %4, %16;\n\t"
      "@!ph add.u32 %5, %5, %16;\n\t"
      // max(M, U, L); bit 1 when NOT max(M, U) >= L
      "max.s16x2 %3, mu, %2;\n\t"
      "mov.b32 {x0, x1}, %3;\n\t"
      "mov.b32 {y0, y1}, mu;\n\t"
      "setp.eq.s16 pl, x0, y0;\n\t"
      "setp.eq.s16 ph, x1, y1;\n\t"
      "@!pl add.u32 %4, %4, %17;\n\t"
      "@!ph add.u32 %5, %5, %17;\n\t"
      // end cell: running column maximum of M, the last row wins (M >= maximum so far)
      "max.s16x2 nb, %0, %6;\n\t"
      "mov.b32 {x0, x1}, nb;\n\t"
      "mov.b32 {y0, y1}, %0;\n\t"
      "setp.eq.s16 pl, x0, y0;\n\t"
      "setp.eq.s16 ph, x1, y1;\n\t"
      "mov.b32 %6, nb;\n\t"
      "@pl prmt.b32 %7, %7, %15, 0x3254;\n\t"
      "@ph prmt.b32 %7, %7, %15, 0x7610;\n\t"
      "}"
      : "=&r"(m), "=&r"(u), "=&r"(l), "=&r"(b), "+r"(wA), "+r"(wB), "+r"(bs), "+r"(br)
      : "r"(diag), "r"(sc), "r"(Mup), "r"(Uup), "r"(left_m), "r"(left_l), "r"(open2), "r"(rr), "n"(1u << (8 * CC)),
        "n"(2u << (8 * CC)), "n"(4u << (8 * CC)), "n"(8u << (8 * CC)), "n"(16u << (8 * CC)), "r"(zero2));
}
#endif

// registers of one lane: CW columns of the block, both pairs
template <int CW>
struct PkLane {
  uint32_t Mup[CW], Uup[CW], Bup[CW];  // M, U and max(M, U, L) of the previous row
  uint32_t bs[CW];                     // running column maximum of M (initial 1: only M > 0 counts)
  uint32_t br[CW];                     // last row (0-based) that reached it, 0xFFFF = none
  uint32_t pubM, pubL, pubB, prevB;    // what the next lane reads; max(M, U, L)[r-1][j0-1]
  PK_HD void init() {
#pragma unroll
    for (int c = 0; c < CW; ++c) {
      Mup[c] = Uup[c] = Bup[c] = 0u;
      bs[c] = 0x00010001u;
      br[c] = 0xFFFFFFFFu;
    }
    pubM = pubL = pubB = prevB = 0u;
  }
};

struct PkBlockArgs {
  const int8_t *profA, *profB;  // block profiles [PK_PROF_ROWS][pcols] of the two subjects
  int pcols;
  int N;                        // rows of the job = max over the two pairs (geometry of both traceback regions)
  uint32_t open2;               // SWAffine.GapOpen in both halves
  uint32_t zero2;               // 0, from a run-time value: kept in a register (a literal 0 is re-materialised per cell)
  uint8_t *dirsA, *dirsB;       // this block's traceback lines: [N + 31][32][CW]
  const uint4 *bnd_in;          // block-boundary column of the previous block: one (M, L, max, -) per row
  uint4 *bnd_out;
};

// profile rows of query position `row` of both pairs: row of pair A | row of pair B << 8 (PK_PAD_ROW past the
// end of a pair; nA / nB = 0 for a pair with illegal letters)
PK_HD uint32_t pk_query_rows(const int8_t *lidx, const uint8_t *qA, int nA, const uint8_t *qB, int nB, int row) {
  const uint32_t a = row < nA ? (uint32_t)lidx[pk_fix_u(qA[row])] : (uint32_t)PK_PAD_ROW;
  const uint32_t b = row < nB ? (uint32_t)lidx[pk_fix_u(qB[row])] : (uint32_t)PK_PAD_ROW;
  return a | (b << 8);
}

// One wavefront step of one lane: row r = t - lane of this lane's CW columns, both pairs.
// inM / inL / inB: pubM / pubL / pubB of lane - 1 after the previous step (warp shuffle on the device);
// qi2 = pk_query_rows of row r (on the device the lanes load 32 rows at a time and pass them round by shuffle).
template <int CW>
PK_HD void pk_step(PkLane<CW> &s, const PkBlockArgs &g, int lane, int t, uint32_t inM, uint32_t inL, uint32_t inB,
                   uint32_t qi2) {
  const int r = t - lane;
  const bool active = r >= 0 && r < g.N;
  if (lane == 0) {
    inM = inL = inB = 0u;
    if (g.bnd_in && active) {
      const uint4 v = pk_ld_l2(g.bnd_in + (uint32_t)r);
      inM = v.x;
      inL = v.y;
      inB = v.z;
    }
  }
  uint32_t diag = s.prevB;
  s.prevB = inB;
  if (!active) return;
  const int qiA = (int)(qi2 & 0xFFu), qiB = (int)(qi2 >> 8);
  uint32_t pwA[CW / 4], pwB[CW / 4];
  {
    const int8_t *pa = g.profA + qiA * g.pcols + lane * CW;
    const int8_t *pb = g.profB + qiB * g.pcols + lane * CW;
    if constexpr (CW == 8) {
      const uint2 a = *reinterpret_cast<const uint2 *>(pa), b = *reinterpret_cast<const uint2 *>(pb);
      pwA[0] = a.x, pwA[1] = a.y;
      pwB[0] = b.x, pwB[1] = b.y;
    } else {
#pragma unroll
      for (int k = 0; k < CW / 4; ++k) {
        pwA[k] = reinterpret_cast<const uint32_t *>(pa)[k];
        pwB[k] = reinterpret_cast<const uint32_t *>(pb)[k];
      }
    }
  }
  const uint32_t rr = (uint32_t)r | ((uint32_t)r << 16);
  uint32_t left_m = inM, left_l = inL;
  uint32_t fwA[CW / 4], fwB[CW / 4];
#pragma unroll
  for (int k = 0; k < CW / 4; ++k) {
    uint32_t wA = 0u, wB = 0u;
    pk_for4([&](auto CC) {
      constexpr int cc = decltype(CC)::value;
      const int c = k * 4 + cc;
      // (int16) score of pair A | (int16) score of pair B << 16, from byte cc of the two profile words
      const uint32_t sel = (uint32_t)cc | ((8u | (uint32_t)cc) << 4) | ((4u + (uint32_t)cc) << 8) | ((12u + (uint32_t)cc) << 12);
      const uint32_t sc = pk_prmt(pwA[k], pwB[k], sel);
      // traceback byte (PK encoding): !(M >= U) | !(max(M,U) >= L) << 1 | M==0 << 2 | U opened << 3 | L opened << 4;
      // the layer is min(byte & 3, 2): the codes 0 / 1 / 2 of the 32-bit kernels, plus 3 where those say 2
      uint32_t m, u, l, b;
#ifdef __CUDA_ARCH__
      pk_cell_ptx<cc>(diag, sc, s.Mup[c], s.Uup[c], left_m, left_l, g.open2, g.zero2, rr, m, u, l, b, wA, wB, s.bs[c], s.br[c]);
#else
      // the same cell with the CUDA header's intrinsics (host definitions): what the CPU emulation test runs
      bool pu_h, pu_l, pl_h, pl_l, p1_h, p1_l, p2_h, p2_l, pz_h, pz_l, pt_h, pt_l;
      m = __vibmax_s16x2(0u, pk_add2(diag, sc), &pz_h, &pz_l);                        // max(0, diag + sc); 0 >= .. <=> M == 0
      u = __vibmax_s16x2(pk_add2(s.Mup[c], g.open2), s.Uup[c], &pu_h, &pu_l);         // "opened": Mup + open >= Uup
      l = __vibmax_s16x2(pk_add2(left_m, g.open2), left_l, &pl_h, &pl_l);
      const uint32_t mu = __vibmax_s16x2(m, u, &p1_h, &p1_l);                         // M, then U, then L on ties
      b = __vibmax_s16x2(mu, l, &p2_h, &p2_l);
      constexpr uint32_t K = 1u << (8 * cc);
      wA += (p1_l ? 0u : K) + (p2_l ? 0u : 2u * K) + (pz_l ? 4u * K : 0u) + (pu_l ? 8u * K : 0u) + (pl_l ? 16u * K : 0u);
      wB += (p1_h ? 0u : K) + (p2_h ? 0u : 2u * K) + (pz_h ? 4u * K : 0u) + (pu_h ? 8u * K : 0u) + (pl_h ? 16u * K : 0u);
      // end cell: running column maximum of M, the last row wins (">=")
      s.bs[c] = __vibmax_s16x2(m, s.bs[c], &pt_h, &pt_l);
      if (pt_l) s.br[c] = pk_prmt(s.br[c], rr, 0x3254u);
      if (pt_h) s.br[c] = pk_prmt(s.br[c], rr, 0x7610u);
#endif
      diag = s.Bup[c];
      s.Mup[c] = m;
      s.Uup[c] = u;
      s.Bup[c] = b;
      left_m = m;
      left_l = l;
    });
    fwA[k] = wA;
    fwB[k] = wB;
  }
  s.pubM = left_m;
  s.pubL = left_l;
  s.pubB = s.Bup[CW - 1];
  {
    const uint32_t at = ((uint32_t)t * 32u + (uint32_t)lane) * (uint32_t)CW;
    uint32_t *da = reinterpret_cast<uint32_t *>(g.dirsA + at), *db = reinterpret_cast<uint32_t *>(g.dirsB + at);
    if constexpr (CW == 8) {
      *reinterpret_cast<uint2 *>(da) = make_uint2(fwA[0], fwA[1]);
      *reinterpret_cast<uint2 *>(db) = make_uint2(fwB[0], fwB[1]);
    } else {
#pragma unroll
      for (int k = 0; k < CW / 4; ++k) {
        da[k] = fwA[k];
        db[k] = fwB[k];
      }
    }
  }
  if (lane == 31 && g.bnd_out) g.bnd_out[(uint32_t)r] = make_uint4(s.pubM, s.pubL, s.pubB, 0u);
}

// this lane's end-cell candidates merged into the running (score, position) of each pair:
// higher score, then later in row-major order; position = (row + 1) << 16 | (column + 1)
template <int CW>
PK_HD void pk_block_end(const PkLane<CW> &s, int j0, int &sA, uint32_t &posA, int &sB, uint32_t &posB) {
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    const int a_s = (int)(int16_t)(s.bs[c] & 0xFFFFu), b_s = (int)(int16_t)(s.bs[c] >> 16);
    const uint32_t a_r = s.br[c] & 0xFFFFu, b_r = s.br[c] >> 16;
    if (a_r != 0xFFFFu) {
      const uint32_t pos = ((a_r + 1u) << 16) | (uint32_t)(j0 + c + 1);
      if (a_s > sA || (a_s == sA && pos > posA)) {
        sA = a_s;
        posA = pos;
      }
    }
    if (b_r != 0xFFFFu) {
      const uint32_t pos = ((b_r + 1u) << 16) | (uint32_t)(j0 + c + 1);
      if (b_s > sB || (b_s == sB && pos > posB)) {
        sB = b_s;
        posB = pos;
      }
    }
  }
}

// block profile of one subject: prof[a][col] = matrix[a][s_col]; columns past the subject end and the row of
// a query position past the end of its pair score PK_PAD_SCORE.  `lane` of 32 fills its columns.
PK_HD void pk_build_profile(int8_t *prof, int pcols, const int8_t *b62, const int8_t *lidx, const uint8_t *s, int m,
                            int col0, int bw, int lane) {
  for (int col = lane; col < bw; col += 32) {
    const int j = col0 + col;
    const int sj = j < m ? (int)lidx[pk_fix_u(s[j])] : -1;
#pragma unroll 1
    for (int aa = 0; aa < 26; ++aa) prof[aa * pcols + col] = sj >= 0 ? b62[aa * 32 + sj] : (int8_t)PK_PAD_SCORE;
    prof[PK_PAD_ROW * pcols + col] = (int8_t)PK_PAD_SCORE;
  }
}

#if defined(__CUDACC__)
// One column block (32 * CW subject columns from column col0) of both pairs swept over all rows by one warp.
// Query rows: every 32 steps each lane translates ONE query position of both pairs (residue -> profile row);
// lane l needs row t - l at step t, which is held by lane (t - l) & 31 of the current or the previous group.
template <int CW>
__device__ __forceinline__ void dp_block_packed(const PkBlockArgs &g, const int8_t *lidx, const uint8_t *qA, int nA,
                                                const uint8_t *qB, int nB, int col0, int &sA, uint32_t &posA, int &sB,
                                                uint32_t &posB) {
  const int lane = (int)(threadIdx.x & 31u);
  PkLane<CW> s;
  s.init();
  const uint32_t pad2 = (uint32_t)PK_PAD_ROW | ((uint32_t)PK_PAD_ROW << 8);
  uint32_t q_prev = pad2, q_cur = pk_query_rows(lidx, qA, nA, qB, nB, lane),
           q_next = pk_query_rows(lidx, qA, nA, qB, nB, 32 + lane);
  const int steps = g.N + 31;
#pragma unroll 2  // (the loop-carried registers rotate through moves otherwise: ~14 of ~290 instructions per step)
  for (int t = 0; t < steps; ++t) {
    if ((t & 31) == 0 && t > 0) {
      q_prev = q_cur;
      q_cur = q_next;
      q_next = pk_query_rows(lidx, qA, nA, qB, nB, (t & ~31) + 32 + lane);
    }
    const int r = t - lane;
    const uint32_t from_cur = __shfl_sync(0xFFFFFFFFu, q_cur, r & 31), from_prev = __shfl_sync(0xFFFFFFFFu, q_prev, r & 31);
    const uint32_t qi2 = (r >> 5) == (t >> 5) ? from_cur : from_prev;
    const uint32_t inM = __shfl_up_sync(0xFFFFFFFFu, s.pubM, 1);
    const uint32_t inL = __shfl_up_sync(0xFFFFFFFFu, s.pubL, 1);
    const uint32_t inB = __shfl_up_sync(0xFFFFFFFFu, s.pubB, 1);
    pk_step<CW>(s, g, lane, t, inM, inL, inB, qi2);
  }
  pk_block_end<CW>(s, col0 + lane * CW, sA, posA, sB, posB);
}
#endif

}  // namespace kaamer
