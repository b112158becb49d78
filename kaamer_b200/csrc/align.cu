// align.cu — Smith-Waterman re-alignment of (query, subject) pairs on the device.
//
// Replaces align.Align (pkg/align/align.go:46-161): biogo `align.SWAffine{Matrix: BLOSUM62,
// GapOpen: -11}.Align` (call site align.go:62-67; biogo v1.0.1 is not vendored in the
// reference — DESIGN.md §Alignment states the restated definition and why parity with biogo
// itself is unpinned) followed by kaamer's own post-processing of the segment list (identity /
// similarity / mismatches in float32 :72-101, raw score and the `score == -GapOpen` gap rule
// :116-132, bitscore :136, e-value :141, coordinates :153-156).
//
// The DP (integer, exact):  three layers over a zero-initialised (n+1) x (m+1) table
//     M[i][j] = max(0, max(M,U,L)[i-1][j-1] + B62[q_i][s_j])
//     U[i][j] = max(M[i-1][j] - 11, U[i-1][j])         (query residue against a gap)
//     L[i][j] = max(M[i][j-1] - 11, L[i][j-1])         (subject residue against a gap)
// (the gap row/column of the matrix is 0: extension is free in the DP and kaamer charges
// GapExtend afterwards, align.go:127-131).  End cell = last cell in row-major order holding the
// maximum of M.  Traceback preferences: into the best of (M,U,L)[i-1][j-1] preferring M, then
// U, then L; inside U / L "open" before "extend".
//
// One WARP per pair.  The subject is cut into blocks of 32*CW columns, lane l owns CW
// consecutive columns of the block and sweeps the query rows as a wavefront (lane l is one
// row behind lane l-1); the values crossing a lane boundary travel by warp shuffle, the ones
// crossing a block boundary through a per-pair scratch column.  Substitution scores come from
// a per-warp shared-memory block profile (one 4/8-byte load per row gives the CW scores).
// Per cell one byte of traceback state (best layer:2 | M==0 | U opened | L opened) is written,
// wavefront-major so that the 32 lanes store one contiguous line per step; lane 0 walks the
// path back and accumulates every statistic of align.go on the way.  No tensor cores: this is
// integer max/add work, reported in GCUPS (cell updates per second).
#include <algorithm>
#include <cmath>
#include <type_traits>

#include "internal.cuh"
#include "align_packed.cuh"
#include "align_plan.hpp"

namespace kaamer {

constexpr int ALN_WARPS = 4;
constexpr int PROF_COLS = 512;  // 32 lanes x CW(max 16)
constexpr int GAP_OPEN_DP = -11;  // align.go:64 (hard-coded in the reference, options ignored): the default model

// NCBI BLOSUM62, 24-letter order ARNDCQEGHILKMFPSTWYVBZX*, + the J (I/L) row of BLAST+.
static const char NCBI_ORDER[] = "ARNDCQEGHILKMFPSTWYVBZX*";
static const int8_t NCBI_B62[24][24] = {
    {4, -1, -2, -2, 0, -1, -1, 0, -2, -1, -1, -1, -1, -2, -1, 1, 0, -3, -2, 0, -2, -1, 0, -4},
    {-1, 5, 0, -2, -3, 1, 0, -2, 0, -3, -2, 2, -1, -3, -2, -1, -1, -3, -2, -3, -1, 0, -1, -4},
    {-2, 0, 6, 1, -3, 0, 0, 0, 1, -3, -3, 0, -2, -3, -2, 1, 0, -4, -2, -3, 3, 0, -1, -4},
    {-2, -2, 1, 6, -3, 0, 2, -1, -1, -3, -4, -1, -3, -3, -1, 0, -1, -4, -3, -3, 4, 1, -1, -4},
    {0, -3, -3, -3, 9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2, -4},
    {-1, 1, 0, 0, -3, 5, 2, -2, 0, -3, -2, 1, 0, -3, -1, 0, -1, -2, -1, -2, 0, 3, -1, -4},
    {-1, 0, 0, 2, -4, 2, 5, -2, 0, -3, -3, 1, -2, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4},
    {0, -2, 0, -1, -3, -2, -2, 6, -2, -4, -4, -2, -3, -3, -2, 0, -2, -2, -3, -3, -1, -2, -1, -4},
    {-2, 0, 1, -1, -3, 0, 0, -2, 8, -3, -3, -1, -2, -1, -2, -1, -2, -2, 2, -3, 0, 0, -1, -4},
    {-1, -3, -3, -3, -1, -3, -3, -4, -3, 4, 2, -3, 1, 0, -3, -2, -1, -3, -1, 3, -3, -3, -1, -4},
    {-1, -2, -3, -4, -1, -2, -3, -4, -3, 2, 4, -2, 2, 0, -3, -2, -1, -2, -1, 1, -4, -3, -1, -4},
    {-1, 2, 0, -1, -3, 1, 1, -2, -1, -3, -2, 5, -1, -3, -1, 0, -1, -3, -2, -2, 0, 1, -1, -4},
    {-1, -1, -2, -3, -1, 0, -2, -3, -2, 1, 2, -1, 5, 0, -2, -1, -1, -1, -1, 1, -3, -1, -1, -4},
    {-2, -3, -3, -3, -2, -3, -3, -3, -1, 0, 0, -3, 0, 6, -4, -2, -2, 1, 3, -1, -3, -3, -1, -4},
    {-1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7, -1, -1, -4, -3, -2, -2, -1, -2, -4},
    {1, -1, 1, 0, -1, 0, 0, 0, -1, -2, -2, 0, -1, -2, -1, 4, 1, -3, -2, -2, 0, 0, 0, -4},
    {0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1, 5, -2, -2, 0, -1, -1, 0, -4},
    {-3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1, -4, -3, -2, 11, 2, -3, -4, -3, -2, -4},
    {-2, -2, -2, -3, -2, -1, -2, -3, 2, -1, -1, -2, -1, 3, -3, -2, -2, 2, 7, -1, -3, -2, -1, -4},
    {0, -3, -3, -3, -1, -2, -2, -3, -3, 3, 1, -2, 1, -1, -2, -2, 0, -3, -1, 4, -3, -2, -1, -4},
    {-2, -1, 3, 4, -3, 0, 1, -1, 0, -3, -4, 0, -3, -3, -2, 0, -1, -4, -3, -3, 4, 1, -1, -4},
    {-1, 0, 0, 1, -3, 3, 4, -2, 0, -3, -3, 1, -1, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4},
    {0, -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0, 0, -2, -1, -1, -1, -1, -1, -4},
    {-4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, 1}};
static const int8_t NCBI_J[24] = {-1, -2, -3, -3, -1, -2, -3, -4, -3, 3, 3, -3, 2, 0, -3, -2, -1, -2, -1, 2, -3, -3, -1, -4};
// biogo alphabet.Protein / kaamer AAPosInMatrix order (pkg/align/matrixScores.go:107)
static const char BIOGO_ORDER[] = "-ABCDEFGHIJKLMNPQRSTVWXYZ*";

struct AlnTables {
  int8_t b62[26 * 32];       // [i][j], row stride 32; row/column 0 = per-residue gap cost (0 in the default model)
  int8_t letter_index[256];  // alphabet.Protein.LetterIndex(): case-insensitive, -1 = illegal letter
  int8_t aa_pos[256];        // AAPosInMatrix: exact (upper-case) letters only, miss -> 0 (Go map zero value)
};

// the reference's model: biogo matrix.BLOSUM62 with GapOpen -11 (align.go:62-65); gap row/column 0
void default_align_model(kaamer_aln_model *m) {
  memset(m, 0, sizeof *m);
  auto ncbi = [&](char c) -> int { return (int)(strchr(NCBI_ORDER, c) - NCBI_ORDER); };
  for (int i = 1; i < 26; ++i)
    for (int j = 1; j < 26; ++j) {
      const char a = BIOGO_ORDER[i], b = BIOGO_ORDER[j];
      int v;
      if (a == 'J' && b == 'J') v = 3;
      else if (a == 'J') v = NCBI_J[ncbi(b)];
      else if (b == 'J') v = NCBI_J[ncbi(a)];
      else v = NCBI_B62[ncbi(a)][ncbi(b)];
      m->matrix[i * 26 + j] = (int8_t)v;
    }
  m->gap_open = GAP_OPEN_DP;
}

static void make_tables(AlnTables *t, const kaamer_aln_model *model) {
  memset(t, 0, sizeof *t);
  for (int i = 0; i < 26; ++i)
    for (int j = 0; j < 26; ++j) t->b62[i * 32 + j] = model->matrix[i * 26 + j];
  memset(t->letter_index, -1, sizeof t->letter_index);
  for (int i = 0; i < 26; ++i) {
    const unsigned char c = (unsigned char)BIOGO_ORDER[i];
    t->letter_index[c] = (int8_t)i;
    t->letter_index[(unsigned char)tolower(c)] = (int8_t)i;
    t->aa_pos[c] = (int8_t)i;
  }
}

struct AlnArgs {
  const uint8_t *q_res;
  const uint64_t *q_off;
  const uint8_t *p_res;
  const uint64_t *p_off;
  const AlnPair *pairs;
  uint32_t n_pairs;
  uint8_t *scratch;
  kaamer_aln *out;
  double lambda, K;
  int gap_open_opt, gap_extend_opt;
  double number_of_aa;
  const AlnTables *tables;   // per handle (global memory): the model's scores + letter tables
  int open;                  // SWAffine.GapOpen of the model (-11)
  int zero_gap;              // the model's gap row / column is all zero: the specialised cell update
  // AlnString (align.go:69-103): the traceback leaves the alignment columns, last column first, in
  // rev[rev_off[pair] ..): (query char) | (subject char) << 8 per column; nullptr: not wanted
  uint16_t *rev;
  const uint64_t *rev_off;
};

__device__ __forceinline__ uint8_t fix_u(uint8_t c) { return (c == 'u' || c == 'U') ? (uint8_t)'*' : c; }  // align.go:54-55


// region layout: [nblocks][n+31 steps][32 lanes][CW bytes], then 3 x int32[n] block-boundary column
__device__ __forceinline__ size_t block_stride(int n, int cw) { return (size_t)(n + 31) * 32 * cw; }

// One DP cell, in PTX so that every traceback bit costs one predicated add (the compiler turned the
// C++ form into branches and moves).
//   m = max(diag + sc, 0); u = max(Mup - 11, Uup); l = max(left_m - 11, left_l); b = max(m, u, l)
//   byte CC of w += best layer (M, then U, then L on ties) | M==0 <<2 | U opened <<3 | L opened <<4
//   (bs, br): running column maximum of M and the last row that reached it
template <int CC>
__device__ __forceinline__ void sw_cell(int diag, int sc, int Mup, int Uup, int left_m, int left_l, int r, int &m,
                                        int &u, int &l, int &b, uint32_t &w, int &bs, int &br, int open = GAP_OPEN_DP) {
  asm("{\n\t"
      ".reg .pred pu, pl, p1, p2, pz, pt;\n\t"
      ".reg .s32 uo, lo, mu, dd;\n\t"
      ".reg .u32 f;\n\t"
      "add.s32 dd, %7, %8;\n\t"
      "max.s32 %0, dd, 0;\n\t"
      "add.s32 uo, %9, %19;\n\t"
      "setp.ge.s32 pu, uo, %10;\n\t"
      "max.s32 %1, uo, %10;\n\t"
      "add.s32 lo, %11, %19;\n\t"
      "setp.ge.s32 pl, lo, %12;\n\t"
      "max.s32 %2, lo, %12;\n\t"
      "setp.ge.s32 p1, %0, %1;\n\t"
      "max.s32 mu, %0, %1;\n\t"
      "setp.ge.s32 p2, mu, %2;\n\t"
      "max.s32 %3, mu, %2;\n\t"
      "setp.eq.s32 pz, %0, 0;\n\t"
      "selp.u32 f, 0, %13, p1;\n\t"
      "selp.u32 f, f, %14, p2;\n\t"
      "@pz add.u32 f, f, %15;\n\t"
      "@pu add.u32 f, f, %16;\n\t"
      "@pl add.u32 f, f, %17;\n\t"
      "add.u32 %4, %4, f;\n\t"
      "setp.ge.s32 pt, %0, %5;\n\t"
      "max.s32 %5, %5, %0;\n\t"
      "@pt mov.s32 %6, %18;\n\t"
      "}"
      : "=&r"(m), "=&r"(u), "=&r"(l), "=&r"(b), "+r"(w), "+r"(bs), "+r"(br)
      : "r"(diag), "r"(sc), "r"(Mup), "r"(Uup), "r"(left_m), "r"(left_l), "n"(1u << (8 * CC)), "n"(2u << (8 * CC)),
        "n"(4u << (8 * CC)), "n"(8u << (8 * CC)), "n"(16u << (8 * CC)), "r"(r), "r"(open));
}

// The same cell for a model with per-residue gap costs (biogo adds Matrix[r][gap] for every gap residue):
//   u = max(Mup + ogu, Uup + gu), ogu = open + gap cost of the query residue, gu = that gap cost
//   l = max(left_m + ogl, left_l + gl), same with the subject residue of the column
template <int CC>
__device__ __forceinline__ void sw_cell_gap(int diag, int sc, int Mup, int Uup, int left_m, int left_l, int r, int ogu,
                                            int gu, int ogl, int gl, int &m, int &u, int &l, int &b, uint32_t &w,
                                            int &bs, int &br) {
  const int dd = diag + sc;
  m = dd > 0 ? dd : 0;
  const int uo = Mup + ogu, ue = Uup + gu;
  const int lo = left_m + ogl, le = left_l + gl;
  u = uo > ue ? uo : ue;
  l = lo > le ? lo : le;
  const int mu = m > u ? m : u;
  b = mu > l ? mu : l;
  uint32_t f = m >= u ? 0u : 1u;
  f = mu >= l ? f : 2u;
  f |= (m == 0 ? 4u : 0u) | (uo >= ue ? 8u : 0u) | (lo >= le ? 16u : 0u);
  w += f << (8 * CC);
  if (m >= bs) br = r;
  bs = bs > m ? bs : m;
}

template <class F>
__device__ __forceinline__ void static_for4(F f) {
  f(std::integral_constant<int, 0>());
  f(std::integral_constant<int, 1>());
  f(std::integral_constant<int, 2>());
  f(std::integral_constant<int, 3>());
}

// One column block (32*CW subject columns) swept over all query rows by one warp.
// bnd_in / bnd_out: the values crossing the block boundary, one entry per row (M, L and
// max(M,U,L) of the block's last column).  In the multi-warp kernel the previous block is being
// produced by another warp of the CTA at the same time: prog_in counts its finished rows,
// prog_out publishes ours (shared memory, volatile; data in global memory, fenced).
template <int CW, bool PIPE, int PCOLS, bool ZG>
__device__ __forceinline__ void dp_block(const int8_t *prof, const int8_t *lidx, const int8_t *b62, int open,
                                         const uint8_t *q, int n, int blk,
                                         uint8_t *dirs, const int *bnd_in, int *bnd_out,
                                         const volatile int *prog_in, volatile int *prog_out, int &out_s,
                                         uint32_t &out_pos) {
  const unsigned lane = threadIdx.x & 31;
  // gap costs of this lane's subject columns (row 0 of the block profile); unused when the gap row is zero
  uint32_t gpw[CW / 4];
#pragma unroll
  for (int g = 0; g < CW / 4; ++g) gpw[g] = ZG ? 0u : reinterpret_cast<const uint32_t *>(prof + lane * CW)[g];
  // per-column running maximum of M and the last row that reached it (">=": the last row wins);
  // 1 as the initial maximum implements the `M > 0` condition of the end-cell rule
  int bs[CW], br[CW];
  int Mup[CW], Uup[CW], Bup[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    Mup[c] = Uup[c] = Bup[c] = 0;
    bs[c] = 1;
    br[c] = -1;
  }
  int pubM = 0, pubL = 0, pubB = 0, prevB = 0;
  int avail = (PIPE && prog_in) ? 0 : n;  // rows of the previous block known to be finished
  const int j0 = blk * 32 * CW + (int)lane * CW;  // first column (0-based) of this lane
  const int steps = n + 31;
  for (int t = 0; t < steps; ++t) {
    const int r = t - (int)lane;
    const bool active = r >= 0 && r < n;
    int inM = __shfl_up_sync(0xFFFFFFFFu, pubM, 1);
    int inL = __shfl_up_sync(0xFFFFFFFFu, pubL, 1);
    int inB = __shfl_up_sync(0xFFFFFFFFu, pubB, 1);
    if (lane == 0) {
      inM = inL = inB = 0;
      if (bnd_in && active) {
        if constexpr (PIPE) {
          if (avail <= r) {
            do avail = *prog_in;
            while (avail <= r);
            __threadfence_block();
          }
        }
        inM = __ldcg(bnd_in + (uint32_t)r);
        inL = __ldcg(bnd_in + (uint32_t)(n + r));
        inB = __ldcg(bnd_in + (uint32_t)(2 * n + r));
      }
    }
    int diag = prevB;  // max(M,U,L)[r-1][j0-1]
    prevB = inB;
    if (active) {
      const int qi = lidx[fix_u(q[r])];
      const int gu = ZG ? 0 : (int)b62[qi * 32], ogu = open + gu;
      uint32_t pw[CW / 4];
      {
        const uint32_t *pp = reinterpret_cast<const uint32_t *>(prof + qi * PCOLS + lane * CW);
        if constexpr (CW == 16) {
          const uint4 p = *reinterpret_cast<const uint4 *>(pp);
          pw[0] = p.x;
          pw[1] = p.y;
          pw[2] = p.z;
          pw[3] = p.w;
        } else if constexpr (CW == 8) {
          const uint2 p = *reinterpret_cast<const uint2 *>(pp);
          pw[0] = p.x;
          pw[1] = p.y;
        } else {
#pragma unroll
          for (int g = 0; g < CW / 4; ++g) pw[g] = pp[g];
        }
      }
      int left_m = inM, left_l = inL;
      uint32_t fw[CW / 4];
#pragma unroll
      for (int g = 0; g < CW / 4; ++g) {
        uint32_t w = 0;
        static_for4([&](auto CC) {
          constexpr int cc = decltype(CC)::value;
          const int c = g * 4 + cc;
          const int sc = (int)(pw[g] << (24 - 8 * cc)) >> 24;  // sign-extended byte cc
          // One cell, in PTX so that every traceback bit costs one predicated add (the compiler
          // turned the C++ form into branches).
          //   m = max(diag + sc, 0); u = max(Mup - 11, Uup); l = max(left_m - 11, left_l)
          //   byte = best layer (M, then U, then L on ties) | M==0 <<2 | U opened <<3 | L opened <<4
          int m, u, l, b;
          if constexpr (ZG) {
            sw_cell<cc>(diag, sc, Mup[c], Uup[c], left_m, left_l, r, m, u, l, b, w, bs[c], br[c], open);
          } else {
            const int gl = (int)(gpw[g] << (24 - 8 * cc)) >> 24;
            sw_cell_gap<cc>(diag, sc, Mup[c], Uup[c], left_m, left_l, r, ogu, gu, open + gl, gl, m, u, l, b, w, bs[c],
                            br[c]);
          }
          diag = Bup[c];
          Mup[c] = m;
          Uup[c] = u;
          Bup[c] = b;
          left_m = m;
          left_l = l;
        });
        fw[g] = w;
      }
      pubM = left_m;
      pubL = left_l;
      pubB = Bup[CW - 1];
      {
        uint32_t *dp = reinterpret_cast<uint32_t *>(dirs + ((uint32_t)t * 32u + lane) * (uint32_t)CW);
        if constexpr (CW == 16) {
          *reinterpret_cast<uint4 *>(dp) = make_uint4(fw[0], fw[1], fw[2], fw[3]);
        } else if constexpr (CW == 8) {
          *reinterpret_cast<uint2 *>(dp) = make_uint2(fw[0], fw[1]);
        } else {
#pragma unroll
          for (int g = 0; g < CW / 4; ++g) dp[g] = fw[g];
        }
      }
      if (lane == 31 && bnd_out) {
        bnd_out[(uint32_t)r] = pubM;
        bnd_out[(uint32_t)(n + r)] = pubL;
        bnd_out[(uint32_t)(2 * n + r)] = pubB;
        if constexpr (PIPE) {
          if (prog_out) {
            __threadfence_block();
            *prog_out = r + 1;
          }
        }
      }
    }
  }
  // this lane's end-cell candidate, then the merge with the earlier column blocks: higher score,
  // then later in row-major order
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    if (br[c] < 0) continue;
    const uint32_t pos = ((uint32_t)(br[c] + 1) << 16) | (uint32_t)(j0 + c + 1);
    if (bs[c] > out_s || (bs[c] == out_s && pos > out_pos)) {
      out_s = bs[c];
      out_pos = pos;
    }
  }
}

// tail_from: first column (0-based) of a last block of 4 columns per lane (packed jobs, align_packed.cuh
// pk_geo), 0x7FFFFFFF when every block has cw columns per lane
__device__ __forceinline__ uint32_t dir_at(const uint8_t *scratch, int n, int cw, int tail_from, int i, int j) {
  // cell (i, j), 1-based
  int col = j - 1;
  const int r = i - 1;
  size_t base = 0;
  if (col >= tail_from) {
    base = (size_t)(tail_from / (32 * cw)) * block_stride(n, cw);
    col -= tail_from;
    cw = 4;
  }
  const int bw = 32 * cw;
  const int blk = col / bw, in = col - blk * bw;
  const int lane = in / cw, c = in - lane * cw;
  return scratch[base + (size_t)blk * block_stride(n, cw) + ((size_t)(r + lane) * 32 + lane) * cw + c];
}

// block profile: prof[a][col] = B62[a][s_col]; columns past the subject end score -100 so that
// nothing positive ever lives there
template <int PCOLS>
__device__ __forceinline__ void build_profile(int8_t *prof, const int8_t *b62, const int8_t *lidx, const uint8_t *s,
                                              int m, int blk, int bw) {
  const unsigned lane = threadIdx.x & 31;
  for (int col = lane; col < bw; col += 32) {
    const int j = blk * bw + col;
    const int sj = j < m ? (int)lidx[fix_u(s[j])] : -1;
#pragma unroll 1
    for (int aa = 0; aa < 26; ++aa) prof[aa * PCOLS + col] = sj >= 0 ? b62[aa * 32 + sj] : (int8_t)-100;
  }
}

// Traceback + the post-processing of align.go:72-157, executed by one whole warp: the 32 lanes
// fetch a window of 32 traceback bytes (and residues) ahead of the path in its current direction
// (diagonal, up or left) with ONE memory round trip.  A diagonal run is consumed whole -- votes find its end,
// the statistics are population counts and one warp sum --, gap runs are followed cell by cell through shuffles;
// a new window is fetched when the layer changes or the window is used up.
__device__ __forceinline__ void traceback_and_emit(const AlnArgs &a, const AlnPair &pr, const uint8_t *scratch,
                                                   int nrows, int tail_from, const uint8_t *q, int n, const uint8_t *s, int cw,
                                                   int best_s,
                                                   uint32_t best_pos, bool bad, const int8_t *s_b62,
                                                   const int8_t *s_lidx, const int8_t *s_apos) {
  const unsigned lane = threadIdx.x & 31;
  float identity = 0.f, similarity = 0.f, nb_pos = 0.f;
  int mismatches = 0, raw = 0, gap_openings = 0, aln_len = 0;
  int q_start = 0, q_end = 0, s_start = 0, s_end = 0;
  if (best_s > 0) {
    int i = (int)(best_pos >> 16), j = (int)(best_pos & 0xFFFFu);
    q_end = i;
    s_end = j;
    int layer = 0, cur_kind = -1, cur_score = 0, cur_lq = 0, cur_ls = 0;
    auto flush = [&]() {
      if (cur_kind < 0) return;
      raw += cur_score;
      if (cur_score == -a.gap_open_opt) {  // align.go:127: the test is on the score VALUE
        gap_openings += 1;
        const int gl = cur_lq > cur_ls ? cur_lq : cur_ls;
        raw -= (gl - 1) * a.gap_extend_opt;
      }
      cur_kind = -1;
    };
    auto open_seg = [&](int kind) {
      if (cur_kind != kind) {
        flush();
        cur_kind = kind;
        cur_score = 0;
        cur_lq = cur_ls = 0;
      }
    };
    uint16_t *rev = a.rev ? a.rev + a.rev_off[pr.out_index] : nullptr;
    bool done = false;
    while (!done) {
      const int di = layer == 2 ? 0 : 1, dj = layer == 1 ? 0 : 1;
      const int wi = i - di * (int)lane, wj = j - dj * (int)lane;
      uint32_t F = 0, QA = 0, SB = 0;
      const bool inb = wi > 0 && wj > 0;
      if (inb) {
        F = dir_at(scratch, nrows, cw, tail_from, wi, wj);
        QA = fix_u(q[wi - 1]);
        SB = fix_u(s[wj - 1]);
      }
      if (layer == 0) {
        // The diagonal run inside this window, all its cells at once.  Cell k (lane k) is consumed when the cells
        // before it are and it lies inside the table, its M is not 0 and -- for k > 0 -- the path enters it in
        // layer M: lay = layer of the predecessor stored in the byte (0 / 1 / 2 from the 32-bit kernels; the
        // packed kernel stores the two raw comparisons, so 3 also means L).  Cell 31 is left to the next window.
        uint32_t lay = F & 3u;
        lay = lay > 2u ? 2u : lay;
        const bool stop = (F & 4u) != 0;
        const bool ok = inb && !stop && (lane == 0 || lay == 0u) && lane < 31;
        const uint32_t okm = __ballot_sync(0xFFFFFFFFu, ok), inbm = __ballot_sync(0xFFFFFFFFu, inb);
        const int kf = __ffs((int)~okm) - 1;  // first cell that is not consumed: 0 .. 31
        if (kf > 0) {
          open_seg(0);
          const bool mine = (int)lane < kf;
          const bool eq = QA == SB;
          int v = 0;
          bool sim = false;
          if (mine) {
            v = s_b62[s_lidx[QA] * 32 + s_lidx[SB]];
            sim = eq || s_b62[s_apos[SB] * 32 + s_apos[QA]] > 0;  // GetAlnScoreAA (align.go:91)
            if (rev) rev[aln_len + (int)lane] = (uint16_t)(QA | (SB << 8));
          }
          cur_score += __reduce_add_sync(0xFFFFFFFFu, v);
          identity += (float)__popc(__ballot_sync(0xFFFFFFFFu, mine && eq));  // align.go:82-86 (counts: exact in float32)
          similarity += (float)__popc(__ballot_sync(0xFFFFFFFFu, sim));
          mismatches += __popc(__ballot_sync(0xFFFFFFFFu, mine && !eq && SB != '-' && QA != '-'));  // align.go:88-90
          nb_pos += (float)kf;
          cur_lq += kf;
          cur_ls += kf;
          aln_len += kf;
          i -= kf;
          j -= kf;
        }
        // why the run ended at cell kf: outside the table, the path turns into U / L there, or its M is 0
        const uint32_t lay_f = __shfl_sync(0xFFFFFFFFu, lay, kf);
        if (kf > 0 && !((inbm >> kf) & 1u)) done = true;
        else if (kf > 0 && lay_f != 0u) layer = (int)lay_f;
        else if (kf < 31) done = true;
        continue;
      }
      const int layer0 = layer;
      for (int k = 0; k < 31; ++k) {  // window cell k == current cell (i, j); k+1 stays inside the window
        const uint32_t f = __shfl_sync(0xFFFFFFFFu, F, k);
        const uint32_t ca = __shfl_sync(0xFFFFFFFFu, QA, k), cb = __shfl_sync(0xFFFFFFFFu, SB, k);
        if (layer == 1) {
          open_seg(1);
          if (rev && lane == 0) rev[aln_len] = (uint16_t)(ca | ('-' << 8));
          cur_score += s_b62[s_lidx[ca] * 32];  // the model's gap cost of this query residue (0 by default)
          cur_lq++;
          if (ca == '-') {  // a literal '-' residue equals the gap character (align.go:82)
            identity += 1.f;
            similarity += 1.f;
          }
          nb_pos += 1.f;
          aln_len++;
          if (f & 8u) {
            cur_score += a.open;
            layer = 0;
          }
          --i;
          if (i == 0) {
            done = true;
            break;
          }
        } else {
          open_seg(2);
          if (rev && lane == 0) rev[aln_len] = (uint16_t)('-' | (cb << 8));
          cur_score += s_b62[s_lidx[cb]];  // b62[gap][s_j]
          cur_ls++;
          if (cb == '-') {
            identity += 1.f;
            similarity += 1.f;
          }
          nb_pos += 1.f;
          aln_len++;
          if (f & 16u) {
            cur_score += a.open;
            layer = 0;
          }
          --j;
          if (j == 0) {
            done = true;
            break;
          }
        }
        if (layer != layer0) break;  // the gap closed: fetch a window along the diagonal
      }
    }
    flush();
    q_start = i;
    s_start = j;
  }
  if (lane != 0) return;
  kaamer_aln r;
  r.identity = __fmul_rn(__fdiv_rn(identity, nb_pos), 100.f);  // NaN for an empty alignment, as in Go
  r.similarity = __fmul_rn(__fdiv_rn(similarity, nb_pos), 100.f);
  r.length = aln_len;
  r.mismatches = mismatches;
  r.gap_openings = gap_openings;
  r.raw = raw;
  r.bitscore = __ddiv_rn(__dsub_rn(__dmul_rn(a.lambda, (double)raw), log(a.K)), log(2.0));  // align.go:136
  r.evalue = __ddiv_rn(__dmul_rn((double)n, a.number_of_aa), pow(2.0, r.bitscore));         // align.go:141
  r.query_start = q_start + 1;  // align.go:153-156
  r.query_end = q_end;
  r.subject_start = s_start + 1;
  r.subject_end = s_end;
  r.dp_score = best_s;
  r.status = bad ? 1 : 0;
  a.out[pr.out_index] = r;
}

__device__ __forceinline__ void load_tables(const AlnTables *t, int8_t *s_b62, int8_t *s_lidx, int8_t *s_apos) {
  for (int i = threadIdx.x; i < 26 * 32; i += blockDim.x) s_b62[i] = t->b62[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_lidx[i] = t->letter_index[i];
    s_apos[i] = t->aa_pos[i];
  }
  __syncthreads();
}

// ---- one warp per pair (pairs below BIG_CELLS) ------------------------------------------------
constexpr size_t WARP_SMEM = (size_t)ALN_WARPS * 26 * PROF_COLS;
__global__ void __launch_bounds__(ALN_WARPS * 32) k_sw_affine(AlnArgs a) {
  extern __shared__ __align__(16) int8_t warp_prof[];  // [ALN_WARPS][26 * PROF_COLS]
  __shared__ int8_t s_b62[26 * 32];
  __shared__ int8_t s_lidx[256];
  __shared__ int8_t s_apos[256];
  load_tables(a.tables, s_b62, s_lidx, s_apos);
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t pi = blockIdx.x * ALN_WARPS + w;
  if (pi >= a.n_pairs) return;
  const AlnPair pr = a.pairs[pi];
  const uint8_t *q = a.q_res + a.q_off[pr.q];
  const int n = (int)(a.q_off[pr.q + 1] - a.q_off[pr.q]);
  const uint8_t *s = a.p_res + a.p_off[pr.s];
  const int m = (int)(a.p_off[pr.s + 1] - a.p_off[pr.s]);
  // illegal letters: biogo returns an error that kaamer ignores (align.go:67) -> empty alignment
  bool bad = false;
  for (int i = lane; i < n; i += 32) bad |= s_lidx[fix_u(q[i])] < 0;
  for (int j = lane; j < m; j += 32) bad |= s_lidx[fix_u(s[j])] < 0;
  bad = __any_sync(0xFFFFFFFFu, bad);
  int best_s = 0;
  uint32_t best_pos = 0;
  const int cw = (int)pr.cw;
  uint8_t *scratch = a.scratch + pr.scratch;
  if (!bad && n > 0 && m > 0) {
    const int bw = 32 * cw;
    const int nblk = (m + bw - 1) / bw;
    int *bnd = reinterpret_cast<int *>(scratch + (size_t)nblk * block_stride(n, cw));  // in place: 3 x int[n]
    int8_t *prof = warp_prof + (size_t)w * 26 * PROF_COLS;
    for (int blk = 0; blk < nblk; ++blk) {
      __syncwarp();
      build_profile<PROF_COLS>(prof, s_b62, s_lidx, s, m, blk, bw);
      __syncwarp();
      uint8_t *dirs = scratch + (size_t)blk * block_stride(n, cw);
      const int *bin = blk > 0 ? bnd : nullptr;
      int *bout = blk + 1 < nblk ? bnd : nullptr;
      if (!a.zero_gap) {
        // a model with gap costs (not the reference default): the general cell, 8 columns per lane
        dp_block<8, false, PROF_COLS, false>(prof, s_lidx, s_b62, a.open, q, n, blk, dirs, bin, bout, nullptr, nullptr, best_s, best_pos);
      } else if (cw == 16) dp_block<16, false, PROF_COLS, true>(prof, s_lidx, s_b62, a.open, q, n, blk, dirs, bin, bout, nullptr, nullptr, best_s, best_pos);
      else if (cw == 12) dp_block<12, false, PROF_COLS, true>(prof, s_lidx, s_b62, a.open, q, n, blk, dirs, bin, bout, nullptr, nullptr, best_s, best_pos);
      else if (cw == 8) dp_block<8, false, PROF_COLS, true>(prof, s_lidx, s_b62, a.open, q, n, blk, dirs, bin, bout, nullptr, nullptr, best_s, best_pos);
      else dp_block<4, false, PROF_COLS, true>(prof, s_lidx, s_b62, a.open, q, n, blk, dirs, bin, bout, nullptr, nullptr, best_s, best_pos);
    }
    // end cell: maximum score, then last in row-major order (larger i, then larger j)
    for (int o = 16; o > 0; o >>= 1) {
      const int os = __shfl_xor_sync(0xFFFFFFFFu, best_s, o);
      const uint32_t op = __shfl_xor_sync(0xFFFFFFFFu, best_pos, o);
      if (os > best_s || (os == best_s && op > best_pos)) {
        best_s = os;
        best_pos = op;
      }
    }
  }
  __syncwarp();
  traceback_and_emit(a, pr, scratch, n, 0x7FFFFFFF, q, n, s, cw, best_s, best_pos, bad, s_b62, s_lidx, s_apos);
}

// ---- one warp per TWO pairs: int16x2 lanes, DPX (align_packed.cuh) ------------------------------
// a.pairs holds 2 * a.n_pairs entries: job k = pairs (2k, 2k+1), same cw (4 or 8), traceback regions of the
// job's geometry (rows = max n, columns = max m, pk_geo), the block-boundary column (16 bytes per row) behind the second region.
__global__ void __launch_bounds__(ALN_WARPS * 32, 4) k_sw_affine_pk(AlnArgs a, int pcols) {
  extern __shared__ __align__(16) int8_t pk_prof[];  // [ALN_WARPS][2][PK_PROF_ROWS * pcols]
  __shared__ int8_t s_b62[26 * 32];
  __shared__ int8_t s_lidx[256];
  __shared__ int8_t s_apos[256];
  load_tables(a.tables, s_b62, s_lidx, s_apos);
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t ji = blockIdx.x * ALN_WARPS + w;
  if (ji >= a.n_pairs) return;
  const AlnPair pa = a.pairs[2 * ji], pb = a.pairs[2 * ji + 1];
  const uint8_t *qA = a.q_res + a.q_off[pa.q], *qB = a.q_res + a.q_off[pb.q];
  const int nA = (int)(a.q_off[pa.q + 1] - a.q_off[pa.q]), nB = (int)(a.q_off[pb.q + 1] - a.q_off[pb.q]);
  const uint8_t *sA = a.p_res + a.p_off[pa.s], *sB = a.p_res + a.p_off[pb.s];
  const int mA = (int)(a.p_off[pa.s + 1] - a.p_off[pa.s]), mB = (int)(a.p_off[pb.s + 1] - a.p_off[pb.s]);
  // illegal letters: biogo returns an error that kaamer ignores (align.go:67) -> empty alignment
  bool badA = false, badB = false;
  for (int i = lane; i < nA; i += 32) badA |= s_lidx[fix_u(qA[i])] < 0;
  for (int j = lane; j < mA; j += 32) badA |= s_lidx[fix_u(sA[j])] < 0;
  for (int i = lane; i < nB; i += 32) badB |= s_lidx[fix_u(qB[i])] < 0;
  for (int j = lane; j < mB; j += 32) badB |= s_lidx[fix_u(sB[j])] < 0;
  badA = __any_sync(0xFFFFFFFFu, badA);
  badB = __any_sync(0xFFFFFFFFu, badB);
  const int N = nA > nB ? nA : nB, Mx = mA > mB ? mA : mB;  // the geometry the host sized the regions for
  const int cw = (int)pa.cw;
  const PkGeo geo = pk_geo(Mx, cw);
  const int tail_from = pk_tail_from(Mx, cw);
  uint8_t *scrA = a.scratch + pa.scratch, *scrB = a.scratch + pb.scratch;
  int bsA = 0, bsB = 0;
  uint32_t bpA = 0, bpB = 0;
  if (N > 0 && Mx > 0 && !(badA && badB)) {
    const int nblk = geo.blocks();
    uint4 *bnd = reinterpret_cast<uint4 *>(scrB + pk_flags_bytes((uint64_t)N, (uint64_t)Mx, cw));  // in place: one uint4 per row
    int8_t *profA = pk_prof + (size_t)w * 2 * PK_PROF_ROWS * pcols, *profB = profA + (size_t)PK_PROF_ROWS * pcols;
    // a pair with illegal letters: pad rows and pad columns only, nothing positive
    const int nAd = badA ? 0 : nA, nBd = badB ? 0 : nB, mAd = badA ? 0 : mA, mBd = badB ? 0 : mB;
    PkBlockArgs g;
    g.profA = profA;
    g.profB = profB;
    g.pcols = pcols;
    g.N = N;
    g.open2 = ((uint32_t)(uint16_t)(int16_t)a.open) * 0x00010001u;
    g.zero2 = a.zero_gap ? 0u : 0x00010001u;  // always 0 here (the packed path is the zero-gap-row model)
    for (int blk = 0; blk < nblk; ++blk) {
      const int bcw = blk < geo.nfull ? cw : geo.tail_cw, bw = 32 * bcw;
      const int col0 = blk * 32 * cw;  // every block before this one has cw columns per lane
      __syncwarp();
      pk_build_profile(profA, pcols, s_b62, s_lidx, sA, mAd, col0, bw, (int)lane);
      pk_build_profile(profB, pcols, s_b62, s_lidx, sB, mBd, col0, bw, (int)lane);
      __syncwarp();
      g.dirsA = scrA + (size_t)blk * block_stride(N, cw);
      g.dirsB = scrB + (size_t)blk * block_stride(N, cw);
      g.bnd_in = blk > 0 ? bnd : nullptr;
      g.bnd_out = blk + 1 < nblk ? bnd : nullptr;
      if (bcw == 8) dp_block_packed<8>(g, s_lidx, qA, nAd, qB, nBd, col0, bsA, bpA, bsB, bpB);
      else dp_block_packed<4>(g, s_lidx, qA, nAd, qB, nBd, col0, bsA, bpA, bsB, bpB);
    }
    // end cells: maximum score, then last in row-major order (larger i, then larger j)
    for (int o = 16; o > 0; o >>= 1) {
      const int osA = __shfl_xor_sync(0xFFFFFFFFu, bsA, o), osB = __shfl_xor_sync(0xFFFFFFFFu, bsB, o);
      const uint32_t opA = __shfl_xor_sync(0xFFFFFFFFu, bpA, o), opB = __shfl_xor_sync(0xFFFFFFFFu, bpB, o);
      if (osA > bsA || (osA == bsA && opA > bpA)) {
        bsA = osA;
        bpA = opA;
      }
      if (osB > bsB || (osB == bsB && opB > bpB)) {
        bsB = osB;
        bpB = opB;
      }
    }
  }
  __syncwarp();
  traceback_and_emit(a, pa, scrA, N, tail_from, qA, nA, sA, cw, badA ? 0 : bsA, bpA, badA, s_b62, s_lidx, s_apos);
  traceback_and_emit(a, pb, scrB, N, tail_from, qB, nB, sB, cw, badB ? 0 : bsB, bpB, badB, s_b62, s_lidx, s_apos);
}

// ---- one CTA per pair (long pairs): the column blocks are pipelined over the warps --------------
// Warp w sweeps blocks w, w+BIG_WARPS, ...; block b starts as soon as block b-1 has produced its
// first boundary rows, so up to BIG_WARPS*32 lanes form one wavefront over the DP table.
constexpr int BIG_WARPS = 4;
constexpr int BIG_MAX_BLOCKS = 256;  // 65535 columns / 256
constexpr int BIG_PCOLS = 256;       // 32 lanes x 8 columns
constexpr size_t BIG_SMEM = (size_t)BIG_WARPS * 26 * BIG_PCOLS;

__global__ void __launch_bounds__(BIG_WARPS * 32, 4) k_sw_affine_cta(AlnArgs a) {
  extern __shared__ __align__(16) int8_t big_prof[];  // [BIG_WARPS][26 * BIG_PCOLS]
  __shared__ int8_t s_b62[26 * 32];
  __shared__ int8_t s_lidx[256];
  __shared__ int8_t s_apos[256];
  __shared__ volatile int prog[BIG_MAX_BLOCKS];
  __shared__ int s_best[BIG_WARPS];
  __shared__ uint32_t s_pos[BIG_WARPS];
  __shared__ int s_bad;
  load_tables(a.tables, s_b62, s_lidx, s_apos);
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const AlnPair pr = a.pairs[blockIdx.x];
  const uint8_t *q = a.q_res + a.q_off[pr.q];
  const int n = (int)(a.q_off[pr.q + 1] - a.q_off[pr.q]);
  const uint8_t *s = a.p_res + a.p_off[pr.s];
  const int m = (int)(a.p_off[pr.s + 1] - a.p_off[pr.s]);
  if (threadIdx.x == 0) s_bad = 0;
  for (int i = threadIdx.x; i < BIG_MAX_BLOCKS; i += blockDim.x) prog[i] = 0;
  __syncthreads();
  bool bad = false;
  for (int i = threadIdx.x; i < n; i += blockDim.x) bad |= s_lidx[fix_u(q[i])] < 0;
  for (int j = threadIdx.x; j < m; j += blockDim.x) bad |= s_lidx[fix_u(s[j])] < 0;
  if (bad) s_bad = 1;
  __syncthreads();
  bad = s_bad != 0;
  constexpr int cw = 8, bw = 256;
  uint8_t *scratch = a.scratch + pr.scratch;
  int best_s = 0;
  uint32_t best_pos = 0;
  if (!bad && n > 0 && m > 0) {
    const int nblk = (m + bw - 1) / bw;
    int *bnd = reinterpret_cast<int *>(scratch + (size_t)nblk * block_stride(n, cw));  // [nblk][3][n]
    int8_t *prof = big_prof + (size_t)w * 26 * BIG_PCOLS;
    for (int blk = (int)w; blk < nblk; blk += BIG_WARPS) {
      __syncwarp();
      build_profile<BIG_PCOLS>(prof, s_b62, s_lidx, s, m, blk, bw);
      __syncwarp();
      uint8_t *dirs = scratch + (size_t)blk * block_stride(n, cw);
      const int *bin = blk > 0 ? bnd + (size_t)(blk - 1) * 3 * n : nullptr;
      int *bout = blk + 1 < nblk ? bnd + (size_t)blk * 3 * n : nullptr;
      if (a.zero_gap)
        dp_block<8, true, BIG_PCOLS, true>(prof, s_lidx, s_b62, a.open, q, n, blk, dirs, bin, bout,
                                           blk > 0 ? &prog[blk - 1] : nullptr, blk + 1 < nblk ? &prog[blk] : nullptr,
                                           best_s, best_pos);
      else
        dp_block<8, true, BIG_PCOLS, false>(prof, s_lidx, s_b62, a.open, q, n, blk, dirs, bin, bout,
                                            blk > 0 ? &prog[blk - 1] : nullptr, blk + 1 < nblk ? &prog[blk] : nullptr,
                                            best_s, best_pos);
    }
    for (int o = 16; o > 0; o >>= 1) {
      const int os = __shfl_xor_sync(0xFFFFFFFFu, best_s, o);
      const uint32_t op = __shfl_xor_sync(0xFFFFFFFFu, best_pos, o);
      if (os > best_s || (os == best_s && op > best_pos)) {
        best_s = os;
        best_pos = op;
      }
    }
  }
  if (lane == 0) {
    s_best[w] = best_s;
    s_pos[w] = best_pos;
  }
  __threadfence();  // traceback bytes of every warp visible before warp 0 walks them
  __syncthreads();
  if (w != 0) return;
  for (int k = 1; k < BIG_WARPS; ++k) {
    const int os = s_best[k];
    const uint32_t op = s_pos[k];
    if (os > best_s || (os == best_s && op > best_pos)) {
      best_s = os;
      best_pos = op;
    }
  }
  traceback_and_emit(a, pr, scratch, n, 0x7FFFFFFF, q, n, s, cw, best_s, best_pos, bad, s_b62, s_lidx, s_apos);
}

// ---- AlnString (align.go:69-103): one warp per pair turns the reversed columns into the three lines ----
__global__ void __launch_bounds__(128) k_aln_text(const AlnTables *t, const kaamer_aln *out, const uint16_t *rev,
                                                  const uint64_t *rev_off, const uint64_t *text_off, uint32_t n_pairs,
                                                  char *text) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (i >= n_pairs) return;
  const int L = out[i].length;
  const uint16_t *r = rev + rev_off[i];
  char *dst = text + text_off[i];
  for (int c = lane; c < L; c += 32) {
    const uint32_t col = r[L - 1 - c];
    const uint32_t ca = col & 0xFFu, cb = col >> 8;
    char mc;
    if (cb == ca) mc = (char)cb;                                                  // align.go:82-85
    else mc = t->b62[t->aa_pos[cb] * 32 + t->aa_pos[ca]] > 0 ? '+' : ' ';         // GetAlnScoreAA (:91-96)
    dst[c] = (char)ca;
    dst[L + 1 + c] = mc;
    dst[2 * L + 2 + c] = (char)cb;
  }
  if (lane == 0) {
    dst[L] = '\n';
    dst[2 * L + 1] = '\n';
  }
}

static thread_local uint32_t t_last_plan[3] = {0, 0, 0};  // long pairs, single pairs, packed jobs of the last call

int align_pairs(kaamer_gpu *h, const uint8_t *q_res, const uint64_t *q_off, const uint32_t *pair_q,
                const uint32_t *pair_s, uint32_t n_pairs, const kaamer_aln_opts *o, kaamer_aln *out,
                kaamer_aln_text **text_out) {
  DevIndex &ix = h->idx;
  cudaStream_t st = h->stream;
  if (!ix.has_proteins) {
    set_error("kaamer_gpu_align: the index was opened without a protein table");
    return KAAMER_ERR_ARG;
  }
  kaamer_aln_text *text = nullptr;
  HitsOwner *text_owner = nullptr;
  if (text_out) {
    *text_out = nullptr;
    text = new kaamer_aln_text();
    memset(text, 0, sizeof *text);
    text_owner = new HitsOwner();
    text->_owner = text_owner;
    if (text_owner->alloc(&text->off, (size_t)n_pairs + 1) != KAAMER_OK) {
      delete text_owner;
      delete text;
      return KAAMER_ERR_NOMEM;
    }
    text->off[0] = 0;
  }
  auto drop_text = [&]() {
    delete text_owner;
    delete text;
    text = nullptr;
    text_owner = nullptr;
  };
  struct TextGuard {  // every early return below releases the half-built text
    decltype(drop_text) &f;
    ~TextGuard() { f(); }
  } text_guard{drop_text};
  if (n_pairs == 0) {
    if (text_out) {
      *text_out = text;
      text = nullptr;
      text_owner = nullptr;
    }
    return KAAMER_OK;
  }
  kaamer_aln_model model;
  if (h->aln_model_set) model = h->aln_model;
  else default_align_model(&model);
  bool zero_gap = true;
  for (int i = 0; i < 26; ++i) zero_gap = zero_gap && model.matrix[i * 26] == 0 && model.matrix[i] == 0;
  if (!h->aln_ready) {
    // score tables and function attributes belong to the handle / device context (a process may drive
    // several GPUs, and two handles of one device may carry different models)
    AlnTables t;
    make_tables(&t, &model);
    if (h->ws.a_tables.ensure(sizeof t) != KAAMER_OK) {
      drop_text();
      return KAAMER_ERR_NOMEM;
    }
    cudaError_t e = cudaMemcpyAsync(h->ws.a_tables.p, &t, sizeof t, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `t` is a stack object
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sw_affine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WARP_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sw_affine_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIG_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sw_affine_pk, cudaFuncAttributeMaxDynamicSharedMemorySize, ALN_WARPS * 2 * PK_PROF_ROWS * 256);
    if (e != cudaSuccess) {
      set_error("alignment tables: %s", cudaGetErrorString(e));
      drop_text();
      return KAAMER_ERR_CUDA;
    }
    h->aln_ready = true;
  }
  uint32_t nq = 0;
  for (uint32_t i = 0; i < n_pairs; ++i) nq = pair_q[i] + 1 > nq ? pair_q[i] + 1 : nq;
  const uint64_t n_qres = q_off[nq];
  std::vector<uint64_t> cost(n_pairs);
  std::vector<uint32_t> dim_n(n_pairs), dim_m(n_pairs);
  for (uint32_t i = 0; i < n_pairs; ++i) {
    if (pair_s[i] > ix.max_protein_id) {
      set_error("pair %u: subject id %u not in the protein table (max %u)", i, pair_s[i], ix.max_protein_id);
      drop_text();
      return KAAMER_ERR_ARG;
    }
    const uint64_t n = q_off[pair_q[i] + 1] - q_off[pair_q[i]];
    const uint64_t m = ix.h_prot_off[pair_s[i] + 1] - ix.h_prot_off[pair_s[i]];
    if (n >= 65535 || m >= 65535) {
      set_error("pair %u: sequences longer than 65534 residues are not supported (%llu x %llu)", i,
                (unsigned long long)n, (unsigned long long)m);
      drop_text();
      return KAAMER_ERR_LIMIT;
    }
    cost[i] = n * m;
    dim_n[i] = (uint32_t)n;
    dim_m[i] = (uint32_t)m;
  }
  // device buffers live in the handle's workspace (allocating tens of GB per call costs more than
  // the kernels)
  SearchWorkspace &ws = h->ws;
  uint8_t *d_q = nullptr, *d_scratch = nullptr;
  uint64_t *d_qoff = nullptr;
  AlnPair *d_pairs = nullptr;
  kaamer_aln *d_out = nullptr;
  auto cleanup = [&]() {
    if (text) drop_text();
  };
#define ACUDA(call)                                                                      \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));   \
      cleanup();                                                                         \
      return KAAMER_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)
  KCHECK(ws.residues.ensure((size_t)n_qres + 16));
  KCHECK(ws.seq_off.ensure((size_t)nq + 1));
  KCHECK(ws.a_pairs.ensure((size_t)n_pairs * sizeof(AlnPair)));
  KCHECK(ws.a_out.ensure((size_t)n_pairs * sizeof(kaamer_aln)));
  d_q = ws.residues.p;
  d_qoff = ws.seq_off.p;
  d_pairs = reinterpret_cast<AlnPair *>(ws.a_pairs.p);
  d_out = reinterpret_cast<kaamer_aln *>(ws.a_out.p);
  ACUDA(cudaMemcpyAsync(d_q, q_res, (size_t)n_qres, cudaMemcpyHostToDevice, st));
  ACUDA(cudaMemcpyAsync(d_qoff, q_off, ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, st));
  // Three kinds of work: long pairs (one CTA each), packed jobs (two pairs per warp, int16x2) and single pairs
  // (one warp each), in chunks under the traceback-memory budget (align_plan.hpp)
  size_t free_b = 0, total_b = 0;
  ACUDA(cudaMemGetInfo(&free_b, &total_b));
  // traceback state of a chunk: at most half of what is free, and at most 32 GB (the C5 batch needs 27 GB: one
  // chunk, so the long-pair kernel runs underneath all the jobs); KAAMER_ALIGN_SCRATCH_GB: measurement hook
  uint64_t budget = (free_b + ws.a_scratch.n) / 2, cap = 32ull << 30;
  if (const char *e = getenv("KAAMER_ALIGN_SCRATCH_GB")) {
    const long long v = atoll(e);
    if (v >= 1 && v <= 128) cap = (uint64_t)v << 30;
  }
  if (budget > cap) budget = cap;
  AlnPlan plan;
  if (!build_align_plan(n_pairs, pair_q, pair_s, dim_n, dim_m, cost, zero_gap, packed_config(model, zero_gap), budget, plan)) {
    const uint32_t i = (uint32_t)plan.too_large_pair;
    set_error("pair %u (%u x %u) needs %llu bytes of traceback state, more than the device has free", i, dim_n[i],
              dim_m[i], (unsigned long long)plan.too_large_bytes);
    cleanup();
    return KAAMER_ERR_NOMEM;
  }
  const std::vector<AlnPair> &big_pairs = plan.big_pairs, &small_pairs = plan.small_pairs, &job_pairs = plan.job_pairs;
  const std::vector<AlnChunk> &chunks = plan.chunks;
  const uint64_t max_used = plan.max_used;
  const int pk_maxcw_used = plan.pk_maxcw_used;
  KCHECK(ws.a_scratch.ensure((size_t)max_used + 256));
  d_scratch = ws.a_scratch.p;
  AlnPair *d_big = d_pairs, *d_small = d_pairs + big_pairs.size(), *d_jobs = d_small + small_pairs.size();
  if (!big_pairs.empty())
    ACUDA(cudaMemcpyAsync(d_big, big_pairs.data(), big_pairs.size() * sizeof(AlnPair), cudaMemcpyHostToDevice, st));
  if (!small_pairs.empty())
    ACUDA(cudaMemcpyAsync(d_small, small_pairs.data(), small_pairs.size() * sizeof(AlnPair), cudaMemcpyHostToDevice, st));
  if (!job_pairs.empty())
    ACUDA(cudaMemcpyAsync(d_jobs, job_pairs.data(), job_pairs.size() * sizeof(AlnPair), cudaMemcpyHostToDevice, st));
  t_last_plan[0] = (uint32_t)big_pairs.size();
  t_last_plan[1] = (uint32_t)small_pairs.size();
  t_last_plan[2] = (uint32_t)(job_pairs.size() / 2);
  AlnArgs a{};
  a.q_res = d_q;
  a.q_off = d_qoff;
  a.p_res = ix.prot_res;
  a.p_off = ix.prot_off;
  a.scratch = d_scratch;
  a.out = d_out;
  a.lambda = o->lambda;
  a.K = o->K;
  a.gap_open_opt = o->gap_open;
  a.gap_extend_opt = o->gap_extend;
  a.number_of_aa = (double)(o->number_of_aa ? o->number_of_aa : ix.n_aa);
  a.tables = reinterpret_cast<const AlnTables *>(ws.a_tables.p);
  a.open = model.gap_open;
  a.zero_gap = zero_gap ? 1 : 0;
  a.rev = nullptr;
  a.rev_off = nullptr;
  std::vector<uint64_t> rev_off;
  if (text) {
    // room for the reversed columns of every pair: an alignment has at most n + m columns
    rev_off.resize((size_t)n_pairs + 1);
    rev_off[0] = 0;
    for (uint32_t i = 0; i < n_pairs; ++i) rev_off[i + 1] = rev_off[i] + dim_n[i] + dim_m[i];
    if (ws.a_rev.ensure((size_t)rev_off[n_pairs] * 2 + 16) != KAAMER_OK ||
        ws.a_revoff.ensure((size_t)2 * n_pairs + 2) != KAAMER_OK) {
      cleanup();
      return KAAMER_ERR_NOMEM;
    }
    ACUDA(cudaMemcpyAsync(ws.a_revoff.p, rev_off.data(), ((size_t)n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    a.rev = reinterpret_cast<uint16_t *>(ws.a_rev.p);
    a.rev_off = ws.a_revoff.p;
  }
  // Per chunk: the long pairs (one CTA each) on `st`, the jobs and the single pairs (one warp each) on the
  // second stream so that the kernels share the GPU; the chunk's scratch is reused only after all ended.
  cudaStream_t st2 = h->copy_stream;
  const int pk_pcols = 32 * (pk_maxcw_used ? pk_maxcw_used : 4);
  const size_t pk_smem = (size_t)ALN_WARPS * 2 * PK_PROF_ROWS * pk_pcols;
  AlnChunk prev{0, 0, 0};
  profile_begin(h, st, 3);
  for (const AlnChunk &ch : chunks) {
    if (ch.big_end > prev.big_end || ch.job_end > prev.job_end || ch.small_end > prev.small_end) {
      ACUDA(cudaEventRecord(h->chunk_ev[0], st));
      ACUDA(cudaStreamWaitEvent(st2, h->chunk_ev[0], 0));
      if (ch.big_end > prev.big_end) {
        a.pairs = d_big + prev.big_end;
        a.n_pairs = ch.big_end - prev.big_end;
        k_sw_affine_cta<<<a.n_pairs, BIG_WARPS * 32, BIG_SMEM, st>>>(a);
        h->prof_all_launches += 1;
      }
      if (ch.job_end > prev.job_end) {
        a.pairs = d_jobs + 2 * (size_t)prev.job_end;
        a.n_pairs = ch.job_end - prev.job_end;  // jobs
        const unsigned grid = (a.n_pairs + ALN_WARPS - 1) / ALN_WARPS;
        k_sw_affine_pk<<<grid, ALN_WARPS * 32, pk_smem, st2>>>(a, pk_pcols);
        h->prof_all_launches += 1;
      }
      if (ch.small_end > prev.small_end) {
        a.pairs = d_small + prev.small_end;
        a.n_pairs = ch.small_end - prev.small_end;
        k_sw_affine<<<(a.n_pairs + ALN_WARPS - 1) / ALN_WARPS, ALN_WARPS * 32, WARP_SMEM, st2>>>(a);
        h->prof_all_launches += 1;
      }
      ACUDA(cudaGetLastError());
      ACUDA(cudaEventRecord(h->chunk_ev[1], st2));
      ACUDA(cudaStreamWaitEvent(st, h->chunk_ev[1], 0));
    }
    prev = ch;
  }
  profile_end(h, st);
  ACUDA(cudaMemcpyAsync(out, d_out, (size_t)n_pairs * sizeof(kaamer_aln), cudaMemcpyDeviceToHost, st));
  ACUDA(cudaStreamSynchronize(st));
  if (text) {
    // offsets of the three-line texts from the alignment lengths, then one warp per pair writes them
    for (uint32_t i = 0; i < n_pairs; ++i) text->off[i + 1] = text->off[i] + 3ull * (uint64_t)out[i].length + 2ull;
    const uint64_t total = text->off[n_pairs];
    if (text_owner->alloc(&text->text, (size_t)total + 1) != KAAMER_OK || ws.a_text.ensure((size_t)total + 16) != KAAMER_OK) {
      cleanup();
      return KAAMER_ERR_NOMEM;
    }
    uint64_t *d_toff = ws.a_revoff.p + (size_t)n_pairs + 1;
    ACUDA(cudaMemcpyAsync(d_toff, text->off, ((size_t)n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    k_aln_text<<<(unsigned)(((uint64_t)n_pairs * 32 + 127) / 128), 128, 0, st>>>(
        a.tables, d_out, a.rev, a.rev_off, d_toff, n_pairs, reinterpret_cast<char *>(ws.a_text.p));
    h->prof_all_launches += 1;
    ACUDA(cudaGetLastError());
    ACUDA(cudaMemcpyAsync(text->text, ws.a_text.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    ACUDA(cudaStreamSynchronize(st));
    *text_out = text;
    text = nullptr;  // handed over
    text_owner = nullptr;
  }
#undef ACUDA
  return KAAMER_OK;
}

}  // namespace kaamer

using namespace kaamer;

static int align_entry(kaamer_gpu_t *h, const uint8_t *q_residues, const uint64_t *q_off, const uint32_t *pair_query,
                       const uint32_t *pair_subject, uint32_t n_pairs, const kaamer_aln_opts *opts, kaamer_aln *out,
                       kaamer_aln_text **text) {
  if (!h || !opts || (n_pairs && (!q_residues || !q_off || !pair_query || !pair_subject || !out))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  return guarded([&]() -> int { return align_pairs(h, q_residues, q_off, pair_query, pair_subject, n_pairs, opts, out, text); });
}

extern "C" {

int kaamer_gpu_align(kaamer_gpu_t *h, const uint8_t *q_residues, const uint64_t *q_off, const uint32_t *pair_query,
                     const uint32_t *pair_subject, uint32_t n_pairs, const kaamer_aln_opts *opts, kaamer_aln *out) {
  return align_entry(h, q_residues, q_off, pair_query, pair_subject, n_pairs, opts, out, nullptr);
}

int kaamer_gpu_align_text(kaamer_gpu_t *h, const uint8_t *q_residues, const uint64_t *q_off, const uint32_t *pair_query,
                          const uint32_t *pair_subject, uint32_t n_pairs, const kaamer_aln_opts *opts, kaamer_aln *out,
                          kaamer_aln_text **text) {
  if (!text) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  return align_entry(h, q_residues, q_off, pair_query, pair_subject, n_pairs, opts, out, text);
}

void kaamer_gpu_free_aln_text(kaamer_aln_text *t) {
  if (!t) return;
  delete (HitsOwner *)t->_owner;
  delete t;
}

void kaamer_gpu_align_last_plan(uint32_t out[3]) {
  if (!out) return;
  for (int i = 0; i < 3; ++i) out[i] = t_last_plan[i];
}

int kaamer_gpu_default_align_model(kaamer_aln_model *out) {
  if (!out) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  default_align_model(out);
  return KAAMER_OK;
}

int kaamer_gpu_set_align_model(kaamer_gpu_t *h, const kaamer_aln_model *model) {
  if (!h) {
    set_error("null handle");
    return KAAMER_ERR_ARG;
  }
  if (model && (model->gap_open > 0 || model->gap_open < -1000)) {
    set_error("align model: gap_open must be in [-1000, 0]");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  if (model) h->aln_model = *model;
  h->aln_model_set = model != nullptr;
  h->aln_ready = false;  // the score tables are rebuilt by the next call
  return KAAMER_OK;
}

}  // extern "C"
