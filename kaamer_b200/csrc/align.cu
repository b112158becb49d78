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

#include "internal.cuh"

namespace kaamer {

constexpr int ALN_WARPS = 4;
constexpr int PROF_COLS = 256;  // 32 lanes x CW(max 8)
constexpr int GAP_OPEN_DP = -11;  // align.go:64 (hard-coded in the reference, options ignored)

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
  int8_t b62[26 * 32];       // [i][j], row stride 32; row/column 0 (gap) = 0
  int8_t letter_index[256];  // alphabet.Protein.LetterIndex(): case-insensitive, -1 = illegal letter
  int8_t aa_pos[256];        // AAPosInMatrix: exact (upper-case) letters only, miss -> 0 (Go map zero value)
};
__constant__ AlnTables c_aln;

static void make_tables(AlnTables *t) {
  memset(t, 0, sizeof *t);
  auto ncbi = [&](char c) -> int { return (int)(strchr(NCBI_ORDER, c) - NCBI_ORDER); };
  for (int i = 1; i < 26; ++i)
    for (int j = 1; j < 26; ++j) {
      const char a = BIOGO_ORDER[i], b = BIOGO_ORDER[j];
      int v;
      if (a == 'J' && b == 'J') v = 3;
      else if (a == 'J') v = NCBI_J[ncbi(b)];
      else if (b == 'J') v = NCBI_J[ncbi(a)];
      else v = NCBI_B62[ncbi(a)][ncbi(b)];
      t->b62[i * 32 + j] = (int8_t)v;
    }
  memset(t->letter_index, -1, sizeof t->letter_index);
  for (int i = 0; i < 26; ++i) {
    const unsigned char c = (unsigned char)BIOGO_ORDER[i];
    t->letter_index[c] = (int8_t)i;
    t->letter_index[(unsigned char)tolower(c)] = (int8_t)i;
    t->aa_pos[c] = (int8_t)i;
  }
}

struct AlnPair {
  uint32_t out_index;  // position in the caller's pair list
  uint32_t q;          // query index
  uint32_t s;          // subject protein id
  uint32_t cw;         // columns per lane (4 or 8)
  uint64_t scratch;    // byte offset of the pair's traceback region
};

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
};

__device__ __forceinline__ uint8_t fix_u(uint8_t c) { return (c == 'u' || c == 'U') ? (uint8_t)'*' : c; }  // align.go:54-55

struct WarpShared {
  int8_t prof[26 * PROF_COLS];  // prof[a][col] = B62[a][s_col] of the current column block
};

// region layout: [nblocks][n+31 steps][32 lanes][CW bytes], then 3 x int32[n] block-boundary column
__device__ __forceinline__ size_t block_stride(int n, int cw) { return (size_t)(n + 31) * 32 * cw; }

template <int CW>
__device__ __forceinline__ void dp_block(const int8_t *prof, const int8_t *lidx, const uint8_t *q, int n, int blk,
                                         int nblk, uint8_t *dirs, int *bndM, int *bndL, int *bndB, int &out_s,
                                         uint32_t &out_pos) {
  const unsigned lane = threadIdx.x & 31;
  int best_s = 0;
  uint32_t best_pos = 0;
  int Mup[CW], Uup[CW], Bup[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) Mup[c] = Uup[c] = Bup[c] = 0;
  int pubM = 0, pubL = 0, pubB = 0, prevB = 0;
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
      if (blk > 0 && active) {
        inM = bndM[r];
        inL = bndL[r];
        inB = bndB[r];
      }
    }
    int diag = prevB;  // max(M,U,L)[r-1][j0-1]
    prevB = inB;
    if (active) {
      const int qi = lidx[fix_u(q[r])];
      int sc[CW];
      if constexpr (CW == 8) {
        const uint2 p = *reinterpret_cast<const uint2 *>(prof + qi * PROF_COLS + lane * 8);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          sc[c] = (int)(int8_t)(p.x >> (8 * c));
          sc[4 + c] = (int)(int8_t)(p.y >> (8 * c));
        }
      } else {
        const uint32_t p = *reinterpret_cast<const uint32_t *>(prof + qi * PROF_COLS + lane * 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) sc[c] = (int)(int8_t)(p >> (8 * c));
      }
      int left_m = inM, left_l = inL;
      uint32_t flags[CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const int d = diag + sc[c];
        const int m = d > 0 ? d : 0;
        const int uo = Mup[c] + GAP_OPEN_DP;
        const int u = uo > Uup[c] ? uo : Uup[c];
        const int lo = left_m + GAP_OPEN_DP;
        const int l = lo > left_l ? lo : left_l;
        // traceback state of this cell
        uint32_t f = (m >= u && m >= l) ? 0u : (u >= l ? 1u : 2u);
        f |= (m == 0 ? 4u : 0u) | (u == uo ? 8u : 0u) | (l == lo ? 16u : 0u);
        flags[c] = f;
        const int b = max(m, max(u, l));
        if (m > 0 && m >= best_s) {  // row-major visiting order inside the lane: last maximum wins
          best_s = m;
          best_pos = ((uint32_t)(r + 1) << 16) | (uint32_t)(j0 + c + 1);
        }
        diag = Bup[c];
        Mup[c] = m;
        Uup[c] = u;
        Bup[c] = b;
        left_m = m;
        left_l = l;
      }
      pubM = left_m;
      pubL = left_l;
      pubB = Bup[CW - 1];
      if constexpr (CW == 8) {
        uint2 w;
        w.x = flags[0] | (flags[1] << 8) | (flags[2] << 16) | (flags[3] << 24);
        w.y = flags[4] | (flags[5] << 8) | (flags[6] << 16) | (flags[7] << 24);
        *reinterpret_cast<uint2 *>(dirs + ((size_t)t * 32 + lane) * 8) = w;
      } else {
        const uint32_t w = flags[0] | (flags[1] << 8) | (flags[2] << 16) | (flags[3] << 24);
        *reinterpret_cast<uint32_t *>(dirs + ((size_t)t * 32 + lane) * 4) = w;
      }
      if (lane == 31 && blk + 1 < nblk) {
        bndM[r] = pubM;
        bndL[r] = pubL;
        bndB[r] = pubB;
      }
    }
  }
  // merge with the earlier column blocks: higher score, then later in row-major order
  if (best_s > out_s || (best_s == out_s && best_pos > out_pos)) {
    out_s = best_s;
    out_pos = best_pos;
  }
}

__device__ __forceinline__ uint32_t dir_at(const uint8_t *scratch, int n, int cw, int i, int j) {
  // cell (i, j), 1-based
  const int col = j - 1, r = i - 1;
  const int bw = 32 * cw;
  const int blk = col / bw, in = col - blk * bw;
  const int lane = in / cw, c = in - lane * cw;
  return scratch[(size_t)blk * block_stride(n, cw) + ((size_t)(r + lane) * 32 + lane) * cw + c];
}

__global__ void __launch_bounds__(ALN_WARPS * 32) k_sw_affine(AlnArgs a) {
  __shared__ __align__(16) WarpShared ws[ALN_WARPS];
  __shared__ int8_t s_b62[26 * 32];
  __shared__ int8_t s_lidx[256];
  __shared__ int8_t s_apos[256];
  for (int i = threadIdx.x; i < 26 * 32; i += blockDim.x) s_b62[i] = c_aln.b62[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_lidx[i] = c_aln.letter_index[i];
    s_apos[i] = c_aln.aa_pos[i];
  }
  __syncthreads();
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
    int *bnd = reinterpret_cast<int *>(scratch + (size_t)nblk * block_stride(n, cw));
    int8_t *prof = ws[w].prof;
    for (int blk = 0; blk < nblk; ++blk) {
      __syncwarp();
      // block profile: prof[a][col] = B62[a][s_col]; columns past the subject end score -100 so
      // that nothing positive ever lives there
      for (int col = lane; col < bw; col += 32) {
        const int j = blk * bw + col;
        const int sj = j < m ? (int)s_lidx[fix_u(s[j])] : -1;
#pragma unroll 1
        for (int aa = 0; aa < 26; ++aa) prof[aa * PROF_COLS + col] = sj >= 0 ? s_b62[aa * 32 + sj] : (int8_t)-100;
      }
      __syncwarp();
      uint8_t *dirs = scratch + (size_t)blk * block_stride(n, cw);
      if (cw == 8) dp_block<8>(prof, s_lidx, q, n, blk, nblk, dirs, bnd, bnd + n, bnd + 2 * n, best_s, best_pos);
      else dp_block<4>(prof, s_lidx, q, n, blk, nblk, dirs, bnd, bnd + n, bnd + 2 * n, best_s, best_pos);
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
  if (lane != 0) return;
  // ---- traceback + align.go post-processing (lane 0) --------------------------------------
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
    uint32_t f = dir_at(scratch, n, cw, i, j);
    while (i > 0 && j > 0) {
      if (layer == 0) {
        if (f & 4u) break;
        const uint8_t ca = fix_u(q[i - 1]), cb = fix_u(s[j - 1]);
        if (cur_kind != 0) {
          flush();
          cur_kind = 0;
          cur_score = 0;
          cur_lq = cur_ls = 0;
        }
        cur_score += s_b62[s_lidx[ca] * 32 + s_lidx[cb]];
        cur_lq++;
        cur_ls++;
        if (cb == ca) {  // align.go:82-86
          identity += 1.f;
          similarity += 1.f;
        } else {
          if (cb != '-' && ca != '-') mismatches += 1;                      // align.go:88-90
          if (s_b62[s_apos[cb] * 32 + s_apos[ca]] > 0) similarity += 1.f;  // GetAlnScoreAA (:91)
        }
        nb_pos += 1.f;
        aln_len++;
        --i;
        --j;
        if (i > 0 && j > 0) {
          f = dir_at(scratch, n, cw, i, j);
          layer = (int)(f & 3u);
        }
      } else if (layer == 1) {
        if (cur_kind != 1) {
          flush();
          cur_kind = 1;
          cur_score = 0;
          cur_lq = cur_ls = 0;
        }
        cur_lq++;
        if (fix_u(q[i - 1]) == '-') {  // a literal '-' residue equals the gap character (align.go:82)
          identity += 1.f;
          similarity += 1.f;
        }
        nb_pos += 1.f;
        aln_len++;
        if (f & 8u) {
          cur_score += GAP_OPEN_DP;
          layer = 0;
        }
        --i;
        if (i > 0) f = dir_at(scratch, n, cw, i, j);
      } else {
        if (cur_kind != 2) {
          flush();
          cur_kind = 2;
          cur_score = 0;
          cur_lq = cur_ls = 0;
        }
        cur_ls++;
        if (fix_u(s[j - 1]) == '-') {
          identity += 1.f;
          similarity += 1.f;
        }
        nb_pos += 1.f;
        aln_len++;
        if (f & 16u) {
          cur_score += GAP_OPEN_DP;
          layer = 0;
        }
        --j;
        if (j > 0) f = dir_at(scratch, n, cw, i, j);
      }
    }
    flush();
    q_start = i;
    s_start = j;
  }
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

// ---------------------------------------------------------------------------------------
static bool g_tables_ready[64] = {false};

static int choose_cw(uint64_t m) {
  const uint64_t p8 = (m + 255) / 256 * 256, p4 = (m + 127) / 128 * 128;
  return p4 < p8 ? 4 : 8;
}

static uint64_t pair_scratch_bytes(uint64_t n, uint64_t m, int cw) {
  const uint64_t bw = 32ull * cw;
  const uint64_t nblk = (m + bw - 1) / bw;
  uint64_t b = nblk * (n + 31) * 32 * cw + 3 * 4 * n;
  return (b + 255) & ~255ull;
}

int align_pairs(kaamer_gpu *h, const uint8_t *q_res, const uint64_t *q_off, const uint32_t *pair_q,
                const uint32_t *pair_s, uint32_t n_pairs, const kaamer_aln_opts *o, kaamer_aln *out) {
  DevIndex &ix = h->idx;
  cudaStream_t st = h->stream;
  if (!ix.has_proteins) {
    set_error("kaamer_gpu_align: the index was opened without a protein table");
    return KAAMER_ERR_ARG;
  }
  if (n_pairs == 0) return KAAMER_OK;
  if (h->device < 64 && !g_tables_ready[h->device]) {
    AlnTables t;
    make_tables(&t);
    KCUDA(cudaMemcpyToSymbol(c_aln, &t, sizeof t));
    g_tables_ready[h->device] = true;
  }
  uint32_t nq = 0;
  for (uint32_t i = 0; i < n_pairs; ++i) nq = pair_q[i] + 1 > nq ? pair_q[i] + 1 : nq;
  const uint64_t n_qres = q_off[nq];
  // cost-ordered pair list (long pairs first: short tail, similar pairs share a CTA)
  std::vector<uint64_t> cost(n_pairs);
  uint64_t max_cost = 1;
  for (uint32_t i = 0; i < n_pairs; ++i) {
    if (pair_s[i] > ix.max_protein_id) {
      set_error("pair %u: subject id %u not in the protein table (max %u)", i, pair_s[i], ix.max_protein_id);
      return KAAMER_ERR_ARG;
    }
    const uint64_t n = q_off[pair_q[i] + 1] - q_off[pair_q[i]];
    const uint64_t m = ix.h_prot_off[pair_s[i] + 1] - ix.h_prot_off[pair_s[i]];
    if (n >= 65535 || m >= 65535) {
      set_error("pair %u: sequences longer than 65534 residues are not supported (%llu x %llu)", i,
                (unsigned long long)n, (unsigned long long)m);
      return KAAMER_ERR_LIMIT;
    }
    cost[i] = n * m;
    max_cost = cost[i] > max_cost ? cost[i] : max_cost;
  }
  constexpr int NB = 1024;
  std::vector<uint32_t> bucket_start(NB + 1, 0), order(n_pairs);
  auto bucket_of = [&](uint64_t c) { return (uint32_t)(NB - 1 - (c * (NB - 1)) / max_cost); };  // descending cost
  for (uint32_t i = 0; i < n_pairs; ++i) bucket_start[bucket_of(cost[i]) + 1]++;
  for (int b = 0; b < NB; ++b) bucket_start[b + 1] += bucket_start[b];
  {
    std::vector<uint32_t> cur(bucket_start.begin(), bucket_start.end() - 1);
    for (uint32_t i = 0; i < n_pairs; ++i) order[cur[bucket_of(cost[i])]++] = i;
  }
  // device buffers
  uint8_t *d_q = nullptr, *d_scratch = nullptr;
  uint64_t *d_qoff = nullptr;
  AlnPair *d_pairs = nullptr;
  kaamer_aln *d_out = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_q);
    cudaFree(d_qoff);
    cudaFree(d_pairs);
    cudaFree(d_out);
    cudaFree(d_scratch);
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
  ACUDA(cudaMalloc((void **)&d_q, (size_t)n_qres + 16));
  ACUDA(cudaMalloc((void **)&d_qoff, ((size_t)nq + 1) * 8));
  ACUDA(cudaMalloc((void **)&d_pairs, (size_t)n_pairs * sizeof(AlnPair)));
  ACUDA(cudaMalloc((void **)&d_out, (size_t)n_pairs * sizeof(kaamer_aln)));
  ACUDA(cudaMemcpyAsync(d_q, q_res, (size_t)n_qres, cudaMemcpyHostToDevice, st));
  ACUDA(cudaMemcpyAsync(d_qoff, q_off, ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, st));
  // chunks under the traceback-memory budget
  size_t free_b = 0, total_b = 0;
  ACUDA(cudaMemGetInfo(&free_b, &total_b));
  uint64_t budget = free_b / 2;
  if (budget > (24ull << 30)) budget = 24ull << 30;
  std::vector<AlnPair> pairs(n_pairs);
  std::vector<uint32_t> chunk_end;
  uint64_t used = 0, max_used = 0;
  for (uint32_t k = 0; k < n_pairs; ++k) {
    const uint32_t i = order[k];
    const uint64_t n = q_off[pair_q[i] + 1] - q_off[pair_q[i]];
    const uint64_t m = ix.h_prot_off[pair_s[i] + 1] - ix.h_prot_off[pair_s[i]];
    const int cw = choose_cw(m);
    const uint64_t b = pair_scratch_bytes(n, m, cw);
    if (b > budget) {
      set_error("pair %u (%llu x %llu) needs %llu bytes of traceback state, more than the device has free", i,
                (unsigned long long)n, (unsigned long long)m, (unsigned long long)b);
      cleanup();
      return KAAMER_ERR_NOMEM;
    }
    if (used + b > budget) {
      chunk_end.push_back(k);
      used = 0;
    }
    pairs[k] = AlnPair{i, pair_q[i], pair_s[i], (uint32_t)cw, used};
    used += b;
    max_used = used > max_used ? used : max_used;
  }
  chunk_end.push_back(n_pairs);
  ACUDA(cudaMalloc((void **)&d_scratch, (size_t)max_used + 256));
  ACUDA(cudaMemcpyAsync(d_pairs, pairs.data(), (size_t)n_pairs * sizeof(AlnPair), cudaMemcpyHostToDevice, st));
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
  uint32_t begin = 0;
  for (uint32_t end : chunk_end) {
    if (end > begin) {
      a.pairs = d_pairs + begin;
      a.n_pairs = end - begin;
      profile_begin(h, st, 3);
      k_sw_affine<<<(a.n_pairs + ALN_WARPS - 1) / ALN_WARPS, ALN_WARPS * 32, 0, st>>>(a);
      profile_end(h, st);
      h->prof_all_launches += 1;
      ACUDA(cudaGetLastError());
    }
    begin = end;
  }
  ACUDA(cudaMemcpyAsync(out, d_out, (size_t)n_pairs * sizeof(kaamer_aln), cudaMemcpyDeviceToHost, st));
  ACUDA(cudaStreamSynchronize(st));
  cleanup();
#undef ACUDA
  return KAAMER_OK;
}

}  // namespace kaamer

using namespace kaamer;

extern "C" int kaamer_gpu_align(kaamer_gpu_t *h, const uint8_t *q_residues, const uint64_t *q_off,
                                const uint32_t *pair_query, const uint32_t *pair_subject, uint32_t n_pairs,
                                const kaamer_aln_opts *opts, kaamer_aln *out) {
  if (!h || !opts || (n_pairs && (!q_residues || !q_off || !pair_query || !pair_subject || !out))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  return align_pairs(h, q_residues, q_off, pair_query, pair_subject, n_pairs, opts, out);
}
