// translate.cu — six-frame translation and ORF extraction on the device.
//
// Replaces GetORFs / GetFrame / ReverseComplement (pkg/search/dna.go:55-196) with table 11
// (gcodeBacteria, pkg/search/gcode.go:36-101; the geneticCode argument is ignored by the
// reference, dna.go:106).
//
//   k_translate6   one thread per nucleotide position p of a contig: the triple dna[p..p+2]
//                  is the codon p/3 of plus frame p%3 AND (complemented, reversed) the codon
//                  (L-3-p)/3 of minus frame (L-3-p)%3.  One byte per codon is written to the
//                  frame-major codon arrays: bits 0..6 amino-acid letter (0 = codon not in the
//                  table, i.e. it holds a non-acgt byte: the Go map miss appends nothing,
//                  dna.go:106,123), bit 7 = start codon.
//   k_orf_ends     every stop codon and every frame-final codon closes at most one ORF: the
//                  thread walks back to the previous stop, remembering the left-most start
//                  codon of the stretch (the frame-initial stretch starts inside an ORF,
//                  dna.go:98), counts residues and alternative starts, and emits the ORF when
//                  it has >= 21 residues (dna.go:26,128,155).
//   radix sort     by (contig, End (+) / Start (-), strand): the order of sort.Slice at
//                  dna.go:167-177 with ties broken by emission order (ties only occur between
//                  a plus and a minus ORF, the plus one was emitted first).
//   k_orf_write    one warp per ORF: residues (empty codons skipped) and StartsAlternative.
//
// HBM traffic: 1 B read per nucleotide (+2 halo), 2 B written per nucleotide (six frames x 1/3),
// the codon arrays are re-read once by k_orf_ends and once by k_orf_write.
#include <cub/cub.cuh>

#include "internal.cuh"

namespace kaamer {

constexpr int MIN_LEN_CDS = 21;  // dna.go:26

// table 11 (gcodeBacteria — the table GetORFs always uses, dna.go:106), codon index = 16*b0 + 4*b1 + b2
// with t=0 c=1 a=2 g=3; start codons ttg ctg att atc ata atg gtg (gcode.go:40,56,69-72,88).  The table is a
// per-handle parameter (kaamer_gpu_set_genetic_code, SURVEY §8f-4: the other tables of gcode.go behind an
// explicit call); the default is table 11.
static const char TABLE11_AAS[65] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
constexpr uint64_t TABLE11_STARTS = (1ull << 3) | (1ull << 19) | (1ull << 32) | (1ull << 33) | (1ull << 34) |
                                    (1ull << 35) | (1ull << 51);
struct GCodeArg {
  char aas[64];
  uint64_t starts;
};

__device__ __forceinline__ int base_code(uint32_t c) {
  c |= 0x20u;  // strings.ToLower (dna.go:68) as far as a/c/g/t are concerned
  return c == 't' ? 0 : c == 'c' ? 1 : c == 'a' ? 2 : c == 'g' ? 3 : -1;
}

// contig of global nucleotide index g: last c with coff[c] <= g
__device__ __forceinline__ uint32_t find_contig(const uint64_t *__restrict__ coff, uint32_t nc, uint64_t g) {
  uint32_t lo = 0, hi = nc;  // invariant coff[lo] <= g < coff[hi]
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (coff[mid] <= g) lo = mid;
    else hi = mid;
  }
  return lo;
}

struct TranslateArgs {
  GCodeArg gc;
  const uint8_t *nt;
  const uint64_t *coff;   // [nc+1] nucleotide offsets
  const uint64_t *cbase;  // [nc+1] codon-array offsets: contig c owns 6 frames of stride L/3+1
  uint32_t nc;
  uint64_t total_nt;
  uint8_t *cod;
  // unsorted ORF records
  unsigned long long *n_orfs;  // [0] ORF records, [1] end codons seen, [2] end list overflowed
  uint64_t cap;
  // compacted list of "end" codons (stops and frame-final codons) as global codon indices: the ORF
  // walk runs one thread per END instead of one per codon (5 % of them), with all lanes busy
  uint32_t *ends;
  uint64_t ends_cap;
  uint64_t *key;      // sort key
  uint32_t *r_contig;
  int32_t *r_b, *r_e, *r_cnt, *r_nalt;
  uint8_t *r_frame;
};

__global__ void __launch_bounds__(256) k_translate6(TranslateArgs a) {
  __shared__ char aas[64];
  if (threadIdx.x < 64) aas[threadIdx.x] = a.gc.aas[threadIdx.x];
  __syncthreads();
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool end_p = false, end_m = false;
  uint64_t gi_p = 0, gi_m = 0;
  uint32_t c = 0;
  int64_t L = 0, p = 0;
  if (g < a.total_nt) {
    c = find_contig(a.coff, a.nc, g);
    const uint64_t cb = a.coff[c];
    L = (int64_t)(a.coff[c + 1] - cb);
    p = (int64_t)(g - cb);
  }
  // Go would panic on len < 2 (dna.go:183-196): nothing emitted
  if (g < a.total_nt && L >= 2 && p + 3 <= L) {
  const int b0 = base_code(a.nt[g]), b1 = base_code(a.nt[g + 1]), b2 = base_code(a.nt[g + 2]);
  const bool ok = (b0 | b1 | b2) >= 0;
  const int64_t S = L / 3 + 1;
  uint8_t *cod = a.cod + a.cbase[c];
  {
    uint8_t v = 0;
    if (ok) {
      const int idx = b0 * 16 + b1 * 4 + b2;
      v = (uint8_t)aas[idx] | (uint8_t)(((a.gc.starts >> idx) & 1ull) << 7);
    }
    cod[(p % 3) * S + p / 3] = v;
    end_p = (v & 0x7F) == '*' || p / 3 == (L - p % 3) / 3 - 1;
    gi_p = a.cbase[c] + (uint64_t)((p % 3) * S + p / 3);
  }
  {
    // reverse complement (dna.go:55-63): only a<->t, c<->g are swapped = code ^ 2
    const int64_t j = L - 3 - p;
    uint8_t v = 0;
    if (ok) {
      const int idx = (b2 ^ 2) * 16 + (b1 ^ 2) * 4 + (b0 ^ 2);
      v = (uint8_t)aas[idx] | (uint8_t)(((a.gc.starts >> idx) & 1ull) << 7);
    }
    cod[(3 + j % 3) * S + j / 3] = v;
    end_m = (v & 0x7F) == '*' || j / 3 == (L - j % 3) / 3 - 1;
    gi_m = a.cbase[c] + (uint64_t)((3 + j % 3) * S + j / 3);
  }
  }
  // CTA-aggregated append of the end codons: one global atomic per block (one per warp made
  // the kernel 3.5x slower: ~600 k atomics on one address)
  __shared__ uint32_t s_cnt[2][8];
  __shared__ unsigned long long s_base;
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned mp = __ballot_sync(0xFFFFFFFFu, end_p), mm = __ballot_sync(0xFFFFFFFFu, end_m);
  if (lane == 0) {
    s_cnt[0][w] = __popc(mp);
    s_cnt[1][w] = __popc(mm);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int t = 0; t < 2; ++t)
      for (int i = 0; i < 8; ++i) {
        const uint32_t c2 = s_cnt[t][i];
        s_cnt[t][i] = tot;  // exclusive offset of (strand t, warp i) inside the block
        tot += c2;
      }
    s_base = tot ? atomicAdd(a.n_orfs + 1, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  const unsigned below = (1u << lane) - 1u;
  if (end_p) {
    const unsigned long long slot = s_base + s_cnt[0][w] + __popc(mp & below);
    if (slot < a.ends_cap) a.ends[slot] = (uint32_t)gi_p;
    else a.n_orfs[2] = 1;
  }
  if (end_m) {
    const unsigned long long slot = s_base + s_cnt[1][w] + __popc(mm & below);
    if (slot < a.ends_cap) a.ends[slot] = (uint32_t)gi_m;
    else a.n_orfs[2] = 1;
  }
}

struct OrfRec {
  int64_t b, e;
  int32_t cnt, nalt;
  bool valid;
};

__device__ __forceinline__ OrfRec orf_close(int64_t e, const uint8_t *__restrict__ fr) {
  // Walk back from e to the previous stop (exclusive) or to the frame start, eight codons per
  // aligned 64-bit load; the bytes are classified with SWAR arithmetic (all codon bytes have
  // their letter in bits 0..6, so `x + 0x7f` sets bit 7 exactly for non-zero letters).
  constexpr unsigned long long LO7 = 0x7F7F7F7F7F7F7F7Full, HI1 = 0x8080808080808080ull,
                               STOPS = 0x2A2A2A2A2A2A2A2Aull;  // '*'
  int32_t cnt = 0, nalt = 0, cnt_b = 0, nalt_b = 0;
  int64_t b = -1, k = e;
  bool hit_stop = false;
  while (k >= 0) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(fr + k);
    const int hi = (int)(addr & 7);          // byte of codon k inside its aligned word
    const int64_t k0 = k - hi;               // codon of byte 0
    const int lo = k0 < 0 ? (int)(-k0) : 0;  // bytes below belong to another array
    unsigned long long w = *reinterpret_cast<const unsigned long long *>(addr - hi);
    unsigned long long valid = (hi == 7 ? ~0ull : ((1ull << (8 * (hi + 1))) - 1ull)) & (~0ull << (8 * lo));
    w &= valid;
    unsigned long long aa = w & LO7;
    unsigned long long stopm = ~((aa ^ STOPS) + LO7) & HI1 & valid;
    if (k == e) stopm &= ~(0x80ull << (8 * hi));  // the closing codon itself is not a boundary
    if (stopm) {                                  // nearest stop below k: keep only the bytes above it
      const int js = (63 - __clzll((long long)stopm)) >> 3;
      const unsigned long long above = js == 7 ? 0ull : (~0ull << (8 * (js + 1)));
      w &= above;
      aa &= above;
      hit_stop = true;
    }
    const unsigned long long nz = (aa + LO7) & HI1;
    const unsigned long long st = w & HI1;
    const int nst = __popcll(st);
    if (st) {  // left-most start codon seen so far
      const int jb = (__ffsll((long long)st) - 1) >> 3;
      b = k0 + jb;
      cnt_b = cnt + __popcll(nz >> (8 * jb));
      nalt_b = nalt + nst;
    }
    cnt += __popcll(nz);
    nalt += nst;
    if (hit_stop) break;
    k = k0 - 1;
  }
  if (!hit_stop) {  // frame-initial stretch: insideORF starts true (dna.go:98)
    b = 0;
    cnt_b = cnt;
    nalt_b = nalt;
  }
  OrfRec r;
  r.b = b;
  r.e = e;
  r.cnt = cnt_b;
  r.nalt = nalt_b;
  r.valid = b >= 0 && cnt_b >= MIN_LEN_CDS;
  return r;
}

// warp-aggregated append of the ORF records (one atomic per warp instead of one per ORF: 200 k
// same-address atomics were the whole cost of the kernel)
__device__ __forceinline__ void orf_emit(const TranslateArgs &a, const OrfRec &r, uint32_t c, int64_t L, int frame) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned mask = __ballot_sync(0xFFFFFFFFu, r.valid);
  if (mask == 0) return;
  const int leader = __ffs(mask) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(a.n_orfs, (unsigned long long)__popc(mask));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (!r.valid) return;
  const unsigned long long slot = base + __popc(mask & ((1u << lane) - 1u));
  if (slot >= a.cap) return;  // cannot happen (cap = codons/21 + slack); checked on the host
  int64_t start, end, poskey;
  if (frame < 3) {
    start = frame + 3 * r.b + 1;  // dna.go:84,111
    end = 3 * r.e + 3 + frame;    // dna.go:129,156
    poskey = end;
  } else {
    start = L - (frame - 3) - 3 * r.b;       // dna.go:80-82,112-114
    end = start - 3 * (int64_t)r.cnt + 1;   // dna.go:131,158
    poskey = start;
  }
  (void)end;
  a.key[slot] = ((uint64_t)c << 40) | ((uint64_t)poskey << 1) | (frame >= 3 ? 1ull : 0ull);
  a.r_contig[slot] = c;
  a.r_frame[slot] = (uint8_t)frame;
  a.r_b[slot] = (int32_t)r.b;
  a.r_e[slot] = (int32_t)r.e;
  a.r_cnt[slot] = r.cnt;
  a.r_nalt[slot] = r.nalt;
}

// one thread per end codon of the compacted list
__global__ void __launch_bounds__(256) k_orf_ends_list(TranslateArgs a) {
  if (a.n_orfs[2]) return;  // list overflowed: the dense kernel below does the work
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  OrfRec r;
  r.valid = false;
  uint32_t c = 0;
  int64_t L = 0;
  int frame = 0;
  if (i < a.n_orfs[1]) {
    const uint64_t gi = a.ends[i];
    c = find_contig(a.cbase, a.nc, gi);  // cbase is increasing like coff
    L = (int64_t)(a.coff[c + 1] - a.coff[c]);
    const int64_t S = L / 3 + 1;
    const uint64_t o = gi - a.cbase[c];
    frame = (int)(o / (uint64_t)S);
    const int64_t k = (int64_t)(o % (uint64_t)S);
    r = orf_close(k, a.cod + a.cbase[c] + (int64_t)frame * S);
  }
  orf_emit(a, r, c, L, frame);
}

// one thread per codon (fallback when the end list overflowed: > 12.5 % of the codons are ends)
__global__ void __launch_bounds__(256) k_orf_ends(TranslateArgs a) {
  if (!a.n_orfs[2]) return;
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  OrfRec rp, rm;
  rp.valid = rm.valid = false;
  uint32_t c = 0;
  int64_t L = 0;
  int fp = 0, fm = 3;
  if (g < a.total_nt) {
    c = find_contig(a.coff, a.nc, g);
    const uint64_t cb = a.coff[c];
    L = (int64_t)(a.coff[c + 1] - cb);
    const int64_t p = (int64_t)(g - cb);
    if (L >= 2 && p + 3 <= L) {
      const int64_t S = L / 3 + 1;
      const uint8_t *cod = a.cod + a.cbase[c];
      {
        fp = (int)(p % 3);
        const int64_t k = p / 3, ncod = (L - fp) / 3;
        const uint8_t *fr = cod + fp * S;
        if ((fr[k] & 0x7F) == '*' || k == ncod - 1) rp = orf_close(k, fr);
      }
      {
        const int64_t j = L - 3 - p;
        const int f = (int)(j % 3);
        fm = 3 + f;
        const int64_t k = j / 3, ncod = (L - f) / 3;
        const uint8_t *fr = cod + (3 + f) * S;
        if ((fr[k] & 0x7F) == '*' || k == ncod - 1) rm = orf_close(k, fr);
      }
    }
  }
  orf_emit(a, rp, c, L, fp);
  orf_emit(a, rm, c, L, fm);
}

__global__ void k_iota(uint32_t *p, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}

struct OrfWriteArgs {
  const uint8_t *cod;
  const uint64_t *coff, *cbase;
  const uint32_t *perm;  // sorted rank -> record slot
  const uint32_t *r_contig;
  const int32_t *r_b, *r_e, *r_cnt, *r_nalt;
  const uint8_t *r_frame;
  uint64_t n;
  // outputs (sorted order)
  uint32_t *contig;
  int64_t *start, *end;
  uint8_t *plus;
  uint64_t *len, *nalt;  // per-ORF sizes (scanned into seq_off / alts_off afterwards)
  const uint64_t *seq_off, *alts_off;
  uint8_t *seq;
  int32_t *alts;
};

__global__ void k_orf_meta(OrfWriteArgs a) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n) return;
  const uint32_t s = a.perm[r];
  const uint32_t c = a.r_contig[s];
  const int frame = a.r_frame[s];
  const int64_t L = (int64_t)(a.coff[c + 1] - a.coff[c]);
  const int64_t b = a.r_b[s], e = a.r_e[s], cnt = a.r_cnt[s];
  int64_t start, end;
  if (frame < 3) {
    start = frame + 3 * b + 1;
    end = 3 * e + 3 + frame;
  } else {
    start = L - (frame - 3) - 3 * b;
    end = start - 3 * cnt + 1;
  }
  a.contig[r] = c;
  a.start[r] = start;
  a.end[r] = end;
  a.plus[r] = frame < 3 ? 1 : 0;
  a.len[r] = (uint64_t)cnt;
  a.nalt[r] = (uint64_t)a.r_nalt[s];
}

__global__ void __launch_bounds__(256) k_orf_write(OrfWriteArgs a) {
  const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (r >= a.n) return;
  const uint32_t s = a.perm[r];
  const uint32_t c = a.r_contig[s];
  const int frame = a.r_frame[s];
  const int64_t L = (int64_t)(a.coff[c + 1] - a.coff[c]);
  const int64_t S = L / 3 + 1;
  const uint8_t *fr = a.cod + a.cbase[c] + (int64_t)frame * S;
  const int64_t b = a.r_b[s], e = a.r_e[s];
  uint8_t *seq = a.seq + a.seq_off[r];
  int32_t *alts = a.alts + a.alts_off[r];
  uint32_t nseq = 0, nal = 0;
  for (int64_t kb = b; kb <= e; kb += 32) {
    const int64_t k = kb + lane;
    const uint8_t v = k <= e ? fr[k] : 0;
    const unsigned has = __ballot_sync(0xFFFFFFFFu, (v & 0x7F) != 0);
    const unsigned st = __ballot_sync(0xFFFFFFFFu, (v & 0x80) != 0);
    const unsigned below = (1u << lane) - 1u;
    if (v & 0x7F) seq[nseq + __popc(has & below)] = v & 0x7F;
    if (v & 0x80) alts[nal + __popc(st & below)] = (int32_t)(k - b);  // currentAAPos (dna.go:115,118)
    nseq += __popc(has);
    nal += __popc(st);
  }
}

// ---------------------------------------------------------------------------------------
// the ORF arrays live in the handle's arena (valid until the next call resets it)
void orfset_release(OrfSet *o) { *o = OrfSet(); }

// d_nt: nucleotides on the device (+2 readable bytes of slack are NOT required: the last two
// positions of a contig are never dereferenced beyond the contig end);  h_coff: host copy.
int orfs_device(kaamer_gpu *h, const uint8_t *d_nt, const uint64_t *h_coff, uint32_t nc, OrfSet *out,
                cudaStream_t st) {
  *out = OrfSet();
  const uint64_t total_nt = nc ? h_coff[nc] : 0;
  if (nc >= (1u << 24)) {
    set_error("too many contigs in one batch (%u >= 2^24)", nc);
    return KAAMER_ERR_LIMIT;
  }
  std::vector<uint64_t> cbase((size_t)nc + 1, 0);
  for (uint32_t c = 0; c < nc; ++c) {
    if (h_coff[c + 1] < h_coff[c]) {
      set_error("contig offsets must be non-decreasing");
      return KAAMER_ERR_ARG;
    }
    const uint64_t L = h_coff[c + 1] - h_coff[c];
    if (L >= (1ull << 38)) {
      set_error("contig %u too long", c);
      return KAAMER_ERR_LIMIT;
    }
    cbase[c + 1] = cbase[c] + 6 * (L / 3 + 1);
  }
  const uint64_t total_cod = cbase[nc];
  const uint64_t cap = total_cod / MIN_LEN_CDS + 6ull * nc + 64;
  uint64_t *d_coff = nullptr, *d_cbase = nullptr, *d_key = nullptr, *d_key2 = nullptr;
  uint8_t *d_cod = nullptr, *r_frame = nullptr;
  unsigned long long *d_n = nullptr;
  uint32_t *r_contig = nullptr, *d_iota = nullptr, *d_perm = nullptr;
  int32_t *r_b = nullptr, *r_e = nullptr, *r_cnt = nullptr, *r_nalt = nullptr;
  uint64_t *d_len = nullptr, *d_nalt = nullptr;
  void *d_tmp = nullptr;
  int rc = KAAMER_OK;
  auto cleanup = [&]() {};
#define TCHECK(x)                \
  do {                           \
    rc = (x);                    \
    if (rc != KAAMER_OK) {       \
      cleanup();                 \
      orfset_release(out);       \
      return rc;                 \
    }                            \
  } while (0)
#define TCUDA(call)                                                                      \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));   \
      cleanup();                                                                         \
      orfset_release(out);                                                               \
      return KAAMER_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)
  TCHECK(h->arena.get(&d_coff, (size_t)nc + 1));
  TCHECK(h->arena.get(&d_cbase, (size_t)nc + 1));
  TCHECK(h->arena.get(&d_cod, (size_t)total_cod));
  TCHECK(h->arena.get(&d_n, 4));
  if (total_cod >= (1ull << 32)) {
    set_error("batch too large: %llu codons (split the contigs over several calls)", (unsigned long long)total_cod);
    return KAAMER_ERR_LIMIT;
  }
  const uint64_t ends_cap = total_cod / 8 + 1024;
  uint32_t *d_ends = nullptr;
  TCHECK(h->arena.get(&d_ends, (size_t)ends_cap));
  TCHECK(h->arena.get(&d_key, (size_t)cap));
  TCHECK(h->arena.get(&r_contig, (size_t)cap));
  TCHECK(h->arena.get(&r_frame, (size_t)cap));
  TCHECK(h->arena.get(&r_b, (size_t)cap));
  TCHECK(h->arena.get(&r_e, (size_t)cap));
  TCHECK(h->arena.get(&r_cnt, (size_t)cap));
  TCHECK(h->arena.get(&r_nalt, (size_t)cap));
  TCUDA(cudaMemcpyAsync(d_coff, h_coff, ((size_t)nc + 1) * 8, cudaMemcpyHostToDevice, st));
  TCUDA(cudaMemcpyAsync(d_cbase, cbase.data(), ((size_t)nc + 1) * 8, cudaMemcpyHostToDevice, st));
  TCUDA(cudaMemsetAsync(d_n, 0, 32, st));
  TranslateArgs ta{};
  if (h->gcode_set) {
    memcpy(ta.gc.aas, h->gcode_aas, 64);
    ta.gc.starts = h->gcode_starts;
  } else {
    memcpy(ta.gc.aas, TABLE11_AAS, 64);
    ta.gc.starts = TABLE11_STARTS;
  }
  ta.nt = d_nt;
  ta.coff = d_coff;
  ta.cbase = d_cbase;
  ta.nc = nc;
  ta.total_nt = total_nt;
  ta.cod = d_cod;
  ta.n_orfs = d_n;
  ta.cap = cap;
  ta.ends = d_ends;
  ta.ends_cap = ends_cap;
  ta.key = d_key;
  ta.r_contig = r_contig;
  ta.r_frame = r_frame;
  ta.r_b = r_b;
  ta.r_e = r_e;
  ta.r_cnt = r_cnt;
  ta.r_nalt = r_nalt;
  unsigned long long n_orfs = 0;
  if (total_nt) {
    const unsigned grid = (unsigned)((total_nt + 255) / 256);
    profile_begin(h, st, 7);
    k_translate6<<<grid, 256, 0, st>>>(ta);
    k_orf_ends_list<<<(unsigned)((ends_cap + 255) / 256), 256, 0, st>>>(ta);
    k_orf_ends<<<grid, 256, 0, st>>>(ta);
    profile_end(h, st);
    h->prof_all_launches += 3;
    TCUDA(cudaGetLastError());
    TCUDA(cudaMemcpyAsync(&n_orfs, d_n, 8, cudaMemcpyDeviceToHost, st));
    TCUDA(cudaStreamSynchronize(st));
  }
  if (n_orfs > cap) {
    set_error("internal: ORF record capacity exceeded (%llu > %llu)", n_orfs, (unsigned long long)cap);
    cleanup();
    return KAAMER_ERR_LIMIT;
  }
  const uint64_t n = n_orfs;
  out->n = n;
  TCHECK(h->arena.get(&out->contig, (size_t)n));
  TCHECK(h->arena.get(&out->start, (size_t)n));
  TCHECK(h->arena.get(&out->end, (size_t)n));
  TCHECK(h->arena.get(&out->plus, (size_t)n));
  TCHECK(h->arena.get(&out->seq_off, (size_t)n + 1));
  TCHECK(h->arena.get(&out->alts_off, (size_t)n + 1));
  if (n == 0) {
    TCUDA(cudaMemsetAsync(out->seq_off, 0, 8, st));
    TCUDA(cudaMemsetAsync(out->alts_off, 0, 8, st));
    TCHECK(h->arena.get(&out->seq, 16));
    TCHECK(h->arena.get(&out->alts, 1));
    TCUDA(cudaStreamSynchronize(st));
    cleanup();
    return KAAMER_OK;
  }
  // sort the records
  TCHECK(h->arena.get(&d_key2, (size_t)n));
  TCHECK(h->arena.get(&d_iota, (size_t)n));
  TCHECK(h->arena.get(&d_perm, (size_t)n));
  TCHECK(h->arena.get(&d_len, (size_t)n + 1));
  TCHECK(h->arena.get(&d_nalt, (size_t)n + 1));
  profile_begin(h, st, 7);
  k_iota<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_iota, n);
  TCUDA(cudaGetLastError());
  size_t need = 0, need2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, d_key, d_key2, d_iota, d_perm, (int64_t)n, 0, 64, st);
  cub::DeviceScan::ExclusiveSum(nullptr, need2, d_len, out->seq_off, (int64_t)n + 1, st);
  if (need2 > need) need = need2;
  TCHECK(h->arena.get((uint8_t **)&d_tmp, need + 16));
  size_t tb = need;
  TCUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_key, d_key2, d_iota, d_perm, (int64_t)n, 0, 64, st));
  OrfWriteArgs wa{};
  wa.cod = d_cod;
  wa.coff = d_coff;
  wa.cbase = d_cbase;
  wa.perm = d_perm;
  wa.r_contig = r_contig;
  wa.r_frame = r_frame;
  wa.r_b = r_b;
  wa.r_e = r_e;
  wa.r_cnt = r_cnt;
  wa.r_nalt = r_nalt;
  wa.n = n;
  wa.contig = out->contig;
  wa.start = out->start;
  wa.end = out->end;
  wa.plus = out->plus;
  wa.len = d_len;
  wa.nalt = d_nalt;
  TCUDA(cudaMemsetAsync(d_len + n, 0, 8, st));
  TCUDA(cudaMemsetAsync(d_nalt + n, 0, 8, st));
  k_orf_meta<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(wa);
  TCUDA(cudaGetLastError());
  tb = need;
  TCUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_len, out->seq_off, (int64_t)n + 1, st));
  tb = need;
  TCUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_nalt, out->alts_off, (int64_t)n + 1, st));
  uint64_t tot[2];
  TCUDA(cudaMemcpyAsync(&tot[0], out->seq_off + n, 8, cudaMemcpyDeviceToHost, st));
  TCUDA(cudaMemcpyAsync(&tot[1], out->alts_off + n, 8, cudaMemcpyDeviceToHost, st));
  TCUDA(cudaStreamSynchronize(st));
  out->n_seq = tot[0];
  out->n_alts = tot[1];
  TCHECK(h->arena.get(&out->seq, (size_t)tot[0] + 16));
  TCHECK(h->arena.get(&out->alts, (size_t)tot[1] + 1));
  wa.seq_off = out->seq_off;
  wa.alts_off = out->alts_off;
  wa.seq = out->seq;
  wa.alts = out->alts;
  k_orf_write<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(wa);
  profile_end(h, st);
  h->prof_all_launches += 2;
  TCUDA(cudaGetLastError());
  TCUDA(cudaStreamSynchronize(st));
  cleanup();
#undef TCHECK
#undef TCUDA
  return KAAMER_OK;
}

}  // namespace kaamer

using namespace kaamer;

extern "C" {

static int kaamer_gpu_get_orfs_impl(kaamer_gpu_t *h, const uint8_t *nt, const uint64_t *contig_off, uint32_t n_contigs,
                        kaamer_orfs **out) {
  if (!h || !out || (n_contigs && (!nt || !contig_off))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  h->arena.reset();
  const uint64_t zero = 0;
  const uint64_t *coff = n_contigs ? contig_off : &zero;
  if (coff[0] != 0) {
    set_error("contig_off[0] must be 0");
    return KAAMER_ERR_ARG;
  }
  const uint64_t total = coff[n_contigs];
  KCHECK(h->ws.residues.ensure((size_t)total + 16));
  if (total) KCUDA(cudaMemcpyAsync(h->ws.residues.p, nt, (size_t)total, cudaMemcpyHostToDevice, st));
  OrfSet os;
  KCHECK(orfs_device(h, h->ws.residues.p, coff, n_contigs, &os, st));
  auto *o = new kaamer_orfs();
  memset(o, 0, sizeof *o);
  auto *owner = new HitsOwner();
  o->_owner = owner;
  o->n_orfs = os.n;
  int rc = KAAMER_OK;
  const size_t n = (size_t)os.n;
  if ((rc = owner->alloc(&o->contig, n)) == KAAMER_OK && (rc = owner->alloc(&o->start, n)) == KAAMER_OK &&
      (rc = owner->alloc(&o->end, n)) == KAAMER_OK && (rc = owner->alloc(&o->plus, n)) == KAAMER_OK &&
      (rc = owner->alloc(&o->seq_off, n + 1)) == KAAMER_OK && (rc = owner->alloc(&o->seq, (size_t)os.n_seq)) == KAAMER_OK &&
      (rc = owner->alloc(&o->alts_off, n + 1)) == KAAMER_OK &&
      (rc = owner->alloc(&o->alts, (size_t)os.n_alts)) == KAAMER_OK) {
    cudaError_t e = cudaSuccess;
    auto cp = [&](void *dst, const void *src, size_t bytes) {
      if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
    };
    cp(o->contig, os.contig, n * 4);
    cp(o->start, os.start, n * 8);
    cp(o->end, os.end, n * 8);
    cp(o->plus, os.plus, n);
    cp(o->seq_off, os.seq_off, (n + 1) * 8);
    cp(o->seq, os.seq, (size_t)os.n_seq);
    cp(o->alts_off, os.alts_off, (n + 1) * 8);
    cp(o->alts, os.alts, (size_t)os.n_alts * 4);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("get_orfs: D2H failed: %s", cudaGetErrorString(e));
      rc = KAAMER_ERR_CUDA;
    }
  }
  orfset_release(&os);
  if (rc != KAAMER_OK) {
    delete owner;
    delete o;
    return rc;
  }
  *out = o;
  return KAAMER_OK;
}
int kaamer_gpu_get_orfs(kaamer_gpu_t *h, const uint8_t *nt, const uint64_t *contig_off, uint32_t n_contigs,
                        kaamer_orfs **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_get_orfs_impl(h, nt, contig_off, n_contigs, out); });
}

void kaamer_gpu_free_orfs(kaamer_orfs *o) {
  if (!o) return;
  delete (HitsOwner *)o->_owner;
  delete o;
}

}  // extern "C"

extern "C" int kaamer_gpu_set_genetic_code(kaamer_gpu_t *h, const char *aas64, uint64_t start_mask) {
  if (!h) {
    kaamer::set_error("null handle");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  if (!aas64) {
    h->gcode_set = false;
    return KAAMER_OK;
  }
  for (int i = 0; i < 64; ++i) {
    const char c = aas64[i];
    if (!((c >= 'A' && c <= 'Z') || c == '*')) {
      kaamer::set_error("genetic code: amino-acid letter %d is not A-Z or '*'", i);
      return KAAMER_ERR_ARG;
    }
  }
  memcpy(h->gcode_aas, aas64, 64);
  h->gcode_starts = start_mask;
  h->gcode_set = true;
  return KAAMER_OK;
}
