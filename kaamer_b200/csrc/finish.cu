// finish.cu — everything after the count pass that needs query positions.
//
//   k_pos_need     words of position bit-sets per query (n_hits x ceil(SizeInKmer/32))
//   k_positions    one warp per query that kept hits: PositionHits of every kept hit
//                  (StoreMatchPositions, pkg/search/search.go:442-452) recomputed as "is the
//                  subject in the posting list of the query k-mer at position k" — one table
//                  probe per position, shared by all hits of the query; then, for nucleotide /
//                  reads queries, SetBestStartCodon (pkg/search/dna.go:198-272) and the final
//                  FilterResults (search.go:189-220) with the SizeInKmer it leaves
//                  (search_nucleotide.go:116-120)
//   k_row_sizes + 5 scans + k_assemble    CSR rows in query order: hits, position bytes, and
//                  for nucleotide rows Location / trimmed Query.Sequence
//
// The reference keeps a []bool per (query, subject) for EVERY subject seen; only the bit-sets of
// hits that survive FilterResults are observable (search.go:216-218 deletes the rest), so only
// those are built here.
#include <cub/cub.cuh>

#include "internal.cuh"

namespace kaamer {

struct FinishArgs {
  // index
  const uint64_t *table;
  uint64_t d_lo, d_hi;
  const uint32_t *postings;
  const PeerView *peer;  // mode P: shards of all ranks (nullptr otherwise)
  // queries + count-pass results
  const uint8_t *res;
  const uint64_t *off;
  uint32_t nq;
  const int32_t *size_in_kmer;
  const uint32_t *n_hits, *hit_base;
  const uint64_t *pool;
  // options
  long long min_kmatch;
  double min_kratio;
  int nt_mode;
  int want_pos;
  const uint8_t *any0;
  // ORF metadata (nt mode)
  const uint32_t *o_contig;
  const int64_t *o_start, *o_end;
  const uint8_t *o_plus;
  const uint64_t *o_alts_off;
  const int32_t *o_alts;
  // per-query scratch
  uint64_t *sc_words;  // [nq+1] -> scanned: first word of the query's bit-sets
  uint64_t *sc_rows, *sc_hits, *sc_pos, *sc_seq;  // [nq+1] each, scanned in place
  uint32_t *posbits;
  uint32_t *keep, *trim;
  // outputs (device staging, layout of kaamer_hits)
  uint64_t *hit_off;
  uint32_t *subject, *kmatch;
  int32_t *out_size;
  uint64_t *pos_off;
  uint8_t *pos;
  uint32_t *row_contig;
  int64_t *row_start, *row_end;
  uint8_t *row_plus;
  uint64_t *row_seq_off;
  uint8_t *row_seq;
};

__device__ __forceinline__ uint64_t ldg_entry_f(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__global__ void k_pos_need(FinishArgs a) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > a.nq) return;
  uint64_t w = 0;
  if (q < a.nq) {
    const uint32_t H = a.n_hits[q];
    const int32_t K = a.size_in_kmer[q];
    if (H && K > 0) w = (uint64_t)H * (uint64_t)((K + 31) / 32);
  }
  a.sc_words[q] = w;
}

// posting lists are sorted by id, descending (pkg/kvstore/kv_store.go:284-305)
__device__ __forceinline__ bool list_contains(const uint32_t *__restrict__ list, uint32_t n, uint32_t id) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t v = __ldg(list + mid);
    if (v == id) return true;
    if (v > id) lo = mid + 1;
    else hi = mid;
  }
  return false;
}

__global__ void __launch_bounds__(256) k_positions(FinishArgs a) {
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (q >= a.nq) return;
  const uint32_t H = a.n_hits[q];
  const int32_t K = a.size_in_kmer[q];
  if (H == 0 || K <= 0) {
    if (lane == 0) {
      a.keep[q] = 0;
      a.trim[q] = 0;
    }
    return;
  }
  const uint8_t *seq = a.res + a.off[q];
  const uint32_t wpr = (uint32_t)(K + 31) / 32;
  uint32_t *words = a.posbits + a.sc_words[q];
  const uint64_t *hits = a.pool + a.hit_base[q];
  int32_t p1 = -1;  // first matched position of the best hit
  for (uint32_t h0 = 0; h0 < H; h0 += 32) {
    const uint32_t nh = H - h0 < 32u ? H - h0 : 32u;
    const uint32_t my_id = lane < nh ? (uint32_t)hits[h0 + lane] : 0u;
    for (int32_t kb = 0; kb < K; kb += 32) {
      const int32_t k = kb + (int32_t)lane;
      uint32_t cnt = 0;
      uint64_t val = 0;
      const uint32_t *post = a.postings;
      if (k < K) {
        const uint8_t *s = seq + k;
        const uint32_t d = dense_from_codes(aa_code(s[0]), aa_code(s[1]), aa_code(s[2]), aa_code(s[3]),
                                            aa_code(s[4]), aa_code(s[5]), aa_code(s[6]));
        const uint64_t *tab = a.table;
        uint64_t lo = a.d_lo, hi = a.d_hi;
        if (a.peer && a.peer->full_table != nullptr) {  // replicated table: the entry names the shard of its list
          const uint64_t e = ldg_entry_f(a.peer->full_table + d);
          cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);
          val = e & ENTRY_VALUE_MASK;
          if (cnt >= 2) {
            post = a.peer->postings[val >> PEER_SHARD_SHIFT];
            val &= PEER_LOCAL_MASK;
          }
          hi = 0;  // (probed already)
        } else if (a.peer) {  // owner shard of d: its table and postings, local or behind NVLink
          uint32_t sh = 0;
          for (int i = 1; i < MAX_PEER_SHARDS; ++i) sh += d >= a.peer->fence[i] ? 1u : 0u;
          tab = a.peer->table[sh];
          post = a.peer->postings[sh];
          lo = a.peer->fence[sh];
          hi = DENSE_SPACE;
          if (a.peer->presence != nullptr && (int32_t)sh != a.peer->self &&
              ((__ldg(a.peer->presence + (d >> 5)) >> (d & 31u)) & 1u) == 0u)
            hi = 0;  // absent on its (remote) owner: no probe
        }
        if (d >= lo && d < hi) {
          const uint64_t e = ldg_entry_f(tab + (d - lo));
          cnt = (uint32_t)(e >> ENTRY_VALUE_BITS);
          val = e & ENTRY_VALUE_MASK;
        }
      }
      for (uint32_t hh = 0; hh < nh; ++hh) {
        const uint32_t id = __shfl_sync(0xFFFFFFFFu, my_id, hh);
        bool m = false;
        if (cnt == 1) m = (uint32_t)val == id;
        else if (cnt >= 2) m = list_contains(post + val, cnt, id);
        const unsigned w = __ballot_sync(0xFFFFFFFFu, m);
        if (lane == 0) words[(size_t)(h0 + hh) * wpr + (uint32_t)kb / 32] = w;
        if (h0 + hh == 0 && p1 < 0 && w) p1 = kb + __ffs(w) - 1;
      }
    }
  }
  uint32_t trim = 0, keep = H;
  if (a.nt_mode) {
    // SetBestStartCodon (dna.go:198-272): first matched position of the first best hit; any
    // later hit tied with it is inspected at position 0 only (`exit` is never reset, :224-237)
    const int32_t fb = a.any0[q] ? 0 : p1;
    const uint64_t ao = a.o_alts_off[q];
    const uint32_t na = (uint32_t)(a.o_alts_off[q + 1] - ao);
    if (na > 0) {
      const int32_t first = a.o_alts[ao];
      int32_t best = first;
      for (uint32_t i = 0; i < na; ++i) {  // dna.go:240-249 (alternatives are ascending)
        const int32_t s = a.o_alts[ao + i];
        if (s <= fb) best = s;
        else break;
      }
      if (best != first) trim = (uint32_t)best;  // dna.go:252-268
    }
    const int32_t knew = K - (int32_t)trim;
    const uint32_t kmin = filter_kmin(a.min_kmatch, a.min_kratio, knew);
    uint32_t kept = 0;
    for (uint32_t hb = 0; hb < H; hb += 32) {
      const uint32_t hi = hb + lane;
      const bool ok = hi < H && (uint32_t)(hits[hi] >> 32) >= kmin;
      kept += __popc(__ballot_sync(0xFFFFFFFFu, ok));
    }
    keep = kept;  // the pool is in rank order and the test is monotone: a prefix survives
  }
  if (lane == 0) {
    a.keep[q] = keep;
    a.trim[q] = trim;
  }
}

__global__ void k_row_sizes(FinishArgs a) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > a.nq) return;
  uint64_t row = 0, hits = 0, pos = 0, seq = 0;
  if (q < a.nq) {
    const uint32_t keep = a.keep[q];
    const int64_t knew = (int64_t)a.size_in_kmer[q] - (int64_t)a.trim[q];
    row = a.nt_mode ? (keep > 0) : 1;
    hits = keep;
    if (a.want_pos && keep && knew > 0) pos = (uint64_t)keep * (uint64_t)knew;
    if (a.nt_mode && row) seq = (a.off[q + 1] - a.off[q]) - a.trim[q];
  }
  a.sc_rows[q] = row;
  a.sc_hits[q] = hits;
  a.sc_pos[q] = pos;
  a.sc_seq[q] = seq;
}

__global__ void __launch_bounds__(256) k_assemble(FinishArgs a) {
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (q >= a.nq) return;
  const uint64_t r = a.sc_rows[q];
  if (a.sc_rows[q + 1] == r) return;  // not a row
  const uint32_t keep = a.keep[q], trim = a.trim[q];
  const int32_t K = a.size_in_kmer[q];
  const int32_t knew = K - (int32_t)trim;
  const uint64_t hoff = a.sc_hits[q];
  if (lane == 0) {
    a.hit_off[r] = hoff;
    a.out_size[r] = knew;
  }
  if (a.nt_mode) {
    const uint64_t so = a.sc_seq[q];
    const uint64_t len = a.off[q + 1] - a.off[q];
    if (lane == 0) {
      a.row_contig[r] = a.o_contig[q];
      const int64_t st = a.o_start[q];
      a.row_start[r] = a.o_plus[q] ? st + 3 * (int64_t)trim : st - 3 * (int64_t)trim;  // dna.go:253-257
      a.row_end[r] = a.o_end[q];
      a.row_plus[r] = a.o_plus[q];
      a.row_seq_off[r] = so;
    }
    const uint8_t *src = a.res + a.off[q] + trim;
    for (uint64_t i = lane; i + trim < len; i += 32) a.row_seq[so + i] = src[i];
  }
  const uint64_t *hits = a.pool + a.hit_base[q];
  for (uint32_t h = lane; h < keep; h += 32) {
    const uint64_t v = hits[h];
    a.subject[hoff + h] = (uint32_t)v;
    a.kmatch[hoff + h] = (uint32_t)(v >> 32);
  }
  if (a.want_pos && keep) {
    const uint64_t po = a.sc_pos[q];
    const uint32_t wpr = K > 0 ? (uint32_t)(K + 31) / 32 : 0;
    const uint32_t *words = a.posbits + a.sc_words[q];
    const int64_t per = knew > 0 ? knew : 0;
    for (uint32_t h = 0; h < keep; ++h) {
      if (lane == 0) a.pos_off[hoff + h] = po + (uint64_t)h * (uint64_t)per;
      for (int64_t i = lane; i < per; i += 32) {
        const uint32_t k = (uint32_t)i + trim;  // PositionHits[_k] = _positions[bestStart:] (dna.go:261-263)
        a.pos[po + (uint64_t)h * (uint64_t)per + (uint64_t)i] = (words[(size_t)h * wpr + (k >> 5)] >> (k & 31)) & 1u;
      }
    }
  }
}

int finish_rows(kaamer_gpu *h, const uint8_t *d_res, const uint64_t *d_off, uint32_t nq, const kaamer_opts *o,
                int nt_mode, const uint8_t *d_any0, const OrfSet *orfs, kaamer_hits *hits, HitsOwner *owner,
                cudaStream_t st) {
  SearchWorkspace &ws = h->ws;
  const int want_pos = (nt_mode || o->want_positions) ? 1 : 0;  // search.go:416
  const size_t n1 = (size_t)nq + 1;
  KCHECK(ws.f_scan.ensure(5 * n1));
  KCHECK(ws.f_keep.ensure(n1));
  KCHECK(ws.f_trim.ensure(n1));
  FinishArgs a{};
  a.table = h->idx.table;
  a.d_lo = h->idx.d_lo;
  a.d_hi = h->idx.d_hi;
  a.postings = h->idx.postings;
  a.peer = h->idx.peer.n > 0 ? h->idx.d_peer : nullptr;
  a.res = d_res;
  a.off = d_off;
  a.nq = nq;
  a.size_in_kmer = ws.size_in_kmer.p;
  a.n_hits = ws.n_hits.p;
  a.hit_base = ws.hit_base.p;
  a.pool = ws.pool.p;
  a.min_kmatch = o->min_kmatch;
  a.min_kratio = o->min_kratio;
  a.nt_mode = nt_mode;
  a.want_pos = want_pos;
  a.any0 = d_any0;
  if (orfs) {
    a.o_contig = orfs->contig;
    a.o_start = orfs->start;
    a.o_end = orfs->end;
    a.o_plus = orfs->plus;
    a.o_alts_off = orfs->alts_off;
    a.o_alts = orfs->alts;
  }
  a.sc_words = ws.f_scan.p;
  a.sc_rows = ws.f_scan.p + n1;
  a.sc_hits = ws.f_scan.p + 2 * n1;
  a.sc_pos = ws.f_scan.p + 3 * n1;
  a.sc_seq = ws.f_scan.p + 4 * n1;
  a.keep = ws.f_keep.p;
  a.trim = ws.f_trim.p;
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, a.sc_words, a.sc_words, (int64_t)n1, st);
  KCHECK(ws.f_tmp.ensure(tmp_bytes + 16));
  const unsigned g1 = (unsigned)((n1 + 255) / 256);
  const unsigned gw = (unsigned)(((uint64_t)nq * 32 + 255) / 256);
  uint64_t totals[5] = {0, 0, 0, 0, 0};
  HostPhase ph_pos(h, 5);
  if (nq) {
    // 1. bit-set space
    k_pos_need<<<g1, 256, 0, st>>>(a);
    size_t tb = tmp_bytes;
    KCUDA(cub::DeviceScan::ExclusiveSum(ws.f_tmp.p, tb, a.sc_words, a.sc_words, (int64_t)n1, st));
    KCUDA(cudaMemcpyAsync(&totals[0], a.sc_words + nq, 8, cudaMemcpyDeviceToHost, st));
    KCUDA(cudaStreamSynchronize(st));
    KCHECK(ws.f_posbits.ensure((size_t)totals[0] + 1));
    a.posbits = ws.f_posbits.p;
    // 2. positions, start-codon correction, final filter
    k_positions<<<gw, 256, 0, st>>>(a);
    // 3. row / hit / position / sequence offsets
    k_row_sizes<<<g1, 256, 0, st>>>(a);
    uint64_t *arrs[4] = {a.sc_rows, a.sc_hits, a.sc_pos, a.sc_seq};
    for (int i = 0; i < 4; ++i) {
      tb = tmp_bytes;
      KCUDA(cub::DeviceScan::ExclusiveSum(ws.f_tmp.p, tb, arrs[i], arrs[i], (int64_t)n1, st));
      KCUDA(cudaMemcpyAsync(&totals[1 + i], arrs[i] + nq, 8, cudaMemcpyDeviceToHost, st));
    }
    h->prof_all_launches += 3;
    KCUDA(cudaGetLastError());
    KCUDA(cudaStreamSynchronize(st));
  }
  ph_pos.stop();
  HostPhase ph_alloc(h, 6);
  const uint64_t n_rows = totals[1], n_hits = totals[2], n_pos = totals[3], n_seq = totals[4];
  if (n_rows > 0xFFFFFFFFull) {
    set_error("too many result rows");
    return KAAMER_ERR_LIMIT;
  }
  hits->n_rows = (uint32_t)n_rows;
  hits->n_hits = n_hits;
  // host outputs
  KCHECK(owner->alloc(&hits->hit_off, (size_t)n_rows + 1));
  KCHECK(owner->alloc(&hits->size_in_kmer, (size_t)n_rows));
  KCHECK(owner->alloc(&hits->subject_id, (size_t)n_hits));
  KCHECK(owner->alloc(&hits->kmatch, (size_t)n_hits));
  if (want_pos) {
    KCHECK(owner->alloc(&hits->pos_off, (size_t)n_hits + 1));
    KCHECK(owner->alloc(&hits->pos, (size_t)n_pos));
  }
  if (nt_mode) {
    KCHECK(owner->alloc(&hits->row_contig, (size_t)n_rows));
    KCHECK(owner->alloc(&hits->row_start, (size_t)n_rows));
    KCHECK(owner->alloc(&hits->row_end, (size_t)n_rows));
    KCHECK(owner->alloc(&hits->row_plus, (size_t)n_rows));
    KCHECK(owner->alloc(&hits->row_seq_off, (size_t)n_rows + 1));
    KCHECK(owner->alloc(&hits->row_seq, (size_t)n_seq));
  }
  hits->hit_off[n_rows] = n_hits;
  if (want_pos) hits->pos_off[n_hits] = n_pos;
  if (nt_mode) hits->row_seq_off[n_rows] = n_seq;
  if (nq == 0 || n_rows == 0) {
    hits->hit_off[0] = 0;
    if (want_pos) hits->pos_off[0] = 0;
    if (nt_mode) hits->row_seq_off[0] = 0;
    return KAAMER_OK;
  }
  ph_alloc.stop();
  HostPhase ph_d2h(h, 7);
  // device staging: one allocation, 64-byte aligned sections
  auto up = [](size_t x) { return (x + 63) & ~(size_t)63; };
  size_t o_hit_off = 0, cur = up(n_rows * 8);
  size_t o_size = cur; cur += up(n_rows * 4);
  size_t o_subj = cur; cur += up(n_hits * 4);
  size_t o_km = cur; cur += up(n_hits * 4);
  size_t o_poff = cur; cur += up(n_hits * 8);
  size_t o_pos = cur; cur += up(n_pos);
  size_t o_rc = cur; cur += up(n_rows * 4);
  size_t o_rs = cur; cur += up(n_rows * 8);
  size_t o_re = cur; cur += up(n_rows * 8);
  size_t o_rp = cur; cur += up(n_rows);
  size_t o_rso = cur; cur += up(n_rows * 8);
  size_t o_rseq = cur; cur += up(n_seq);
  uint8_t *stage = nullptr;
  KCHECK(h->arena.get(&stage, cur + 64));
  a.hit_off = (uint64_t *)(stage + o_hit_off);
  a.out_size = (int32_t *)(stage + o_size);
  a.subject = (uint32_t *)(stage + o_subj);
  a.kmatch = (uint32_t *)(stage + o_km);
  a.pos_off = (uint64_t *)(stage + o_poff);
  a.pos = stage + o_pos;
  a.row_contig = (uint32_t *)(stage + o_rc);
  a.row_start = (int64_t *)(stage + o_rs);
  a.row_end = (int64_t *)(stage + o_re);
  a.row_plus = stage + o_rp;
  a.row_seq_off = (uint64_t *)(stage + o_rso);
  a.row_seq = stage + o_rseq;
  k_assemble<<<gw, 256, 0, st>>>(a);
  h->prof_all_launches += 1;
  cudaError_t e = cudaGetLastError();
  auto cp = [&](void *dst, const void *src, size_t bytes) {
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
  };
  cp(hits->hit_off, a.hit_off, n_rows * 8);
  cp(hits->size_in_kmer, a.out_size, n_rows * 4);
  cp(hits->subject_id, a.subject, n_hits * 4);
  cp(hits->kmatch, a.kmatch, n_hits * 4);
  if (want_pos) {
    cp(hits->pos_off, a.pos_off, n_hits * 8);
    cp(hits->pos, a.pos, n_pos);
  }
  if (nt_mode) {
    cp(hits->row_contig, a.row_contig, n_rows * 4);
    cp(hits->row_start, a.row_start, n_rows * 8);
    cp(hits->row_end, a.row_end, n_rows * 8);
    cp(hits->row_plus, a.row_plus, n_rows);
    cp(hits->row_seq_off, a.row_seq_off, n_rows * 8);
    cp(hits->row_seq, a.row_seq, n_seq);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("finish_rows: %s", cudaGetErrorString(e));
    return KAAMER_ERR_CUDA;
  }
  return KAAMER_OK;
}

}  // namespace kaamer
