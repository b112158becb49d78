// build_stream.cu — streaming index builder (kaamer_gpu_builder_*): the makedb + indexdb step
// (pkg/makedb/inputFASTA.go:195-250, pkg/indexdb/indexdb.go:92-128, pkg/kvstore/kcomb_store.go:42-85)
// for databases whose records do not fit the device at once.  The key space of the handle is covered
// by consecutive PASSES over sub-ranges [pass_lo, pass_hi) of the dense code space; in every pass the
// caller feeds ALL records again, in device-resident chunks (a FASTA reader chunk, or — for the
// synthetic C4 database — the output of the on-device generator, synth.cu), and the builder keeps
// only the windows whose code falls into the pass range:
//
//   pass:  records -> (dense code << 32 | protein id) of the in-range windows (append)
//          -> radix sort -> unique (a protein holding a k-mer twice counts once: set semantics)
//          -> run-length encode by code -> postings appended (ids descending, kv_store.go:284-305)
//          -> table[d] = count:27 | (count == 1 ? id : first posting index):37
//
// Peak memory = table + all postings + the pairs of ONE pass (x2 for the sort), so the whole C4 index
// (14.5 GB table + ~59 GB postings) builds on one B200 in 16 passes, and a key-range shard of it in 2.
// The sorted keys[] / offsets[] export form is not kept (kaamer_gpu_index_copy / _save refuse).
#include <cub/cub.cuh>

#include "internal.cuh"

struct kaamer_builder {
  kaamer_gpu *h = nullptr;
  uint64_t max_postings = 0, used = 0, n_keys = 0;
  uint64_t pass_lo = 0, pass_hi = 0, pairs_cap = 0;
  bool in_pass = false;
  int passes_done = 0;
  uint64_t *pairs = nullptr, *pairs2 = nullptr;
  unsigned long long *d_cur = nullptr;  // [0] pair cursor [1] proteins [2] aa [3] k-mers [4] max id [5] overflow
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  uint64_t stats[5] = {0, 0, 0, 0, 0};
};

namespace kaamer {

// one warp per record; only windows whose dense code lies in [lo, hi) are kept (warp-aggregated append)
__global__ void __launch_bounds__(256) k_stream_emit(const uint8_t *__restrict__ res, const uint64_t *__restrict__ seq_off,
                                                     const uint32_t *__restrict__ ids, uint32_t id_base,
                                                     uint64_t n_records, uint64_t lo, uint64_t hi, uint64_t *pairs,
                                                     uint64_t cap, unsigned long long *cur, int want_stats) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= n_records) return;
  const uint64_t b = seq_off[warp], e = seq_off[warp + 1];
  const uint64_t len = e - b;
  if (len < KAAMER_KMER_SIZE) return;  // inputFASTA.go:226-228
  const uint64_t nwin = len - KAAMER_KMER_SIZE + 1;
  const uint64_t id = ids ? ids[warp] : (uint64_t)id_base + warp;
  if (want_stats && lane == 0) {
    atomicAdd(cur + 1, 1ull);
    atomicAdd(cur + 2, (unsigned long long)len);  // KStats.NumberOfAA (inputFASTA.go:142-145)
    atomicAdd(cur + 3, (unsigned long long)nwin);
    atomicMax(cur + 4, (unsigned long long)id);
  }
  const uint8_t *s = res + b;
  for (uint64_t i0 = 0; i0 < nwin; i0 += 32) {
    const uint64_t i = i0 + lane;
    bool in = false;
    uint32_t d = 0;
    if (i < nwin) {
      uint32_t c[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) c[j] = aa_code(s[i + j]);
      d = dense_from_codes(c[0], c[1], c[2], c[3], c[4], c[5], c[6]);
      in = d >= lo && d < hi;
    }
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, in);
    if (mask == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cur, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (in) {
      const uint64_t slot = base + __popc(mask & ((1u << lane) - 1u));
      if (slot < cap) pairs[slot] = ((uint64_t)d << 32) | id;
      else cur[5] = 1ull;
    }
  }
}

struct Hi32 {
  __host__ __device__ uint32_t operator()(uint64_t x) const { return (uint32_t)(x >> 32); }
};
struct Widen32 {
  __host__ __device__ uint64_t operator()(uint32_t x) const { return x; }
};

// one thread per key run: table entry, and the postings of short runs; long runs go to the warp kernel
constexpr uint32_t RUN_LONG = 64;
__global__ void k_stream_runs(const uint64_t *__restrict__ pairs, const uint32_t *__restrict__ codes,
                              const uint64_t *__restrict__ run_off, uint64_t n_runs, uint64_t post_base,
                              uint32_t *postings, uint64_t *table, uint64_t d_lo, uint32_t *long_runs,
                              unsigned long long *n_long, unsigned long long *bad) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const uint64_t b = run_off[r], e = run_off[r + 1];
  const uint64_t cnt = e - b;
  if (cnt > ENTRY_MAX_COUNT) {
    atomicAdd(bad, 1ull);
    return;
  }
  const uint64_t val = cnt == 1 ? (uint64_t)(uint32_t)pairs[b] : post_base + b;
  table[(uint64_t)codes[r] - d_lo] = (cnt << ENTRY_VALUE_BITS) | val;
  if (cnt >= RUN_LONG) {
    long_runs[atomicAdd(n_long, 1ull)] = (uint32_t)r;
    return;
  }
  // ids ascending inside the run -> descending in postings (kv_store.go:284-305)
  for (uint64_t i = b; i < e; ++i) postings[post_base + b + (e - 1 - i)] = (uint32_t)pairs[i];
}
__global__ void k_stream_long_runs(const uint64_t *__restrict__ pairs, const uint64_t *__restrict__ run_off,
                                   const uint32_t *__restrict__ long_runs, uint64_t n_long, uint64_t post_base,
                                   uint32_t *postings) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (w >= n_long) return;
  const uint64_t r = long_runs[w];
  const uint64_t b = run_off[r], e = run_off[r + 1];
  for (uint64_t i = b + lane; i < e; i += 32) postings[post_base + b + (e - 1 - i)] = (uint32_t)pairs[i];
}

static void builder_free_pass(kaamer_builder *b) {
  cudaFree(b->pairs);
  cudaFree(b->pairs2);
  cudaFree(b->tmp);
  b->pairs = b->pairs2 = nullptr;
  b->tmp = nullptr;
  b->pairs_cap = 0;
  b->tmp_bytes = 0;
}

static int ensure_tmp(kaamer_builder *b, size_t need) {
  if (need <= b->tmp_bytes && b->tmp) return KAAMER_OK;
  cudaFree(b->tmp);
  b->tmp = nullptr;
  b->tmp_bytes = 0;
  cudaError_t e = cudaMalloc(&b->tmp, need + 256);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(build scratch, %zu bytes): %s", need, cudaGetErrorString(e));
    return KAAMER_ERR_NOMEM;
  }
  b->tmp_bytes = need;
  return KAAMER_OK;
}

int new_handle_for_builder(int device, kaamer_gpu **out);  // api.cu
void destroy_handle_for_builder(kaamer_gpu *h);            // api.cu
int alloc_index_storage(kaamer_gpu *h, uint64_t d_lo, uint64_t d_hi, uint64_t n_postings, bool shareable);  // index.cu

}  // namespace kaamer

using namespace kaamer;

extern "C" {

int kaamer_gpu_builder_open(int device, uint64_t shard_lo, uint64_t shard_hi, uint64_t max_postings, int shareable,
                            kaamer_builder_t **out) {
  if (!out) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  if (shard_lo == 0 && shard_hi == 0) shard_hi = DENSE_SPACE;
  if (shard_hi > DENSE_SPACE || shard_lo >= shard_hi) {
    set_error("builder: bad key range");
    return KAAMER_ERR_ARG;
  }
  if (max_postings > ENTRY_VALUE_MASK || (shareable && max_postings > PEER_LOCAL_MASK)) {
    set_error("builder: too many postings");
    return KAAMER_ERR_LIMIT;
  }
  kaamer_gpu *h = nullptr;
  KCHECK(new_handle_for_builder(device, &h));
  int rc = alloc_index_storage(h, shard_lo, shard_hi, max_postings, shareable != 0);
  auto *b = new kaamer_builder();
  b->h = h;
  b->max_postings = max_postings;
  if (rc == KAAMER_OK && cudaMalloc((void **)&b->d_cur, 8 * sizeof(unsigned long long)) != cudaSuccess) {
    set_error("cudaMalloc(builder counters) failed");
    rc = KAAMER_ERR_NOMEM;
  }
  if (rc == KAAMER_OK && cudaMemset(b->d_cur, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) rc = KAAMER_ERR_CUDA;
  if (rc != KAAMER_OK) {
    cudaFree(b->d_cur);
    delete b;
    destroy_handle_for_builder(h);
    return rc;
  }
  *out = b;
  return KAAMER_OK;
}

void kaamer_gpu_builder_abort(kaamer_builder_t *b) {
  if (!b) return;
  cudaSetDevice(b->h->device);
  builder_free_pass(b);
  cudaFree(b->d_cur);
  destroy_handle_for_builder(b->h);
  delete b;
}

int kaamer_gpu_builder_pass_begin(kaamer_builder_t *b, uint64_t pass_lo, uint64_t pass_hi, uint64_t max_pairs) {
  if (!b || b->in_pass) {
    set_error("builder: pass_begin out of order");
    return KAAMER_ERR_ARG;
  }
  if (pass_lo < b->h->idx.d_lo || pass_hi > b->h->idx.d_hi || pass_lo >= pass_hi) {
    set_error("builder: pass range outside the handle's key range");
    return KAAMER_ERR_ARG;
  }
  if (max_pairs >= (1ull << 31) - 1) {
    set_error("builder: at most 2^31-2 pairs per pass (use more passes)");
    return KAAMER_ERR_LIMIT;
  }
  KCUDA(cudaSetDevice(b->h->device));
  if (max_pairs + 1 > b->pairs_cap) {
    cudaFree(b->pairs);
    cudaFree(b->pairs2);
    b->pairs = b->pairs2 = nullptr;
    b->pairs_cap = 0;
    cudaError_t e = cudaMalloc((void **)&b->pairs, (size_t)(max_pairs + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&b->pairs2, (size_t)(max_pairs + 1) * 8);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(2 x %llu pass pairs): %s", (unsigned long long)max_pairs, cudaGetErrorString(e));
      return KAAMER_ERR_NOMEM;
    }
    b->pairs_cap = max_pairs + 1;
  }
  KCUDA(cudaMemsetAsync(b->d_cur, 0, sizeof(unsigned long long), b->h->stream));
  KCUDA(cudaMemsetAsync(b->d_cur + 5, 0, sizeof(unsigned long long), b->h->stream));
  KCUDA(cudaStreamSynchronize(b->h->stream));
  b->pass_lo = pass_lo;
  b->pass_hi = pass_hi;
  b->in_pass = true;
  return KAAMER_OK;
}

int kaamer_gpu_builder_add_device(kaamer_builder_t *b, const uint8_t *d_residues, const uint64_t *d_seq_off,
                                  const uint32_t *d_ids, uint32_t id_base, uint64_t n_records, void *stream) {
  if (!b || !b->in_pass || (n_records && (!d_residues || !d_seq_off))) {
    set_error("builder: add outside a pass, or null argument");
    return KAAMER_ERR_ARG;
  }
  if (n_records == 0) return KAAMER_OK;
  KCUDA(cudaSetDevice(b->h->device));
  k_stream_emit<<<(unsigned)((n_records * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_residues, d_seq_off, d_ids, id_base, n_records, b->pass_lo, b->pass_hi, b->pairs, b->pairs_cap - 1, b->d_cur,
      b->passes_done == 0 ? 1 : 0);
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

int kaamer_gpu_builder_pass_end(kaamer_builder_t *b, void *stream_) {
  if (!b || !b->in_pass) {
    set_error("builder: pass_end out of order");
    return KAAMER_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  kaamer_gpu *h = b->h;
  DevIndex &ix = h->idx;
  KCUDA(cudaSetDevice(h->device));
  unsigned long long cur[8];
  KCUDA(cudaMemcpyAsync(cur, b->d_cur, sizeof cur, cudaMemcpyDeviceToHost, st));
  KCUDA(cudaStreamSynchronize(st));
  b->in_pass = false;
  if (cur[5] || cur[0] > b->pairs_cap - 1) {
    set_error("builder: pass [%llu, %llu) produced %llu pairs, capacity %llu", (unsigned long long)b->pass_lo,
              (unsigned long long)b->pass_hi, cur[0], (unsigned long long)(b->pairs_cap - 1));
    return KAAMER_ERR_LIMIT;
  }
  if (b->passes_done == 0)
    for (int i = 0; i < 5; ++i) b->stats[i] = cur[i];
  b->passes_done++;
  const uint64_t n_pairs = cur[0];
  if (n_pairs == 0) return KAAMER_OK;
  // sort (code, id); only the bits that can differ: id < 2^32, code < pass_hi
  int code_bits = 1;
  while (code_bits < 32 && (b->pass_hi - 1) >> code_bits) ++code_bits;
  cub::DoubleBuffer<uint64_t> db(b->pairs, b->pairs2);
  size_t need = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, need, db, n_pairs, 0, 32 + code_bits, st);
  KCHECK(ensure_tmp(b, need));
  size_t tb = b->tmp_bytes;
  KCUDA(cub::DeviceRadixSort::SortKeys(b->tmp, tb, db, n_pairs, 0, 32 + code_bits, st));
  uint64_t *sorted = db.Current(), *other = db.Alternate();
  // unique pairs -> `other`
  unsigned long long *d_n = b->d_cur + 6;
  need = 0;
  cub::DeviceSelect::Unique(nullptr, need, sorted, other, d_n, n_pairs, st);
  KCHECK(ensure_tmp(b, need));
  tb = b->tmp_bytes;
  KCUDA(cub::DeviceSelect::Unique(b->tmp, tb, sorted, other, d_n, n_pairs, st));
  unsigned long long n_uniq = 0;
  KCUDA(cudaMemcpyAsync(&n_uniq, d_n, 8, cudaMemcpyDeviceToHost, st));
  KCUDA(cudaStreamSynchronize(st));
  if (b->used + n_uniq > b->max_postings) {
    set_error("builder: more than max_postings = %llu postings", (unsigned long long)b->max_postings);
    return KAAMER_ERR_LIMIT;
  }
  // run-length encode by code; the `sorted` buffer is free now and holds the run arrays:
  // codes u32[n] | counts u32[n] | offsets u64[n+1]  (n <= n_uniq: 16 B per run <= 8 B per pair x 2)
  uint64_t *uniq = other;
  uint32_t *codes = reinterpret_cast<uint32_t *>(sorted);
  uint32_t *counts = codes + n_uniq;
  cub::TransformInputIterator<uint32_t, Hi32, const uint64_t *> key_in(uniq, Hi32());
  need = 0;
  cub::DeviceRunLengthEncode::Encode(nullptr, need, key_in, codes, counts, d_n, (int)n_uniq, st);
  KCHECK(ensure_tmp(b, need));
  tb = b->tmp_bytes;
  KCUDA(cub::DeviceRunLengthEncode::Encode(b->tmp, tb, key_in, codes, counts, d_n, (int)n_uniq, st));
  unsigned long long n_runs = 0;
  KCUDA(cudaMemcpyAsync(&n_runs, d_n, 8, cudaMemcpyDeviceToHost, st));
  KCUDA(cudaStreamSynchronize(st));
  // offsets + long-run list live in scratch of their own (the pair buffers are both in use)
  const size_t off_bytes = (size_t)(n_runs + 1) * 8, long_bytes = (size_t)(n_uniq / RUN_LONG + 1) * 4;
  cub::TransformInputIterator<uint64_t, Widen32, const uint32_t *> cnt_in(counts, Widen32());
  need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, cnt_in, (uint64_t *)nullptr, (int64_t)n_runs, st);
  const size_t scan_bytes = (need + 255) & ~(size_t)255;
  KCHECK(ensure_tmp(b, scan_bytes + off_bytes + 256 + long_bytes + 256));
  uint64_t *run_off = reinterpret_cast<uint64_t *>((uint8_t *)b->tmp + scan_bytes);
  uint32_t *long_runs = reinterpret_cast<uint32_t *>((uint8_t *)run_off + ((off_bytes + 255) & ~(size_t)255));
  tb = scan_bytes;
  KCUDA(cub::DeviceScan::ExclusiveSum(b->tmp, tb, cnt_in, run_off, (int64_t)n_runs, st));
  KCUDA(cudaMemcpyAsync(run_off + n_runs, &n_uniq, 8, cudaMemcpyHostToDevice, st));
  KCUDA(cudaMemsetAsync(b->d_cur + 6, 0, 16, st));  // [6] long-run count, [7] bad runs
  k_stream_runs<<<(unsigned)((n_runs + 255) / 256), 256, 0, st>>>(uniq, codes, run_off, n_runs, b->used, ix.postings,
                                                                  ix.table, ix.d_lo, long_runs, b->d_cur + 6,
                                                                  b->d_cur + 7);
  KCUDA(cudaGetLastError());
  unsigned long long tail[2];
  KCUDA(cudaMemcpyAsync(tail, b->d_cur + 6, 16, cudaMemcpyDeviceToHost, st));
  KCUDA(cudaStreamSynchronize(st));
  if (tail[1]) {
    set_error("%llu posting lists exceed %llu entries", tail[1], (unsigned long long)ENTRY_MAX_COUNT);
    return KAAMER_ERR_LIMIT;
  }
  if (tail[0]) {
    k_stream_long_runs<<<(unsigned)((tail[0] * 32 + 255) / 256), 256, 0, st>>>(uniq, run_off, long_runs, tail[0],
                                                                               b->used, ix.postings);
    KCUDA(cudaGetLastError());
    KCUDA(cudaStreamSynchronize(st));
  }
  b->used += n_uniq;
  b->n_keys += n_runs;
  return KAAMER_OK;
}

int kaamer_gpu_builder_finish(kaamer_builder_t *b, kaamer_gpu_t **out) {
  if (!b || !out || b->in_pass) {
    set_error("builder: finish out of order");
    return KAAMER_ERR_ARG;
  }
  kaamer_gpu *h = b->h;
  cudaSetDevice(h->device);
  DevIndex &ix = h->idx;
  ix.n_keys = b->n_keys;
  ix.n_postings = b->used;
  ix.lists_sorted = true;  // k_stream_runs / k_stream_long_runs: ids descending inside every run
  ix.n_proteins = b->stats[1];
  ix.n_aa = b->stats[2];
  ix.n_kmers = b->stats[3];
  ix.max_protein_id = (uint32_t)b->stats[4];
  builder_free_pass(b);
  cudaFree(b->d_cur);
  delete b;
  *out = h;
  return KAAMER_OK;
}

}  // extern "C"
