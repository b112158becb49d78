// index.cu — the device-resident k-mer index.
//
// Replaces on the search path the two badger stores the reference probes per k-mer
// (kmer_store -> kcomb_store, pkg/search/search.go:421-429):
//   * table[d]   : direct-address array over the dense 7-mer code space (internal.cuh),
//                  one 8-byte entry per possible k-mer = (count:27 | value:37);
//                  a singleton posting list is inlined in the entry, so the common
//                  lookup is ONE 32-byte-sector HBM access;
//   * postings[] : CSR protein-id lists (ids unique per k-mer, descending — the order
//                  CreateKCKeyValue produces, pkg/kvstore/kcomb_store.go:42-85).
// and, as §8f-1, pkg/makedb + pkg/indexdb for the device index: records -> (k-mer, id)
// pairs -> radix sort -> unique -> CSR, all on the GPU (cub is used for the sort/scan/
// select plumbing of this one-shot build; the search kernels are hand-written).
#include <cstdlib>

#include <cub/cub.cuh>

#include "internal.cuh"

namespace kaamer {

// ---------------------------------------------------------------------------------------
// table fill from the sorted CSR form
// ---------------------------------------------------------------------------------------
__global__ void k_fill_table(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ offsets,
                             const uint32_t *__restrict__ postings, uint64_t n_keys, uint64_t *table,
                             uint64_t d_lo, uint64_t d_hi, uint32_t *filter, unsigned long long *bad) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  uint32_t d;
  if (!dense_from_key(keys[i], &d)) {
    atomicAdd(bad, 1ull);
    return;
  }
  if (d < d_lo || d >= d_hi) return;
  uint64_t b = offsets[i], e = offsets[i + 1];
  uint64_t cnt = e - b;
  if (cnt == 0) return;
  if (cnt > ENTRY_MAX_COUNT) {
    atomicAdd(bad + 1, 1ull);
    return;
  }
  uint64_t val = cnt == 1 ? (uint64_t)postings[b] : b;
  table[d - d_lo] = (cnt << ENTRY_VALUE_BITS) | val;
  if (filter) atomicOr(filter + ((d & FILTER_MASK) >> 5), 1u << (d & 31u));
}

// A key-range shard (shard_hi != 0) keeps its table and postings in shareable memory (vmm.cu):
// the other GPUs of the node map them and probe them through NVLink (mode P).  A full index
// uses plain cudaMalloc.
static int alloc_postings(kaamer_gpu *h, uint64_t n_postings, bool shareable) {
  DevIndex &ix = h->idx;
  const size_t bytes = (size_t)(n_postings + 1) * sizeof(uint32_t);
  if (shareable) {
    KCHECK(vmm_alloc(h->device, bytes, &ix.vm_postings));
    ix.postings = (uint32_t *)ix.vm_postings.ptr;
    return KAAMER_OK;
  }
  KCUDA(cudaMalloc((void **)&ix.postings, bytes));
  return KAAMER_OK;
}

static int alloc_table(kaamer_gpu *h, uint64_t d_lo, uint64_t d_hi, bool shareable) {
  DevIndex &ix = h->idx;
  ix.d_lo = d_lo;
  ix.d_hi = d_hi;
  size_t bytes = (size_t)(d_hi - d_lo) * sizeof(uint64_t);
  if (shareable) {
    KCHECK(vmm_alloc(h->device, bytes, &ix.vm_table));
    ix.table = (uint64_t *)ix.vm_table.ptr;
  } else {
    cudaError_t e = cudaMalloc((void **)&ix.table, bytes);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(table, %zu bytes): %s", bytes, cudaGetErrorString(e));
      return KAAMER_ERR_NOMEM;
    }
  }
  KCUDA(cudaMemsetAsync(ix.table, 0, bytes, h->stream));
  return KAAMER_OK;
}

static int fill_table(kaamer_gpu *h) {
  DevIndex &ix = h->idx;
  unsigned long long *d_bad;
  KCUDA(cudaMalloc((void **)&d_bad, 2 * sizeof(unsigned long long)));
  KCUDA(cudaMemsetAsync(d_bad, 0, 2 * sizeof(unsigned long long), h->stream));
  // the L2-resident presence filter (internal.cuh), unless the key space is too densely occupied
  if (ix.n_keys <= FILTER_MAX_KEYS && getenv("KAAMER_NO_L2_FILTER") == nullptr) {
    KCUDA(cudaMalloc((void **)&ix.filter, FILTER_WORDS * sizeof(uint32_t)));
    KCUDA(cudaMemsetAsync(ix.filter, 0, FILTER_WORDS * sizeof(uint32_t), h->stream));
  }
  if (ix.n_keys) {
    unsigned grid = (unsigned)((ix.n_keys + 255) / 256);
    k_fill_table<<<grid, 256, 0, h->stream>>>(ix.keys, ix.offsets, ix.postings, ix.n_keys, ix.table, ix.d_lo,
                                               ix.d_hi, ix.filter, d_bad);
    KCUDA(cudaGetLastError());
  }
  unsigned long long bad[2];
  KCUDA(cudaMemcpyAsync(bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, h->stream));
  KCUDA(cudaStreamSynchronize(h->stream));
  cudaFree(d_bad);
  if (bad[0]) {
    set_error("index holds %llu keys that are not EncodeKmer outputs", bad[0]);
    return KAAMER_ERR_FORMAT;
  }
  if (bad[1]) {
    set_error("%llu posting lists exceed %llu entries", bad[1], (unsigned long long)ENTRY_MAX_COUNT);
    return KAAMER_ERR_LIMIT;
  }
  return KAAMER_OK;
}

// table + postings of a handle that a streaming builder fills pass by pass (build_stream.cu)
int alloc_index_storage(kaamer_gpu *h, uint64_t d_lo, uint64_t d_hi, uint64_t n_postings, bool shareable) {
  KCHECK(alloc_table(h, d_lo, d_hi, shareable));
  KCHECK(alloc_postings(h, n_postings, shareable));
  KCUDA(cudaStreamSynchronize(h->stream));
  return KAAMER_OK;
}

void index_release(kaamer_gpu *h) {
  DevIndex &ix = h->idx;
  if (ix.vm_table.ptr) vmm_free(&ix.vm_table);
  else cudaFree(ix.table);
  if (ix.vm_postings.ptr) vmm_free(&ix.vm_postings);
  else cudaFree(ix.postings);
  cudaFree(ix.filter);
  cudaFree(ix.keys);
  cudaFree(ix.offsets);
  cudaFree(ix.prot_off);
  cudaFree(ix.prot_res);
  ix = DevIndex();
}

static int upload_proteins(kaamer_gpu *h, const uint64_t *off, const uint8_t *res, uint32_t max_id) {
  DevIndex &ix = h->idx;
  size_t n_off = (size_t)max_id + 2;
  uint64_t n_res = off[n_off - 1];
  KCUDA(cudaMalloc((void **)&ix.prot_off, n_off * sizeof(uint64_t)));
  KCUDA(cudaMalloc((void **)&ix.prot_res, (size_t)n_res + 16));
  KCUDA(cudaMemcpyAsync(ix.prot_off, off, n_off * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
  KCUDA(cudaMemcpyAsync(ix.prot_res, res, (size_t)n_res, cudaMemcpyHostToDevice, h->stream));
  KCUDA(cudaStreamSynchronize(h->stream));
  ix.n_prot_res = n_res;
  ix.has_proteins = true;
  ix.h_prot_off.assign(off, off + n_off);
  return KAAMER_OK;
}

int index_from_view(kaamer_gpu *h, const kaamer_index_view *v) {
  DevIndex &ix = h->idx;
  if (!v || (v->n_keys && (!v->keys || !v->offsets)) || (v->n_postings && !v->postings)) {
    set_error("index view: null section");
    return KAAMER_ERR_ARG;
  }
  if (v->n_keys && v->offsets[v->n_keys] != v->n_postings) {
    set_error("index view: offsets[n_keys]=%llu != n_postings=%llu", (unsigned long long)v->offsets[v->n_keys],
              (unsigned long long)v->n_postings);
    return KAAMER_ERR_FORMAT;
  }
  if (v->n_postings > ENTRY_VALUE_MASK) {
    set_error("index view: too many postings");
    return KAAMER_ERR_LIMIT;
  }
  uint64_t lo = v->shard_lo, hi = v->shard_hi;
  if (lo == 0 && hi == 0) hi = DENSE_SPACE;
  if (hi > DENSE_SPACE || lo >= hi) {
    set_error("index view: bad shard range");
    return KAAMER_ERR_ARG;
  }
  ix.n_keys = v->n_keys;
  ix.n_postings = v->n_postings;
  // are the lists in the reference's order (ids strictly descending, kv_store.go:284-305)?  Class D then
  // verifies its candidates by binary search; any other order is searched correctly, a little slower
  ix.lists_sorted = true;
  for (uint64_t k = 0; k < v->n_keys && ix.lists_sorted; ++k)
    for (uint64_t i = v->offsets[k] + 1; i < v->offsets[k + 1]; ++i)
      if (v->postings[i] >= v->postings[i - 1]) {
        ix.lists_sorted = false;
        break;
      }
  ix.n_proteins = v->n_proteins;
  ix.n_aa = v->n_aa;
  ix.n_kmers = v->n_kmers;
  ix.max_protein_id = v->max_protein_id;
  KCHECK(alloc_table(h, lo, hi, v->shard_hi != 0));
  KCUDA(cudaMalloc((void **)&ix.keys, (size_t)(v->n_keys + 1) * sizeof(uint32_t)));
  KCUDA(cudaMalloc((void **)&ix.offsets, (size_t)(v->n_keys + 1) * sizeof(uint64_t)));
  KCHECK(alloc_postings(h, v->n_postings, v->shard_hi != 0));
  if (v->n_keys) {
    KCUDA(cudaMemcpyAsync(ix.keys, v->keys, (size_t)v->n_keys * 4, cudaMemcpyHostToDevice, h->stream));
    KCUDA(cudaMemcpyAsync(ix.offsets, v->offsets, (size_t)(v->n_keys + 1) * 8, cudaMemcpyHostToDevice, h->stream));
  } else {
    KCUDA(cudaMemsetAsync(ix.offsets, 0, 8, h->stream));
  }
  if (v->n_postings)
    KCUDA(cudaMemcpyAsync(ix.postings, v->postings, (size_t)v->n_postings * 4, cudaMemcpyHostToDevice, h->stream));
  KCHECK(fill_table(h));
  if (v->prot_seq_off && v->prot_residues) KCHECK(upload_proteins(h, v->prot_seq_off, v->prot_residues, v->max_protein_id));
  return KAAMER_OK;
}

// ---------------------------------------------------------------------------------------
// on-device build (makedb + indexdb semantics)
// ---------------------------------------------------------------------------------------
// one warp per record (records average ~350 residues); lanes stride over the windows.  Only
// windows whose dense code lies in [d_lo, d_hi) are kept (key-range shards build their own part).
__global__ void k_emit_pairs(const uint8_t *__restrict__ res, const uint64_t *__restrict__ seq_off,
                             const uint64_t *__restrict__ win_off, const uint32_t *__restrict__ ids,
                             uint64_t n_records, uint64_t d_lo, uint64_t d_hi, uint64_t *pairs) {
  uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  if (warp >= n_records) return;
  uint64_t b = seq_off[warp], e = seq_off[warp + 1];
  uint64_t len = e - b;
  if (len < KAAMER_KMER_SIZE) return;  // inputFASTA.go:226-228
  uint64_t nwin = len - KAAMER_KMER_SIZE + 1;
  uint64_t wo = win_off[warp];
  uint64_t id = ids[warp];
  const uint8_t *s = res + b;
  for (uint64_t i0 = 0; i0 < nwin; i0 += 32) {
    const uint64_t i = i0 + lane;
    bool in = false;
    uint32_t key = 0;
    if (i < nwin) {
      uint32_t c[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) c[j] = aa_code(s[i + j]);
      const uint32_t d = dense_from_codes(c[0], c[1], c[2], c[3], c[4], c[5], c[6]);
      in = d >= d_lo && d < d_hi;
      key = key_from_codes(c[0], c[1], c[2], c[3], c[4], c[5], c[6]);
    }
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, in);
    if (in) pairs[wo + __popc(mask & ((1u << lane) - 1u))] = ((uint64_t)key << 32) | id;
    wo += __popc(mask);
  }
}

// windows of each record inside the shard's code range + KStats of the WHOLE input
__global__ void k_window_counts(const uint8_t *__restrict__ res, const uint64_t *__restrict__ seq_off,
                                uint64_t n_records, uint64_t d_lo, uint64_t d_hi, uint64_t *win,
                                unsigned long long *stats) {
  uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  if (warp >= n_records) return;
  uint64_t b = seq_off[warp];
  uint64_t len = seq_off[warp + 1] - b;
  uint64_t nwin = len >= KAAMER_KMER_SIZE ? len - KAAMER_KMER_SIZE + 1 : 0;
  uint64_t cnt = 0;
  const uint8_t *s = res + b;
  const bool whole = d_lo == 0 && d_hi >= DENSE_SPACE;
  if (whole) {
    cnt = nwin;
  } else {
    for (uint64_t i0 = 0; i0 < nwin; i0 += 32) {
      const uint64_t i = i0 + lane;
      bool in = false;
      if (i < nwin) {
        uint32_t c[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) c[j] = aa_code(s[i + j]);
        const uint32_t d = dense_from_codes(c[0], c[1], c[2], c[3], c[4], c[5], c[6]);
        in = d >= d_lo && d < d_hi;
      }
      cnt += __popc(__ballot_sync(0xFFFFFFFFu, in));
    }
  }
  if (lane == 0) {
    win[warp] = cnt;
    if (nwin) {
      atomicAdd(stats + 0, 1ull);                       // NumberOfProteins
      atomicAdd(stats + 1, (unsigned long long)len);    // NumberOfAA   (inputFASTA.go:142-145)
      atomicAdd(stats + 2, (unsigned long long)nwin);   // NumberOfKmers
    }
  }
}

// heads of key runs in the sorted unique pair array
__global__ void k_flag_heads(const uint64_t *__restrict__ pairs, uint64_t n, uint32_t *head) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || (pairs[i] >> 32) != (pairs[i - 1] >> 32)) ? 1u : 0u;
}
// rank[i] = inclusive scan of head; key k starts where head==1
__global__ void k_write_keys(const uint64_t *__restrict__ pairs, const uint32_t *__restrict__ head,
                             const uint64_t *__restrict__ rank, uint64_t n, uint32_t *keys, uint64_t *offsets,
                             uint64_t n_keys) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (head[i]) {
    uint64_t k = rank[i] - 1;
    keys[k] = (uint32_t)(pairs[i] >> 32);
    offsets[k] = i;
  }
  if (i == 0) offsets[n_keys] = n;
}
// ids ascending inside a run -> descending in postings (kv_store.go:284-305)
__global__ void k_write_postings(const uint64_t *__restrict__ pairs, const uint64_t *__restrict__ rank,
                                 const uint64_t *__restrict__ offsets, uint64_t n, uint32_t *postings) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = rank[i] - 1;
  uint64_t b = offsets[k], e = offsets[k + 1];
  postings[b + (e - 1 - i)] = (uint32_t)pairs[i];
}

int index_build(kaamer_gpu *h, const uint8_t *residues, const uint64_t *seq_off, const uint32_t *ids,
                uint64_t n_records, int keep_proteins, uint64_t shard_lo, uint64_t shard_hi) {
  const bool shareable = !(shard_lo == 0 && shard_hi == 0);  // an explicit key range: mode S / mode P shard
  if (!shareable) shard_hi = DENSE_SPACE;
  if (shard_hi > DENSE_SPACE || shard_lo >= shard_hi) {
    set_error("build: bad shard range");
    return KAAMER_ERR_ARG;
  }
  DevIndex &ix = h->idx;
  cudaStream_t st = h->stream;
  uint64_t n_res = n_records ? seq_off[n_records] : 0;
  uint8_t *d_res = nullptr;
  uint64_t *d_off = nullptr, *d_win = nullptr, *d_woff = nullptr;
  uint32_t *d_ids = nullptr;
  unsigned long long *d_stats = nullptr;
  void *d_tmp = nullptr;
  uint64_t *d_pairs = nullptr, *d_pairs2 = nullptr, *d_rank = nullptr, *d_nsel = nullptr;
  uint32_t *d_head = nullptr;
  int rc = KAAMER_OK;
  auto cleanup = [&]() {
    cudaFree(d_res); cudaFree(d_off); cudaFree(d_win); cudaFree(d_woff); cudaFree(d_ids);
    cudaFree(d_stats); cudaFree(d_tmp); cudaFree(d_pairs); cudaFree(d_pairs2); cudaFree(d_rank);
    cudaFree(d_nsel); cudaFree(d_head);
  };
#define BCUDA(call)                                                                      \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));   \
      cleanup();                                                                         \
      return KAAMER_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)
  BCUDA(cudaMalloc((void **)&d_res, (size_t)n_res + 16));
  BCUDA(cudaMalloc((void **)&d_off, (size_t)(n_records + 1) * 8));
  BCUDA(cudaMalloc((void **)&d_win, (size_t)(n_records + 1) * 8));
  BCUDA(cudaMalloc((void **)&d_woff, (size_t)(n_records + 1) * 8));
  BCUDA(cudaMalloc((void **)&d_ids, (size_t)(n_records + 1) * 4));
  BCUDA(cudaMalloc((void **)&d_stats, 4 * 8));
  BCUDA(cudaMemsetAsync(d_stats, 0, 4 * 8, st));
  BCUDA(cudaMemsetAsync(d_win, 0, (size_t)(n_records + 1) * 8, st));
  if (n_records) {
    BCUDA(cudaMemcpyAsync(d_res, residues, (size_t)n_res, cudaMemcpyHostToDevice, st));
    BCUDA(cudaMemcpyAsync(d_off, seq_off, (size_t)(n_records + 1) * 8, cudaMemcpyHostToDevice, st));
    BCUDA(cudaMemcpyAsync(d_ids, ids, (size_t)n_records * 4, cudaMemcpyHostToDevice, st));
    k_window_counts<<<(unsigned)((n_records * 32 + 255) / 256), 256, 0, st>>>(d_res, d_off, n_records, shard_lo, shard_hi,
                                                                             d_win, d_stats);
    BCUDA(cudaGetLastError());
  } else {
    BCUDA(cudaMemsetAsync(d_off, 0, 8, st));
  }
  size_t tmp_bytes = 0, need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, d_win, d_woff, n_records + 1, st);
  tmp_bytes = need;
  BCUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
  BCUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_win, d_woff, n_records + 1, st));
  uint64_t n_pairs = 0;
  unsigned long long stats[4];
  BCUDA(cudaMemcpyAsync(&n_pairs, d_woff + n_records, 8, cudaMemcpyDeviceToHost, st));
  BCUDA(cudaMemcpyAsync(stats, d_stats, sizeof stats, cudaMemcpyDeviceToHost, st));
  BCUDA(cudaStreamSynchronize(st));
  ix.n_proteins = stats[0];
  ix.n_aa = stats[1];
  ix.n_kmers = stats[2];
  uint32_t max_id = 0;
  for (uint64_t i = 0; i < n_records; ++i) max_id = ids[i] > max_id ? ids[i] : max_id;
  ix.max_protein_id = max_id;

  BCUDA(cudaMalloc((void **)&d_pairs, (size_t)(n_pairs + 1) * 8));
  BCUDA(cudaMalloc((void **)&d_pairs2, (size_t)(n_pairs + 1) * 8));
  BCUDA(cudaMalloc((void **)&d_nsel, 8));
  uint64_t n_uniq = 0;
  if (n_pairs) {
    uint64_t warps = n_records;
    unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    k_emit_pairs<<<grid, 256, 0, st>>>(d_res, d_off, d_woff, d_ids, n_records, shard_lo, shard_hi, d_pairs);
    BCUDA(cudaGetLastError());
    // radix sort (key,id) as one u64
    need = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, need, d_pairs, d_pairs2, n_pairs, 0, 64, st);
    if (need > tmp_bytes) {
      cudaFree(d_tmp);
      d_tmp = nullptr;
      tmp_bytes = need;
      BCUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
    }
    size_t tb = tmp_bytes;
    BCUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tb, d_pairs, d_pairs2, n_pairs, 0, 64, st));
    // unique pairs: a protein holding the same k-mer twice counts once (set semantics)
    need = 0;
    cub::DeviceSelect::Unique(nullptr, need, d_pairs2, d_pairs, d_nsel, n_pairs, st);
    if (need > tmp_bytes) {
      cudaFree(d_tmp);
      d_tmp = nullptr;
      tmp_bytes = need;
      BCUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
    }
    tb = tmp_bytes;
    BCUDA(cub::DeviceSelect::Unique(d_tmp, tb, d_pairs2, d_pairs, d_nsel, n_pairs, st));
    BCUDA(cudaMemcpyAsync(&n_uniq, d_nsel, 8, cudaMemcpyDeviceToHost, st));
    BCUDA(cudaStreamSynchronize(st));
  }
  cudaFree(d_pairs2);
  d_pairs2 = nullptr;
  if (n_uniq > ENTRY_VALUE_MASK) {
    set_error("too many postings");
    cleanup();
    return KAAMER_ERR_LIMIT;
  }
  // key runs -> keys / offsets / postings
  uint64_t n_keys = 0;
  BCUDA(cudaMalloc((void **)&d_head, (size_t)(n_uniq + 1) * 4));
  BCUDA(cudaMalloc((void **)&d_rank, (size_t)(n_uniq + 1) * 8));
  if (n_uniq) {
    unsigned grid = (unsigned)((n_uniq + 255) / 256);
    k_flag_heads<<<grid, 256, 0, st>>>(d_pairs, n_uniq, d_head);
    BCUDA(cudaGetLastError());
    need = 0;
    cub::DeviceScan::InclusiveSum(nullptr, need, d_head, d_rank, n_uniq, st);
    if (need > tmp_bytes) {
      cudaFree(d_tmp);
      d_tmp = nullptr;
      tmp_bytes = need;
      BCUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
    }
    size_t tb = tmp_bytes;
    BCUDA(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_head, d_rank, n_uniq, st));
    BCUDA(cudaMemcpyAsync(&n_keys, d_rank + (n_uniq - 1), 8, cudaMemcpyDeviceToHost, st));
    BCUDA(cudaStreamSynchronize(st));
  }
  ix.n_keys = n_keys;
  ix.n_postings = n_uniq;
  ix.lists_sorted = true;  // k_write_postings: ids descending inside every run
  BCUDA(cudaMalloc((void **)&ix.keys, (size_t)(n_keys + 1) * 4));
  BCUDA(cudaMalloc((void **)&ix.offsets, (size_t)(n_keys + 1) * 8));
  rc = alloc_postings(h, n_uniq, shareable);
  if (rc != KAAMER_OK) {
    cleanup();
    return rc;
  }
  if (n_uniq) {
    unsigned grid = (unsigned)((n_uniq + 255) / 256);
    k_write_keys<<<grid, 256, 0, st>>>(d_pairs, d_head, d_rank, n_uniq, ix.keys, ix.offsets, n_keys);
    BCUDA(cudaGetLastError());
    k_write_postings<<<grid, 256, 0, st>>>(d_pairs, d_rank, ix.offsets, n_uniq, ix.postings);
    BCUDA(cudaGetLastError());
  } else {
    BCUDA(cudaMemsetAsync(ix.offsets, 0, 8, st));
  }
  BCUDA(cudaStreamSynchronize(st));
  // free the big temporaries before the 14.5 GB table is allocated
  cudaFree(d_pairs); d_pairs = nullptr;
  cudaFree(d_rank); d_rank = nullptr;
  cudaFree(d_head); d_head = nullptr;
  cudaFree(d_tmp); d_tmp = nullptr;
  rc = alloc_table(h, shard_lo, shard_hi, shareable);
  if (rc == KAAMER_OK) rc = fill_table(h);
  if (rc == KAAMER_OK && keep_proteins && n_records) {
    // protein table indexed by id (later records with the same id overwrite earlier ones,
    // as protein_store[id] does in the reference)
    std::vector<uint64_t> poff((size_t)max_id + 2, 0);
    std::vector<int64_t> rec_of((size_t)max_id + 1, -1);
    for (uint64_t i = 0; i < n_records; ++i)
      if (seq_off[i + 1] - seq_off[i] >= KAAMER_KMER_SIZE) rec_of[ids[i]] = (int64_t)i;
    for (uint32_t id = 0; id <= max_id; ++id) {
      uint64_t len = rec_of[id] >= 0 ? seq_off[rec_of[id] + 1] - seq_off[rec_of[id]] : 0;
      poff[id + 1] = poff[id] + len;
    }
    std::vector<uint8_t> pres((size_t)poff[(size_t)max_id + 1]);
    for (uint32_t id = 0; id <= max_id; ++id)
      if (rec_of[id] >= 0)
        memcpy(pres.data() + poff[id], residues + seq_off[rec_of[id]], (size_t)(poff[id + 1] - poff[id]));
    rc = upload_proteins(h, poff.data(), pres.data(), max_id);
  }
  cleanup();
#undef BCUDA
  return rc;
}

}  // namespace kaamer
