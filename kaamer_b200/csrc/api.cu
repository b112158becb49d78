// api.cu — the extern "C" surface of libkaamer_gpu.so (include/kaamer_gpu.h): handle
// lifetime, `.kidx` I/O, argument checking.  No torch types, no C++ types in signatures.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <algorithm>
#include <exception>
#include <unistd.h>

#include "internal.cuh"

namespace kaamer {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

// ---- pinned block cache ------------------------------------------------------------------
static std::mutex g_pin_mu;
static std::vector<std::pair<void *, size_t>> g_pin_free;
static size_t g_pin_cached = 0;
constexpr size_t PIN_CACHE_MAX = 4ull << 30;

void *pinned_get(size_t bytes, size_t *cap) {
  size_t c = 4096;
  while (c < bytes) c <<= 1;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_pin_free.size(); ++i)
      if (g_pin_free[i].second == c) {
        void *p = g_pin_free[i].first;
        g_pin_free[i] = g_pin_free.back();
        g_pin_free.pop_back();
        g_pin_cached -= c;
        *cap = c;
        return p;
      }
  }
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, c, cudaHostAllocPortable | cudaHostAllocMapped);
  if (e != cudaSuccess) {
    // drop the cache and retry once
    {
      std::lock_guard<std::mutex> lk(g_pin_mu);
      for (auto &b : g_pin_free) cudaFreeHost(b.first);
      g_pin_free.clear();
      g_pin_cached = 0;
    }
    e = cudaHostAlloc(&p, c, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
      set_error("cudaHostAlloc(%zu bytes): %s", c, cudaGetErrorString(e));
      return nullptr;
    }
  }
  *cap = c;
  return p;
}

void pinned_put(void *p, size_t cap) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (g_pin_cached + cap <= PIN_CACHE_MAX) {
      g_pin_free.emplace_back(p, cap);
      g_pin_cached += cap;
      return;
    }
  }
  cudaFreeHost(p);
}

// ---- .kidx container (little-endian; sections 64-byte aligned) ------------------------
struct KidxHeader {
  char magic[8];  // "KIDX0001"
  uint32_t version, k;
  uint64_t n_keys, n_postings, n_proteins, n_aa, n_kmers;
  uint32_t max_protein_id, flags;  // flags bit0: protein table present
  uint64_t n_residues;
  uint64_t annot_bytes;  // flags bit1: annotation sections present (entry offsets, EntryId blob of this size, lengths)
  uint8_t pad[128 - 8 - 8 - 40 - 8 - 8 - 8];
};
static_assert(sizeof(KidxHeader) == 128, "kidx header is 128 bytes");

static size_t pad64(size_t n) { return (n + 63) & ~(size_t)63; }

static int read_section(FILE *f, void *dst, size_t bytes) {
  if (bytes && fread(dst, 1, bytes, f) != bytes) {
    set_error("kidx: short read");
    return KAAMER_ERR_IO;
  }
  size_t skip = pad64(bytes) - bytes;
  if (skip && fseek(f, (long)skip, SEEK_CUR) != 0) {
    set_error("kidx: seek failed");
    return KAAMER_ERR_IO;
  }
  return KAAMER_OK;
}
static int write_section(FILE *f, const void *src, size_t bytes) {
  static const char zeros[64] = {0};
  if (bytes && fwrite(src, 1, bytes, f) != bytes) {
    set_error("kidx: short write");
    return KAAMER_ERR_IO;
  }
  size_t skip = pad64(bytes) - bytes;
  if (skip && fwrite(zeros, 1, skip, f) != skip) {
    set_error("kidx: short write");
    return KAAMER_ERR_IO;
  }
  return KAAMER_OK;
}

static int new_handle(int device, kaamer_gpu **out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libkaamer_gpu has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return KAAMER_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (0..%d)", device, n - 1);
    return KAAMER_ERR_ARG;
  }
  KCUDA(cudaSetDevice(device));
  auto *h = new kaamer_gpu();
  h->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->chunk_ev[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->done_ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
    delete h;
    return KAAMER_ERR_CUDA;
  }
  *out = h;
  return KAAMER_OK;
}

static void destroy_handle(kaamer_gpu *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (auto &a : h->idx.imported) vmm_free(&a);
  h->idx.imported.clear();
  if (h->idx.presence) cudaFree(h->idx.presence);
  h->idx.presence = nullptr;
  if (h->idx.full_table) cudaFree(h->idx.full_table);
  h->idx.full_table = nullptr;
  if (h->idx.repl_postings) cudaFree(h->idx.repl_postings);
  h->idx.repl_postings = nullptr;
  if (h->idx.d_peer) cudaFree(h->idx.d_peer);
  h->idx.d_peer = nullptr;
  release_pending(h);
  index_release(h);
  h->ws.release_all();
  h->arena.release();
  for (auto &p : h->prof_pending) {
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  for (int i = 0; i < 8; ++i)
    if (h->chunk_ev[i]) cudaEventDestroy(h->chunk_ev[i]);
  if (h->done_ev) cudaEventDestroy(h->done_ev);
  delete h;
}

int new_handle_for_builder(int device, kaamer_gpu **out) { return new_handle(device, out); }
void destroy_handle_for_builder(kaamer_gpu *h) { destroy_handle(h); }

}  // namespace kaamer

using namespace kaamer;

extern "C" {

const char *kaamer_gpu_last_error(void) { return g_err; }
const char *kaamer_gpu_version(void) { return "kaamer_b200 0.1 (sm_100a)"; }

static int kaamer_gpu_open_view_impl(const kaamer_index_view *view, int device, kaamer_gpu_t **out) {
  if (!out || !view) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  kaamer_gpu *h = nullptr;
  KCHECK(new_handle(device, &h));
  int rc = index_from_view(h, view);
  if (rc != KAAMER_OK) {
    destroy_handle(h);
    return rc;
  }
  *out = h;
  return KAAMER_OK;
}
int kaamer_gpu_open_view(const kaamer_index_view *view, int device, kaamer_gpu_t **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_open_view_impl(view, device, out); });
}

// the sections of a .kidx file in host memory
struct KidxFile {
  KidxHeader hd;
  std::vector<uint32_t> keys, postings;
  std::vector<uint64_t> offsets, poff, aoff;
  std::vector<uint8_t> pres;
  std::vector<char> aids;
  std::vector<int32_t> alen;
};

// A `.kidx` file is untrusted input: every section size is checked against the file size BEFORE anything
// is allocated, and the contents are checked for the invariants the kernels rely on (keys strictly
// ascending EncodeKmer outputs, offsets non-decreasing and ending at n_postings, protein offsets
// monotonic and ending at n_residues) — a truncated or corrupt file yields KAAMER_ERR_FORMAT, never an
// exception through the C boundary or an out-of-bounds read on the device.
static int read_kidx(const char *path, KidxFile *k, bool want_postings, bool want_proteins) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    set_error("cannot open %s", path);
    return KAAMER_ERR_IO;
  }
  auto bad = [&](const char *what) {
    fclose(f);
    set_error("%s: %s", path, what);
    return KAAMER_ERR_FORMAT;
  };
  KidxHeader &hd = k->hd;
  if (fread(&hd, 1, sizeof hd, f) != sizeof hd || memcmp(hd.magic, "KIDX0001", 8) != 0 || hd.version != 1 ||
      hd.k != KAAMER_KMER_SIZE)
    return bad("not a kidx v1 (k=7) file");
  if (fseek(f, 0, SEEK_END) != 0) return bad("seek failed");
  const long long file_size = ftell(f);
  if (file_size < 0 || fseek(f, (long)sizeof hd, SEEK_SET) != 0) return bad("seek failed");
  const bool has_prot = (hd.flags & 1) != 0;
  // (all counts are bounded by the file size first, so the products below cannot overflow)
  const unsigned long long fs = (unsigned long long)file_size;
  if (hd.n_keys > fs / 4 || hd.n_postings > fs / 4 || hd.n_residues > fs || (has_prot && hd.max_protein_id > fs / 8))
    return bad("header counts exceed the file size (truncated or corrupt)");
  unsigned long long need = sizeof hd + pad64((size_t)hd.n_keys * 4) + pad64(((size_t)hd.n_keys + 1) * 8) +
                            pad64((size_t)hd.n_postings * 4);
  if (has_prot) need += pad64(((size_t)hd.max_protein_id + 2) * 8) + (size_t)hd.n_residues;
  const bool has_annot = (hd.flags & 2) != 0;
  if (has_annot) {
    if (hd.annot_bytes > fs || hd.max_protein_id > fs / 8) return bad("header counts exceed the file size (truncated or corrupt)");
    need = sizeof hd + pad64((size_t)hd.n_keys * 4) + pad64(((size_t)hd.n_keys + 1) * 8) + pad64((size_t)hd.n_postings * 4) +
           (has_prot ? pad64(((size_t)hd.max_protein_id + 2) * 8) + pad64((size_t)hd.n_residues) : 0) +
           pad64(((size_t)hd.max_protein_id + 2) * 8) + pad64((size_t)hd.annot_bytes) + ((size_t)hd.max_protein_id + 1) * 4;
  }
  if (need > fs) return bad("sections exceed the file size (truncated or corrupt)");
  if (hd.n_postings > ENTRY_VALUE_MASK) return bad("too many postings");
  int rc = KAAMER_OK;
  try {
    k->keys.resize((size_t)hd.n_keys);
    k->offsets.resize((size_t)hd.n_keys + 1);
    rc = read_section(f, k->keys.data(), k->keys.size() * 4);
    if (rc == KAAMER_OK) rc = read_section(f, k->offsets.data(), k->offsets.size() * 8);
    if (rc == KAAMER_OK && want_postings) {
      k->postings.resize((size_t)hd.n_postings);
      rc = read_section(f, k->postings.data(), k->postings.size() * 4);
      if (rc == KAAMER_OK && want_proteins && has_prot) {
        k->poff.resize((size_t)hd.max_protein_id + 2);
        k->pres.resize((size_t)hd.n_residues);
        rc = read_section(f, k->poff.data(), k->poff.size() * 8);
        if (rc == KAAMER_OK) rc = read_section(f, k->pres.data(), k->pres.size());
      } else if (rc == KAAMER_OK && has_prot && has_annot) {
        // skip the protein sections
        if (fseek(f, (long)(pad64(((size_t)hd.max_protein_id + 2) * 8) + pad64((size_t)hd.n_residues)), SEEK_CUR) != 0) {
          set_error("kidx: seek failed");
          rc = KAAMER_ERR_IO;
        }
      }
      if (rc == KAAMER_OK && has_annot) {
        k->aoff.resize((size_t)hd.max_protein_id + 2);
        k->aids.resize((size_t)hd.annot_bytes);
        k->alen.resize((size_t)hd.max_protein_id + 1);
        rc = read_section(f, k->aoff.data(), k->aoff.size() * 8);
        if (rc == KAAMER_OK) rc = read_section(f, k->aids.data(), k->aids.size());
        if (rc == KAAMER_OK && fread(k->alen.data(), 4, k->alen.size(), f) != k->alen.size()) {
          set_error("kidx: short read");
          rc = KAAMER_ERR_IO;
        }
      }
    }
  } catch (const std::exception &e) {
    fclose(f);
    set_error("%s: out of host memory reading the index (%s)", path, e.what());
    return KAAMER_ERR_NOMEM;
  }
  if (rc != KAAMER_OK) {
    fclose(f);
    return rc;
  }
  // content invariants
  uint32_t prev_d = 0;
  for (size_t i = 0; i < k->keys.size(); ++i) {
    uint32_t d = 0;
    if (!dense_from_key(k->keys[i], &d)) return bad("a key is not an EncodeKmer output");
    if (i && d <= prev_d) return bad("keys are not strictly ascending");
    prev_d = d;
    if (k->offsets[i + 1] < k->offsets[i]) return bad("offsets are not non-decreasing");
  }
  if (k->offsets[0] != 0 || k->offsets[k->keys.size()] != hd.n_postings) return bad("offsets do not span [0, n_postings]");
  if (!k->poff.empty()) {
    if (k->poff[0] != 0) return bad("protein offsets do not start at 0");
    for (size_t i = 1; i < k->poff.size(); ++i)
      if (k->poff[i] < k->poff[i - 1]) return bad("protein offsets are not monotonic");
    if (k->poff.back() != hd.n_residues) return bad("protein offsets do not end at n_residues");
  }
  if (want_postings)
    for (size_t i = 0; i < k->postings.size(); ++i)
      if (k->postings[i] > hd.max_protein_id) return bad("a posting exceeds max_protein_id");
  if (!k->aoff.empty()) {
    if (k->aoff[0] != 0) return bad("annotation offsets do not start at 0");
    for (size_t i = 1; i < k->aoff.size(); ++i)
      if (k->aoff[i] < k->aoff[i - 1]) return bad("annotation offsets are not monotonic");
    if (k->aoff.back() != hd.annot_bytes) return bad("annotation offsets do not end at annot_bytes");
  }
  fclose(f);
  return KAAMER_OK;
}

// first key of the (ascending) key array whose dense code is >= d; key order == dense order
static size_t lower_bound_dense(const std::vector<uint32_t> &keys, uint64_t d) {
  size_t lo = 0, hi = keys.size();
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    uint32_t dm = 0;
    if (dense_from_key(keys[mid], &dm) && (uint64_t)dm < d) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

static int open_kidx_range(const char *path, int device, uint64_t shard_lo, uint64_t shard_hi, kaamer_gpu_t **out) {
  if (!out || !path) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  const bool whole = shard_lo == 0 && shard_hi == 0;
  if (!whole && (shard_hi > DENSE_SPACE || shard_lo >= shard_hi)) {
    set_error("bad shard range [%llu, %llu)", (unsigned long long)shard_lo, (unsigned long long)shard_hi);
    return KAAMER_ERR_ARG;
  }
  KidxFile k;
  KCHECK(read_kidx(path, &k, true, true));
  const KidxHeader &hd = k.hd;
  size_t a = 0, b = k.keys.size();
  if (!whole) {
    a = lower_bound_dense(k.keys, shard_lo);
    b = lower_bound_dense(k.keys, shard_hi);
  }
  const uint64_t p0 = k.offsets[a], p1 = k.offsets[b];
  std::vector<uint64_t> offs(b - a + 1);
  for (size_t i = a; i <= b; ++i) offs[i - a] = k.offsets[i] - p0;
  kaamer_index_view v{};
  v.n_keys = b - a;
  v.n_postings = p1 - p0;
  v.keys = k.keys.data() + a;
  v.offsets = offs.data();
  v.postings = k.postings.data() + p0;
  v.n_proteins = hd.n_proteins;
  v.n_aa = hd.n_aa;
  v.n_kmers = hd.n_kmers;
  v.max_protein_id = hd.max_protein_id;
  if (hd.flags & 1) {
    v.prot_seq_off = k.poff.data();
    v.prot_residues = k.pres.data();
  }
  v.shard_lo = whole ? 0 : shard_lo;
  v.shard_hi = whole ? 0 : shard_hi;
  int rc = kaamer_gpu_open_view(&v, device, out);
  if (rc == KAAMER_OK && !k.aoff.empty()) {
    (*out)->idx.annot_off = std::move(k.aoff);
    (*out)->idx.annot_ids = std::move(k.aids);
    (*out)->idx.annot_len = std::move(k.alen);
  }
  return rc;
}

static int kaamer_gpu_open_impl(const char *path, int device, kaamer_gpu_t **out) {
  return open_kidx_range(path, device, 0, 0, out);
}
int kaamer_gpu_open(const char *path, int device, kaamer_gpu_t **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_open_impl(path, device, out); });
}

static int kaamer_gpu_open_shard_impl(const char *path, int device, uint64_t shard_lo, uint64_t shard_hi, kaamer_gpu_t **out) {
  if (shard_lo == 0 && shard_hi == 0) shard_hi = DENSE_SPACE;  // an explicit (shareable) full-range shard
  return open_kidx_range(path, device, shard_lo, shard_hi, out);
}
int kaamer_gpu_open_shard(const char *path, int device, uint64_t shard_lo, uint64_t shard_hi, kaamer_gpu_t **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_open_shard_impl(path, device, shard_lo, shard_hi, out); });
}

static int kaamer_gpu_kidx_fences_impl(const char *path, int n_shards, uint64_t *fences) {
  if (!path || !fences || n_shards < 1) {
    set_error("bad argument");
    return KAAMER_ERR_ARG;
  }
  KidxFile k;
  KCHECK(read_kidx(path, &k, false, false));
  fences[0] = 0;
  fences[n_shards] = DENSE_SPACE;
  const size_t n = k.keys.size();
  if (n == 0) {
    for (int s = 1; s < n_shards; ++s) fences[s] = DENSE_SPACE * (uint64_t)s / (uint64_t)n_shards;
    return KAAMER_OK;
  }
  // mass of a key = its postings + 1 (the probe itself); cut at equal cumulative mass
  const uint64_t total = k.offsets[n] + n;
  size_t i = 0;
  for (int s = 1; s < n_shards; ++s) {
    const double target = (double)total * s / n_shards;
    while (i < n - 1 && (double)(k.offsets[i + 1] + i + 1) < target) ++i;
    uint32_t d = 0;
    dense_from_key(k.keys[i], &d);
    fences[s] = d > fences[s - 1] ? d : fences[s - 1];
  }
  return KAAMER_OK;
}
int kaamer_gpu_kidx_fences(const char *path, int n_shards, uint64_t *fences) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_kidx_fences_impl(path, n_shards, fences); });
}

int kaamer_gpu_build(const uint8_t *residues, const uint64_t *seq_off, const uint32_t *ids, uint64_t n_records,
                     int keep_proteins, int device, kaamer_gpu_t **out) {
  return kaamer_gpu_build_shard(residues, seq_off, ids, n_records, keep_proteins, device, 0, 0, out);
}

static int kaamer_gpu_build_shard_impl(const uint8_t *residues, const uint64_t *seq_off, const uint32_t *ids, uint64_t n_records,
                           int keep_proteins, int device, uint64_t shard_lo, uint64_t shard_hi, kaamer_gpu_t **out) {
  if (!out || (n_records && (!residues || !seq_off || !ids))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  kaamer_gpu *h = nullptr;
  KCHECK(new_handle(device, &h));
  int rc = index_build(h, residues, seq_off, ids, n_records, keep_proteins, shard_lo, shard_hi);
  if (rc != KAAMER_OK) {
    destroy_handle(h);
    return rc;
  }
  *out = h;
  return KAAMER_OK;
}
int kaamer_gpu_build_shard(const uint8_t *residues, const uint64_t *seq_off, const uint32_t *ids, uint64_t n_records,
                           int keep_proteins, int device, uint64_t shard_lo, uint64_t shard_hi, kaamer_gpu_t **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_build_shard_impl(residues, seq_off, ids, n_records, keep_proteins, device, shard_lo, shard_hi, out); });
}

void kaamer_gpu_close(kaamer_gpu_t *h) { destroy_handle(h); }

// ---- peer-mapped shards (mode P) -----------------------------------------------------------
int kaamer_gpu_shard_export(kaamer_gpu_t *h, kaamer_shard_handle *out) {
  if (!h || !out) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  DevIndex &ix = h->idx;
  if (!ix.table) {
    set_error("no index resident");
    return KAAMER_ERR_ARG;
  }
  memset(out, 0, sizeof *out);
  out->shard_lo = ix.d_lo;
  out->shard_hi = ix.d_hi;
  out->n_postings = ix.n_postings;
  out->table_ptr = (uint64_t)(uintptr_t)ix.table;
  out->postings_ptr = (uint64_t)(uintptr_t)ix.postings;
  out->device = h->device;
  out->pid = (int32_t)getpid();
  out->table_fd = out->postings_fd = -1;
  if (ix.n_postings > PEER_LOCAL_MASK) {
    set_error("shard holds %llu postings: more than 2^%d per shard cannot be peer-mapped",
              (unsigned long long)ix.n_postings, PEER_SHARD_SHIFT);
    return KAAMER_ERR_LIMIT;
  }
  KCUDA(cudaStreamSynchronize(h->stream));  // the build is complete before anyone maps it
  if (ix.vm_table.ptr && ix.vm_postings.ptr) {
    out->table_bytes = ix.vm_table.bytes;
    out->postings_bytes = ix.vm_postings.bytes;
    int ft = -1, fp = -1;
    KCHECK(vmm_export_fd(ix.vm_table, &ft));
    int rc = vmm_export_fd(ix.vm_postings, &fp);
    if (rc != KAAMER_OK) {
      close(ft);
      return rc;
    }
    out->table_fd = ft;
    out->postings_fd = fp;
  }
  return KAAMER_OK;
}

static void detach_shards_locked(kaamer_gpu *h) {
  for (auto &a : h->idx.imported) vmm_free(&a);
  h->idx.imported.clear();
  if (h->idx.presence) cudaFree(h->idx.presence);
  h->idx.presence = nullptr;
  if (h->idx.full_table) cudaFree(h->idx.full_table);
  h->idx.full_table = nullptr;
  if (h->idx.repl_postings) cudaFree(h->idx.repl_postings);
  h->idx.repl_postings = nullptr;
  h->idx.flat_view = false;
  h->idx.peer = PeerView{};
}

int kaamer_gpu_detach_shards(kaamer_gpu_t *h) {
  if (!h) {
    set_error("null handle");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  KCUDA(cudaStreamSynchronize(h->stream));
  KCUDA(cudaStreamSynchronize(h->copy_stream));
  KCUDA(cudaStreamSynchronize(h->side_stream));
  detach_shards_locked(h);
  return KAAMER_OK;
}

// pointer to a shard's array as seen from h->device
static int map_shard_array(kaamer_gpu *h, const kaamer_shard_handle &s, bool same_process, uint64_t ptr, uint64_t bytes,
                           int fd, const void **out) {
  if (same_process) {
    if (s.device != h->device) {
      if (fd >= 0) {
        KCHECK(vmm_grant((void *)(uintptr_t)ptr, (size_t)bytes, h->device));  // shareable memory: add this device
      } else {
        int can = 0;
        KCUDA(cudaDeviceCanAccessPeer(&can, h->device, s.device));
        if (!can) {
          set_error("device %d cannot access device %d (no NVLink / P2P path)", h->device, s.device);
          return KAAMER_ERR_CUDA;
        }
        cudaError_t e = cudaDeviceEnablePeerAccess(s.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) {
          cudaGetLastError();
        } else if (e != cudaSuccess) {
          set_error("cudaDeviceEnablePeerAccess(%d): %s", s.device, cudaGetErrorString(e));
          return KAAMER_ERR_CUDA;
        }
      }
    }
    *out = (const void *)(uintptr_t)ptr;
    return KAAMER_OK;
  }
  if (fd < 0) {
    set_error("shard of process %d carries no shareable handle (a full index built without a key range cannot be "
              "mapped by another process)", s.pid);
    return KAAMER_ERR_ARG;
  }
  VmmAlloc a;
  KCHECK(vmm_import_fd(fd, (size_t)bytes, h->device, &a));
  h->idx.imported.push_back(a);
  *out = a.ptr;
  return KAAMER_OK;
}

static int kaamer_gpu_attach_shards_impl(kaamer_gpu_t *h, const kaamer_shard_handle *shards, int n_shards, int flags) {
  if (!h || !shards) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  if (n_shards < 1 || n_shards > MAX_PEER_SHARDS) {
    set_error("n_shards %d out of range (1..%d)", n_shards, MAX_PEER_SHARDS);
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  KCUDA(cudaStreamSynchronize(h->stream));
  KCUDA(cudaStreamSynchronize(h->copy_stream));
  KCUDA(cudaStreamSynchronize(h->side_stream));
  detach_shards_locked(h);
  // order by shard_lo and check that the ranges tile the dense code space
  std::vector<int> order(n_shards);
  for (int i = 0; i < n_shards; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return shards[x].shard_lo < shards[y].shard_lo; });
  uint64_t expect = 0;
  for (int i = 0; i < n_shards; ++i) {
    const kaamer_shard_handle &s = shards[order[i]];
    if (s.shard_lo != expect || s.shard_hi <= s.shard_lo) {
      set_error("shard ranges do not tile the key space: shard %d covers [%llu, %llu), expected to start at %llu", i,
                (unsigned long long)s.shard_lo, (unsigned long long)s.shard_hi, (unsigned long long)expect);
      return KAAMER_ERR_ARG;
    }
    if (s.n_postings > PEER_LOCAL_MASK) {
      set_error("shard %d holds more than 2^%d postings", i, PEER_SHARD_SHIFT);
      return KAAMER_ERR_LIMIT;
    }
    expect = s.shard_hi;
  }
  if (expect != DENSE_SPACE) {
    set_error("shard ranges do not tile the key space: they end at %llu, not at %llu", (unsigned long long)expect,
              (unsigned long long)DENSE_SPACE);
    return KAAMER_ERR_ARG;
  }
  PeerView pv{};
  for (int i = 0; i <= MAX_PEER_SHARDS; ++i) pv.fence[i] = 0xFFFFFFFFu;
  pv.n = n_shards;
  pv.self = -1;
  pv.presence = nullptr;
  pv.full_table = nullptr;
  const int32_t me = (int32_t)getpid();
  for (int i = 0; i < n_shards; ++i) {
    const kaamer_shard_handle &s = shards[order[i]];
    pv.fence[i] = (uint32_t)s.shard_lo;
    const void *pt = nullptr, *pp = nullptr;
    int rc = map_shard_array(h, s, s.pid == me, s.table_ptr, s.table_bytes, s.table_fd, &pt);
    if (rc == KAAMER_OK) rc = map_shard_array(h, s, s.pid == me, s.postings_ptr, s.postings_bytes, s.postings_fd, &pp);
    if (rc != KAAMER_OK) {
      detach_shards_locked(h);
      return rc;
    }
    pv.table[i] = (const uint64_t *)pt;
    pv.postings[i] = (const uint32_t *)pp;
    if (s.pid == me && s.device == h->device && s.table_ptr == (uint64_t)(uintptr_t)h->idx.table) pv.self = i;
  }
  if (n_shards > 1 && (flags & KAAMER_ATTACH_REPLICATE_TABLE)) {
    cudaError_t e = cudaMalloc((void **)&h->idx.full_table, (size_t)DENSE_SPACE * sizeof(uint64_t));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(replicated table, %llu bytes): %s", (unsigned long long)(DENSE_SPACE * 8),
                cudaGetErrorString(e));
      detach_shards_locked(h);
      return KAAMER_ERR_NOMEM;
    }
    int rc = replicate_table(h, pv, h->idx.full_table, h->stream);
    if (rc != KAAMER_OK) {
      detach_shards_locked(h);
      return rc;
    }
    pv.full_table = h->idx.full_table;
    if (flags & KAAMER_ATTACH_REPLICATE_POSTINGS) {
      // built sharded, searched replicated: the posting lists of every shard are copied into one local
      // array (peer copies at NVLink streaming speed, once); the per-shard pointers of the view then point
      // into it, so the search kernels — which already find a list through the shard tag of its entry —
      // never leave this GPU's HBM
      uint64_t total = 0;
      std::vector<uint64_t> base(n_shards);
      for (int i = 0; i < n_shards; ++i) {
        base[i] = total;
        total += (shards[order[i]].n_postings + 3) & ~3ull;  // 16-byte aligned starts
      }
      cudaError_t e2 = cudaMalloc((void **)&h->idx.repl_postings, (size_t)(total + 4) * sizeof(uint32_t));
      if (e2 != cudaSuccess) {
        set_error("cudaMalloc(replicated postings, %llu bytes): %s", (unsigned long long)(total * 4),
                  cudaGetErrorString(e2));
        detach_shards_locked(h);
        return KAAMER_ERR_NOMEM;
      }
      for (int i = 0; i < n_shards; ++i) {
        const uint64_t n = shards[order[i]].n_postings;
        if (n == 0) continue;
        e2 = cudaMemcpyAsync(h->idx.repl_postings + base[i], pv.postings[i], (size_t)n * 4, cudaMemcpyDefault, h->stream);
        if (e2 != cudaSuccess) {
          set_error("copy of shard %d's postings: %s", i, cudaGetErrorString(e2));
          detach_shards_locked(h);
          return KAAMER_ERR_CUDA;
        }
      }
      e2 = cudaStreamSynchronize(h->stream);
      if (e2 != cudaSuccess) {
        set_error("copy of the shards' postings: %s", cudaGetErrorString(e2));
        detach_shards_locked(h);
        return KAAMER_ERR_CUDA;
      }
      for (int i = 0; i < n_shards; ++i) pv.postings[i] = h->idx.repl_postings + base[i];
      if (total <= PEER_LOCAL_MASK) {
        // offsets instead of (shard, local) pairs: the PEER kernels decode them as shard 0 + offset, and the
        // protein search runs its non-PEER kernels on the replica (search.cu flat view)
        int rcf = flatten_table(h, h->idx.full_table, base.data(), n_shards, h->stream);
        if (rcf != KAAMER_OK) {
          detach_shards_locked(h);
          return rcf;
        }
        for (int i = 0; i < n_shards; ++i) pv.postings[i] = h->idx.repl_postings;
        h->idx.flat_view = true;
      }
    }
  } else if (n_shards > 1 && !(flags & KAAMER_ATTACH_NO_PRESENCE_FILTER)) {
    // local replica of "which k-mers exist": 1 bit per dense code, built by streaming every shard once
    const size_t words = (size_t)((DENSE_SPACE + 31) / 32);
    cudaError_t e = cudaMalloc((void **)&h->idx.presence, words * sizeof(uint32_t));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(presence filter, %zu bytes): %s", words * 4, cudaGetErrorString(e));
      detach_shards_locked(h);
      return KAAMER_ERR_NOMEM;
    }
    int rc = build_presence(h, pv, h->idx.presence, h->stream);
    if (rc != KAAMER_OK) {
      detach_shards_locked(h);
      return rc;
    }
    pv.presence = h->idx.presence;
  }
  if (!h->idx.d_peer) KCUDA(cudaMalloc((void **)&h->idx.d_peer, sizeof(PeerView)));
  KCUDA(cudaMemcpy(h->idx.d_peer, &pv, sizeof pv, cudaMemcpyHostToDevice));
  h->idx.peer = pv;
  return KAAMER_OK;
}
int kaamer_gpu_attach_shards(kaamer_gpu_t *h, const kaamer_shard_handle *shards, int n_shards, int flags) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_attach_shards_impl(h, shards, n_shards, flags); });
}

int kaamer_gpu_dbstats(kaamer_gpu_t *h, uint64_t *n_proteins, uint64_t *n_aa, uint64_t *n_kmers) {
  if (!h) {
    set_error("null handle");
    return KAAMER_ERR_ARG;
  }
  if (n_proteins) *n_proteins = h->idx.n_proteins;
  if (n_aa) *n_aa = h->idx.n_aa;
  if (n_kmers) *n_kmers = h->idx.n_kmers;
  return KAAMER_OK;
}

int kaamer_gpu_index_sizes(kaamer_gpu_t *h, uint64_t *n_keys, uint64_t *n_postings) {
  if (!h) {
    set_error("null handle");
    return KAAMER_ERR_ARG;
  }
  if (n_keys) *n_keys = h->idx.n_keys;
  if (n_postings) *n_postings = h->idx.n_postings;
  return KAAMER_OK;
}

static int kaamer_gpu_index_copy_impl(kaamer_gpu_t *h, uint32_t *keys, uint64_t *offsets, uint32_t *postings) {
  if (!h) {
    set_error("null handle");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  DevIndex &ix = h->idx;
  if (ix.n_keys && !ix.keys) {
    set_error("this index was built by the streaming builder: the sorted keys/offsets export form was not kept");
    return KAAMER_ERR_ARG;
  }
  if (keys && ix.n_keys) KCUDA(cudaMemcpy(keys, ix.keys, (size_t)ix.n_keys * 4, cudaMemcpyDeviceToHost));
  if (offsets) KCUDA(cudaMemcpy(offsets, ix.offsets, (size_t)(ix.n_keys + 1) * 8, cudaMemcpyDeviceToHost));
  if (postings && ix.n_postings)
    KCUDA(cudaMemcpy(postings, ix.postings, (size_t)ix.n_postings * 4, cudaMemcpyDeviceToHost));
  return KAAMER_OK;
}
int kaamer_gpu_index_copy(kaamer_gpu_t *h, uint32_t *keys, uint64_t *offsets, uint32_t *postings) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_index_copy_impl(h, keys, offsets, postings); });
}

static int kaamer_gpu_save_impl(kaamer_gpu_t *h, const char *path) {
  if (!h || !path) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  DevIndex &ix = h->idx;
  std::vector<uint32_t> keys((size_t)ix.n_keys), postings((size_t)ix.n_postings);
  std::vector<uint64_t> offsets((size_t)ix.n_keys + 1), poff;
  std::vector<uint8_t> pres;
  KCHECK(kaamer_gpu_index_copy(h, keys.data(), offsets.data(), postings.data()));
  if (ix.has_proteins) {
    poff.resize((size_t)ix.max_protein_id + 2);
    pres.resize((size_t)ix.n_prot_res);
    KCUDA(cudaMemcpy(poff.data(), ix.prot_off, poff.size() * 8, cudaMemcpyDeviceToHost));
    if (ix.n_prot_res) KCUDA(cudaMemcpy(pres.data(), ix.prot_res, pres.size(), cudaMemcpyDeviceToHost));
  }
  KidxHeader hd;
  memset(&hd, 0, sizeof hd);
  memcpy(hd.magic, "KIDX0001", 8);
  hd.version = 1;
  hd.k = KAAMER_KMER_SIZE;
  hd.n_keys = ix.n_keys;
  hd.n_postings = ix.n_postings;
  hd.n_proteins = ix.n_proteins;
  hd.n_aa = ix.n_aa;
  hd.n_kmers = ix.n_kmers;
  hd.max_protein_id = ix.max_protein_id;
  const bool has_annot = !ix.annot_off.empty() && ix.annot_off.size() == (size_t)ix.max_protein_id + 2;
  hd.flags = (ix.has_proteins ? 1u : 0u) | (has_annot ? 2u : 0u);
  hd.n_residues = ix.n_prot_res;
  hd.annot_bytes = has_annot ? ix.annot_ids.size() : 0;
  FILE *f = fopen(path, "wb");
  if (!f) {
    set_error("cannot create %s", path);
    return KAAMER_ERR_IO;
  }
  int rc = fwrite(&hd, 1, sizeof hd, f) == sizeof hd ? KAAMER_OK : KAAMER_ERR_IO;
  if (rc == KAAMER_OK) rc = write_section(f, keys.data(), keys.size() * 4);
  if (rc == KAAMER_OK) rc = write_section(f, offsets.data(), offsets.size() * 8);
  if (rc == KAAMER_OK) rc = write_section(f, postings.data(), postings.size() * 4);
  if (rc == KAAMER_OK && ix.has_proteins) {
    rc = write_section(f, poff.data(), poff.size() * 8);
    if (rc == KAAMER_OK) rc = write_section(f, pres.data(), pres.size());
  }
  if (rc == KAAMER_OK && has_annot) {
    rc = write_section(f, ix.annot_off.data(), ix.annot_off.size() * 8);
    if (rc == KAAMER_OK) rc = write_section(f, ix.annot_ids.data(), ix.annot_ids.size());
    if (rc == KAAMER_OK) rc = write_section(f, ix.annot_len.data(), ix.annot_len.size() * 4);
  }
  if (fclose(f) != 0 && rc == KAAMER_OK) rc = KAAMER_ERR_IO;
  if (rc != KAAMER_OK) set_error("write to %s failed", path);
  return rc;
}
int kaamer_gpu_save(kaamer_gpu_t *h, const char *path) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_save_impl(h, path); });
}

static int check_search_args(kaamer_gpu_t *h, const void *a, const void *b, uint32_t n, const kaamer_opts *o,
                             const void *out) {
  if (!h || !o || !out || (n && (!a || !b))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  return KAAMER_OK;
}

static int kaamer_gpu_search_proteins_impl(kaamer_gpu_t *h, const uint8_t *residues, const uint64_t *seq_off, uint32_t nq,
                               const kaamer_opts *opts, kaamer_hits **out) {
  KCHECK(check_search_args(h, residues, seq_off, nq, opts, out));
  *out = nullptr;
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  HostPhase whole(h, 4);
  // one stream pass, one synchronisation (search.cu: submit / wait); batches that cannot be bounded
  // beforehand (positions, huge MaxResults) take the path with the retry loops
  int slot = -1;
  KCHECK(search_proteins_submit(h, residues, seq_off, nq, opts, &slot));
  if (slot < 0) return search_proteins_host(h, residues, seq_off, nq, opts, out);
  return search_proteins_wait(h, slot, out);
}
int kaamer_gpu_search_proteins(kaamer_gpu_t *h, const uint8_t *residues, const uint64_t *seq_off, uint32_t nq,
                               const kaamer_opts *opts, kaamer_hits **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_search_proteins_impl(h, residues, seq_off, nq, opts, out); });
}

// Pipelined form of kaamer_gpu_search_proteins: submit returns as soon as the batch is enqueued, wait blocks
// until its hits are in host memory.  At most two batches in flight per handle; the caller's buffers must
// stay valid and unchanged until wait returns.
static int kaamer_gpu_search_proteins_submit_impl(kaamer_gpu_t *h, const uint8_t *residues, const uint64_t *seq_off,
                                                  uint32_t nq, const kaamer_opts *opts, int32_t *ticket) {
  if (!h || !opts || !ticket || (nq && (!residues || !seq_off))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *ticket = -1;
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  int slot = -1;
  KCHECK(search_proteins_submit(h, residues, seq_off, nq, opts, &slot));
  if (slot < 0) {
    set_error("this batch cannot be pipelined (positions wanted, empty, or MaxResults too large): use kaamer_gpu_search_proteins");
    return KAAMER_ERR_ARG;
  }
  *ticket = slot;
  return KAAMER_OK;
}
int kaamer_gpu_search_proteins_submit(kaamer_gpu_t *h, const uint8_t *residues, const uint64_t *seq_off, uint32_t nq,
                                      const kaamer_opts *opts, int32_t *ticket) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_search_proteins_submit_impl(h, residues, seq_off, nq, opts, ticket); });
}
int kaamer_gpu_search_proteins_wait(kaamer_gpu_t *h, int32_t ticket, kaamer_hits **out) {
  if (!h || !out) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  return ::kaamer::guarded([&]() -> int { return search_proteins_wait(h, ticket, out); });
}

int kaamer_gpu_search_proteins_device(kaamer_gpu_t *h, const uint8_t *d_residues, const uint64_t *d_seq_off,
                                      uint32_t nq, const kaamer_opts *opts, const kaamer_dev_result *d_out,
                                      void *stream) {
  KCHECK(check_search_args(h, d_residues, d_seq_off, nq, opts, d_out));
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  return search_proteins_device(h, d_residues, d_seq_off, nq, opts, d_out, (cudaStream_t)stream);
}

static int kaamer_gpu_search_nucleotide_impl(kaamer_gpu_t *h, const uint8_t *nt, const uint64_t *contig_off, uint32_t n_contigs,
                                 const kaamer_opts *opts, kaamer_hits **out) {
  KCHECK(check_search_args(h, nt, contig_off, n_contigs, opts, out));
  *out = nullptr;
  std::lock_guard<std::mutex> lk(h->mu);
  KCUDA(cudaSetDevice(h->device));
  if (!h->idx.table) {
    set_error("no index resident");
    return KAAMER_ERR_ARG;
  }
  HostPhase whole(h, 3);
  return search_nucleotide_host(h, nt, contig_off, n_contigs, opts, out);
}
int kaamer_gpu_search_nucleotide(kaamer_gpu_t *h, const uint8_t *nt, const uint64_t *contig_off, uint32_t n_contigs,
                                 const kaamer_opts *opts, kaamer_hits **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_gpu_search_nucleotide_impl(h, nt, contig_off, n_contigs, opts, out); });
}

void kaamer_gpu_free_hits(kaamer_hits *hits) {
  if (!hits) return;
  delete (HitsOwner *)hits->_owner;
  delete hits;
}

int kaamer_gpu_pinned_alloc(uint64_t bytes, void **out) {
  if (!out) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    set_error("cudaMallocHost(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    return KAAMER_ERR_NOMEM;
  }
  return KAAMER_OK;
}
void kaamer_gpu_pinned_free(void *p) {
  if (p) cudaFreeHost(p);
}

int kaamer_gpu_profile_enable(kaamer_gpu_t *h, int on) {
  if (!h) {
    set_error("null handle");
    return KAAMER_ERR_ARG;
  }
  h->profile = on != 0;
  return KAAMER_OK;
}

int kaamer_gpu_profile_read(kaamer_gpu_t *h, double *kernel_ms, uint64_t *kernel_launches, uint64_t *all_launches,
                            int reset) {
  if (!h) {
    set_error("null handle");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  for (auto &p : h->prof_pending) {
    cudaEventSynchronize(p.b);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      h->prof_ms[p.cls] += ms;
      h->prof_launches[p.cls]++;
    }
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  h->prof_pending.clear();
  for (int c = 0; c < 8; ++c) {
    if (kernel_ms) kernel_ms[c] = h->prof_ms[c];
    if (kernel_launches) kernel_launches[c] = h->prof_launches[c];
  }
  if (all_launches) *all_launches = h->prof_all_launches;
  if (reset) {
    for (int c = 0; c < 8; ++c) {
      h->prof_ms[c] = 0;
      h->prof_launches[c] = 0;
    }
    h->prof_all_launches = 0;
  }
  return KAAMER_OK;
}

int kaamer_gpu_profile_host_read(kaamer_gpu_t *h, double *phase_ms, int reset) {
  if (!h || !phase_ms) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  for (int i = 0; i < 8; ++i) {
    phase_ms[i] = h->prof_host_ms[i];
    if (reset) h->prof_host_ms[i] = 0;
  }
  return KAAMER_OK;
}

}  // extern "C"
