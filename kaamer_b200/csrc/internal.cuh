// internal.cuh — shared declarations of libkaamer_gpu (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <chrono>
#include <exception>
#include <new>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/kaamer_gpu.h"

namespace kaamer {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char *fmt, ...);
#define KCUDA(call)                                                                         \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::kaamer::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return KAAMER_ERR_CUDA;                                                               \
    }                                                                                       \
  } while (0)
#define KCHECK(call)          \
  do {                        \
    int _r = (call);          \
    if (_r != KAAMER_OK) return _r; \
  } while (0)

// no C++ exception may cross the C boundary (the host is a Go process): the allocating entry points run
// their bodies through this
template <class F>
int guarded(F &&f) noexcept {
  try {
    return f();
  } catch (const std::bad_alloc &) {
    set_error("out of host memory");
    return KAAMER_ERR_NOMEM;
  } catch (const std::exception &e) {
    set_error("internal error: %s", e.what());
    return KAAMER_ERR_ARG;
  } catch (...) {
    set_error("internal error");
    return KAAMER_ERR_ARG;
  }
}

// ---- k-mer code space -----------------------------------------------------------------
// EncodeKmer (pkg/kvstore/k_store.go:91-117): key = p01<<23 | p23<<14 | p45<<5 | s6 with
// pair codes p in {0} U [22,462] and single codes s in [0,20].  The device table is
// direct-addressed by the DENSE code  d = ((p01'*442 + p23')*442 + p45')*21 + s6,
// p' = (p==0 ? 0 : p-21) in [0,441]:  442^3*21 = 1,813,366,968 slots.
constexpr uint32_t PAIR_RADIX = 442;
constexpr uint64_t DENSE_SPACE = (uint64_t)PAIR_RADIX * PAIR_RADIX * PAIR_RADIX * 21ull;
constexpr uint32_t CODE_UNKNOWN = 0x80;

// table entry: count:27 | value:37.  count==0 empty; count==1 value = protein id (posting
// inlined); count>=2 value = first index of the posting list in postings[] (mode P: shard:3 | index:34 —
// 2^34 postings per shard hold the whole C4 database, 17 G postings, in ONE shard).
constexpr int ENTRY_VALUE_BITS = 37;
constexpr uint64_t ENTRY_VALUE_MASK = (1ull << ENTRY_VALUE_BITS) - 1;
constexpr uint64_t ENTRY_MAX_COUNT = (1ull << 27) - 1;

// L2-resident presence filter in front of the table: bit (d mod 2^29) is set when dense code d has
// postings.  64 MB stay resident in the 126 MB L2, so a query k-mer that is absent from the database
// (and not a false positive of the fold) is answered without touching HBM.  Measured on B200
// (csrc/tools/l2_filter_bench.cu, profiles/r1_l2_filter.log): 35.6 G lookups/s without the filter,
// 42.5 G/s when 68 % of the lookups pass it, 82.5 G/s when 30 % pass.  Built only for databases of up
// to 2^26 distinct k-mers (bitmap density <= 12 %): on the Swiss-Prot-scale C3 database (186 M keys)
// the composition-skewed query k-mers pass the fold so often (~85-90 %) that the extra L2 round trip
// costs more than the saved probes (measured: class W 0.745 ms with the filter, 0.705 ms without),
// while on the 10 k-protein C2 database the same kernel goes from 0.55 ms to 0.25 ms.
constexpr int FILTER_LOG2_BITS = 29;
constexpr uint32_t FILTER_MASK = (1u << FILTER_LOG2_BITS) - 1u;
constexpr uint64_t FILTER_WORDS = (1ull << FILTER_LOG2_BITS) / 32;
constexpr uint64_t FILTER_MAX_KEYS = 1ull << 26;

#ifdef __CUDACC__
// residue byte -> code 0..20 ("ACDEFGHIKLMNPQRSTUVWY", k_store.go:41) or CODE_UNKNOWN.
// Pure ALU (two packed 5-bit tables), no memory lookup.
__host__ __device__ __forceinline__ uint32_t aa_code(uint32_t c) {
  //            A  B  C  D  E  F  G  H  I  J  K  L | M   N   O   P   Q   R   S   T   U   V   W   X   Y
  // code       0  -  1  2  3  4  5  6  7  -  8  9 | 10  11  -   12  13  14  15  16  17  18  19  -   20
  constexpr uint64_t LO = (0ull) | (31ull << 5) | (1ull << 10) | (2ull << 15) | (3ull << 20) | (4ull << 25) |
                          (5ull << 30) | (6ull << 35) | (7ull << 40) | (31ull << 45) | (8ull << 50) | (9ull << 55);
  constexpr uint64_t HI = (10ull) | (11ull << 5) | (31ull << 10) | (12ull << 15) | (13ull << 20) |
                          (14ull << 25) | (15ull << 30) | (16ull << 35) | (17ull << 40) | (18ull << 45) |
                          (19ull << 50) | (31ull << 55);  // M..X; 'Y' handled below (25 letters > 2x12 slots)
  uint32_t i = c - 'A';
  if (i > 24u) return CODE_UNKNOWN;
  if (i == 24u) return 20u;
  uint32_t v = i < 12u ? (uint32_t)(LO >> (5u * i)) & 31u : (uint32_t)(HI >> (5u * (i - 12u))) & 31u;
  return v == 31u ? CODE_UNKNOWN : v;
}
// p' of two codes: 0 if either is unknown (Go map miss -> 0, k_store.go:100-103), else 1+21x+y
__host__ __device__ __forceinline__ uint32_t pair_dense(uint32_t x, uint32_t y) {
  return ((x | y) & CODE_UNKNOWN) ? 0u : 1u + 21u * x + y;
}
__host__ __device__ __forceinline__ uint32_t single_dense(uint32_t x) {
  return (x & CODE_UNKNOWN) ? 0u : x;  // unknown last residue aliases 'A' (k_store.go:108-110)
}
__host__ __device__ __forceinline__ uint32_t dense_from_codes(uint32_t c0, uint32_t c1, uint32_t c2,
                                                              uint32_t c3, uint32_t c4, uint32_t c5,
                                                              uint32_t c6) {
  return ((pair_dense(c0, c1) * PAIR_RADIX + pair_dense(c2, c3)) * PAIR_RADIX + pair_dense(c4, c5)) * 21u +
         single_dense(c6);
}
// reference u32 key (k_store.go:100-110) from 7 codes
__host__ __device__ __forceinline__ uint32_t key_from_codes(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                            uint32_t c4, uint32_t c5, uint32_t c6) {
  uint32_t p0 = pair_dense(c0, c1), p1 = pair_dense(c2, c3), p2 = pair_dense(c4, c5);
  p0 = p0 ? p0 + 21u : 0u;
  p1 = p1 ? p1 + 21u : 0u;
  p2 = p2 ? p2 + 21u : 0u;
  return (p0 << 23) | (p1 << 14) | (p2 << 5) | single_dense(c6);
}
// reference key -> dense code; returns false if the key is not an EncodeKmer output
__host__ __device__ __forceinline__ bool dense_from_key(uint32_t key, uint32_t *d) {
  uint32_t p0 = (key >> 23) & 0x1FF, p1 = (key >> 14) & 0x1FF, p2 = (key >> 5) & 0x1FF, s = key & 0x1F;
  if ((p0 != 0 && (p0 < 22 || p0 > 462)) || (p1 != 0 && (p1 < 22 || p1 > 462)) ||
      (p2 != 0 && (p2 < 22 || p2 > 462)) || s > 20)
    return false;
  p0 = p0 ? p0 - 21u : 0u;
  p1 = p1 ? p1 - 21u : 0u;
  p2 = p2 ? p2 - 21u : 0u;
  *d = ((p0 * PAIR_RADIX + p1) * PAIR_RADIX + p2) * 21u + s;
  return true;
}
__host__ __device__ __forceinline__ uint32_t key_from_dense(uint32_t d) {
  uint32_t s = d % 21u;
  d /= 21u;
  uint32_t p2 = d % PAIR_RADIX;
  d /= PAIR_RADIX;
  uint32_t p1 = d % PAIR_RADIX, p0 = d / PAIR_RADIX;
  p0 = p0 ? p0 + 21u : 0u;
  p1 = p1 ? p1 + 21u : 0u;
  p2 = p2 ? p2 + 21u : 0u;
  return (p0 << 23) | (p1 << 14) | (p2 << 5) | s;
}
#endif

// ---- device buffers -------------------------------------------------------------------
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;  // capacity in elements
  int ensure(size_t want) {
    if (want <= n && p) return KAAMER_OK;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    size_t cap = want + want / 4 + 64;
    cudaError_t e = cudaMalloc((void **)&p, cap * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes): %s", cap * sizeof(T), cudaGetErrorString(e));
      p = nullptr;
      return KAAMER_ERR_NOMEM;
    }
    n = cap;
    return KAAMER_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

template <class T>
struct PinBuf {
  T *p = nullptr;
  size_t n = 0;
  int ensure(size_t want) {
    if (want <= n && p) return KAAMER_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    size_t cap = want + want / 4 + 64;
    cudaError_t e = cudaMallocHost((void **)&p, cap * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMallocHost(%zu bytes): %s", cap * sizeof(T), cudaGetErrorString(e));
      p = nullptr;
      return KAAMER_ERR_NOMEM;
    }
    n = cap;
    return KAAMER_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
};

// bump allocator over device slabs that persist for the life of the handle: the temporaries of a
// call (translation, row assembly) are carved out of it and the whole arena is reset at the start
// of the next call, so steady-state calls do no cudaMalloc / cudaFree at all
struct Arena {
  struct Slab {
    uint8_t *p;
    size_t cap, used;
  };
  std::vector<Slab> slabs;
  void reset() {
    for (auto &s : slabs) s.used = 0;
  }
  void *alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    for (auto &s : slabs)
      if (s.cap - s.used >= bytes) {
        void *r = s.p + s.used;
        s.used += bytes;
        return r;
      }
    size_t cap = bytes + bytes / 4;
    if (cap < ((size_t)16 << 20)) cap = (size_t)16 << 20;
    uint8_t *p = nullptr;
    cudaError_t e = cudaMalloc((void **)&p, cap);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes): %s", cap, cudaGetErrorString(e));
      return nullptr;
    }
    slabs.push_back(Slab{p, cap, bytes});
    return p;
  }
  template <class T>
  int get(T **out, size_t n) {
    *out = (T *)alloc((n ? n : 1) * sizeof(T));
    return *out ? KAAMER_OK : KAAMER_ERR_NOMEM;
  }
  void release() {
    for (auto &s : slabs) cudaFree(s.p);
    slabs.clear();
  }
};

// ---- peer-mapped shards (mode P) ------------------------------------------
// The dense code space is tiled by up to MAX_PEER_SHARDS key ranges; shard s lives in the HBM of
// the GPU that built it and is mapped into this process (same process: plain pointers with
// peer access enabled; other process: cudaIpcOpenMemHandle).  The search kernels resolve the
// owner of a dense code with MAX_PEER_SHARDS-1 compares and read the 8-byte entry (and the
// posting list) straight through NVLink.  A posting-list reference carries its shard in the top
// bits of the 37-bit entry value, so per-shard posting indices are limited to 2^34.
constexpr int MAX_PEER_SHARDS = 8;
constexpr int PEER_SHARD_SHIFT = 34;
constexpr uint64_t PEER_LOCAL_MASK = (1ull << PEER_SHARD_SHIFT) - 1;
struct PeerView {
  uint32_t fence[MAX_PEER_SHARDS + 1];  // fence[s] = first dense code of shard s; unused = 0xFFFFFFFF
  int32_t n;
  int32_t self;  // the shard that lives in this GPU's own HBM (-1: none)
  // presence filter: one bit per dense code, set when the k-mer has postings on its owner shard.
  // A local replica (227 MB) that keeps the k-mers absent from the database off NVLink: only
  // probes that will find something cross to the owner GPU.  nullptr: filter off.
  const uint32_t *presence;
  // replicated table (KAAMER_ATTACH_REPLICATE_TABLE): the direct-address table is 14.5 GB whatever the
  // size of the database, so every GPU can hold ALL of it; only the posting lists — the part of the
  // index that grows with the database — stay sharded.  Entries are copied from the owners at attach
  // time with the owner shard already folded into the value of multi-posting entries: the first probe
  // of every lookup is local, NVLink carries posting lists only.  nullptr: probe the owner's table.
  const uint64_t *full_table;
  const uint64_t *table[MAX_PEER_SHARDS];
  const uint32_t *postings[MAX_PEER_SHARDS];
};

// shareable device allocation (vmm.cu): CUDA VMM memory with 2 MiB pages that can be exported to
// other processes as a POSIX file descriptor and mapped for other devices
struct VmmAlloc {
  void *ptr = nullptr;
  size_t bytes = 0;  // rounded up to the allocation granularity
  size_t granularity = 0;
  unsigned long long handle = 0;  // CUmemGenericAllocationHandle
  int device = 0;
  bool imported = false;
};
int vmm_alloc(int device, size_t bytes, VmmAlloc *out);
int vmm_export_fd(const VmmAlloc &a, int *fd);
int vmm_import_fd(int fd, size_t bytes, int device, VmmAlloc *out);
int vmm_grant(void *ptr, size_t bytes, int device);  // access for one more device of this process
void vmm_free(VmmAlloc *a);

// ---- the resident index ---------------------------------------------------------------
struct DevIndex {
  uint64_t *table = nullptr;  // [d_hi - d_lo] direct-address entries
  uint64_t d_lo = 0, d_hi = 0;
  uint32_t *postings = nullptr;  // [n_postings]
  uint64_t n_postings = 0;
  uint32_t *filter = nullptr;    // [FILTER_WORDS] folded presence bits of this handle's keys (or nullptr)
  // sorted form kept for export / save (keys ascending, offsets, postings descending)
  uint32_t *keys = nullptr;
  uint64_t *offsets = nullptr;
  uint64_t n_keys = 0;
  // protein table (optional)
  uint64_t *prot_off = nullptr;
  uint8_t *prot_res = nullptr;
  std::vector<uint64_t> h_prot_off;  // host copy (pair sizing in align.cu)
  uint64_t n_prot_res = 0;
  uint32_t max_protein_id = 0;
  bool has_proteins = false;
  bool lists_sorted = false;  // every posting list strictly descending (the builders; views after a check)
  uint64_t n_proteins = 0, n_aa = 0, n_kmers = 0;
  // annotation table for the row formatter (format.cu), host memory, indexed by protein id:
  // Protein.EntryId and Protein.Length (pkg/kvstore/protein.proto); empty = not loaded
  std::vector<uint64_t> annot_off;
  std::vector<char> annot_ids;
  std::vector<int32_t> annot_len;
  // mode P: the shards of the other ranks, mapped into this process (api.cu kaamer_gpu_attach_shards)
  PeerView peer{};               // host copy; peer.n == 0: not attached
  PeerView *d_peer = nullptr;    // device copy read by the kernels
  std::vector<VmmAlloc> imported;  // mappings of the other processes' shards
  uint32_t *presence = nullptr;    // presence filter over the whole key space (mode P, api.cu)
  uint64_t *full_table = nullptr;  // replicated table over the whole key space (mode P, api.cu)
  uint32_t *repl_postings = nullptr;  // local copy of every shard's postings (KAAMER_ATTACH_REPLICATE_POSTINGS)
  // table AND postings replicated and the entries rewritten to offsets into repl_postings: the handle is a plain
  // whole index again (full_table / repl_postings) and the protein search runs its non-PEER kernels on it
  bool flat_view = false;
  // key-range shards keep table and postings in shareable memory (vmm.cu); a full index uses cudaMalloc
  VmmAlloc vm_table, vm_postings;
};

struct SearchWorkspace {
  DevBuf<uint8_t> residues;
  DevBuf<uint64_t> seq_off;
  DevBuf<uint32_t> n_hits, hit_base, lists, kmin;
  DevBuf<int32_t> size_in_kmer;
  DevBuf<uint64_t> pool, hit_off, out_hits;
  DevBuf<uint64_t> counters;   // see search.cu
  DevBuf<uint32_t> ghash;      // global-memory hash scratch (class G)
  DevBuf<uint8_t> any0;        // nucleotide mode (search.cu)
  PinBuf<uint64_t> h_counters, h_packed;
  // finish.cu
  DevBuf<uint64_t> f_scan;     // [5][nq+1] sizes, scanned in place
  DevBuf<uint32_t> f_posbits, f_keep, f_trim;
  DevBuf<uint8_t> f_tmp;
  DevBuf<uint8_t> a_pairs, a_out, a_scratch, a_tables, a_rev, a_text;  // align.cu
  DevBuf<uint64_t> a_revoff;
  void release_all() {
    residues.release(); seq_off.release(); n_hits.release(); hit_base.release(); lists.release();
    kmin.release(); size_in_kmer.release(); pool.release(); hit_off.release(); out_hits.release();
    counters.release(); ghash.release(); any0.release(); h_counters.release(); h_packed.release();
    f_scan.release(); f_posbits.release(); f_keep.release(); f_trim.release(); f_tmp.release();
    a_pairs.release(); a_out.release(); a_scratch.release(); a_tables.release(); a_rev.release(); a_text.release();
    a_revoff.release();
  }
};

// ORFs of a batch of contigs on the device, in GetORFs order per contig (translate.cu)
struct OrfSet {
  uint64_t n = 0, n_seq = 0, n_alts = 0;
  uint32_t *contig = nullptr;
  int64_t *start = nullptr, *end = nullptr;
  uint8_t *plus = nullptr;
  uint64_t *seq_off = nullptr;   // [n+1]
  uint8_t *seq = nullptr;        // amino acids
  uint64_t *alts_off = nullptr;  // [n+1]
  int32_t *alts = nullptr;       // StartsAlternative
};

struct ProfSpan {
  cudaEvent_t a = nullptr, b = nullptr;
  int cls = 0;
};

// process-wide cache of pinned host blocks (api.cu): cudaHostAlloc / cudaFreeHost cost
// hundreds of microseconds each, result buffers are recycled instead
void *pinned_get(size_t bytes, size_t *cap);
void pinned_put(void *p, size_t cap);

// owner of the pinned host buffers behind a kaamer_hits / kaamer_orfs
struct HitsOwner {
  std::vector<std::pair<void *, size_t>> pinned;
  ~HitsOwner() {
    for (auto &p : pinned) pinned_put(p.first, p.second);
  }
  template <class T>
  int alloc(T **out, size_t n) {
    size_t cap = 0;
    void *p = pinned_get((n ? n : 1) * sizeof(T), &cap);
    if (!p) return KAAMER_ERR_NOMEM;  // error text set by pinned_get
    pinned.emplace_back(p, cap);
    *out = (T *)p;
    return KAAMER_OK;
  }
};

}  // namespace kaamer

struct kaamer_gpu {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  kaamer::DevIndex idx;
  kaamer::SearchWorkspace ws;
  kaamer::Arena arena;
  // profiling
  bool profile = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // [0..2] search size classes W, M, G; [3] Smith-Waterman; [4] H2D of a host call; [5] CSR
  // compaction + D2H; [6] search class D (dense databases); [7] translation/ORF kernels
  double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t prof_launches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // host-call pipeline: H2D of chunk c+1 on copy_stream overlaps the search of chunk c
  cudaStream_t copy_stream = nullptr;
  cudaStream_t side_stream = nullptr;  // class G / long-query kernels run underneath the main search stream
  cudaEvent_t chunk_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done_ev = nullptr;
  uint64_t prof_all_launches = 0;
  std::vector<kaamer::ProfSpan> prof_pending;
  double prof_host_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // per-handle (= per-device context) one-time setup: constant tables and kernel attributes
  // genetic code of the translated search (translate.cu); gcode_set false: table 11, the reference's
  char gcode_aas[64] = {0};
  uint64_t gcode_starts = 0;
  bool gcode_set = false;
  void *pending = nullptr;          // PendingSlot[2] of the submit / wait pipeline (search.cu)
  kaamer_aln_model aln_model{};     // alignment DP model (align.cu); aln_model_set false: the reference default
  bool aln_model_set = false;
  uint32_t ghash_slots = 1u << 20;  // class G histogram slots per CTA; grown on ST_GHASH_OVERFLOW (search.cu)
  size_t dense_smem_set = 0;        // dynamic shared memory the class-D kernels are configured for
  bool aln_ready = false, shard_attrs_ready = false;  // wall clock of host-call phases (kaamer_gpu_profile_host_read)
};

namespace kaamer {
// index.cu
int index_from_view(kaamer_gpu *h, const kaamer_index_view *v);
int index_build(kaamer_gpu *h, const uint8_t *residues, const uint64_t *seq_off, const uint32_t *ids,
                uint64_t n_records, int keep_proteins, uint64_t shard_lo = 0, uint64_t shard_hi = 0);
void index_release(kaamer_gpu *h);
// search.cu
int search_proteins_device(kaamer_gpu *h, const uint8_t *d_res, const uint64_t *d_off, uint32_t nq,
                           const kaamer_opts *o, const kaamer_dev_result *out, cudaStream_t st, int nt_mode = 0,
                           uint8_t *d_any0 = nullptr, const uint64_t *d_prev_counters = nullptr);
int search_counted(kaamer_gpu *h, const uint8_t *d_res, const uint64_t *d_off, uint32_t nq, const kaamer_opts *o,
                   int nt_mode, uint8_t *d_any0, cudaStream_t st);
int search_proteins_host(kaamer_gpu *h, const uint8_t *res, const uint64_t *off, uint32_t nq,
                         const kaamer_opts *o, kaamer_hits **out);
int search_nucleotide_host(kaamer_gpu *h, const uint8_t *nt, const uint64_t *coff, uint32_t nc,
                           const kaamer_opts *o, kaamer_hits **out);
int search_proteins_submit(kaamer_gpu *h, const uint8_t *res, const uint64_t *off, uint32_t nq, const kaamer_opts *o,
                           int *slot_out);
int search_proteins_wait(kaamer_gpu *h, int slot, kaamer_hits **out);
void release_pending(kaamer_gpu *h);
int grow_ghash(kaamer_gpu *h, uint64_t need);
#ifdef __CUDACC__
// smallest Kmatch that survives FilterResults (search.go:195): the hit is dropped when
// float64(Kmatch)/float64(SizeInKmer) < MinKRatio || Kmatch < MinKMatch; both tests are
// monotone in Kmatch, so the kept set is {Kmatch >= kmin}.
__device__ __forceinline__ uint32_t filter_kmin(long long min_kmatch, double ratio, int32_t size) {
  long long k = min_kmatch > 1 ? min_kmatch : 1;
  double ds = (double)size;
  if (ratio != ratio) {
    // NaN: `x < NaN` is false, the ratio test never drops a hit
  } else if (ratio > 0.0) {
    double g = ceil(ratio * ds);
    if (!(g < 4.0e9)) return 0xFFFFFFFFu;  // nothing can pass (Kmatch <= SizeInKmer < 2^31)
    long long kr = (long long)g;
    while (kr > 0 && !((double)(kr - 1) / ds < ratio)) --kr;
    while ((double)kr / ds < ratio) ++kr;
    if (kr > k) k = kr;
  }
  return k > 0xFFFFFFFEll ? 0xFFFFFFFFu : (uint32_t)k;
}
#endif
// search.cu: presence bitmap of all attached shards (streams every shard table once)
int build_presence(kaamer_gpu *h, const PeerView &pv, uint32_t *d_bits, cudaStream_t st);
// search.cu: local copy of every shard's table range, multi-posting entries tagged with their shard
int replicate_table(kaamer_gpu *h, const PeerView &pv, uint64_t *d_full, cudaStream_t st);
int flatten_table(kaamer_gpu *h, uint64_t *d_full, const uint64_t *base, int n_shards, cudaStream_t st);
// align.cu
int align_pairs(kaamer_gpu *h, const uint8_t *q_res, const uint64_t *q_off, const uint32_t *pair_q,
                const uint32_t *pair_s, uint32_t n_pairs, const kaamer_aln_opts *o, kaamer_aln *out,
                kaamer_aln_text **text = nullptr);
void default_align_model(kaamer_aln_model *m);
// translate.cu
int orfs_device(kaamer_gpu *h, const uint8_t *d_nt, const uint64_t *h_coff, uint32_t nc, OrfSet *out,
                cudaStream_t st);
void orfset_release(OrfSet *o);
// finish.cu: positions, SetBestStartCodon, final FilterResults, row assembly, D2H
int finish_rows(kaamer_gpu *h, const uint8_t *d_res, const uint64_t *d_off, uint32_t nq, const kaamer_opts *o,
                int nt_mode, const uint8_t *d_any0, const OrfSet *orfs, kaamer_hits *hits, HitsOwner *owner,
                cudaStream_t st);
// wall-clock phase timer: adds the elapsed milliseconds to h->prof_host_ms[phase] when profiling is on
struct HostPhase {
  kaamer_gpu *h;
  int phase;
  std::chrono::steady_clock::time_point t0;
  HostPhase(kaamer_gpu *h_, int phase_) : h(h_), phase(phase_), t0(std::chrono::steady_clock::now()) {}
  void stop() {
    if (h && h->profile)
      h->prof_host_ms[phase] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    h = nullptr;
  }
  ~HostPhase() { stop(); }
};
void profile_begin(kaamer_gpu *h, cudaStream_t st, int cls);
void profile_end(kaamer_gpu *h, cudaStream_t st);
}  // namespace kaamer
