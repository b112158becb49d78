/*
 * kaamer_gpu.h — C ABI of libkaamer_gpu.so, the B200 (sm_100a) implementation of the
 * zorino/kaamer search hot path.  Plain C types only: this is what a cgo shim in
 * pkg/search binds.  All `file:line` citations are relative to the
 * reference repository root.
 *
 * Conventions
 *  - every function returns KAAMER_OK (0) or a negative error code and never aborts the
 *    process; kaamer_gpu_last_error() returns a thread-local message for the last failure;
 *  - input pointers are caller-owned and only read during the call (cgo pointer rule);
 *  - outputs are library-owned (pinned host memory) and released by the matching *_free;
 *  - a handle may be used from several threads; calls on one handle are serialised;
 *  - hits are ordered (Kmatch desc, subject id asc): the reference's tie order is random
 *    (pkg/search/search.go:132-152), this is the canonical representative;
 *  - there is NO CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef KAAMER_GPU_H
#define KAAMER_GPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KAAMER_OK 0
#define KAAMER_ERR_CUDA (-1)
#define KAAMER_ERR_ARG (-2)
#define KAAMER_ERR_IO (-3)
#define KAAMER_ERR_FORMAT (-4)
#define KAAMER_ERR_NOMEM (-5)
#define KAAMER_ERR_LIMIT (-6)

#define KAAMER_KMER_SIZE 7 /* pkg/search/search.go:45 */

typedef struct kaamer_gpu kaamer_gpu_t;

/* Flat index (".kidx"): what pkg/makedb exports once from the badger
 * stores by walking kmer_store -> kcomb_store exactly as KmerSearch does
 * (pkg/search/search.go:421-429).  keys ascending (== badger's big-endian byte order,
 * pkg/kvstore/k_store.go:69-70); postings per key = unique protein ids, descending
 * (pkg/kvstore/kv_store.go:284-305). */
typedef struct kaamer_index_view {
  uint64_t n_keys;
  uint64_t n_postings;
  const uint32_t *keys;      /* [n_keys] */
  const uint64_t *offsets;   /* [n_keys+1] */
  const uint32_t *postings;  /* [n_postings] */
  uint64_t n_proteins;       /* KStats.NumberOfProteins (pkg/kvstore/kstats.proto) */
  uint64_t n_aa;             /* KStats.NumberOfAA */
  uint64_t n_kmers;          /* KStats.NumberOfKmers */
  uint32_t max_protein_id;
  uint32_t _pad;
  /* optional protein table indexed by protein id (needed by kaamer_gpu_align):
   * residues of id i = prot_residues[prot_seq_off[i] .. prot_seq_off[i+1]) */
  const uint64_t *prot_seq_off; /* [max_protein_id+2] or NULL */
  const uint8_t *prot_residues; /* or NULL */
  /* optional key-range shard (mode S): only k-mer keys whose dense code is in
   * [shard_lo, shard_hi) are resident; 0,0 = whole key space */
  uint64_t shard_lo, shard_hi;
} kaamer_index_view;

/* Replaces kvstore.KVStoresNew(dbPath, ..., readOnly=true) at server start
 * (api/server.go:65, pkg/kvstore/kv_stores.go:46-104) for the search path. */
int kaamer_gpu_open(const char *kidx_path, int device, kaamer_gpu_t **out);
int kaamer_gpu_open_view(const kaamer_index_view *view, int device, kaamer_gpu_t **out);
/* Key-range shard of a `.kidx` file: only the k-mers whose dense code lies in [shard_lo, shard_hi)
 * become resident (in shareable memory, see kaamer_gpu_shard_export); KStats cover the whole file.
 * kaamer_gpu_kidx_fences cuts the key space of the file into n_shards contiguous ranges of equal
 * posting mass: fences[n_shards+1], fences[0] = 0, fences[n_shards] = kaamer_gpu_dense_space(). */
int kaamer_gpu_open_shard(const char *kidx_path, int device, uint64_t shard_lo, uint64_t shard_hi,
                          kaamer_gpu_t **out);
int kaamer_gpu_kidx_fences(const char *kidx_path, int n_shards, uint64_t *fences);
/* Replaces pkg/makedb (processProteinInputFASTA, inputFASTA.go:195-250) + pkg/indexdb
 * (IndexStore, indexdb.go:68-150) for the device index: records -> sorted unique
 * (k-mer, protein id) -> CSR, built on the GPU.  ids[i] is the protein id of record i. */
int kaamer_gpu_build(const uint8_t *residues, const uint64_t *seq_off, const uint32_t *ids,
                     uint64_t n_records, int keep_proteins, int device, kaamer_gpu_t **out);
/* same, keeping only the k-mers whose dense code lies in [shard_lo, shard_hi) (mode S: every rank
 * builds its own key range from the full record set; 0,0 = whole key space).  KStats cover the
 * whole input. */
int kaamer_gpu_build_shard(const uint8_t *residues, const uint64_t *seq_off, const uint32_t *ids,
                           uint64_t n_records, int keep_proteins, int device, uint64_t shard_lo,
                           uint64_t shard_hi, kaamer_gpu_t **out);
void kaamer_gpu_close(kaamer_gpu_t *h);

/* ---- streaming builder: the same makedb + indexdb step for record sets that do not fit the device
 * at once (a 50 M-protein FASTA read in chunks; the synthetic C4 database generated on the device).
 * The handle's key range [shard_lo, shard_hi) (0,0 = everything) is covered by consecutive PASSES over
 * sub-ranges; in every pass the caller feeds ALL records again, in device-resident chunks, and the
 * builder keeps the k-mers that fall into the pass range (build_stream.cu).  Peak memory = table +
 * max_postings * 4 B + 16 B per pair of ONE pass.  KStats are taken from the first pass.  d_ids NULL:
 * record i of the chunk gets protein id id_base + i.  The sorted keys/offsets export form is not kept
 * (kaamer_gpu_index_copy / _save refuse on such a handle).  shareable != 0: table and postings live in
 * shareable memory (kaamer_gpu_shard_export). ---- */
typedef struct kaamer_builder kaamer_builder_t;
int kaamer_gpu_builder_open(int device, uint64_t shard_lo, uint64_t shard_hi, uint64_t max_postings, int shareable,
                            kaamer_builder_t **out);
int kaamer_gpu_builder_pass_begin(kaamer_builder_t *b, uint64_t pass_lo, uint64_t pass_hi, uint64_t max_pairs);
int kaamer_gpu_builder_add_device(kaamer_builder_t *b, const uint8_t *d_residues, const uint64_t *d_seq_off,
                                  const uint32_t *d_ids, uint32_t id_base, uint64_t n_records, void *stream);
int kaamer_gpu_builder_pass_end(kaamer_builder_t *b, void *stream);
int kaamer_gpu_builder_finish(kaamer_builder_t *b, kaamer_gpu_t **out); /* consumes b */
void kaamer_gpu_builder_abort(kaamer_builder_t *b);

/* KStats (api/server.go:125-132 /api/dbinfo; feeds the e-value, pkg/align/align.go:141) */
int kaamer_gpu_dbstats(kaamer_gpu_t *h, uint64_t *n_proteins, uint64_t *n_aa, uint64_t *n_kmers);
/* export of the resident index (sizes, then copy into caller buffers; also `.kidx` writer) */
int kaamer_gpu_index_sizes(kaamer_gpu_t *h, uint64_t *n_keys, uint64_t *n_postings);
int kaamer_gpu_index_copy(kaamer_gpu_t *h, uint32_t *keys, uint64_t *offsets, uint32_t *postings);
int kaamer_gpu_save(kaamer_gpu_t *h, const char *kidx_path);

/* SearchOptions subset that reaches the hot path (pkg/search/search.go:56-71;
 * defaults api/server.go:194-207: 10 / 0.05 / 10). */
typedef struct kaamer_opts {
  int64_t min_kmatch;     /* MinKMatch */
  double min_kratio;      /* MinKRatio */
  int32_t max_results;    /* MaxResults */
  uint8_t want_positions; /* ExtractPositions (always on for nucleotide, search.go:416) */
  uint8_t _pad[3];
} kaamer_opts;

/* CSR result of a batch. For protein search rows == queries (row i = query i). */
typedef struct kaamer_hits {
  uint32_t n_rows;
  uint32_t _pad;
  uint64_t n_hits;
  uint64_t *hit_off;     /* [n_rows+1] */
  uint32_t *subject_id;  /* [n_hits]  Hit.Key   (search.go:112) */
  uint32_t *kmatch;      /* [n_hits]  Hit.Kmatch (search.go:113) */
  int32_t *size_in_kmer; /* [n_rows]  Query.SizeInKmer (search.go:290-293) */
  /* want_positions: PositionHits[hit] as one byte per query k-mer position (search.go:442-452) */
  uint64_t *pos_off; /* [n_hits+1] or NULL */
  uint8_t *pos;      /* or NULL */
  /* nucleotide search only (rows == surviving ORFs, search_nucleotide.go:78-123) */
  uint32_t *row_contig;  /* [n_rows] index of the contig in the batch */
  int64_t *row_start;    /* Location.StartPosition after SetBestStartCodon (dna.go:252-257) */
  int64_t *row_end;      /* Location.EndPosition */
  uint8_t *row_plus;     /* Location.PlusStrand */
  uint64_t *row_seq_off; /* [n_rows+1] */
  uint8_t *row_seq;      /* Query.Sequence (amino acids, after start-codon trimming) */
  /* work counters of the batch */
  uint64_t n_lookups;    /* query k-mers looked up */
  uint64_t n_increments; /* (query, subject) counter increments */
  void *_owner;
} kaamer_hits;

/* Replaces, per batch of protein queries, the key producer loop (search_protein.go:94-98),
 * KmerSearch (search.go:414-440), sortMapByValue (:132-152) and FilterResults (:189-220).
 * residues: query sequences as read by GetQueriesFasta (already upper-cased by the host
 * reader, search.go:295); seq_off[nq+1].  Queries with SizeInKmer < 7 yield no hits
 * (search_protein.go:74-76 kills the worker instead; documented deviation).
 *
 * Limits (every one answered with an error code, never with a partial result):
 *  - a posting list holds at most 2^27 - 1 proteins, a shard at most 2^34 postings, protein ids < 2^32 - 1
 *    (KAAMER_ERR_LIMIT / _FORMAT when the index is built or opened);
 *  - the kernel is chosen from the density of the database: shared-memory histograms (classes W / M) up to ~3.5
 *    background postings per query k-mer, the filter kernel of search_dense3.cuh (class D) above; a query whose
 *    repeated subjects overflow a class is searched again by the next one, last by class G (global-memory
 *    histogram), which GROWS its table and repeats the batch when a query touches more subjects than it holds —
 *    the reference returns results for such queries, so does this call (KAAMER_ERR_LIMIT only beyond 2^30 slots);
 *  - queries of more than 60000 k-mers on a dense database go to class G directly. */
int kaamer_gpu_search_proteins(kaamer_gpu_t *h, const uint8_t *residues, const uint64_t *seq_off,
                               uint32_t nq, const kaamer_opts *opts, kaamer_hits **out);

/* The same call in two halves, for hosts that keep the GPU busy across requests (the reference runs
 * nbOfThreads queries concurrently, api/server.go:55-59): submit enqueues the whole batch and returns, wait
 * blocks until its hits are in host memory (one event synchronisation).  Up to TWO batches may be in flight
 * per handle — submit batch k+1, then wait for batch k.  residues / seq_off must stay valid and unchanged
 * until the matching wait returns (cgo: C-allocated memory, e.g. kaamer_gpu_pinned_alloc).  Batches that want
 * positions, are empty, or allow more than 2^25 hits in total are refused: use the blocking call. */
int kaamer_gpu_search_proteins_submit(kaamer_gpu_t *h, const uint8_t *residues, const uint64_t *seq_off,
                                      uint32_t nq, const kaamer_opts *opts, int32_t *ticket);
int kaamer_gpu_search_proteins_wait(kaamer_gpu_t *h, int32_t ticket, kaamer_hits **out);

/* Replaces GetORFs (dna.go:65-181) + the per-ORF loop of NucleotideSearch / FastqSearch
 * (search_nucleotide.go:76-130, search_fastq.go:78-140) incl. SetBestStartCodon (dna.go:198-272). */
int kaamer_gpu_search_nucleotide(kaamer_gpu_t *h, const uint8_t *nt, const uint64_t *contig_off,
                                 uint32_t n_contigs, const kaamer_opts *opts, kaamer_hits **out);
void kaamer_gpu_free_hits(kaamer_hits *);

/* The genetic code of GetORFs.  The reference ALWAYS translates with table 11 (gcodeBacteria, dna.go:106: the
 * geneticCode argument is ignored); that is the default here.  Honouring -g is a behaviour change and only
 * happens through this explicit call: aas64 = the amino-acid letter ('*' = stop) of the 64
 * codons in TCAG order (index 16*b0 + 4*b1 + b2, t=0 c=1 a=2 g=3), start_mask = bit i set when codon i is a
 * start codon — the tables of pkg/search/gcode.go.  NULL restores table 11. */
int kaamer_gpu_set_genetic_code(kaamer_gpu_t *h, const char *aas64, uint64_t start_mask);

/* GetORFs alone (dna.go:65-181), for parity tests of the translation kernels. */
typedef struct kaamer_orfs {
  uint64_t n_orfs;
  uint32_t *contig;
  int64_t *start, *end;
  uint8_t *plus;
  uint64_t *seq_off; /* [n_orfs+1] */
  uint8_t *seq;
  uint64_t *alts_off; /* [n_orfs+1] */
  int32_t *alts;      /* StartsAlternative */
  void *_owner;
} kaamer_orfs;
int kaamer_gpu_get_orfs(kaamer_gpu_t *h, const uint8_t *nt, const uint64_t *contig_off,
                        uint32_t n_contigs, kaamer_orfs **out);
void kaamer_gpu_free_orfs(kaamer_orfs *);

/* align.Align (pkg/align/align.go:46-161) for a batch of (query, subject) pairs.
 * Subjects are protein ids resolved in the resident protein table. */
typedef struct kaamer_aln_opts {
  double lambda, K;        /* MatrixScores.Lambda/K (matrixScores.go:59: 0.267 / 0.041) */
  int32_t gap_open;        /* MatrixScores.GapOpen: only used by the `score == -GapOpen` test (align.go:127) */
  int32_t gap_extend;      /* MatrixScores.GapExtend (align.go:130) */
  uint64_t number_of_aa;   /* KStats.NumberOfAA (align.go:141); 0 = use the index' value */
} kaamer_aln_opts;
typedef struct kaamer_aln {
  float identity, similarity;                 /* float32 (align.go:72-101) */
  int32_t length, mismatches, gap_openings, raw;
  double bitscore, evalue;                    /* float64 (align.go:136,141) */
  int32_t query_start, query_end, subject_start, subject_end;
  int32_t dp_score;                           /* optimum of the DP (sum of segment scores) */
  int32_t status;                             /* 0 ok, 1 illegal letter (biogo error ignored, align.go:67) */
} kaamer_aln;
int kaamer_gpu_align(kaamer_gpu_t *h, const uint8_t *q_residues, const uint64_t *q_off,
                     const uint32_t *pair_query, const uint32_t *pair_subject, uint32_t n_pairs,
                     const kaamer_aln_opts *opts, kaamer_aln *out /* [n_pairs], caller-owned */);
/* Same, and AlignmentResult.AlnString (align.go:69-103) of every pair: text of pair i =
 * text[off[i] .. off[i+1]) = "<gapped query>\n<match line>\n<gapped subject>" (no terminator; "\n\n" for an
 * empty alignment, exactly what fmt.Sprintf("%s\n%s\n%s") gives the JSON writer, search.go:497-503). */
typedef struct kaamer_aln_text {
  uint64_t *off; /* [n_pairs+1] */
  char *text;
  void *_owner;
} kaamer_aln_text;
int kaamer_gpu_align_text(kaamer_gpu_t *h, const uint8_t *q_residues, const uint64_t *q_off,
                          const uint32_t *pair_query, const uint32_t *pair_subject, uint32_t n_pairs,
                          const kaamer_aln_opts *opts, kaamer_aln *out, kaamer_aln_text **text);
void kaamer_gpu_free_aln_text(kaamer_aln_text *t);

/* The DP model of the alignment stage.  The reference hard-wires align.SWAffine{Matrix: matrix.BLOSUM62,
 * GapOpen: -11} whatever -mat / -gop / -gex say (align.go:62-65); that is the default here and nothing
 * changes it implicitly.  A host that wants to honour the options (a behaviour change, behind
 * this explicit call only), or that has to match a biogo whose BLOSUM62 carries a non-zero gap row, sets the
 * model: matrix[26*26] in biogo alphabet.Protein order "-ABCDEFGHIJKLMNPQRSTVWXYZ*" (pkg/align/
 * matrixScores.go:107), row / column 0 = the per-residue gap cost SWAffine adds for every gap residue,
 * gap_open = SWAffine.GapOpen (negative).  NULL restores the default. */
typedef struct kaamer_aln_model {
  int8_t matrix[26 * 26];
  int32_t gap_open;
} kaamer_aln_model;
int kaamer_gpu_default_align_model(kaamer_aln_model *out);
int kaamer_gpu_set_align_model(kaamer_gpu_t *h, const kaamer_aln_model *model);
/* How the calling thread's last kaamer_gpu_align[_text] call was scheduled (measurement / test aid, no
 * reference counterpart): out[0] = long pairs (one CTA each), out[1] = pairs run one warp each (32-bit lanes),
 * out[2] = packed jobs (two pairs per warp in int16x2 lanes with the DPX instructions: taken for the default
 * zero-gap-row model when both sequences are shorter than 16384 and largest matrix entry x min(n, m) <= 16000,
 * i.e. min(n, m) <= 1454 with BLOSUM62; results are identical either way). */
void kaamer_gpu_align_last_plan(uint32_t out[3]);

/* ---- device-resident entry points (inputs already in HBM; used by bench.py `value`, by the
 * multi-GPU drivers and by callers that keep query batches on the device).  All pointers are
 * device pointers; work is enqueued on `stream` (a cudaStream_t) and is asynchronous.  The
 * handle's scratch buffers are shared by these calls: enqueue all device-resident calls of one
 * handle on ONE stream (or order them with events); use one handle per stream otherwise. ---- */
typedef struct kaamer_dev_result {
  uint32_t *n_hits;      /* [nq]   hits kept per query */
  uint32_t *hit_base;    /* [nq]   first slot of the query's hits in `pool` */
  int32_t *size_in_kmer; /* [nq] */
  uint64_t *pool;        /* [pool_cap] (subject_id | (uint64)kmatch << 32), rank order per query */
  uint64_t pool_cap;
  uint64_t *counters;    /* [16]: pool demand, n_lookups, n_increments, status flags,
                            [4..6] lookups per size class S/M/G, [8..10] increments per class */
} kaamer_dev_result;
int kaamer_gpu_search_proteins_device(kaamer_gpu_t *h, const uint8_t *d_residues,
                                      const uint64_t *d_seq_off, uint32_t nq,
                                      const kaamer_opts *opts, const kaamer_dev_result *d_out,
                                      void *stream);

/* ---- key-range sharded search (mode S; new with respect to the single-process
 * reference: Kmatch is a sum over query k-mers, search.go:431-436, so it decomposes over a
 * partition of the key space).  Device pointers, asynchronous on `stream`; the all-to-all
 * exchanges between the steps belong to the caller (NCCL / torch.distributed). ---- */
uint64_t kaamer_gpu_dense_space(void); /* size of the dense 7-mer code space the fences partition */
/* step 1, home rank.  fences[n_shards+1] (host): shard s owns dense codes [fences[s], fences[s+1]).
 * Count pass (d_codes == NULL): d_counts[s*nq + q] = k-mers of query q owned by shard s, and
 * d_size_in_kmer[q].  Fill pass (d_codes != NULL): d_offsets = exclusive scan of d_counts;
 * d_codes receives the dense codes in (shard, query) order. */
int kaamer_gpu_shard_route(kaamer_gpu_t *h, const uint8_t *d_residues, const uint64_t *d_seq_off, uint32_t nq,
                           const uint64_t *fences, int n_shards, uint32_t *d_counts, const uint64_t *d_offsets,
                           uint32_t *d_codes, int32_t *d_size_in_kmer, void *stream);
/* step 2, owner shard.  Segment i = d_codes[d_seg_off[i] .. d_seg_off[i+1]) = one query's k-mers
 * on this shard.  Emits every (subject | partial count << 32) of the segment into d_pool at
 * d_part_base[i], d_part_n[i] of them.  d_counters[16]: [0] pool demand, [1] lookups,
 * [2] increments, [3] status (1 = pool overflow: retry with pool_cap >= demand). */
int kaamer_gpu_shard_count(kaamer_gpu_t *h, const uint32_t *d_codes, const uint64_t *d_seg_off, uint32_t n_segments,
                           uint32_t *d_part_n, uint64_t *d_part_base, uint64_t *d_pool, uint64_t pool_cap,
                           uint64_t *d_counters, void *stream);
/* pool -> segment order: d_out[d_part_off[i] + j] = pool[d_part_base[i] + j] */
int kaamer_gpu_shard_gather(kaamer_gpu_t *h, const uint32_t *d_part_n, const uint64_t *d_part_base,
                            const uint64_t *d_part_off, const uint64_t *d_pool, uint32_t n_segments, uint64_t *d_out,
                            void *stream);
/* step 3, home rank.  Partial entries of (shard s, query q) = d_part[d_part_off[s*(nq+1)+q] ..
 * d_part_off[s*(nq+1)+q+1]).  Sums the partial counts per subject, applies FilterResults
 * (search.go:189-220) and ranks: same output as kaamer_gpu_search_proteins_device. */
int kaamer_gpu_shard_merge(kaamer_gpu_t *h, const uint64_t *d_part, const uint64_t *d_part_off, int n_shards,
                           uint32_t nq, const int32_t *d_size_in_kmer, const kaamer_opts *opts,
                           const kaamer_dev_result *d_out, void *stream);

/* ---- peer-mapped shards (mode P): the key-range shards of all GPUs of one NVSwitch
 * domain form ONE index.  Every rank builds (or loads) its own range, exports it, and attaches the
 * exports of all ranks; from then on the ordinary entry points of the handle
 * (kaamer_gpu_search_proteins[_device], _search_nucleotide) see the whole key space: the search
 * kernels resolve the owner of each k-mer and read its table entry and posting list from that
 * GPU's HBM through NVLink (peer loads inside the kernel; no all-to-all, no merge step — Kmatch
 * is accumulated at the query's home GPU exactly as in the single-GPU path, search.go:431-436).
 * Shards are allocated in shareable device memory (CUDA virtual memory management, 2 MiB pages).
 * Same-process shards (one Go server process driving several GPUs) are attached by pointer.
 * Shards of other processes travel as two POSIX file descriptors: send table_fd / postings_fd
 * over a Unix-domain socket with SCM_RIGHTS (Go: syscall.UnixRights) together with the struct,
 * and store the received descriptor numbers in the copy handed to kaamer_gpu_attach_shards.
 * Descriptors belong to the caller: close(2) them after sending / after attaching.  (The legacy
 * cudaIpc* mapping is deliberately not offered: random probes into it were measured 130x slower
 * than into a large-page mapping, profiles/r1_peer_gather.log.) ---- */
typedef struct kaamer_shard_handle {
  uint64_t shard_lo, shard_hi; /* dense-code range [lo, hi) held by the exporting handle */
  uint64_t n_postings;
  uint64_t table_ptr, postings_ptr;     /* device pointers in the exporting process */
  uint64_t table_bytes, postings_bytes; /* sizes of the two shareable allocations */
  int32_t device;                       /* CUDA ordinal in the exporting process */
  int32_t pid;                          /* exporting process */
  int32_t table_fd, postings_fd;        /* shareable handles (-1: the index is not shareable) */
} kaamer_shard_handle;
#define KAAMER_MAX_PEER_SHARDS 8
int kaamer_gpu_shard_export(kaamer_gpu_t *h, kaamer_shard_handle *out);
/* shards[n_shards]: the exports of ALL ranks (this handle's own included), in any order; their
 * ranges must tile [0, kaamer_gpu_dense_space()).  The exporting handles must stay open while
 * attached, and every rank must have finished building before anyone attaches (the attach reads
 * every shard once to build the presence filter: one bit per possible k-mer, 227 MB, replicated in
 * local HBM, so that query k-mers absent from the database never cross NVLink).  Re-attaching
 * replaces the previous set.  flags: KAAMER_ATTACH_* */
#define KAAMER_ATTACH_NO_PRESENCE_FILTER 1 /* saturated key spaces: every k-mer exists, skip the filter */
/* Replicate the direct-address table: it is 14.5 GB whatever the size of the database, so every GPU
 * copies ALL shards' table ranges into its own HBM at attach time (multi-posting entries tagged with
 * their owner shard) and only the posting lists — the part that grows with the database — stay
 * sharded.  The first probe of every lookup is then local; NVLink carries posting lists only. */
#define KAAMER_ATTACH_REPLICATE_TABLE 2
/* (with KAAMER_ATTACH_REPLICATE_TABLE) also copy the posting lists of every shard into local HBM: the index
 * was BUILT sharded (every GPU sorted its own key range) but is SEARCHED replicated — what the north star
 * prescribes whenever the whole index fits one GPU (C4: 14.5 GB table + ~59 GB postings of 180 GB).
 * After the attach no search touches NVLink. */
#define KAAMER_ATTACH_REPLICATE_POSTINGS 4
int kaamer_gpu_attach_shards(kaamer_gpu_t *h, const kaamer_shard_handle *shards, int n_shards, int flags);
int kaamer_gpu_detach_shards(kaamer_gpu_t *h);

/* ---- host-side query readers: GetQueriesFasta / GetQueriesFastq (pkg/search/search.go:222-412)
 * with the reference's exact semantics (content sniffing on 32 bytes, 1 MiB line limit, last FASTA
 * record not upper-cased, SizeInKmer rules, FASTQ '@' / sequence-line rules), filling the flat
 * batch layout of the search entry points.  pinned != 0: residues and offsets live in page-locked
 * memory (the search kernels then read the residues in place over PCIe); needs a CUDA device.
 * A file that the reference would silently ignore yields a batch of 0 queries. ---- */
typedef struct kaamer_query_batch {
  uint32_t n_queries;
  uint32_t _pad;
  uint64_t n_residues;
  uint8_t *residues;      /* Query.Sequence, concatenated */
  uint64_t *seq_off;      /* [n_queries+1] */
  char *names;            /* Query.Name, concatenated (not NUL-separated) */
  uint64_t *name_off;     /* [n_queries+1] */
  int32_t *size_in_kmer;  /* Query.SizeInKmer as the reader computes it */
  void *_owner;
} kaamer_query_batch;
/* FormatPositionsToString (pkg/search/search.go:694-742): the `-pos` column of the TSV / JSON output from one
 * PositionHits row (kaamer_hits.pos, one byte per query k-mer position).  Writes a NUL-terminated string into
 * out[cap]; returns its length, or the negative length needed when cap is too small. */
int64_t kaamer_host_format_positions(const uint8_t *positions, uint64_t n, int with_alignment, char *out, uint64_t cap);
int kaamer_host_read_fasta(const char *path, int is_protein, int pinned, kaamer_query_batch **out);
int kaamer_host_read_fastq(const char *path, int pinned, kaamer_query_batch **out);
void kaamer_host_free_queries(kaamer_query_batch *b);

/* ---- result rows off the critical path.  FetchHitsInformation (pkg/search/search.go:454-470)
 * does one protein_store get per hit and the handler goroutines build every TSV row with fmt.Sprintf
 * (search.go:505-606); after a millisecond GPU batch that is the whole request time.  The handle can keep the
 * two Protein fields the rows need — EntryId and Length, indexed by protein id (they travel in the `.kidx`
 * file) — and kaamer_host_format_tsv writes the rows of a whole batch into ONE buffer, byte for byte what the
 * reference sends to its writer: one line per hit, rows in batch order, hits in rank order (with `aln`:
 * re-sorted by BitScore descending, search.go:491-493, stable).
 *   aln      NULL: layout without alignment (search.go:507-553); else one kaamer_aln per hit, in hits order:
 *            alignment layout (search.go:556-604)
 *   names / name_off   Query.Name of the query a row belongs to (row i of a protein batch; row_contig[i] of
 *            a nucleotide batch); the first blank-separated token is printed (strings.Split(Name, " ")[0])
 *   seq_off  protein batches: the batch's query offsets (QStart = 1, QEnd = len(Sequence), search.go:297,290)
 *   with_annotations   print Protein.Length (needs the table); the database's feature columns are not held
 * The text is malloc'ed: kaamer_host_free_text.  Without a table the protein id stands in for EntryId. ---- */
int kaamer_gpu_set_annotations(kaamer_gpu_t *h, const char *entry_ids, const uint64_t *entry_off /* [max_id+2] */,
                               const int32_t *length /* [max_id+1] */, uint32_t max_protein_id);
int kaamer_host_format_tsv(kaamer_gpu_t *h, const kaamer_hits *hits, const kaamer_aln *aln, const char *names,
                           const uint64_t *name_off, const uint64_t *seq_off, int is_protein, int with_positions,
                           int with_annotations, char **out, uint64_t *out_len);
void kaamer_host_free_text(char *p);

/* ---- synthetic C4 workload (bench / tests; include/kaamer_synth_spec.h is the specification, the CPU twin
 * is oracle/synth_oracle.cpp).  Counter-based: record i and query j are pure functions of (seed, index).
 * All pointers are device pointers of the current device; d_seq_off[n+1] is relative to d_res. ---- */
typedef struct kaamer_synth_cfg {
  uint64_t seed;
  uint64_t n_proteins;
} kaamer_synth_cfg;
int kaamer_synth_record_lengths(const kaamer_synth_cfg *cfg, uint64_t first, uint64_t n, uint32_t *d_len,
                                void *stream);
int kaamer_synth_record_residues(const kaamer_synth_cfg *cfg, uint64_t first, uint64_t n, const uint64_t *d_seq_off,
                                 uint8_t *d_res, void *stream);
int kaamer_synth_query_lengths(const kaamer_synth_cfg *cfg, uint32_t batch, uint64_t first, uint64_t n,
                               uint32_t *d_len, void *stream);
int kaamer_synth_query_residues(const kaamer_synth_cfg *cfg, uint32_t batch, uint64_t first, uint64_t n,
                                const uint64_t *d_seq_off, uint8_t *d_res, void *stream);

/* pinned host buffers for callers that want zero-staging H2D (cgo: C.kaamer_gpu_pinned_alloc) */
int kaamer_gpu_pinned_alloc(uint64_t bytes, void **out);
void kaamer_gpu_pinned_free(void *p);

/* per-handle timing: CUDA events on the launching stream around each hot stage.
 * kernel_ms[8], kernel_launches[8]: [0..2] search size classes W, M, G; [3] Smith-Waterman;
 * [4] host->device copies of a host-buffer call; [5] CSR compaction + device->host; [6] reserved;
 * [7] translation / ORF kernels.  all_launches counts every kernel the library launched since
 * the last reset (bench.py roofline / gpu_launches) */
int kaamer_gpu_profile_enable(kaamer_gpu_t *h, int on);
int kaamer_gpu_profile_read(kaamer_gpu_t *h, double *kernel_ms, uint64_t *kernel_launches,
                            uint64_t *all_launches, int reset);
/* host wall-clock of the phases of the host-buffer calls while profiling is enabled, summed since the
 * last reset.  phase_ms[8]: [0] H2D + six-frame translation + ORF extraction, [1] count pass (search
 * kernels, incl. the pool retry), [2] positions / SetBestStartCodon / row assembly / D2H, [3] whole
 * kaamer_gpu_search_nucleotide calls, [4] whole kaamer_gpu_search_proteins calls; inside [2]:
 * [5] position / start-codon kernels and offset scans, [6] pinned result buffers, [7] row assembly + D2H */
int kaamer_gpu_profile_host_read(kaamer_gpu_t *h, double *phase_ms, int reset);

const char *kaamer_gpu_last_error(void);
const char *kaamer_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
