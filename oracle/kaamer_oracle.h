/*
 * kaamer_oracle.h — CPU ORACLE for the kaamer search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement of the reference
 * (zorino/kaamer, Go) algorithm for the search path.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it,
 * and only as the checker / reported baseline — never as the product path.
 *
 * Pinning status:
 *   - encoding, SizeInKmer, ORF finder, filter, bitscore/e-value, FASTA ids:
 *     pinned against the known-answer vectors derived by hand from the Go source
 *      — the reference ships no tests and cannot be built here (no Go).
 *   - alignment (biogo v1.0.1 SWAffine, not vendored in the reference):
 *     **parity unpinned** — documented deterministic Gotoh restatement.
 *
 * All `file:line` citations are relative to the reference repository root.
 */
#ifndef KAAMER_ORACLE_H
#define KAAMER_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ko_index ko_index;
typedef struct ko_result ko_result;
typedef struct ko_orfs ko_orfs;

typedef struct {
  int64_t min_kmatch;    /* SearchOptions.MinKMatch  pkg/search/search.go:69 */
  double min_kratio;     /* SearchOptions.MinKRatio  pkg/search/search.go:70 */
  int32_t max_results;   /* SearchOptions.MaxResults pkg/search/search.go:62 */
  int32_t want_positions;/* SearchOptions.ExtractPositions search.go:64 */
} ko_opts;

typedef struct {
  float identity, similarity;           /* float32, align.go:72-101 */
  int32_t length, mismatches, gap_openings, raw;
  double bitscore, evalue;              /* float64, align.go:136,141 */
  int32_t query_start, query_end, subject_start, subject_end;
  int32_t dp_score;                     /* raw DP optimum (sum of segment scores) */
  int32_t n_segments;
  int32_t illegal;                      /* 1 if biogo would have returned an error */
} ko_aln;

typedef struct {
  double lambda, K;         /* matrixScores.go:22-105 */
  int32_t gap_open_opt;     /* MatrixScores.GapOpen (option; used by the -11 test, align.go:127) */
  int32_t gap_extend_opt;   /* MatrixScores.GapExtend (align.go:130) */
  uint64_t number_of_aa;    /* KStats.NumberOfAA (align.go:141) */
} ko_aln_params;

/* a2: pkg/kvstore/k_store.go:39-117 */
uint32_t ko_encode_kmer(const uint8_t *kmer7);
void ko_decode_kmer(uint32_t key, uint8_t *out7);

/* a1: SizeInKmer, pkg/search/search.go:290-293 */
int32_t ko_size_in_kmer(const uint8_t *seq, uint64_t len);

/* a9: FASTA protein-id assignment quirk, pkg/makedb/inputFASTA.go:96-124.
 * ids_out[j] = id of the j-th '>' record (0-based j), n = number of records. */
void ko_fasta_ids(uint64_t n, uint32_t *ids_out);

/* a9: index semantics (makedb + indexdb). records: residues[off[i]..off[i+1]) with id ids[i].
 * Records shorter than 7 contribute nothing (inputFASTA.go:226-228). */
ko_index *ko_index_build(const uint8_t *residues, const uint64_t *off, const uint32_t *ids,
                         uint64_t n_records, int n_threads);
ko_index *ko_index_from_arrays(const uint32_t *keys, const uint64_t *offsets,
                               const uint32_t *postings, uint64_t n_keys, uint64_t n_proteins,
                               uint64_t n_aa, uint64_t n_kmers);
void ko_index_free(ko_index *);
uint64_t ko_index_n_keys(const ko_index *);
uint64_t ko_index_n_postings(const ko_index *);
const uint32_t *ko_index_keys(const ko_index *);
const uint64_t *ko_index_offsets(const ko_index *);
const uint32_t *ko_index_postings(const ko_index *);
void ko_index_stats(const ko_index *, uint64_t *n_proteins, uint64_t *n_aa, uint64_t *n_kmers);

/* a3-a6: protein search (search_protein.go:70-114, search.go:132-152,189-220,414-452) */
ko_result *ko_search_proteins(const ko_index *, const uint8_t *residues, const uint64_t *off,
                              uint32_t nq, const ko_opts *, int n_threads);

/* a7: GetORFs (dna.go:65-181) for one contig. */
ko_orfs *ko_get_orfs(const uint8_t *dna, uint64_t len);
void ko_orfs_free(ko_orfs *);
uint64_t ko_orfs_n(const ko_orfs *);
const uint8_t *ko_orfs_seq(const ko_orfs *);       /* concatenated aa */
const uint64_t *ko_orfs_seq_off(const ko_orfs *);  /* n+1 */
const int64_t *ko_orfs_start(const ko_orfs *);
const int64_t *ko_orfs_end(const ko_orfs *);
const uint8_t *ko_orfs_plus(const ko_orfs *);
const int32_t *ko_orfs_alts(const ko_orfs *);      /* concatenated StartsAlternative */
const uint64_t *ko_orfs_alts_off(const ko_orfs *); /* n+1 */

/* a7+a8: nucleotide search over contigs (search_nucleotide.go:61-140). The result rows are
 * the ORFs that survive (>=1 hit after FilterResults), in (contig, GetORFs order).
 * `reads` = 1 restates search_fastq.go (same per-ORF logic). */
ko_result *ko_search_nucleotide(const ko_index *, const uint8_t *nt, const uint64_t *off,
                                uint32_t n_contigs, const ko_opts *, int n_threads);

/* result accessors. For protein search rows == queries (nq rows, rows without hits have
 * hit_off[i]==hit_off[i+1]); for nucleotide search rows == surviving ORFs. */
void ko_result_free(ko_result *);
uint64_t ko_result_n_rows(const ko_result *);
const uint64_t *ko_result_hit_off(const ko_result *);   /* n_rows+1 */
const uint32_t *ko_result_subject(const ko_result *);
const int64_t *ko_result_kmatch(const ko_result *);
const int32_t *ko_result_size_in_kmer(const ko_result *);
/* positions (only if want_positions or nucleotide): one byte per query position per hit */
const uint64_t *ko_result_pos_off(const ko_result *);   /* n_hits+1 */
const uint8_t *ko_result_pos(const ko_result *);
/* nucleotide rows only */
const uint32_t *ko_result_row_contig(const ko_result *);
const int64_t *ko_result_row_start(const ko_result *);
const int64_t *ko_result_row_end(const ko_result *);
const uint8_t *ko_result_row_plus(const ko_result *);
const uint8_t *ko_result_row_seq(const ko_result *);
const uint64_t *ko_result_row_seq_off(const ko_result *);
/* work counters (for the bench's algorithmic-bytes figure) */
uint64_t ko_result_n_lookups(const ko_result *);
uint64_t ko_result_n_increments(const ko_result *);

/* a6 in isolation: FilterResults on an already sorted Kmatch list; returns kept count. */
int32_t ko_filter_count(const int64_t *kmatch_sorted, int32_t n, int32_t size_in_kmer,
                        const ko_opts *);

/* a12: align.Align (pkg/align/align.go:46-161) on top of a restated biogo SWAffine. */
int ko_align(const uint8_t *q, int32_t qlen, const uint8_t *s, int32_t slen,
             const ko_aln_params *, ko_aln *out, char *aln_a, char *aln_b, int32_t aln_cap);
int32_t ko_blosum62(int32_t i, int32_t j); /* 26x26, order "-ABCDEFGHIJKLMNPQRSTVWXYZ*" */
double ko_bitscore(double lambda, double K, int32_t raw);
double ko_evalue(int32_t qlen, uint64_t number_of_aa, double bitscore);

/* a11: FormatPositionsToString (search.go:694-742); returns strlen, writes NUL-terminated. */
int32_t ko_format_positions(const uint8_t *pos, int32_t n, int32_t with_alignment, char *out,
                            int32_t cap);

#ifdef __cplusplus
}
#endif
#endif
