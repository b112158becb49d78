// kaamer_search.cpp — the search path driven from compiled host code through the C ABI alone
// (include/kaamer_gpu.h; no Python, no torch): what the cgo shim of go/pkg/gpusearch does, in C++.
//
//   kaamer_search <index.kidx> <queries.fasta> [MaxResults MinKMatch MinKRatio] [--pos]
//
// Reads the queries with the reference's reader semantics (kaamer_host_read_fasta, page-locked batch),
// searches them in one call and prints one line per hit in the reference's TSV layout for
// `-fmt tsv` without alignment, positions or annotations (pkg/search/search.go:507-553):
//   QueryId  EntryId  %KMatchIdentity  SizeInKmer  KMatch  N/A  QStart  QEnd  1  N/A
// and with --pos (ExtractPositions): the number of position ranges instead of the first N/A and the
// FormatPositionsToString column at the end.
// EntryId is the protein id here: entry names and annotations live in the reference's protein_store
// (FetchHitsInformation, search.go:454-470), which stays with the Go host.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kaamer_gpu.h"

static int fail(const char *what, int rc) {
  fprintf(stderr, "%s failed (%d): %s\n", what, rc, kaamer_gpu_last_error());
  return 1;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <index.kidx> <queries.fasta> [MaxResults MinKMatch MinKRatio]\n", argv[0]);
    return 2;
  }
  bool want_pos = false;
  if (argc > 3 && strcmp(argv[argc - 1], "--pos") == 0) {
    want_pos = true;
    --argc;
  }
  kaamer_opts opts;
  memset(&opts, 0, sizeof opts);
  opts.want_positions = want_pos ? 1 : 0;
  opts.max_results = argc > 3 ? atoi(argv[3]) : 10;  // defaults: api/server.go:200-203
  opts.min_kmatch = argc > 4 ? atoll(argv[4]) : 10;
  opts.min_kratio = argc > 5 ? atof(argv[5]) : 0.05;
  kaamer_gpu_t *h = nullptr;
  int rc = kaamer_gpu_open(argv[1], 0, &h);
  if (rc != KAAMER_OK) return fail("kaamer_gpu_open", rc);
  kaamer_query_batch *q = nullptr;
  rc = kaamer_host_read_fasta(argv[2], 1, 1, &q);
  if (rc != KAAMER_OK) return fail("kaamer_host_read_fasta", rc);
  kaamer_hits *hits = nullptr;
  rc = kaamer_gpu_search_proteins(h, q->residues, q->seq_off, q->n_queries, &opts, &hits);
  if (rc != KAAMER_OK) return fail("kaamer_gpu_search_proteins", rc);
  for (uint32_t i = 0; i < hits->n_rows; ++i) {
    if (hits->hit_off[i] == hits->hit_off[i + 1]) continue;  // queries without hits are not reported (search_protein.go:109)
    std::string name(q->names + q->name_off[i], q->names + q->name_off[i + 1]);
    const std::string query_id = name.substr(0, name.find(' '));  // strings.Split(Name, " ")[0]
    const long long q_end = (long long)(q->seq_off[i + 1] - q->seq_off[i]);
    for (uint64_t k = hits->hit_off[i]; k < hits->hit_off[i + 1]; ++k) {
      const float ident = (float)hits->kmatch[k] / (float)hits->size_in_kmer[i] * 100.0f;  // float32 as in the reference
      if (!want_pos) {
        printf("%s\t%u\t%.2f\t%d\t%u\tN/A\t1\t%lld\t1\tN/A\n", query_id.c_str(), hits->subject_id[k], (double)ident,
               hits->size_in_kmer[i], hits->kmatch[k], q_end);
      } else {
        const uint64_t np = hits->pos_off[k + 1] - hits->pos_off[k];
        std::vector<char> buf(24 * (np / 2 + 2));
        if (kaamer_host_format_positions(hits->pos + hits->pos_off[k], np, 0, buf.data(), buf.size()) < 0)
          return fail("kaamer_host_format_positions", -1);
        int commas = 0;
        for (const char *c = buf.data(); *c; ++c) commas += *c == ',';
        printf("%s\t%u\t%.2f\t%d\t%u\t%d\t1\t%lld\t1\tN/A\t%s\n", query_id.c_str(), hits->subject_id[k], (double)ident,
               hits->size_in_kmer[i], hits->kmatch[k], commas, q_end, buf.data());
      }
    }
  }
  kaamer_gpu_free_hits(hits);
  kaamer_host_free_queries(q);
  kaamer_gpu_close(h);
  return 0;
}
