// format.cu — result rows off the critical path: the per-hit protein_store lookup of
// FetchHitsInformation (pkg/search/search.go:454-470) and the TSV writers of QueryResultHandler
// (search.go:505-606) restated as ONE native call per batch.
//
// After a 1.5 ms GPU batch, 10^5-10^6 badger gets and fmt.Sprintf calls in Go are the whole request time.
// The handle therefore keeps the two annotation fields the TSV rows need — Protein.EntryId and
// Protein.Length (pkg/kvstore/protein.proto), indexed by protein id — next to the index (they travel in the
// `.kidx` file), and kaamer_host_format_tsv writes the rows of a whole kaamer_hits into one buffer, byte for
// byte what the reference's handler goroutines would send to the writer; the Go side then does ONE Write.
// Host code only (this file launches nothing on the device).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "internal.cuh"

namespace kaamer {

// fmt.Sprintf("%.2f") / ("%e") of Go: identical to C's for finite values (correctly rounded, at least two
// exponent digits); NaN and infinities are spelled the Go way
static void go_fixed2(std::string &o, double v) {
  if (std::isnan(v)) {
    o += "NaN";
  } else if (std::isinf(v)) {
    o += v > 0 ? "+Inf" : "-Inf";
  } else {
    char b[64];
    snprintf(b, sizeof b, "%.2f", v);
    o += b;
  }
}
static void go_exp(std::string &o, double v) {
  if (std::isnan(v)) {
    o += "NaN";
  } else if (std::isinf(v)) {
    o += v > 0 ? "+Inf" : "-Inf";
  } else {
    char b[64];
    snprintf(b, sizeof b, "%e", v);
    o += b;
  }
}
static void itoa_append(std::string &o, long long v) {
  char b[32];
  snprintf(b, sizeof b, "%lld", v);
  o += b;
}

// FormatPositionsToString (search.go:694-742)
static void positions_string(std::string &ps, const uint8_t *positions, uint64_t n, bool with_alignment) {
  ps.clear();
  long long current_start = 0, end_pos = 0;
  bool in_sequence = false;
  for (uint64_t pos = 0; pos < n; ++pos) {
    if (positions[pos]) {
      if (!in_sequence) {
        current_start = (long long)pos + 1;
        in_sequence = true;
      }
    } else if (in_sequence) {
      if (!ps.empty()) ps += ",";
      if ((long long)pos + 1 > current_start) {
        end_pos = (long long)pos + 1;
        if (with_alignment) end_pos += KAAMER_KMER_SIZE - 1;
        itoa_append(ps, current_start);
        ps += "-";
        itoa_append(ps, end_pos);
      } else {
        itoa_append(ps, current_start);
      }
      in_sequence = false;
    }
  }
  if (in_sequence) {  // the run reaches the end of the query (search.go:726-740)
    if (!ps.empty()) ps += ",";
    end_pos = (long long)n;  // len(positions), search.go:733
    if (with_alignment) end_pos += KAAMER_KMER_SIZE - 1;
    itoa_append(ps, current_start);
    ps += "-";
    itoa_append(ps, end_pos);
  }
}

}  // namespace kaamer

using namespace kaamer;

extern "C" {

int kaamer_gpu_set_annotations(kaamer_gpu_t *h, const char *entry_ids, const uint64_t *entry_off, const int32_t *length,
                               uint32_t max_protein_id) {
  if (!h || !entry_off || !length || (entry_off[(size_t)max_protein_id + 1] && !entry_ids)) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  return guarded([&]() -> int {
    std::lock_guard<std::mutex> lk(h->mu);
    for (size_t i = 0; i <= (size_t)max_protein_id; ++i)
      if (entry_off[i + 1] < entry_off[i]) {
        set_error("annotations: entry offsets are not monotonic");
        return KAAMER_ERR_ARG;
      }
    DevIndex &ix = h->idx;
    ix.annot_off.assign(entry_off, entry_off + (size_t)max_protein_id + 2);
    ix.annot_ids.assign(entry_ids, entry_ids + entry_off[(size_t)max_protein_id + 1]);
    ix.annot_len.assign(length, length + (size_t)max_protein_id + 1);
    return KAAMER_OK;
  });
}

// One TSV line per hit, rows in batch order, hits of a row in rank order (with `aln`: re-sorted by BitScore
// descending, search.go:491-493; stable, so ties keep the Kmatch order).  aln: NULL (no alignment,
// search.go:507-553) or one kaamer_aln per hit in hits order (alignment layout, search.go:556-604).
// names / name_off: Query.Name of the QUERY a row belongs to (row i for protein batches; hits->row_contig[i]
// for nucleotide batches).  seq_off: the batch's query offsets (protein batches: QEnd = len(Sequence)); for
// nucleotide batches the Location comes from the rows.  *out is malloc'ed (kaamer_host_free_text).
int kaamer_host_format_tsv(kaamer_gpu_t *h, const kaamer_hits *hits, const kaamer_aln *aln, const char *names,
                           const uint64_t *name_off, const uint64_t *seq_off, int is_protein, int with_positions,
                           int with_annotations, char **out, uint64_t *out_len) {
  if (!hits || !names || !name_off || !out || !out_len || (is_protein && !seq_off)) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  *out_len = 0;
  return guarded([&]() -> int {
    const DevIndex *ix = h ? &h->idx : nullptr;
    const bool have_annot = ix && !ix->annot_off.empty();
    if (with_annotations && !have_annot) {
      set_error("annotations wanted but the handle holds no EntryId / Length table (kaamer_gpu_set_annotations)");
      return KAAMER_ERR_ARG;
    }
    if (with_positions && hits->n_hits && (!hits->pos_off || !hits->pos)) {
      set_error("positions wanted but the hits carry none (search with want_positions)");
      return KAAMER_ERR_ARG;
    }
    if (!is_protein && hits->n_rows && (!hits->row_contig || !hits->row_start || !hits->row_end)) {
      set_error("nucleotide rows expected (kaamer_gpu_search_nucleotide result)");
      return KAAMER_ERR_ARG;
    }
    std::string o, ps;
    o.reserve((size_t)hits->n_hits * 96 + 64);
    std::vector<uint64_t> order;
    for (uint32_t i = 0; i < hits->n_rows; ++i) {
      const uint64_t b = hits->hit_off[i], e = hits->hit_off[i + 1];
      if (b == e) continue;  // queries without hits are not reported (search_protein.go:109)
      const uint32_t qn = is_protein ? i : hits->row_contig[i];
      const char *nm = names + name_off[qn];
      size_t nlen = (size_t)(name_off[qn + 1] - name_off[qn]);
      const void *sp = memchr(nm, ' ', nlen);
      if (sp) nlen = (size_t)((const char *)sp - nm);  // strings.Split(Name, " ")[0]
      const long long q_start = is_protein ? 1 : (long long)hits->row_start[i];
      const long long q_end = is_protein ? (long long)(seq_off[i + 1] - seq_off[i]) : (long long)hits->row_end[i];
      order.resize((size_t)(e - b));
      for (uint64_t k = b; k < e; ++k) order[(size_t)(k - b)] = k;
      if (aln)
        std::stable_sort(order.begin(), order.end(),
                         [&](uint64_t x, uint64_t y) { return aln[x].bitscore > aln[y].bitscore; });
      for (uint64_t k : order) {
        const uint32_t sid = hits->subject_id[k];
        o.append(nm, nlen);
        o += '\t';
        if (have_annot && sid + 1ull < ix->annot_off.size())
          o.append(ix->annot_ids.data() + ix->annot_off[sid], (size_t)(ix->annot_off[sid + 1] - ix->annot_off[sid]));
        else if (!have_annot)
          itoa_append(o, sid);  // no table: the protein id stands in for EntryId
        o += '\t';
        if (!aln) {
          // float32 arithmetic as in the reference (search.go:512)
          const float ident = ((float)hits->kmatch[k] / (float)hits->size_in_kmer[i]) * 100.0f;
          go_fixed2(o, (double)ident);
          o += '\t';
          itoa_append(o, hits->size_in_kmer[i]);
          o += '\t';
          itoa_append(o, hits->kmatch[k]);
          o += '\t';
          if (with_positions) {
            positions_string(ps, hits->pos + hits->pos_off[k], hits->pos_off[k + 1] - hits->pos_off[k], false);
            itoa_append(o, (long long)std::count(ps.begin(), ps.end(), ','));
          } else {
            o += "N/A";
          }
          o += '\t';
          itoa_append(o, q_start);
          o += '\t';
          itoa_append(o, q_end);
          o += "\t1\t";  // subject always starts at 1 in kmer mode
          if (with_annotations) itoa_append(o, sid < ix->annot_len.size() ? ix->annot_len[sid] : 0);
          else o += "N/A";
          if (with_positions) {
            o += '\t';
            o += ps;
          }
        } else {
          const kaamer_aln &a = aln[k];
          go_fixed2(o, (double)a.identity);
          o += '\t';
          itoa_append(o, a.length);
          o += '\t';
          itoa_append(o, a.mismatches);
          o += '\t';
          itoa_append(o, a.gap_openings);
          o += '\t';
          itoa_append(o, is_protein ? (long long)a.query_start : q_start);  // search.go:572-582
          o += '\t';
          itoa_append(o, is_protein ? (long long)a.query_end : q_end);
          o += '\t';
          itoa_append(o, a.subject_start);
          o += '\t';
          itoa_append(o, a.subject_end);
          o += '\t';
          go_exp(o, a.evalue);
          o += '\t';
          go_fixed2(o, a.bitscore);
          if (with_positions) {
            positions_string(ps, hits->pos + hits->pos_off[k], hits->pos_off[k + 1] - hits->pos_off[k], true);
            o += '\t';
            o += ps;
          }
        }
        // (dbStats.Features columns, search.go:547-551 / 596-601: the feature strings stay in protein_store;
        //  a database without feature columns — every FASTA-built one — has none to print)
        o += '\n';
      }
    }
    char *buf = (char *)malloc(o.size() + 1);
    if (!buf) {
      set_error("out of host memory");
      return KAAMER_ERR_NOMEM;
    }
    memcpy(buf, o.data(), o.size());
    buf[o.size()] = 0;
    *out = buf;
    *out_len = o.size();
    return KAAMER_OK;
  });
}

void kaamer_host_free_text(char *p) { free(p); }

}  // extern "C"
