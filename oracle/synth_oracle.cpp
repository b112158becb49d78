// synth_oracle.cpp — CPU twin of the device-side synthetic C4 generator (kaamer_b200/csrc/synth.cu) and
// the "restricted index" the C4 parity check needs.
//
// TEST INFRASTRUCTURE ONLY (see kaamer_oracle.h).  SURVEY.md §7 "Scale of config 4": a 50 M-protein
// database can neither be written as FASTA nor indexed by the CPU oracle; instead the oracle checks a
// SAMPLE of queries by streaming the same counter-based generator (include/kaamer_synth_spec.h) over all
// records and keeping only the k-mers those queries look up.  For a query, Kmatch[s] only depends on the
// posting lists of its own k-mers (pkg/search/search.go:421-436), so the search of the sample against
// this restricted index equals the search against the full index.
//
// The index semantics restated here are the reference's: every window of a record of >= 7 residues is
// inserted (pkg/makedb/inputFASTA.go:226-239), key = EncodeKmer (pkg/kvstore/k_store.go:91-117), the
// posting list of a key is the set of protein ids, sorted descending (pkg/kvstore/kv_store.go:284-305).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/kaamer_synth_spec.h"

namespace {

const uint32_t AA_THR[20] = KAAMER_SYNTH_AA_THR;
const uint16_t LEN_Q[1024] = KAAMER_SYNTH_LEN_Q;
const char LETTERS[21] = KAAMER_SYNTH_LETTERS;

// k_store.go:41: "ACDEFGHIKLMNPQRSTUVWY" -> 0..20
int aa_index(uint8_t c) {
  static const char *A = "ACDEFGHIKLMNPQRSTUVWY";
  const char *p = (const char *)memchr(A, c, 21);
  return p ? (int)(p - A) : -1;
}

void record_letters(uint64_t seed, uint64_t i, const ksyn_meta &m, std::vector<uint8_t> &out) {
  out.resize(m.length);
  const uint32_t nb = (m.length + 3) / 4;
  for (uint32_t b = 0; b < nb; ++b) {
    const uint32_t v = ksyn_record_block(seed, i, &m, b, AA_THR);
    for (int k = 0; k < 4; ++k)
      if (b * 4 + k < m.length) out[b * 4 + k] = (uint8_t)LETTERS[(v >> (8 * k)) & 0xFF];
  }
}

// Fast restatement of ksyn_record_block for the streaming loop: letter_index through a 256-entry table
// on the top byte (letter_index is monotone in u), written straight as k_store.go alphabet codes.
struct FastGen {
  uint8_t first[256];  // letter index of u = b << 24
  uint8_t code[20];    // synth letter index -> index in "ACDEFGHIKLMNPQRSTUVWY"
  FastGen() {
    for (int b = 0; b < 256; ++b) first[b] = (uint8_t)ksyn_letter_index(AA_THR, (uint32_t)b << 24);
    for (int k = 0; k < 20; ++k) code[k] = (uint8_t)aa_index((uint8_t)LETTERS[k]);
  }
  inline uint32_t letter(uint32_t u) const {
    uint32_t k = first[u >> 24];
    while (k < 19 && u >= AA_THR[k]) ++k;
    return k;
  }
  void record_codes(uint64_t seed, uint64_t i, const ksyn_meta &m, std::vector<uint8_t> &sc) const {
    const uint32_t nb = (m.length + 3) / 4;
    sc.resize((size_t)nb * 4);
    for (uint32_t b = 0; b < nb; ++b) {
      const ksyn_u4 f = ksyn_draw(seed, m.founder, b, 0, KSYN_STREAM_RES);
      uint32_t u[4] = {f.x, f.y, f.z, f.w};
      if (m.sub_thr) {
        const ksyn_u4 sd = ksyn_draw(seed, i, b, 0, KSYN_STREAM_SUB);
        const uint32_t sv[4] = {sd.x, sd.y, sd.z, sd.w};
        if (sv[0] < m.sub_thr || sv[1] < m.sub_thr || sv[2] < m.sub_thr || sv[3] < m.sub_thr) {
          const ksyn_u4 r = ksyn_draw(seed, i, b, 0, KSYN_STREAM_SUBRES);
          const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
          for (int k = 0; k < 4; ++k)
            if (sv[k] < m.sub_thr) u[k] = rv[k];
        }
      }
      for (int k = 0; k < 4; ++k) sc[(size_t)b * 4 + k] = code[letter(u[k])];
    }
    sc.resize(m.length);
  }
};

}  // namespace

extern "C" {

// record i -> letters; returns the length (out may be NULL to query it)
uint32_t kso_record(uint64_t seed, uint64_t i, uint8_t *out, uint32_t cap) {
  const ksyn_meta m = ksyn_record_meta(seed, i, LEN_Q);
  if (out && cap >= m.length) {
    std::vector<uint8_t> v;
    record_letters(seed, i, m, v);
    memcpy(out, v.data(), m.length);
  }
  return m.length;
}

void kso_record_meta(uint64_t seed, uint64_t i, uint64_t *founder, uint32_t *length, uint32_t *sub_thr) {
  const ksyn_meta m = ksyn_record_meta(seed, i, LEN_Q);
  *founder = m.founder;
  *length = m.length;
  *sub_thr = m.sub_thr;
}

// query j of batch -> letters; returns the length, *rec = the record it was sampled from
uint32_t kso_query(uint64_t seed, uint64_t n_proteins, uint64_t j, uint32_t batch, uint8_t *out, uint32_t cap,
                   uint64_t *rec_out) {
  const uint64_t rec = ksyn_query_record(seed, n_proteins, j, batch);
  const ksyn_meta m = ksyn_record_meta(seed, rec, LEN_Q);
  if (rec_out) *rec_out = rec;
  if (out && cap >= m.length) {
    const uint32_t nb = (m.length + 3) / 4;
    for (uint32_t b = 0; b < nb; ++b) {
      const uint32_t v = ksyn_query_block(seed, j, batch, rec, &m, b, AA_THR);
      for (int k = 0; k < 4; ++k)
        if (b * 4 + k < m.length) out[b * 4 + k] = (uint8_t)LETTERS[(v >> (8 * k)) & 0xFF];
    }
  }
  return m.length;
}

// Stream records [0, n_proteins) (protein id of record i = id_base + i) and collect, for each of the
// n_keys reference keys (ascending, unique), the set of protein ids holding it, descending.
// offsets_out[n_keys+1]; *postings_out is malloc'ed (kso_free).  stats_out[3] = NumberOfProteins,
// NumberOfAA, NumberOfKmers of the whole database (KStats, inputFASTA.go:142-145).
int kso_restricted_index(uint64_t seed, uint64_t n_proteins, uint32_t id_base, const uint32_t *keys, uint64_t n_keys,
                         int n_threads, uint64_t *offsets_out, uint32_t **postings_out, uint64_t *stats_out) {
  if (n_threads < 1) n_threads = 1;
  // open-addressing set: key -> index
  uint64_t cap = 16;
  while (cap < 4 * n_keys + 16) cap <<= 1;
  std::vector<uint32_t> hk(cap, 0xFFFFFFFFu), hv(cap, 0);
  auto slot_of = [&](uint32_t key) { return (uint64_t)(key * 2654435761u) & (cap - 1); };
  for (uint64_t i = 0; i < n_keys; ++i) {
    uint64_t s = slot_of(keys[i]);
    while (hk[s] != 0xFFFFFFFFu) s = (s + 1) & (cap - 1);
    hk[s] = keys[i];
    hv[s] = (uint32_t)i;
  }
  struct Part {
    std::vector<uint64_t> pairs;  // key index << 32 | protein id
    uint64_t n_prot = 0, n_aa = 0, n_kmers = 0;
  };
  std::vector<Part> parts(n_threads);
  const FastGen fast;
  auto work = [&](int t) {
    Part &P = parts[t];
    const uint64_t lo = n_proteins * (uint64_t)t / n_threads, hi = n_proteins * (uint64_t)(t + 1) / n_threads;
    std::vector<uint32_t> pc;  // pair code of (seq[p], seq[p+1])
    std::vector<uint8_t> sc;
    for (uint64_t i = lo; i < hi; ++i) {
      const ksyn_meta m = ksyn_record_meta(seed, i, LEN_Q);
      if (m.length < 7) continue;
      fast.record_codes(seed, i, m, sc);
      const uint32_t L = m.length;
      P.n_prot++;
      P.n_aa += L;
      P.n_kmers += L - 6;
      pc.resize(L);
      for (uint32_t p = 0; p + 1 < L; ++p) pc[p] = 22u + 21u * sc[p] + sc[p + 1];  // k_store.go:46-59
      for (uint32_t p = 0; p + 7 <= L; ++p) {
        const uint32_t key = (pc[p] << 23) | (pc[p + 2] << 14) | (pc[p + 4] << 5) | sc[p + 6];  // k_store.go:100-110
        uint64_t s = slot_of(key);
        while (hk[s] != 0xFFFFFFFFu) {
          if (hk[s] == key) {
            P.pairs.push_back(((uint64_t)hv[s] << 32) | (uint32_t)(id_base + i));
            break;
          }
          s = (s + 1) & (cap - 1);
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
  for (auto &x : th) x.join();
  std::vector<uint64_t> all;
  uint64_t st[3] = {0, 0, 0};
  for (auto &P : parts) {
    all.insert(all.end(), P.pairs.begin(), P.pairs.end());
    st[0] += P.n_prot;
    st[1] += P.n_aa;
    st[2] += P.n_kmers;
  }
  std::sort(all.begin(), all.end());
  all.erase(std::unique(all.begin(), all.end()), all.end());  // a protein holding a k-mer twice counts once
  uint32_t *post = (uint32_t *)malloc((all.size() + 1) * sizeof(uint32_t));
  if (!post) return -1;
  uint64_t a = 0;
  for (uint64_t k = 0; k < n_keys; ++k) {
    offsets_out[k] = a;
    uint64_t b = a;
    while (b < all.size() && (all[b] >> 32) == k) ++b;
    for (uint64_t i = a; i < b; ++i) post[a + (b - 1 - i)] = (uint32_t)all[i];  // descending (kv_store.go:284-305)
    a = b;
  }
  offsets_out[n_keys] = a;
  *postings_out = post;
  if (stats_out) memcpy(stats_out, st, sizeof st);
  return 0;
}

void kso_free(void *p) { free(p); }

}  // extern "C"
