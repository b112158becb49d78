/*
 * kaamer_oracle.cpp — CPU ORACLE (test infrastructure only; see kaamer_oracle.h).
 *
 * A literal CPU restatement of the zorino/kaamer search hot path, function by function,
 * each citing the Go source it follows.  Deliberately simple (maps, strings, std::sort):
 * it is the checker and the reported CPU baseline, never the product path.
 *
 * Alignment (biogo v1.0.1, github.com/biogo/biogo, go.mod:8) is NOT in the reference tree:
 * **parity unpinned** for ko_align — see the comment block above sw_affine().
 */
#include "kaamer_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

constexpr int KMER_SIZE = 7;  // pkg/search/search.go:45, pkg/makedb/makedb.go:30

// ---------------------------------------------------------------------------------------
// a2. NewAATable / EncodeKmer  (pkg/kvstore/k_store.go:39-117)
// ---------------------------------------------------------------------------------------
struct AATable {
  int8_t idx[256];
  AATable() {
    memset(idx, -1, sizeof idx);
    const char *aa = "ACDEFGHIKLMNPQRSTUVWY";  // k_store.go:41 — order matters
    for (int j = 0; aa[j]; ++j) idx[(uint8_t)aa[j]] = (int8_t)j;
  }
};
const AATable g_aa;

// aaTable[[2]rune{a,b}]: 22 + 21*idx(a) + idx(b); Go map miss -> 0  (k_store.go:46-59,100-103)
inline uint32_t pair_code(uint8_t a, uint8_t b) {
  int ia = g_aa.idx[a], ib = g_aa.idx[b];
  if (ia < 0 || ib < 0) return 0;
  return 22u + 21u * (uint32_t)ia + (uint32_t)ib;
}
// aaTable[[2]rune{a,'.'}]: idx(a); miss -> 0  (k_store.go:49-51,108-110)
inline uint32_t single_code(uint8_t a) {
  int ia = g_aa.idx[a];
  return ia < 0 ? 0u : (uint32_t)ia;
}
inline uint32_t encode_kmer(const uint8_t *k) {
  // pairs at shifts 32-9, 32-18, 32-27 (k_store.go:100-105), last aa in the low 5 bits (:108-110)
  return (pair_code(k[0], k[1]) << 23) | (pair_code(k[2], k[3]) << 14) |
         (pair_code(k[4], k[5]) << 5) | single_code(k[6]);
}

inline int32_t size_in_kmer(const uint8_t *seq, uint64_t len) {
  // search.go:290-293 / search_nucleotide.go:81,88-90
  int64_t s = (int64_t)len - KMER_SIZE + 1;
  if (len > 0 && seq[len - 1] == '*') s--;
  return (int32_t)s;
}

}  // namespace

struct ko_index {
  std::vector<uint32_t> keys;      // ascending (badger big-endian byte order == numeric order)
  std::vector<uint64_t> offsets;   // n_keys+1
  std::vector<uint32_t> postings;  // per key: unique protein ids, DESCENDING (kv_store.go:284-305)
  uint64_t n_proteins = 0, n_aa = 0, n_kmers = 0;

  // KmerStore.Get + KCombStore.Get (search.go:421-429): exact-match lookup
  inline bool find(uint32_t key, uint64_t &b, uint64_t &e) const {
    auto it = std::lower_bound(keys.begin(), keys.end(), key);
    if (it == keys.end() || *it != key) return false;
    size_t i = it - keys.begin();
    b = offsets[i];
    e = offsets[i + 1];
    return true;
  }
};

struct ko_result {
  std::vector<uint64_t> hit_off{0};
  std::vector<uint32_t> subject;
  std::vector<int64_t> kmatch;
  std::vector<int32_t> size_in_kmer;
  std::vector<uint64_t> pos_off{0};
  std::vector<uint8_t> pos;
  std::vector<uint32_t> row_contig;
  std::vector<int64_t> row_start, row_end;
  std::vector<uint8_t> row_plus;
  std::vector<uint8_t> row_seq;
  std::vector<uint64_t> row_seq_off{0};
  uint64_t n_lookups = 0, n_increments = 0;
};

namespace {

struct Hit {
  uint32_t key;
  int64_t kmatch;
};

struct SearchRes {
  std::unordered_map<uint32_t, int64_t> counter;             // cnt.CounterBox (search.go:74,432)
  std::vector<Hit> hits;                                     // HitList
  std::unordered_map<uint32_t, std::vector<uint8_t>> posHits;  // PositionHits map[uint32][]bool
  uint64_t n_lookups = 0, n_increments = 0;
};

// KmerSearch + StoreMatchPositions (search.go:414-452) driven by the key-producer loop
// (search_protein.go:94-98 / search_nucleotide.go:105-108).
void kmer_search(const ko_index &idx, const uint8_t *seq, int32_t sizeInKmer, bool extractPos,
                 SearchRes &sr) {
  for (int32_t k = 0; k < sizeInKmer; ++k) {
    uint32_t key = encode_kmer(seq + k);
    sr.n_lookups++;
    uint64_t b, e;
    if (!idx.find(key, b, e)) continue;  // search.go:421 (miss => skip)
    for (uint64_t p = b; p < e; ++p) {
      uint32_t id = idx.postings[p];
      sr.counter[id]++;  // search.go:432
      sr.n_increments++;
      if (extractPos) {  // search.go:433-435, 446-449
        auto &v = sr.posHits[id];
        if (v.empty()) v.assign((size_t)sizeInKmer, 0);
        v[(size_t)k] = 1;
      }
    }
  }
}

// sortMapByValue (search.go:132-152): descending Kmatch.  The reference's tie order is
// nondeterministic (sync.Map range + unstable sort); canonical tie-break = subject id ascending.
void sort_map_by_value(SearchRes &sr) {
  sr.hits.clear();
  sr.hits.reserve(sr.counter.size());
  for (auto &kv : sr.counter) sr.hits.push_back({kv.first, kv.second});
  std::sort(sr.hits.begin(), sr.hits.end(), [](const Hit &a, const Hit &b) {
    if (a.kmatch != b.kmatch) return a.kmatch > b.kmatch;
    return a.key < b.key;
  });
}

// FilterResults (search.go:189-220), literal.
void filter_results(std::vector<Hit> &hits, std::unordered_map<uint32_t, std::vector<uint8_t>> *posHits,
                    int32_t sizeInKmer, const ko_opts &o) {
  std::vector<uint32_t> hitsToDelete;
  int64_t n = (int64_t)hits.size();
  int64_t lastGood = n - 1;
  for (int64_t i = 0; i < n; ++i) {
    const Hit &h = hits[(size_t)i];
    if (((double)h.kmatch / (double)sizeInKmer) < o.min_kratio || h.kmatch < o.min_kmatch) {
      if (lastGood == n - 1) lastGood = i - 1;
      hitsToDelete.push_back(h.key);
    }
  }
  if (lastGood >= (int64_t)o.max_results) {
    lastGood = (int64_t)o.max_results - 1;
    for (int64_t i = lastGood + 1; i < n; ++i)
      if (i >= 0) hitsToDelete.push_back(hits[(size_t)i].key);
  }
  if (lastGood < 0)
    hits.clear();
  else
    hits.resize((size_t)(lastGood + 1));
  if (posHits)
    for (uint32_t k : hitsToDelete) posHits->erase(k);
}

// ---------------------------------------------------------------------------------------
// a7. GetORFs / GetFrame / ReverseComplement (pkg/search/dna.go:55-196), table 11 only
// (gcodeBacteria, gcode.go:36-101; the geneticCode argument is ignored, dna.go:106).
// ---------------------------------------------------------------------------------------
struct AminoAcid {
  char aa = 0;  // 0 == "" (codon not in the map)
  bool start = false, stop = false;
};

struct GCode {
  AminoAcid t[64];
  static int b(char c) { return c == 't' ? 0 : c == 'c' ? 1 : c == 'a' ? 2 : c == 'g' ? 3 : -1; }
  GCode() {
    // standard order TCAG; table 11 amino acids
    const char *aas = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    for (int i = 0; i < 64; ++i) {
      t[i].aa = aas[i];
      t[i].stop = aas[i] == '*';
    }
    // starts: ttg ctg att atc ata atg gtg (gcode.go:40,56,69-72,88)
    const char *starts[] = {"ttg", "ctg", "att", "atc", "ata", "atg", "gtg"};
    for (auto s : starts) t[b(s[0]) * 16 + b(s[1]) * 4 + b(s[2])].start = true;
  }
  AminoAcid get(const char *c) const {
    int x = b(c[0]), y = b(c[1]), z = b(c[2]);
    if (x < 0 || y < 0 || z < 0) return AminoAcid{};  // map miss: zero value
    return t[x * 16 + y * 4 + z];
  }
};
GCode g_gcode;  // ko_set_genetic_code replaces the table (§8f-4); default = table 11

struct Location {
  int64_t start = 0, end = 0;
  bool plus = true;
  std::vector<int32_t> alts;
};
struct ORF {
  std::string seq;
  Location loc;
  int frame = 0;
  int64_t order = 0;
};

std::string reverse_complement(const std::string &dna_lower) {  // dna.go:55-63
  std::string r(dna_lower.rbegin(), dna_lower.rend());
  for (auto &c : r) {
    switch (c) {
      case 'a': c = 't'; break;
      case 't': c = 'a'; break;
      case 'g': c = 'c'; break;
      case 'c': c = 'g'; break;
      default: break;
    }
  }
  return r;
}

std::string get_frame(int frameNumber, const std::string &dna_in) {  // dna.go:183-196
  std::string dna = dna_in;
  if (frameNumber < 0) {
    dna = reverse_complement(dna);
    frameNumber = -frameNumber;
  }
  int64_t startPos = frameNumber - 1;
  int64_t lenFrame = (int64_t)dna.size() - startPos;
  int64_t endPos = (int64_t)dna.size() - (lenFrame % 3);
  if (startPos > (int64_t)dna.size() || endPos > (int64_t)dna.size() || endPos < startPos)
    return std::string();  // the Go code would panic (len < 2); treated as empty
  return dna.substr((size_t)startPos, (size_t)(endPos - startPos));
}

std::vector<ORF> get_orfs(const uint8_t *dna_raw, uint64_t len) {  // dna.go:65-181
  const int minLenCDS = 21;                                        // dna.go:26
  const int frameStartPosition[6] = {0, 1, 2, 0, 1, 2};            // dna.go:25
  std::vector<ORF> orfs;
  std::string dna((const char *)dna_raw, (size_t)len);
  for (auto &c : dna) c = (char)tolower((unsigned char)c);
  if (dna.size() < 2) return orfs;
  std::string frames[6] = {get_frame(1, dna),  get_frame(2, dna),  get_frame(3, dna),
                           get_frame(-1, dna), get_frame(-2, dna), get_frame(-3, dna)};
  const int64_t L = (int64_t)dna.size();
  int64_t order = 0;
  for (int framePos = 0; framePos < 6; ++framePos) {
    const std::string &frameSeq = frames[framePos];
    int64_t startPos = frameStartPosition[framePos];
    bool plusStrand = framePos <= 2;
    int64_t absPos = framePos;
    if (!plusStrand) absPos = L - startPos - 1;
    int64_t currentPos = 0;
    ORF orf;
    orf.loc.start = absPos + 1;
    orf.loc.end = 0;
    orf.loc.plus = plusStrand;
    bool insideORF = true;  // dna.go:98
    std::string cds;
    int32_t currentAAPos = 0;
    int64_t flen = (int64_t)frameSeq.size();
    for (int64_t i = 0; i < flen - (flen % 3); i += 3) {
      currentPos = i;
      AminoAcid cur = g_gcode.get(frameSeq.data() + i);
      if (cur.start) {
        if (!insideORF) {
          insideORF = true;
          currentAAPos = 0;
          orf.loc.start = framePos + i + 1;
          if (!plusStrand) orf.loc.start = L - (framePos + i) + 3;
          orf.loc.alts.push_back(currentAAPos);
        } else {
          orf.loc.alts.push_back(currentAAPos);
        }
      }
      if (insideORF && cur.aa) cds.push_back(cur.aa);
      if (cur.stop) {
        if (insideORF && (int)cds.size() >= minLenCDS) {
          int64_t endPos = i + 3 + framePos;
          if (!plusStrand) endPos = orf.loc.start - ((int64_t)cds.size() * 3) + 1;
          orf.loc.end = endPos;
          orf.seq = cds;
          orf.frame = framePos;
          orf.order = order++;
          orfs.push_back(orf);
        }
        orf = ORF();
        orf.loc.start = 0;
        orf.loc.end = 0;
        orf.loc.plus = plusStrand;
        cds.clear();
        insideORF = false;
      }
      currentAAPos += 1;
    }
    if (insideORF && (int)cds.size() >= minLenCDS) {
      int64_t endPos = currentPos + 3 + framePos;
      if (!plusStrand) endPos = orf.loc.start - ((int64_t)cds.size() * 3) + 1;
      orf.loc.end = endPos;
      orf.seq = cds;
      orf.frame = framePos;
      orf.order = order++;
      orfs.push_back(orf);
    }
  }
  // dna.go:167-177 — sort.Slice (unstable) by End (+) / Start (-).  Canonical: ties keep
  // emission order (frame, position).
  std::stable_sort(orfs.begin(), orfs.end(), [](const ORF &a, const ORF &b) {
    int64_t ia = a.loc.plus ? a.loc.end : a.loc.start;
    int64_t ib = b.loc.plus ? b.loc.end : b.loc.start;
    return ia < ib;
  });
  return orfs;
}

// SetBestStartCodon (dna.go:198-272), literal; Hits in canonical order.
void set_best_start_codon(std::string &seq, Location &loc, int32_t &sizeInKmer, SearchRes &sr) {
  std::vector<Hit> bestHits;
  int64_t bestHitScore = 0;
  for (auto &h : sr.hits) {
    if (h.kmatch >= bestHitScore) {
      bestHitScore = h.kmatch;
      bestHits.push_back(h);
    }
  }
  if (loc.alts.size() < 1) return;
  int32_t bestStart = loc.alts[0];
  int32_t firstStart = loc.alts[0];
  int64_t firstBestHitPos = 999999999;
  bool exit = false;
  for (auto &bh : bestHits) {
    auto it = sr.posHits.find(bh.key);
    if (it == sr.posHits.end()) continue;
    const auto &v = it->second;
    for (size_t i = 0; i < v.size(); ++i) {
      if (v[i]) {
        if ((int64_t)i < firstBestHitPos) firstBestHitPos = (int64_t)i;
        exit = true;
      }
      if (exit) break;  // `exit` is never reset: later tied hits are inspected at i==0 only
    }
  }
  exit = false;
  for (int32_t s : loc.alts) {
    if (s <= firstBestHitPos)
      bestStart = s;
    else
      exit = true;
    if (exit) break;
  }
  if (bestStart != firstStart) {
    if (loc.plus)
      loc.start = loc.start + 3 * (int64_t)bestStart;
    else
      loc.start = loc.start - 3 * (int64_t)bestStart;
    seq = seq.substr((size_t)bestStart);
    for (auto &kv : sr.posHits) {
      auto &v = kv.second;
      if ((size_t)bestStart <= v.size())
        v.erase(v.begin(), v.begin() + bestStart);
      else
        v.clear();
    }
    sizeInKmer = (int32_t)seq.size() - KMER_SIZE + 1;
    if (!seq.empty() && seq.back() == '*') sizeInKmer -= 1;
  }
  loc.alts.clear();
}

template <class F>
void parallel_for(uint64_t n, int n_threads, F f) {
  if (n_threads < 1) n_threads = 1;
  if ((uint64_t)n_threads > n) n_threads = (int)std::max<uint64_t>(1, n);
  if (n_threads == 1) {
    f(0, 0, n);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) {
    uint64_t b = n * (uint64_t)t / (uint64_t)n_threads, e = n * (uint64_t)(t + 1) / (uint64_t)n_threads;
    th.emplace_back([=] { f(t, b, e); });
  }
  for (auto &t : th) t.join();
}

// chunked std::sort + rounds of pairwise std::merge (the badger LSM keeps keys sorted; any
// sort gives the same index)
void parallel_sort(std::vector<uint64_t> &v, int n_threads) {
  size_t n = v.size();
  int T = 1;
  while (T * 2 <= n_threads && T < 64) T *= 2;
  if (T == 1 || n < (1u << 20)) {
    std::sort(v.begin(), v.end());
    return;
  }
  std::vector<size_t> cut((size_t)T + 1);
  for (int t = 0; t <= T; ++t) cut[(size_t)t] = n * (size_t)t / (size_t)T;
  {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] { std::sort(v.begin() + cut[(size_t)t], v.begin() + cut[(size_t)t + 1]); });
    for (auto &x : th) x.join();
  }
  std::vector<uint64_t> tmp(n);
  std::vector<uint64_t> *src = &v, *dst = &tmp;
  for (int width = 1; width < T; width *= 2) {
    std::vector<std::thread> th;
    for (int t = 0; t < T; t += 2 * width) {
      size_t a = cut[(size_t)t], b = cut[(size_t)std::min(t + width, T)], c = cut[(size_t)std::min(t + 2 * width, T)];
      th.emplace_back([=] {
        std::merge(src->begin() + a, src->begin() + b, src->begin() + b, src->begin() + c, dst->begin() + a);
      });
    }
    for (auto &x : th) x.join();
    std::swap(src, dst);
  }
  if (src != &v) v.swap(tmp);
}

void append_hits(ko_result &r, const SearchRes &sr, bool withPos) {
  for (auto &h : sr.hits) {
    r.subject.push_back(h.key);
    r.kmatch.push_back(h.kmatch);
    if (withPos) {
      auto it = sr.posHits.find(h.key);
      if (it != sr.posHits.end()) r.pos.insert(r.pos.end(), it->second.begin(), it->second.end());
      r.pos_off.push_back(r.pos.size());
    }
  }
  r.hit_off.push_back(r.subject.size());
}

void merge_results(ko_result &dst, std::vector<ko_result> &parts, bool nt) {
  for (auto &p : parts) {
    uint64_t hbase = dst.subject.size(), pbase = dst.pos.size(), sbase = dst.row_seq.size();
    dst.subject.insert(dst.subject.end(), p.subject.begin(), p.subject.end());
    dst.kmatch.insert(dst.kmatch.end(), p.kmatch.begin(), p.kmatch.end());
    dst.size_in_kmer.insert(dst.size_in_kmer.end(), p.size_in_kmer.begin(), p.size_in_kmer.end());
    for (size_t i = 1; i < p.hit_off.size(); ++i) dst.hit_off.push_back(hbase + p.hit_off[i]);
    dst.pos.insert(dst.pos.end(), p.pos.begin(), p.pos.end());
    for (size_t i = 1; i < p.pos_off.size(); ++i) dst.pos_off.push_back(pbase + p.pos_off[i]);
    if (nt) {
      dst.row_contig.insert(dst.row_contig.end(), p.row_contig.begin(), p.row_contig.end());
      dst.row_start.insert(dst.row_start.end(), p.row_start.begin(), p.row_start.end());
      dst.row_end.insert(dst.row_end.end(), p.row_end.begin(), p.row_end.end());
      dst.row_plus.insert(dst.row_plus.end(), p.row_plus.begin(), p.row_plus.end());
      dst.row_seq.insert(dst.row_seq.end(), p.row_seq.begin(), p.row_seq.end());
      for (size_t i = 1; i < p.row_seq_off.size(); ++i) dst.row_seq_off.push_back(sbase + p.row_seq_off[i]);
    }
    dst.n_lookups += p.n_lookups;
    dst.n_increments += p.n_increments;
  }
}

// ---------------------------------------------------------------------------------------
// a12. BLOSUM62 in biogo's alphabet.Protein order "-ABCDEFGHIJKLMNPQRSTVWXYZ*"
// (kaamer's AAPosInMatrix, pkg/align/matrixScores.go:107).  Values: NCBI BLOSUM62
// (24-letter file order ARNDCQEGHILKMFPSTWYVBZX*) + the J row of the BLAST+ matrix.
// The gap row/column (index 0) is 0: kaamer's `Score() == -GapOpen` gap test
// (align.go:127) only works if every gap segment scores exactly GapOpen, i.e. biogo's
// per-residue gap cost for BLOSUM62 is 0.  biogo is not vendored => **parity unpinned**.
// ---------------------------------------------------------------------------------------
const char *NCBI_ORDER = "ARNDCQEGHILKMFPSTWYVBZX*";
const int8_t NCBI_B62[24][24] = {
    {4, -1, -2, -2, 0, -1, -1, 0, -2, -1, -1, -1, -1, -2, -1, 1, 0, -3, -2, 0, -2, -1, 0, -4},
    {-1, 5, 0, -2, -3, 1, 0, -2, 0, -3, -2, 2, -1, -3, -2, -1, -1, -3, -2, -3, -1, 0, -1, -4},
    {-2, 0, 6, 1, -3, 0, 0, 0, 1, -3, -3, 0, -2, -3, -2, 1, 0, -4, -2, -3, 3, 0, -1, -4},
    {-2, -2, 1, 6, -3, 0, 2, -1, -1, -3, -4, -1, -3, -3, -1, 0, -1, -4, -3, -3, 4, 1, -1, -4},
    {0, -3, -3, -3, 9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2, -4},
    {-1, 1, 0, 0, -3, 5, 2, -2, 0, -3, -2, 1, 0, -3, -1, 0, -1, -2, -1, -2, 0, 3, -1, -4},
    {-1, 0, 0, 2, -4, 2, 5, -2, 0, -3, -3, 1, -2, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4},
    {0, -2, 0, -1, -3, -2, -2, 6, -2, -4, -4, -2, -3, -3, -2, 0, -2, -2, -3, -3, -1, -2, -1, -4},
    {-2, 0, 1, -1, -3, 0, 0, -2, 8, -3, -3, -1, -2, -1, -2, -1, -2, -2, 2, -3, 0, 0, -1, -4},
    {-1, -3, -3, -3, -1, -3, -3, -4, -3, 4, 2, -3, 1, 0, -3, -2, -1, -3, -1, 3, -3, -3, -1, -4},
    {-1, -2, -3, -4, -1, -2, -3, -4, -3, 2, 4, -2, 2, 0, -3, -2, -1, -2, -1, 1, -4, -3, -1, -4},
    {-1, 2, 0, -1, -3, 1, 1, -2, -1, -3, -2, 5, -1, -3, -1, 0, -1, -3, -2, -2, 0, 1, -1, -4},
    {-1, -1, -2, -3, -1, 0, -2, -3, -2, 1, 2, -1, 5, 0, -2, -1, -1, -1, -1, 1, -3, -1, -1, -4},
    {-2, -3, -3, -3, -2, -3, -3, -3, -1, 0, 0, -3, 0, 6, -4, -2, -2, 1, 3, -1, -3, -3, -1, -4},
    {-1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7, -1, -1, -4, -3, -2, -2, -1, -2, -4},
    {1, -1, 1, 0, -1, 0, 0, 0, -1, -2, -2, 0, -1, -2, -1, 4, 1, -3, -2, -2, 0, 0, 0, -4},
    {0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1, 5, -2, -2, 0, -1, -1, 0, -4},
    {-3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1, -4, -3, -2, 11, 2, -3, -4, -3, -2, -4},
    {-2, -2, -2, -3, -2, -1, -2, -3, 2, -1, -1, -2, -1, 3, -3, -2, -2, 2, 7, -1, -3, -2, -1, -4},
    {0, -3, -3, -3, -1, -2, -2, -3, -3, 3, 1, -2, 1, -1, -2, -2, 0, -3, -1, 4, -3, -2, -1, -4},
    {-2, -1, 3, 4, -3, 0, 1, -1, 0, -3, -4, 0, -3, -3, -2, 0, -1, -4, -3, -3, 4, 1, -1, -4},
    {-1, 0, 0, 1, -3, 3, 4, -2, 0, -3, -3, 1, -1, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4},
    {0, -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0, 0, -2, -1, -1, -1, -1, -1, -4},
    {-4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, 1}};
// J (I/L ambiguity) vs ARNDCQEGHILKMFPSTWYVBZX* ; J/J = 3
const int8_t NCBI_J[24] = {-1, -2, -3, -3, -1, -2, -3, -4, -3, 3, 3, -3, 2, 0, -3, -2, -1, -2, -1, 2, -3, -3, -1, -4};

const char *BIOGO_ORDER = "-ABCDEFGHIJKLMNPQRSTVWXYZ*";  // matrixScores.go:107

struct Blosum62 {
  int32_t m[26][26];
  int8_t letter_index[256];  // alphabet.Protein.LetterIndex(), case-insensitive; -1 illegal
  int8_t aa_pos[256];        // kaamer AAPosInMatrix: upper-case only, miss -> 0
  Blosum62() {
    memset(m, 0, sizeof m);
    auto ncbi = [&](char c) -> int {
      const char *p = strchr(NCBI_ORDER, c);
      return p ? (int)(p - NCBI_ORDER) : -1;
    };
    for (int i = 1; i < 26; ++i)
      for (int j = 1; j < 26; ++j) {
        char a = BIOGO_ORDER[i], b = BIOGO_ORDER[j];
        int v;
        if (a == 'J' && b == 'J') v = 3;
        else if (a == 'J') v = NCBI_J[ncbi(b)];
        else if (b == 'J') v = NCBI_J[ncbi(a)];
        else v = NCBI_B62[ncbi(a)][ncbi(b)];
        m[i][j] = v;
      }
    memset(letter_index, -1, sizeof letter_index);
    memset(aa_pos, 0, sizeof aa_pos);
    for (int i = 0; i < 26; ++i) {
      unsigned char c = (unsigned char)BIOGO_ORDER[i];
      letter_index[c] = (int8_t)i;
      letter_index[(unsigned char)tolower(c)] = (int8_t)i;
      aa_pos[c] = (int8_t)i;
    }
  }
};
Blosum62 g_b62;          // the alignment model; ko_set_align_model replaces the scores (§8f-4, and the biogo gap row)
int32_t g_gap_open_dp = -11;  // SWAffine.GapOpen (align.go:64)

struct Segment {
  int32_t qs, qe, ss, se;  // half-open, 0-based
  int32_t score;
};

/*
 * Restated biogo align.SWAffine{Matrix: BLOSUM62, GapOpen: -11}.Align(query, subject)
 * (call site pkg/align/align.go:62-67).  **PARITY UNPINNED**: biogo v1.0.1 is not in the
 * reference tree.  Definition used here (and matched bit-exactly by the CUDA kernel):
 *   three layers over a zero-initialised (len(q)+1) x (len(s)+1) table
 *     M[i][j] = max(0, max(M,U,L)[i-1][j-1] + B62[q_i][s_j])
 *     U[i][j] = max(M[i-1][j] + open + B62[q_i][-], U[i-1][j] + B62[q_i][-])   (gap in subject)
 *     L[i][j] = max(M[i][j-1] + open + B62[-][s_j], L[i][j-1] + B62[-][s_j])   (gap in query)
 *   open = -11, gap row/column = 0 (free extension; kaamer re-adds GapExtend afterwards).
 *   End cell: the LAST cell in row-major order attaining the maximum of M (>=).
 *   Traceback: from an M cell stop when M == 0, else step diagonally into the best of
 *   (M,U,L)[i-1][j-1] preferring M, then U, then L; inside U (L) prefer "open" (return
 *   to M) over "extend" on ties.  Result = list of segments (runs of diagonal steps, or
 *   one gap run), each with an integer score; their sum is the DP optimum.
 */
int sw_affine(const uint8_t *q, int32_t n, const uint8_t *s, int32_t m, std::vector<Segment> &segs,
              int32_t &dp_score) {
  segs.clear();
  dp_score = 0;
  const int32_t open = g_gap_open_dp;
  for (int32_t i = 0; i < n; ++i)
    if (g_b62.letter_index[q[i]] < 0) return -1;
  for (int32_t j = 0; j < m; ++j)
    if (g_b62.letter_index[s[j]] < 0) return -1;
  if (n == 0 || m == 0) return 0;
  const size_t C = (size_t)m + 1;
  std::vector<int32_t> M((size_t)(n + 1) * C, 0), U((size_t)(n + 1) * C, 0), L((size_t)(n + 1) * C, 0);
  int32_t maxS = 0, maxI = 0, maxJ = 0;
  for (int32_t i = 1; i <= n; ++i) {
    int qi = g_b62.letter_index[q[i - 1]];
    for (int32_t j = 1; j <= m; ++j) {
      int sj = g_b62.letter_index[s[j - 1]];
      size_t p = (size_t)i * C + (size_t)j;
      int32_t best = std::max(M[p - C - 1], std::max(U[p - C - 1], L[p - C - 1]));
      int32_t d = best + g_b62.m[qi][sj];
      M[p] = d > 0 ? d : 0;
      int32_t gq = g_b62.m[qi][0], gs = g_b62.m[0][sj];
      U[p] = std::max(M[p - C] + open + gq, U[p - C] + gq);
      L[p] = std::max(M[p - 1] + open + gs, L[p - 1] + gs);
      if (M[p] > 0 && M[p] >= maxS) {
        maxS = M[p];
        maxI = i;
        maxJ = j;
      }
    }
  }
  dp_score = maxS;
  if (maxS == 0) return 0;
  // traceback
  int32_t i = maxI, j = maxJ;
  int layer = 0;  // 0=M,1=U,2=L
  std::vector<Segment> rev;
  int cur_kind = -1;
  Segment cur{};
  auto flush = [&]() {
    if (cur_kind >= 0) rev.push_back(cur);
    cur_kind = -1;
  };
  while (i > 0 && j > 0) {
    size_t p = (size_t)i * C + (size_t)j;
    if (layer == 0) {
      if (M[p] == 0) break;
      int sc = g_b62.m[g_b62.letter_index[q[i - 1]]][g_b62.letter_index[s[j - 1]]];
      if (cur_kind != 0) {
        flush();
        cur_kind = 0;
        cur = Segment{i, i, j, j, 0};
      }
      cur.qs = i - 1;
      cur.ss = j - 1;
      cur.score += sc;
      size_t pp = p - C - 1;
      int32_t bm = M[pp], bu = U[pp], bl = L[pp];
      if (bm >= bu && bm >= bl) layer = 0;
      else if (bu >= bl) layer = 1;
      else layer = 2;
      --i;
      --j;
    } else if (layer == 1) {
      int32_t gq = g_b62.m[g_b62.letter_index[q[i - 1]]][0];
      if (cur_kind != 1) {
        flush();
        cur_kind = 1;
        cur = Segment{i, i, j, j, 0};
      }
      cur.qs = i - 1;
      if (U[p] == M[p - C] + open + gq) {
        cur.score += open + gq;
        layer = 0;
      } else {
        cur.score += gq;
      }
      --i;
    } else {
      int32_t gs = g_b62.m[0][g_b62.letter_index[s[j - 1]]];
      if (cur_kind != 2) {
        flush();
        cur_kind = 2;
        cur = Segment{i, i, j, j, 0};
      }
      cur.ss = j - 1;
      if (L[p] == M[p - 1] + open + gs) {
        cur.score += open + gs;
        layer = 0;
      } else {
        cur.score += gs;
      }
      --j;
    }
  }
  flush();
  segs.assign(rev.rbegin(), rev.rend());
  return 0;
}

}  // namespace

// =======================================================================================
extern "C" {

uint32_t ko_encode_kmer(const uint8_t *k) { return encode_kmer(k); }

void ko_decode_kmer(uint32_t key, uint8_t *out7) {  // k_store.go:120-145
  const char *aa = "ACDEFGHIKLMNPQRSTUVWY";
  uint32_t f[3] = {(key >> 23) & 0x1FF, (key >> 14) & 0x1FF, (key >> 5) & 0x1FF};
  for (int i = 0; i < 3; ++i) {
    if (f[i] >= 22 && f[i] <= 462) {
      out7[2 * i] = (uint8_t)aa[(f[i] - 22) / 21];
      out7[2 * i + 1] = (uint8_t)aa[(f[i] - 22) % 21];
    } else {
      out7[2 * i] = out7[2 * i + 1] = 0;  // Go: zero rune from a missing map entry
    }
  }
  uint32_t d = key & 0x1F;
  out7[6] = d <= 20 ? (uint8_t)aa[d] : 0;
}

int32_t ko_size_in_kmer(const uint8_t *seq, uint64_t len) { return size_in_kmer(seq, len); }

void ko_fasta_ids(uint64_t n, uint32_t *ids_out) {
  // inputFASTA.go:96-124: proteinNb++ on every '>' BEFORE the previous record is dispatched
  // with id=proteinNb; the last record is dispatched after EOF with the final proteinNb.
  uint64_t proteinNb = 0;
  int64_t pending = -1;
  for (uint64_t j = 0; j < n; ++j) {
    proteinNb += 1;
    if (pending >= 0) ids_out[pending] = (uint32_t)proteinNb;
    pending = (int64_t)j;
  }
  if (pending >= 0) ids_out[pending] = (uint32_t)proteinNb;
}

ko_index *ko_index_build(const uint8_t *residues, const uint64_t *off, const uint32_t *ids,
                         uint64_t n_records, int n_threads) {
  auto *idx = new ko_index();
  // processProteinInputFASTA (inputFASTA.go:219-248): every window i in [0, len-7] -> (kmer, id)
  std::vector<uint64_t> woff(n_records + 1, 0);
  for (uint64_t r = 0; r < n_records; ++r) {
    uint64_t len = off[r + 1] - off[r];
    uint64_t w = len >= KMER_SIZE ? len - KMER_SIZE + 1 : 0;
    woff[r + 1] = woff[r] + w;
    if (len >= KMER_SIZE) {  // skip peptide shorter than kmerSize (:226-228)
      idx->n_proteins++;
      idx->n_aa += len;
      idx->n_kmers += len - KMER_SIZE + 1;  // inputFASTA.go:142-145
    }
  }
  std::vector<uint64_t> pairs(woff[n_records]);
  parallel_for(n_records, n_threads, [&](int, uint64_t b, uint64_t e) {
    for (uint64_t r = b; r < e; ++r) {
      const uint8_t *s = residues + off[r];
      uint64_t w = woff[r + 1] - woff[r];
      for (uint64_t i = 0; i < w; ++i)
        pairs[woff[r] + i] = ((uint64_t)encode_kmer(s + i) << 32) | ids[r];
    }
  });
  parallel_sort(pairs, n_threads);
  // IndexStore + CreateKCKeyValue + RemoveDuplicatesFromSlice (indexdb.go:92-128,
  // kcomb_store.go:42-85, kv_store.go:284-305): per k-mer the SET of ids, sorted descending.
  pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
  size_t n = pairs.size();
  idx->postings.resize(n);
  idx->offsets.push_back(0);
  size_t g = 0;
  while (g < n) {
    uint32_t key = (uint32_t)(pairs[g] >> 32);
    size_t h = g;
    while (h < n && (uint32_t)(pairs[h] >> 32) == key) ++h;
    for (size_t t = g; t < h; ++t) idx->postings[g + (h - 1 - t)] = (uint32_t)pairs[t];  // descending
    idx->keys.push_back(key);
    idx->offsets.push_back(h);
    g = h;
  }
  return idx;
}

ko_index *ko_index_from_arrays(const uint32_t *keys, const uint64_t *offsets, const uint32_t *postings,
                               uint64_t n_keys, uint64_t n_proteins, uint64_t n_aa, uint64_t n_kmers) {
  auto *idx = new ko_index();
  idx->keys.assign(keys, keys + n_keys);
  idx->offsets.assign(offsets, offsets + n_keys + 1);
  idx->postings.assign(postings, postings + offsets[n_keys]);
  idx->n_proteins = n_proteins;
  idx->n_aa = n_aa;
  idx->n_kmers = n_kmers;
  return idx;
}
void ko_index_free(ko_index *i) { delete i; }
uint64_t ko_index_n_keys(const ko_index *i) { return i->keys.size(); }
uint64_t ko_index_n_postings(const ko_index *i) { return i->postings.size(); }
const uint32_t *ko_index_keys(const ko_index *i) { return i->keys.data(); }
const uint64_t *ko_index_offsets(const ko_index *i) { return i->offsets.data(); }
const uint32_t *ko_index_postings(const ko_index *i) { return i->postings.data(); }
void ko_index_stats(const ko_index *i, uint64_t *p, uint64_t *a, uint64_t *k) {
  *p = i->n_proteins;
  *a = i->n_aa;
  *k = i->n_kmers;
}

ko_result *ko_search_proteins(const ko_index *idx, const uint8_t *residues, const uint64_t *off,
                              uint32_t nq, const ko_opts *o, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::vector<ko_result> parts((size_t)std::max(1, std::min<int>(n_threads, (int)std::max<uint32_t>(1, nq))));
  parallel_for(nq, (int)parts.size(), [&](int t, uint64_t b, uint64_t e) {
    ko_result &r = parts[(size_t)t];
    for (uint64_t qi = b; qi < e; ++qi) {
      const uint8_t *seq = residues + off[qi];
      uint64_t len = off[qi + 1] - off[qi];
      int32_t sk = size_in_kmer(seq, len);
      r.size_in_kmer.push_back(sk);
      SearchRes sr;
      // search_protein.go:74-76: `if q.SizeInKmer < 7 { return }` kills the worker goroutine
      // (may deadlock the request).  Documented deviation: the query is skipped (no hits).
      if (sk >= 7) {
        kmer_search(*idx, seq, sk, o->want_positions != 0, sr);
        sort_map_by_value(sr);
        filter_results(sr.hits, &sr.posHits, sk, *o);
      }
      r.n_lookups += sr.n_lookups;
      r.n_increments += sr.n_increments;
      append_hits(r, sr, o->want_positions != 0);
    }
  });
  auto *res = new ko_result();
  merge_results(*res, parts, false);
  return res;
}

struct ko_orfs {
  std::vector<uint8_t> seq;
  std::vector<uint64_t> seq_off{0};
  std::vector<int64_t> start, end;
  std::vector<uint8_t> plus;
  std::vector<int32_t> alts;
  std::vector<uint64_t> alts_off{0};
};

ko_orfs *ko_get_orfs(const uint8_t *dna, uint64_t len) {
  auto v = get_orfs(dna, len);
  auto *o = new ko_orfs();
  for (auto &f : v) {
    o->seq.insert(o->seq.end(), f.seq.begin(), f.seq.end());
    o->seq_off.push_back(o->seq.size());
    o->start.push_back(f.loc.start);
    o->end.push_back(f.loc.end);
    o->plus.push_back(f.loc.plus ? 1 : 0);
    o->alts.insert(o->alts.end(), f.loc.alts.begin(), f.loc.alts.end());
    o->alts_off.push_back(o->alts.size());
  }
  return o;
}
void ko_orfs_free(ko_orfs *o) { delete o; }
uint64_t ko_orfs_n(const ko_orfs *o) { return o->start.size(); }
const uint8_t *ko_orfs_seq(const ko_orfs *o) { return o->seq.data(); }
const uint64_t *ko_orfs_seq_off(const ko_orfs *o) { return o->seq_off.data(); }
const int64_t *ko_orfs_start(const ko_orfs *o) { return o->start.data(); }
const int64_t *ko_orfs_end(const ko_orfs *o) { return o->end.data(); }
const uint8_t *ko_orfs_plus(const ko_orfs *o) { return o->plus.data(); }
const int32_t *ko_orfs_alts(const ko_orfs *o) { return o->alts.data(); }
const uint64_t *ko_orfs_alts_off(const ko_orfs *o) { return o->alts_off.data(); }

ko_result *ko_search_nucleotide(const ko_index *idx, const uint8_t *nt, const uint64_t *off,
                                uint32_t n_contigs, const ko_opts *o, int n_threads) {
  auto *res = new ko_result();
  for (uint32_t c = 0; c < n_contigs; ++c) {  // contigs are serial in the reference (:61)
    auto orfs = get_orfs(nt + off[c], off[c + 1] - off[c]);  // search_nucleotide.go:136
    std::vector<ko_result> parts((size_t)std::max(1, std::min<int>(n_threads, (int)std::max<size_t>(1, orfs.size()))));
    parallel_for(orfs.size(), (int)parts.size(), [&](int t, uint64_t b, uint64_t e) {
      ko_result &r = parts[(size_t)t];
      for (uint64_t oi = b; oi < e; ++oi) {
        ORF &orf = orfs[oi];
        std::string seq = orf.seq;
        Location loc = orf.loc;
        int32_t sk = (int32_t)seq.size() - KMER_SIZE + 1;  // search_nucleotide.go:81
        if (!seq.empty() && seq.back() == '*') sk -= 1;     // :88-90
        SearchRes sr;
        kmer_search(*idx, (const uint8_t *)seq.data(), sk, true, sr);  // positions always (search.go:416)
        r.n_lookups += sr.n_lookups;
        r.n_increments += sr.n_increments;
        sort_map_by_value(sr);
        if (!sr.hits.empty() && sr.hits[0].kmatch >= o->min_kmatch) {  // :116
          set_best_start_codon(seq, loc, sk, sr);                     // :118
          filter_results(sr.hits, &sr.posHits, sk, *o);               // :119
          if (!sr.hits.empty()) {                                     // :120
            r.size_in_kmer.push_back(sk);
            r.row_contig.push_back(c);
            r.row_start.push_back(loc.start);
            r.row_end.push_back(loc.end);
            r.row_plus.push_back(loc.plus ? 1 : 0);
            r.row_seq.insert(r.row_seq.end(), seq.begin(), seq.end());
            r.row_seq_off.push_back(r.row_seq.size());
            append_hits(r, sr, true);
          }
        }
      }
    });
    merge_results(*res, parts, true);
  }
  return res;
}

void ko_result_free(ko_result *r) { delete r; }
uint64_t ko_result_n_rows(const ko_result *r) { return r->hit_off.size() - 1; }
const uint64_t *ko_result_hit_off(const ko_result *r) { return r->hit_off.data(); }
const uint32_t *ko_result_subject(const ko_result *r) { return r->subject.data(); }
const int64_t *ko_result_kmatch(const ko_result *r) { return r->kmatch.data(); }
const int32_t *ko_result_size_in_kmer(const ko_result *r) { return r->size_in_kmer.data(); }
const uint64_t *ko_result_pos_off(const ko_result *r) { return r->pos_off.data(); }
const uint8_t *ko_result_pos(const ko_result *r) { return r->pos.data(); }
const uint32_t *ko_result_row_contig(const ko_result *r) { return r->row_contig.data(); }
const int64_t *ko_result_row_start(const ko_result *r) { return r->row_start.data(); }
const int64_t *ko_result_row_end(const ko_result *r) { return r->row_end.data(); }
const uint8_t *ko_result_row_plus(const ko_result *r) { return r->row_plus.data(); }
const uint8_t *ko_result_row_seq(const ko_result *r) { return r->row_seq.data(); }
const uint64_t *ko_result_row_seq_off(const ko_result *r) { return r->row_seq_off.data(); }
uint64_t ko_result_n_lookups(const ko_result *r) { return r->n_lookups; }
uint64_t ko_result_n_increments(const ko_result *r) { return r->n_increments; }

int32_t ko_filter_count(const int64_t *kmatch_sorted, int32_t n, int32_t sk, const ko_opts *o) {
  std::vector<Hit> hits;
  for (int32_t i = 0; i < n; ++i) hits.push_back({(uint32_t)i, kmatch_sorted[i]});
  filter_results(hits, nullptr, sk, *o);
  return (int32_t)hits.size();
}

int32_t ko_blosum62(int32_t i, int32_t j) { return g_b62.m[i][j]; }

double ko_bitscore(double lambda, double K, int32_t raw) {  // align.go:136
  return ((lambda * (double)raw) - std::log(K)) / std::log(2.0);
}
double ko_evalue(int32_t qlen, uint64_t number_of_aa, double bitscore) {  // align.go:141
  return (double)qlen * (double)number_of_aa / std::pow(2.0, bitscore);
}

int ko_align(const uint8_t *q_in, int32_t qlen, const uint8_t *s_in, int32_t slen,
             const ko_aln_params *prm, ko_aln *out, char *aln_a, char *aln_b, int32_t aln_cap) {
  // align.go:54-55 — [uU] -> '*'
  std::string q((const char *)q_in, (size_t)qlen), s((const char *)s_in, (size_t)slen);
  for (auto &c : q) if (c == 'u' || c == 'U') c = '*';
  for (auto &c : s) if (c == 'u' || c == 'U') c = '*';
  std::vector<Segment> segs;
  int32_t dp = 0;
  int rc = sw_affine((const uint8_t *)q.data(), qlen, (const uint8_t *)s.data(), slen, segs, dp);
  memset(out, 0, sizeof *out);
  out->illegal = rc < 0 ? 1 : 0;  // biogo error is ignored by kaamer (align.go:67)
  out->dp_score = dp;
  out->n_segments = (int32_t)segs.size();
  // align.Format (align.go:69): gapped strings
  std::string a, b;
  for (auto &g : segs) {
    int32_t lq = g.qe - g.qs, ls = g.se - g.ss;
    if (lq > 0 && ls > 0) {
      a.append(q, (size_t)g.qs, (size_t)lq);
      b.append(s, (size_t)g.ss, (size_t)ls);
    } else if (lq > 0) {
      a.append(q, (size_t)g.qs, (size_t)lq);
      b.append((size_t)lq, '-');
    } else {
      a.append((size_t)ls, '-');
      b.append(s, (size_t)g.ss, (size_t)ls);
    }
  }
  // align.go:72-101 — float32 arithmetic
  float identity = 0, similarity = 0, nbPos = 0;
  int32_t mismatches = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    unsigned char ca = (unsigned char)a[i], cb = (unsigned char)b[i];
    if (cb == ca) {
      identity += 1;
      similarity += 1;
    } else {
      if (cb != '-' && ca != '-') mismatches += 1;
      if (g_b62.m[g_b62.aa_pos[cb]][g_b62.aa_pos[ca]] > 0) similarity += 1;  // GetAlnScoreAA
    }
    nbPos += 1;
  }
  identity = (identity / nbPos) * 100;  // NaN when the alignment is empty, as in Go
  similarity = (similarity / nbPos) * 100;
  // align.go:104-132
  int32_t rawScore = 0, gapOpenings = 0;
  int32_t queryStart = 0, queryEnd = 0, subjectStart = 0, subjectEnd = 0;
  for (size_t i = 0; i < segs.size(); ++i) {
    const Segment &g = segs[i];
    if (i == 0) {
      queryStart = g.qs;
      subjectStart = g.ss;
    }
    if (i == segs.size() - 1) {
      queryEnd = g.qe;
      subjectEnd = g.se;
    }
    rawScore += g.score;
    if (g.score == -prm->gap_open_opt) {  // align.go:127 — test on the score VALUE
      gapOpenings += 1;
      int32_t gapLen = std::max(g.qe - g.qs, g.se - g.ss);
      rawScore = rawScore - ((gapLen - 1) * prm->gap_extend_opt);
    }
  }
  out->identity = identity;
  out->similarity = similarity;
  out->length = (int32_t)a.size();
  out->mismatches = mismatches;
  out->gap_openings = gapOpenings;
  out->raw = rawScore;
  out->bitscore = ko_bitscore(prm->lambda, prm->K, rawScore);
  out->evalue = ko_evalue(qlen, prm->number_of_aa, out->bitscore);  // len(querySeq) in bytes
  out->query_start = queryStart + 1;  // align.go:153-156
  out->query_end = queryEnd;
  out->subject_start = subjectStart + 1;
  out->subject_end = subjectEnd;
  if (aln_a && aln_b && aln_cap > 0) {
    size_t nn = std::min<size_t>(a.size(), (size_t)aln_cap - 1);
    memcpy(aln_a, a.data(), nn);
    aln_a[nn] = 0;
    memcpy(aln_b, b.data(), nn);
    aln_b[nn] = 0;
  }
  return 0;
}

// Replace the genetic code of ko_get_orfs / ko_search_nucleotide: aas64 in TCAG codon order ('*' = stop),
// start_mask bit i = codon i is a start codon.  The reference always uses table 11 (dna.go:106).
void ko_set_genetic_code(const char *aas64, uint64_t start_mask) {
  if (!aas64) {
    g_gcode = GCode();
    return;
  }
  for (int i = 0; i < 64; ++i) {
    g_gcode.t[i].aa = aas64[i];
    g_gcode.t[i].stop = aas64[i] == '*';
    g_gcode.t[i].start = ((start_mask >> i) & 1ull) != 0;
  }
}

// AlnString (align.go:69-103): "aString\nalnMatch\nbString" from the two gapped strings of ko_align
int32_t ko_aln_string(const char *a, const char *b, int32_t n, char *out, int32_t cap) {
  std::string m;
  for (int32_t i = 0; i < n; ++i) {
    unsigned char ca = (unsigned char)a[i], cb = (unsigned char)b[i];
    if (cb == ca) m += (char)cb;                                              // align.go:82-85
    else if (g_b62.m[g_b62.aa_pos[cb]][g_b62.aa_pos[ca]] > 0) m += '+';       // :91-93
    else m += ' ';                                                            // :95
  }
  std::string t = std::string(a, (size_t)n) + "\n" + m + "\n" + std::string(b, (size_t)n);
  if ((int32_t)t.size() + 1 > cap) return -(int32_t)t.size();
  memcpy(out, t.data(), t.size());
  out[t.size()] = 0;
  return (int32_t)t.size();
}

// Replace the scores of the alignment model: m[26*26] in biogo alphabet.Protein order
// "-ABCDEFGHIJKLMNPQRSTVWXYZ*" (row/column 0 = per-residue gap cost), gap_open = SWAffine.GapOpen.
// The reference hard-wires BLOSUM62 / -11 (align.go:62-65); this exists so that (a) whichever gap row biogo's
// BLOSUM62 carries is a parameter, not a rewrite, and (b) §8f-4's "other matrices behind an explicit flag".
void ko_set_align_model(const int8_t *m, int32_t gap_open) {
  for (int i = 0; i < 26; ++i)
    for (int j = 0; j < 26; ++j) g_b62.m[i][j] = m[i * 26 + j];
  g_gap_open_dp = gap_open;
}
void ko_reset_align_model() {
  g_b62 = Blosum62();
  g_gap_open_dp = -11;
}

int32_t ko_format_positions(const uint8_t *positions, int32_t n, int32_t withAlignment, char *out,
                            int32_t cap) {  // search.go:694-742
  int currentStart = 0, endPos = 0;
  bool inSequence = false;
  std::string ps;
  for (int pos = 0; pos < n; ++pos) {
    bool match = positions[pos] != 0;
    if (match) {
      if (!inSequence) {
        currentStart = pos + 1;
        inSequence = true;
      }
    } else {
      if (inSequence) {
        if (pos + 1 > currentStart) {
          if (!ps.empty()) ps += ",";
          endPos = pos + 1;
          if (withAlignment) endPos = endPos + KMER_SIZE - 1;
          ps += std::to_string(currentStart) + "-" + std::to_string(endPos);
          inSequence = false;
        } else {
          if (!ps.empty()) ps += ",";
          ps += std::to_string(currentStart);
          inSequence = false;
        }
      }
    }
  }
  if (inSequence) {
    if (!ps.empty()) ps += ",";
    endPos = n;
    if (withAlignment) endPos = endPos + KMER_SIZE - 1;
    ps += std::to_string(currentStart) + "-" + std::to_string(endPos);
  }
  if (out && cap > 0) {
    size_t nn = std::min<size_t>(ps.size(), (size_t)cap - 1);
    memcpy(out, ps.data(), nn);
    out[nn] = 0;
  }
  return (int32_t)ps.size();
}

}  // extern "C"
