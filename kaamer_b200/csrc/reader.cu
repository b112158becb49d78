// reader.cu — host-side query readers with the reference's exact semantics, filling the flat
// (residues, offsets) batches the search entry points take.  Host code only (no kernels).
//
//   kaamer_host_read_fasta   GetQueriesFasta   pkg/search/search.go:222-322
//   kaamer_host_read_fastq   GetQueriesFastq   pkg/search/search.go:324-412
//
// Reproduced on purpose:
//   * the file type is sniffed on the first 32 bytes with http.DetectContentType; anything that is
//     neither gzip nor "text/plain; charset=utf-8" yields NO queries, silently (:255-271) — which
//     includes every file shorter than 32 bytes (the zero padding of the sniff buffer is binary);
//   * bufio.Scanner with a 1 MiB token limit (:273-274): a line of 1 MiB or more ends the input,
//     what was read so far is kept;
//   * FASTA: Name = header without '>', Sequence = concatenation of the trimmed lines, upper-cased
//     for every record EXCEPT THE LAST (:295 vs :313-320); SizeInKmer = len - 7 + 1, minus one when
//     the sequence ends with '*' (:290-293);
//   * FASTQ: every line starting with '@' starts a record (quality lines too), the sequence is the
//     LAST line of the record matching ^[ATGCNatgcn]+$, never upper-cased, no '*' rule (:392-403).
// Go's strings.ToUpper / TrimSpace are Unicode-aware; this reader handles ASCII (and the UTF-8
// encodings of U+0085 / U+00A0 for TrimSpace) — sequence files are ASCII.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "internal.cuh"

namespace kaamer {

namespace {
constexpr size_t MAX_TOKEN = 1024 * 1024;  // scanner.Buffer(buf, 1024*1024)

struct BatchOwner {
  bool pinned = false;
  std::vector<std::pair<void *, size_t>> blocks;  // pinned_get blocks or malloc'ed (cap 0)
  ~BatchOwner() {
    for (auto &b : blocks) {
      if (pinned) pinned_put(b.first, b.second);
      else free(b.first);
    }
  }
  template <class T>
  int alloc(T **out, size_t n) {
    const size_t bytes = (n ? n : 1) * sizeof(T);
    if (pinned) {
      size_t cap = 0;
      void *p = pinned_get(bytes, &cap);
      if (!p) return KAAMER_ERR_NOMEM;
      blocks.emplace_back(p, cap);
      *out = (T *)p;
    } else {
      void *p = malloc(bytes);
      if (!p) {
        set_error("out of host memory (%zu bytes)", bytes);
        return KAAMER_ERR_NOMEM;
      }
      blocks.emplace_back(p, (size_t)0);
      *out = (T *)p;
    }
    return KAAMER_OK;
  }
};

bool ieq_prefix(const uint8_t *d, size_t n, const char *sig) {  // case-insensitive ASCII prefix
  const size_t m = strlen(sig);
  if (n < m) return false;
  for (size_t i = 0; i < m; ++i) {
    uint8_t c = d[i];
    if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
    if (c != (uint8_t)sig[i]) return false;
  }
  return true;
}
bool has_prefix(const uint8_t *d, size_t n, const char *sig, size_t m) { return n >= m && memcmp(d, sig, m) == 0; }

// http.DetectContentType (net/http/sniff.go) restricted to what decides between "gzip",
// "text/plain; charset=utf-8" and "anything else" for a 32-byte buffer.
enum Sniff { SNIFF_GZIP, SNIFF_TEXT, SNIFF_OTHER };
Sniff detect_content_type(const uint8_t *d, size_t n) {
  size_t ws = 0;
  while (ws < n && (d[ws] == '\t' || d[ws] == '\n' || d[ws] == '\x0c' || d[ws] == '\r' || d[ws] == ' ')) ++ws;
  static const char *html[] = {"<!DOCTYPE HTML", "<HTML", "<HEAD", "<SCRIPT", "<IFRAME", "<H1", "<DIV", "<FONT", "<TABLE",
                               "<A", "<STYLE", "<TITLE", "<B", "<BODY", "<BR", "<P", "<!--"};
  for (const char *sig : html) {
    const size_t m = strlen(sig);
    if (ieq_prefix(d + ws, n - ws, sig) && ws + m < n && (d[ws + m] == ' ' || d[ws + m] == '>')) return SNIFF_OTHER;
  }
  if (has_prefix(d + ws, n - ws, "<?xml", 5)) return SNIFF_OTHER;
  if (has_prefix(d, n, "%PDF-", 5) || has_prefix(d, n, "%!PS-Adobe-", 11)) return SNIFF_OTHER;
  if (has_prefix(d, n, "\xFE\xFF", 2) || has_prefix(d, n, "\xFF\xFE", 2)) return SNIFF_OTHER;  // charset=utf-16
  if (has_prefix(d, n, "\xEF\xBB\xBF", 3)) return SNIFF_TEXT;                                    // utf-8 BOM
  // signatures made of bytes that the text test below would accept (the others — PNG, zip, rar, ico,
  // webm, wasm, mp4, ogg, midi ... — contain a byte of the binary class and fall out there anyway)
  static const char *bin[] = {"GIF87a", "GIF89a", "BM", "ID3", ".snd", "wOFF", "wOF2", "OTTO", "ttcf", "\xFF\xD8\xFF"};
  for (const char *sig : bin)
    if (has_prefix(d, n, sig, strlen(sig))) return SNIFF_OTHER;
  if (n >= 14 && has_prefix(d, n, "RIFF", 4) &&
      (memcmp(d + 8, "WEBPVP", 6) == 0 || memcmp(d + 8, "AVI ", 4) == 0 || memcmp(d + 8, "WAVE", 4) == 0))
    return SNIFF_OTHER;
  if (n >= 12 && has_prefix(d, n, "FORM", 4) && memcmp(d + 8, "AIFF", 4) == 0) return SNIFF_OTHER;
  if (has_prefix(d, n, "\x1F\x8B\x08", 3)) return SNIFF_GZIP;
  for (size_t i = ws; i < n; ++i) {
    const uint8_t c = d[i];
    if (c <= 0x08 || c == 0x0B || (c >= 0x0E && c <= 0x1A) || (c >= 0x1C && c <= 0x1F)) return SNIFF_OTHER;
  }
  return SNIFF_TEXT;
}

int read_file(const char *path, std::vector<uint8_t> *out) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    set_error("cannot open %s", path);
    return KAAMER_ERR_IO;
  }
  uint8_t buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->insert(out->end(), buf, buf + n);
  const bool bad = ferror(f) != 0;
  fclose(f);
  if (bad) {
    set_error("read error on %s", path);
    return KAAMER_ERR_IO;
  }
  return KAAMER_OK;
}

// gzip.NewReader + reads to the end (multi-member streams are concatenated, as Go's reader does);
// a corrupt tail ends the stream where it is (the Go scanner stops on the read error).
int gunzip(const std::vector<uint8_t> &in, std::vector<uint8_t> *out) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) {
    set_error("zlib: inflateInit2 failed");
    return KAAMER_ERR_IO;
  }
  zs.next_in = const_cast<Bytef *>(in.data());
  zs.avail_in = (uInt)in.size();  // (query files: < 4 GiB compressed)
  std::vector<uint8_t> buf(1 << 20);
  for (;;) {
    zs.next_out = buf.data();
    zs.avail_out = (uInt)buf.size();
    const int rc = inflate(&zs, Z_NO_FLUSH);
    out->insert(out->end(), buf.data(), buf.data() + (buf.size() - zs.avail_out));
    if (rc == Z_STREAM_END) {
      if (zs.avail_in == 0) break;
      if (inflateReset(&zs) != Z_OK) break;  // next member
      continue;
    }
    if (rc != Z_OK) break;  // truncated / corrupt: keep what was decoded
    if (zs.avail_in == 0 && zs.avail_out != 0) break;
  }
  inflateEnd(&zs);
  return KAAMER_OK;
}

// bufio.ScanLines over `data` with the 1 MiB token limit: calls fn(line, len) per line; stops at a line
// of MAX_TOKEN bytes or more (ErrTooLong ends the `for scanner.Scan()` loop).
template <class F>
void scan_lines(const std::vector<uint8_t> &data, F fn) {
  size_t p = 0;
  const size_t n = data.size();
  while (p < n) {
    const void *nl = memchr(data.data() + p, '\n', n - p);
    size_t e = nl ? (size_t)((const uint8_t *)nl - data.data()) : n;
    if (e - p >= MAX_TOKEN) return;  // the buffer fills up before a newline is seen
    size_t le = e;
    if (le > p && data[le - 1] == '\r') --le;  // dropCR
    fn(data.data() + p, le - p);
    p = nl ? e + 1 : n;
  }
}

bool is_space_at(const uint8_t *s, size_t n, size_t i, size_t *w) {  // unicode.IsSpace on ASCII + U+0085, U+00A0
  const uint8_t c = s[i];
  if (c == ' ' || (c >= '\t' && c <= '\r')) {
    *w = 1;
    return true;
  }
  if (c == 0xC2 && i + 1 < n && (s[i + 1] == 0x85 || s[i + 1] == 0xA0)) {
    *w = 2;
    return true;
  }
  return false;
}
void append_trimmed(std::string *dst, const uint8_t *s, size_t n) {  // dst += strings.TrimSpace(line)
  size_t b = 0, w = 0;
  while (b < n && is_space_at(s, n, b, &w)) b += w;
  size_t e = n;
  for (;;) {
    if (e > b && (s[e - 1] == ' ' || (s[e - 1] >= '\t' && s[e - 1] <= '\r'))) {
      --e;
    } else if (e >= b + 2 && s[e - 2] == 0xC2 && (s[e - 1] == 0x85 || s[e - 1] == 0xA0)) {
      e -= 2;
    } else {
      break;
    }
  }
  dst->append((const char *)s + b, e - b);
}

struct Record {
  std::string name, seq;
  int32_t size_in_kmer;
};

int finish_batch(std::vector<Record> &recs, int pinned, kaamer_query_batch **out) {
  auto *b = new kaamer_query_batch();
  memset(b, 0, sizeof *b);
  auto *own = new BatchOwner();
  own->pinned = pinned != 0;
  b->_owner = own;
  size_t n_res = 0, n_name = 0;
  for (auto &r : recs) {
    n_res += r.seq.size();
    n_name += r.name.size();
  }
  int rc = own->alloc(&b->residues, n_res + 16);
  if (rc == KAAMER_OK) rc = own->alloc(&b->seq_off, recs.size() + 1);
  if (rc == KAAMER_OK) rc = own->alloc(&b->names, n_name + 1);
  if (rc == KAAMER_OK) rc = own->alloc(&b->name_off, recs.size() + 1);
  if (rc == KAAMER_OK) rc = own->alloc(&b->size_in_kmer, recs.size());
  if (rc != KAAMER_OK) {
    delete own;
    delete b;
    return rc;
  }
  b->n_queries = (uint32_t)recs.size();
  b->n_residues = n_res;
  size_t ro = 0, no = 0;
  for (size_t i = 0; i < recs.size(); ++i) {
    b->seq_off[i] = ro;
    b->name_off[i] = no;
    memcpy(b->residues + ro, recs[i].seq.data(), recs[i].seq.size());
    memcpy(b->names + no, recs[i].name.data(), recs[i].name.size());
    ro += recs[i].seq.size();
    no += recs[i].name.size();
    b->size_in_kmer[i] = recs[i].size_in_kmer;
  }
  b->seq_off[recs.size()] = ro;
  b->name_off[recs.size()] = no;
  b->names[no] = 0;
  *out = b;
  return KAAMER_OK;
}

// shared front end: open, sniff, decompress.  *proceed = false: the reference returns without queries.
int load_input(const char *path, std::vector<uint8_t> *text, bool *proceed) {
  *proceed = false;
  std::vector<uint8_t> raw;
  KCHECK(read_file(path, &raw));
  if (raw.empty()) {
    set_error("%s is empty (the reference exits on the failed 32-byte read, search.go:248-252)", path);
    return KAAMER_ERR_FORMAT;
  }
  uint8_t sniff[32];
  memset(sniff, 0, sizeof sniff);  // buff := make([]byte, 32): short files leave zero bytes behind
  memcpy(sniff, raw.data(), raw.size() < 32 ? raw.size() : 32);
  const Sniff t = detect_content_type(sniff, 32);
  if (t == SNIFF_OTHER) return KAAMER_OK;
  if (t == SNIFF_GZIP) {
    KCHECK(gunzip(raw, text));
  } else {
    text->swap(raw);
  }
  *proceed = true;
  return KAAMER_OK;
}

}  // namespace
}  // namespace kaamer

using namespace kaamer;

extern "C" {

static int kaamer_host_read_fasta_impl(const char *path, int is_protein, int pinned, kaamer_query_batch **out) {
  (void)is_protein;  // (only Name/Contig/StartPosition bookkeeping differs, search.go:296-305)
  if (!path || !out) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  std::vector<uint8_t> text;
  bool proceed = false;
  KCHECK(load_input(path, &text, &proceed));
  std::vector<Record> recs;
  if (proceed) {
    Record cur;
    auto emit = [&](bool upper) {
      cur.size_in_kmer = (int32_t)cur.seq.size() - KAAMER_KMER_SIZE + 1;
      if (cur.seq.back() == '*') cur.size_in_kmer--;
      if (upper)
        for (auto &c : cur.seq)
          if (c >= 'a' && c <= 'z') c = (char)(c - 32);
      recs.push_back(cur);
    };
    scan_lines(text, [&](const uint8_t *l, size_t n) {
      if (n < 1) return;
      if (l[0] == '>') {
        if (!cur.seq.empty()) {
          emit(true);
          cur = Record();
        }
        cur.name.assign((const char *)l + 1, n - 1);
      } else {
        append_trimmed(&cur.seq, l, n);
      }
    });
    if (!cur.seq.empty()) emit(false);  // the last record is not upper-cased (search.go:313-320)
  }
  return finish_batch(recs, pinned, out);
}
int kaamer_host_read_fasta(const char *path, int is_protein, int pinned, kaamer_query_batch **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_host_read_fasta_impl(path, is_protein, pinned, out); });
}

static int kaamer_host_read_fastq_impl(const char *path, int pinned, kaamer_query_batch **out) {
  if (!path || !out) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  *out = nullptr;
  std::vector<uint8_t> text;
  bool proceed = false;
  KCHECK(load_input(path, &text, &proceed));
  std::vector<Record> recs;
  if (proceed) {
    Record cur;
    auto emit = [&]() {
      cur.size_in_kmer = (int32_t)cur.seq.size() - KAAMER_KMER_SIZE + 1;
      recs.push_back(cur);
    };
    scan_lines(text, [&](const uint8_t *l, size_t n) {
      if (n < 1) return;
      if (l[0] == '@') {
        if (!cur.seq.empty()) {
          emit();
          cur = Record();
        }
        cur.name.assign((const char *)l + 1, n - 1);
      } else {
        bool is_seq = true;  // ^[ATGCNatgcn]+$
        for (size_t i = 0; i < n && is_seq; ++i) {
          const uint8_t c = l[i] | 0x20;
          is_seq = (l[i] < 0x80) && (c == 'a' || c == 't' || c == 'g' || c == 'c' || c == 'n');
        }
        if (is_seq) cur.seq.assign((const char *)l, n);
      }
    });
    if (!cur.seq.empty()) emit();
  }
  return finish_batch(recs, pinned, out);
}
int kaamer_host_read_fastq(const char *path, int pinned, kaamer_query_batch **out) {
  return ::kaamer::guarded([&]() -> int { return kaamer_host_read_fastq_impl(path, pinned, out); });
}

int64_t kaamer_host_format_positions(const uint8_t *positions, uint64_t n, int with_alignment, char *out, uint64_t cap) {
  // search.go:694-742, statement by statement (the single-position branch can never be taken: at a
  // non-matching position pos the run started at currentStart <= pos, so pos+1 > currentStart)
  std::string ps;
  uint64_t current_start = 0, end_pos = 0;
  bool in_sequence = false;
  for (uint64_t pos = 0; pos < n; ++pos) {
    if (positions[pos]) {
      if (!in_sequence) {
        current_start = pos + 1;
        in_sequence = true;
      }
    } else if (in_sequence) {
      if (!ps.empty()) ps += ",";
      if (pos + 1 > current_start) {
        end_pos = pos + 1;
        if (with_alignment) end_pos = end_pos + KAAMER_KMER_SIZE - 1;
        ps += std::to_string(current_start) + "-" + std::to_string(end_pos);
      } else {
        ps += std::to_string(current_start);
      }
      in_sequence = false;
    }
  }
  if (in_sequence) {
    if (!ps.empty()) ps += ",";
    end_pos = n;
    if (with_alignment) end_pos = end_pos + KAAMER_KMER_SIZE - 1;
    ps += std::to_string(current_start) + "-" + std::to_string(end_pos);
  }
  if (!out || ps.size() + 1 > cap) return -(int64_t)(ps.size() + 1);
  memcpy(out, ps.c_str(), ps.size() + 1);
  return (int64_t)ps.size();
}

void kaamer_host_free_queries(kaamer_query_batch *b) {
  if (!b) return;
  delete (BatchOwner *)b->_owner;
  delete b;
}

}  // extern "C"
