"""A second, independent restatement of the reference's search path: a LITERAL Python transliteration
of the Go functions (strings, maps and loops as in the source, one statement per Go statement where
possible), written without looking at oracle/kaamer_oracle.cpp.  tests/test_oracle_vs_transliteration.py
runs both on the same small random inputs: the C++ oracle the GPU kernels are compared with must agree
with this reading of the Go source bit for bit.  Test infrastructure only (pure-Python loops: small cases).

Every function cites the Go code it follows (paths relative to the reference repository).
"""
from __future__ import annotations

import json
import os

KMER_SIZE = 7  # pkg/search/search.go:45, pkg/makedb/makedb.go

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- pkg/kvstore/k_store.go ---------------------------------------------------------------------
def new_aa_table() -> dict:
    """NewAATable, k_store.go:39-64: single codes 0..20 under (a, '.'), pair codes from 22."""
    aa = list("ACDEFGHIKLMNPQRSTUVWY")
    table = {}
    i = 22
    for j, a in enumerate(aa):
        table[(a, ".")] = j
        for b in aa:
            table[(a, b)] = i
            i += 1
    return table


_AA_TABLE = new_aa_table()


def encode_kmer(kmer: str) -> int:
    """EncodeKmer, k_store.go:91-117.  A Go map lookup of a missing key yields 0."""
    kmer_int = 0
    i = 0
    shift_index = 1
    while (i + 2) < len(kmer):
        key = (kmer[i], kmer[i + 1])
        kmer_int |= (_AA_TABLE.get(key, 0) << (32 - (shift_index * 9))) & 0xFFFFFFFF
        shift_index += 1
        i += 2
    key = (kmer[len(kmer) - 1], ".")
    kmer_int |= _AA_TABLE.get(key, 0)
    return kmer_int


# ---- pkg/makedb/inputFASTA.go + pkg/indexdb + pkg/kvstore/kcomb_store.go ------------------------------
def make_index(fasta_text: str) -> dict:
    """kmer key -> posting list.  Reader loop inputFASTA.go:96-124 (the id quirk: proteinNb is incremented
    before the previous record is dispatched), record processing :195-250, then the index step: per key the
    set of ids sorted descending without duplicates (indexdb.go:92-128, kv_store.go:284-305)."""
    jobs = []
    protein_nb = 0
    protein_entry = ""
    for line in fasta_text.split("\n"):
        if line == "":  # bufio.Scanner yields no token after the final newline
            continue
        if line[0:1] == ">":
            protein_nb += 1
            if protein_entry != "":
                jobs.append((protein_nb, protein_entry))
                protein_entry = ""
        protein_entry += line
        protein_entry += "\n"
    if protein_entry != "":
        jobs.append((protein_nb, protein_entry))

    kmer_store: dict[int, list] = {}
    stats = {"proteins": 0, "aa": 0}
    for protein_id, text_entry in jobs:  # processProteinInputFASTA
        sequence = ""
        protein_name = ""
        for l in text_entry.split("\n"):
            if len(l) < 1:
                continue
            if l[0:1] == ">":
                header_split = l.split(" ")
                protein_name = " ".join(header_split[1:])
            else:
                sequence += l.upper()
        if ", partial" in protein_name:
            continue
        length = len(sequence)
        if length < KMER_SIZE:
            continue
        stats["proteins"] += 1
        stats["aa"] += length
        for i in range(0, length - KMER_SIZE + 1):
            kmer_store.setdefault(encode_kmer(sequence[i:i + KMER_SIZE]), []).append(protein_id)
    index = {}
    for key, ids in kmer_store.items():
        s = sorted(ids, reverse=True)
        out = []
        for e in s:  # RemoveDuplicatesFromSlice on the sorted slice
            if not out or out[-1] != e:
                out.append(e)
        index[key] = out
    index["__stats__"] = stats
    return index


# ---- pkg/search/search.go -----------------------------------------------------------------------
def size_in_kmer(sequence: str) -> int:
    """search.go:290-293"""
    n = len(sequence) - KMER_SIZE + 1
    if len(sequence) > 0 and sequence[len(sequence) - 1:] == "*":
        n -= 1
    return n


def kmer_search(index: dict, sequence: str, q_size: int, extract_pos: bool):
    """key producer loop (search_protein.go:94-98) + KmerSearch (search.go:414-440) +
    StoreMatchPositions (:442-452)."""
    counter: dict[int, int] = {}
    position_hits: dict[int, list] = {}
    for k in range(0, q_size):
        key = encode_kmer(sequence[k:k + KMER_SIZE])
        ids = index.get(key)
        if ids is None or len(ids) < 1:
            continue
        for pid in ids:
            counter[pid] = counter.get(pid, 0) + 1
            if extract_pos:
                if pid not in position_hits:
                    position_hits[pid] = [False] * q_size
                position_hits[pid][k] = True
    return counter, position_hits


def sort_map_by_value(counter: dict) -> list:
    """sortMapByValue, search.go:132-152: descending by Kmatch; the reference's order among equal Kmatch
    is that of a Go map iteration (random) — the canonical representative used everywhere in this
    repository is subject id ascending."""
    return sorted(counter.items(), key=lambda kv: (-kv[1], kv[0]))


def filter_results(hits: list, q_size: int, position_hits: dict, min_kmatch: int, min_kratio: float, max_results: int):
    """FilterResults, search.go:189-220, statement by statement."""
    hits_to_delete = []
    last_good = len(hits) - 1
    for i, (key, kmatch) in enumerate(hits):
        if (float(kmatch) / float(q_size)) < min_kratio or kmatch < min_kmatch:
            if last_good == (len(hits) - 1):
                last_good = i - 1
            hits_to_delete.append(key)
    if last_good >= max_results:
        last_good = max_results - 1
        for key, _ in hits[last_good + 1:]:
            hits_to_delete.append(key)
    hits = [] if last_good < 0 else hits[0:last_good + 1]
    for k in hits_to_delete:
        position_hits.pop(k, None)
    return hits


# ---- pkg/search/dna.go + gcode.go -----------------------------------------------------------------
_GCODE = json.load(open(os.path.join(_GOLDEN, "gcode11.json")))  # gcodeBacteria, extracted from gcode.go:36-101
_FRAME_START = {0: 0, 1: 1, 2: 2, 3: 0, 4: 1, 5: 2}
MIN_LEN_CDS = 21


def reverse_complement(dna: str) -> str:
    """dna.go:55-63: only a<->t and g<->c are swapped."""
    r = {"a": "t", "t": "a", "g": "c", "c": "g"}
    return "".join(r.get(c, c) for c in reversed(dna.lower()))


def get_frame(frame_number: int, dna: str) -> str:
    """dna.go:183-196"""
    if frame_number < 0:
        dna = reverse_complement(dna)
        frame_number = -frame_number
    start_pos = frame_number - 1
    len_frame = len(dna) - start_pos
    end_pos = len(dna) - (len_frame % 3)
    return dna[start_pos:end_pos]


def get_orfs(dna: str) -> list:
    """GetORFs, dna.go:65-181.  Each ORF: dict(seq, start, end, plus, alts).  sort.Slice is unstable in
    Go; ties keep emission order here (as everywhere in this repository)."""
    orfs = []
    dna = dna.lower()
    frames = [get_frame(1, dna), get_frame(2, dna), get_frame(3, dna), get_frame(-1, dna), get_frame(-2, dna),
              get_frame(-3, dna)]
    for frame_pos, frame_seq in enumerate(frames):
        start_pos = _FRAME_START[frame_pos]
        plus = frame_pos <= 2
        abs_pos = frame_pos
        if not plus:
            abs_pos = len(dna) - start_pos - 1
        current_pos = 0
        orf = {"seq": "", "start": abs_pos + 1, "end": 0, "plus": plus, "alts": []}
        inside = True
        cds = ""
        current_aa_pos = 0
        i = 0
        while i < len(frame_seq) - (len(frame_seq) % 3):
            current_pos = i
            aa = _GCODE.get(frame_seq[i:i + 3], {"aa": "", "start": False, "stop": False})
            if aa["start"]:
                if not inside:
                    inside = True
                    current_aa_pos = 0
                    orf["start"] = frame_pos + i + 1
                    if not plus:
                        orf["start"] = len(dna) - (frame_pos + i) + 3
                    orf["alts"] = orf["alts"] + [current_aa_pos]
                else:
                    orf["alts"] = orf["alts"] + [current_aa_pos]
            if inside:
                cds += aa["aa"]
            if aa["stop"]:
                if inside and len(cds) >= MIN_LEN_CDS:
                    end_pos = i + 3 + frame_pos
                    if not plus:
                        end_pos = orf["start"] - (len(cds) * 3) + 1
                    orf["end"] = end_pos
                    orf["seq"] = cds
                    orfs.append(dict(orf))
                orf = {"seq": "", "start": 0, "end": 0, "plus": plus, "alts": []}
                cds = ""
                inside = False
            current_aa_pos += 1
            i += 3
        if inside and len(cds) >= MIN_LEN_CDS:
            end_pos = current_pos + 3 + frame_pos
            if not plus:
                end_pos = orf["start"] - (len(cds) * 3) + 1
            orf["end"] = end_pos
            orf["seq"] = cds
            orfs.append(dict(orf))
    orfs.sort(key=lambda o: o["end"] if o["plus"] else o["start"])
    return orfs


def set_best_start_codon(q: dict, hits: list, position_hits: dict) -> None:
    """SetBestStartCodon, dna.go:198-272; q: dict(seq, size, start, plus, alts), modified in place."""
    best_hits = []
    best_score = 0
    for key, kmatch in hits:
        if kmatch >= best_score:
            best_score = kmatch
            best_hits.append(key)
    if len(q["alts"]) < 1:
        return
    best_start = q["alts"][0]
    first_start = q["alts"][0]
    first_best_hit_pos = 999999999
    exit_ = False
    for key in best_hits:
        for i, is_match in enumerate(position_hits[key]):
            if is_match:
                if i < first_best_hit_pos:
                    first_best_hit_pos = i
                exit_ = True
            if exit_:
                break
    exit_ = False
    for s in q["alts"]:
        if s <= first_best_hit_pos:
            best_start = s
        else:
            exit_ = True
        if exit_:
            break
    if best_start != first_start:
        if q["plus"]:
            q["start"] = q["start"] + 3 * best_start
        else:
            q["start"] = q["start"] - 3 * best_start
        q["seq"] = q["seq"][best_start:]
        for k in list(position_hits.keys()):
            position_hits[k] = position_hits[k][best_start:]
        q["size"] = len(q["seq"]) - KMER_SIZE + 1
        if q["seq"][len(q["seq"]) - 1:] == "*":
            q["size"] = q["size"] - 1
    q["alts"] = []


# ---- drivers --------------------------------------------------------------------------------------
def protein_search(index: dict, queries: list, min_kmatch=10, min_kratio=0.05, max_results=10, extract_positions=False):
    """Per-query body of ProteinSearch (search_protein.go:70-114).  Queries with SizeInKmer < 7 are skipped
    (the reference worker returns; documented deviation).  Returns one entry per query:
    (SizeInKmer, [(subject, Kmatch)], {subject: positions})."""
    out = []
    for seq in queries:
        size = size_in_kmer(seq)
        if size < 7:
            out.append((size, [], {}))
            continue
        counter, pos = kmer_search(index, seq, size, extract_positions)
        hits = sort_map_by_value(counter)
        hits = filter_results(hits, size, pos, min_kmatch, min_kratio, max_results)
        out.append((size, hits, pos))
    return out


def nucleotide_search(index: dict, contigs: list, min_kmatch=10, min_kratio=0.05, max_results=10):
    """GetORFs + per-ORF body of NucleotideSearch (search_nucleotide.go:76-124).  Returns the surviving rows
    in (contig, GetORFs order): dict(contig, seq, size, start, end, plus, hits, pos)."""
    rows = []
    for c, dna in enumerate(contigs):
        for o in get_orfs(dna):
            q = {"seq": o["seq"], "size": len(o["seq"]) - KMER_SIZE + 1, "start": o["start"], "end": o["end"],
                 "plus": o["plus"], "alts": list(o["alts"])}
            if q["seq"][len(q["seq"]) - 1:] == "*":
                q["size"] = q["size"] - 1
            counter, pos = kmer_search(index, q["seq"], q["size"], True)
            hits = sort_map_by_value(counter)
            if len(hits) > 0 and hits[0][1] >= min_kmatch:
                set_best_start_codon(q, hits, pos)
                hits = filter_results(hits, q["size"], pos, min_kmatch, min_kratio, max_results)
                if len(hits) > 0:
                    rows.append({"contig": c, "seq": q["seq"], "size": q["size"], "start": q["start"], "end": q["end"],
                                 "plus": q["plus"], "hits": hits, "pos": {k: pos[k] for k, _ in hits}})
    return rows


# ---- pkg/align/align.go (kaamer's own post-processing of the biogo alignment) -----------------------------
AA_POS_IN_MATRIX = {c: i for i, c in enumerate("-ABCDEFGHIJKLMNPQRSTVWXYZ*")}  # matrixScores.go:107


def segments_from_strings(a: str, b: str, q0: int, s0: int, matrix) -> list:
    """The feat.Pair list biogo's SWAffine returns, rebuilt from the two gapped strings: maximal runs of
    aligned columns (score = sum of matrix cells) and maximal gap runs (score = GapOpen = -11 plus the
    matrix's gap row, which is 0).  Each segment: (score, q_start, q_end, s_start, s_end), 0-based half-open
    as Features()[k].Start()/End()."""
    segs = []
    qi, si = q0, s0
    i = 0
    while i < len(a):
        kind = 0 if (a[i] != "-" and b[i] != "-") else (1 if a[i] == "-" else 2)
        j = i
        score = 0
        qs, ss = qi, si
        while j < len(a):
            k = 0 if (a[j] != "-" and b[j] != "-") else (1 if a[j] == "-" else 2)
            if k != kind:
                break
            if kind == 0:
                score += int(matrix[AA_POS_IN_MATRIX[a[j]]][AA_POS_IN_MATRIX[b[j]]])
                qi += 1
                si += 1
            elif kind == 1:
                si += 1
            else:
                qi += 1
            j += 1
        if kind != 0:
            score = -11
        segs.append((score, qs, qi, ss, si))
        i = j
    return segs


def align_postprocess(a_string: str, b_string: str, segs: list, query_len: int, number_of_aa: int, matrix,
                      lambda_=0.267, K=0.041, gap_open=11, gap_extend=1) -> dict:
    """align.go:72-157, statement by statement (float32 for identity / similarity, float64 for scores)."""
    import math

    import numpy as np

    f32 = np.float32
    identity = f32(0)
    similarity = f32(0)
    nb_pos = f32(0)
    mismatches = 0
    for i, a in enumerate(a_string):
        if b_string[i] == a:
            identity = f32(identity + f32(1))
            similarity = f32(similarity + f32(1))
        else:
            if b_string[i] != "-" and a != "-":
                mismatches += 1
            if int(matrix[AA_POS_IN_MATRIX[b_string[i]]][AA_POS_IN_MATRIX[a]]) > 0:
                similarity = f32(similarity + f32(1))
        nb_pos = f32(nb_pos + f32(1))
    identity = f32(f32(identity / nb_pos) * f32(100))
    similarity = f32(f32(similarity / nb_pos) * f32(100))
    raw = 0
    gap_openings = 0
    q_start = q_end = s_start = s_end = 0
    for i, (score, qs, qe, ss, se) in enumerate(segs):
        if i == 0:
            q_start, s_start = qs, ss
        if i == len(segs) - 1:
            q_end, s_end = qe, se
        raw += score
        if score == -gap_open:
            gap_openings += 1
            gap_len = max(qe - qs, se - ss)
            raw = raw - ((gap_len - 1) * gap_extend)
    bitscore = ((lambda_ * float(raw)) - math.log(K)) / math.log(2)
    evalue = float(query_len) * float(number_of_aa) / math.pow(2, bitscore)
    return {"identity": float(identity), "similarity": float(similarity), "length": len(a_string),
            "mismatches": mismatches, "gap_openings": gap_openings, "raw": raw, "bitscore": bitscore, "evalue": evalue,
            "query_start": q_start + 1, "query_end": q_end, "subject_start": s_start + 1, "subject_end": s_end}


# ---- pkg/search/search.go:222-412 (query readers) ---------------------------------------------------------
def detect_content_type(buf: bytes) -> str:
    """net/http.DetectContentType on the 32-byte sniff buffer, reduced to the three outcomes the readers
    distinguish: 'gzip', 'text' ("text/plain; charset=utf-8") or 'other'."""
    ws = 0
    while ws < len(buf) and buf[ws] in b"\t\n\x0c\r ":
        ws += 1
    d = buf[ws:]
    for sig in (b"<!DOCTYPE HTML", b"<HTML", b"<HEAD", b"<SCRIPT", b"<IFRAME", b"<H1", b"<DIV", b"<FONT", b"<TABLE", b"<A",
                b"<STYLE", b"<TITLE", b"<B", b"<BODY", b"<BR", b"<P", b"<!--"):
        if len(d) >= len(sig) + 1 and d[:len(sig)].upper() == sig and d[len(sig):len(sig) + 1] in (b" ", b">"):
            return "other"
    if d.startswith(b"<?xml") or buf.startswith(b"%PDF-") or buf.startswith(b"%!PS-Adobe-"):
        return "other"
    if buf.startswith(b"\xfe\xff") or buf.startswith(b"\xff\xfe"):
        return "other"
    if buf.startswith(b"\xef\xbb\xbf"):
        return "text"
    for sig in (b"GIF87a", b"GIF89a", b"BM", b"ID3", b".snd", b"wOFF", b"wOF2", b"OTTO", b"ttcf", b"\xff\xd8\xff"):
        if buf.startswith(sig):
            return "other"
    if buf.startswith(b"RIFF") and (buf[8:14] == b"WEBPVP" or buf[8:12] in (b"AVI ", b"WAVE")):
        return "other"
    if buf.startswith(b"FORM") and buf[8:12] == b"AIFF":
        return "other"
    if buf.startswith(b"\x1f\x8b\x08"):
        return "gzip"
    for c in d:
        if c <= 0x08 or c == 0x0B or 0x0E <= c <= 0x1A or 0x1C <= c <= 0x1F:
            return "other"
    return "text"


def _scan_lines(data: bytes):
    """bufio.Scanner / ScanLines with scanner.Buffer(buf, 1024*1024): a line of 1 MiB or more ends the scan."""
    p = 0
    while p < len(data):
        e = data.find(b"\n", p)
        end = len(data) if e < 0 else e
        if end - p >= 1024 * 1024:
            return
        line = data[p:end]
        if line.endswith(b"\r"):
            line = line[:-1]
        yield line
        p = len(data) if e < 0 else e + 1


def _open_queries(path: str):
    import gzip

    raw = open(path, "rb").read()
    buff = (raw[:32] + bytes(32))[:32]  # buff := make([]byte, 32); file.Read(buff)
    kind = detect_content_type(buff)
    if kind == "gzip":
        return gzip.decompress(raw)
    if kind == "text":
        return raw
    return None


# unicode.IsSpace (Go): what strings.TrimSpace removes
_GO_SPACE = "\t\n\v\f\r \u0085\u00a0\u1680\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008\u2009\u200a\u2028\u2029\u202f\u205f\u3000"


def _go_len(s: str) -> int:
    """len() of a Go string = its UTF-8 byte length"""
    return len(s.encode("utf-8", "surrogateescape"))


def get_queries_fasta(path: str) -> list:
    """GetQueriesFasta, search.go:222-322 -> [(Name, Sequence, SizeInKmer)].  Go strings are UTF-8 byte
    strings: lines are decoded as UTF-8 here (invalid bytes preserved) so that TrimSpace / ToUpper see the
    same characters; the returned strings are latin-1 views of the bytes."""
    data = _open_queries(path)
    if data is None:
        return []
    out = []
    name, seq = "", ""

    def emit(upper):
        size = _go_len(seq) - KMER_SIZE + 1
        if seq[len(seq) - 1:] == "*":
            size -= 1
        s2 = "".join(c.upper() if c.isascii() else c for c in seq) if upper else seq
        out.append((name.encode("utf-8", "surrogateescape").decode("latin-1"),
                    s2.encode("utf-8", "surrogateescape").decode("latin-1"), size))

    for raw in _scan_lines(data):
        if len(raw) < 1:
            continue
        l = raw.decode("utf-8", "surrogateescape")
        if raw[0:1] == b">":
            if seq != "":
                emit(True)
                name, seq = "", ""
            name = l[1:]
        else:
            seq += l.strip(_GO_SPACE)
    if seq != "":
        emit(False)
    return out


def get_queries_fastq(path: str) -> list:
    """GetQueriesFastq, search.go:324-412 -> [(Name, Sequence, SizeInKmer)]"""
    import re

    data = _open_queries(path)
    if data is None:
        return []
    is_sequence = re.compile(r"^[ATGCNatgcn]+$").match
    out = []
    name, seq = "", ""
    for raw in _scan_lines(data):
        l = raw.decode("latin-1")
        if len(l) < 1:
            continue
        if l[0] == "@":
            if seq != "":
                out.append((name, seq, len(seq) - KMER_SIZE + 1))
                name, seq = "", ""
            name = l[1:]
        elif is_sequence(l):
            seq = l
    if seq != "":
        out.append((name, seq, len(seq) - KMER_SIZE + 1))
    return out


# ---- pkg/makedb/inputTSV.go:94-142, 221-239 ---------------------------------------------------------------
def make_index_tsv(tsv_text: str) -> dict:
    """kmer key -> posting list for a TSV database: header line, then rows; accepted rows are numbered from
    0 in file order; sequences are indexed as they are (no upper-casing)."""
    features = None
    protein_nb = 0
    kmer_store: dict[int, list] = {}
    stats = {"proteins": 0, "aa": 0, "entries": []}
    for line in tsv_text.split("\n"):
        if line == "" and features is not None:
            # bufio.Scanner yields no token after the final newline; an empty row has no EntryID and is skipped
            continue
        if features is None:
            features = line.split("\t")
            continue
        cols = line.split("\t")
        entry, sequence, length = "", "", 0
        for i, f in enumerate(cols):
            if features[i].lower() == "entryid":
                entry = f
            elif features[i].lower() == "sequence":
                sequence = f
                length = len(f)
        if length < KMER_SIZE or sequence == "" or entry == "":
            continue
        stats["proteins"] += 1
        stats["aa"] += length
        stats["entries"].append(entry)
        for i in range(0, length - KMER_SIZE + 1):
            kmer_store.setdefault(encode_kmer(sequence[i:i + KMER_SIZE]), []).append(protein_nb)
        protein_nb += 1
    index = {}
    for key, ids in kmer_store.items():
        s = sorted(ids, reverse=True)
        out = []
        for e in s:
            if not out or out[-1] != e:
                out.append(e)
        index[key] = out
    index["__stats__"] = stats
    return index


def format_positions_to_string(positions: list, with_alignment: bool) -> str:
    """FormatPositionsToString, search.go:694-742"""
    current_start = 0
    in_sequence = False
    end_pos = 0
    ps = ""
    for pos, match in enumerate(positions):
        if match:
            if not in_sequence:
                current_start = pos + 1
                in_sequence = True
        else:
            if in_sequence:
                if pos + 1 > current_start:
                    if ps != "":
                        ps += ","
                    end_pos = pos + 1
                    if with_alignment:
                        end_pos = end_pos + KMER_SIZE - 1
                    ps += str(current_start) + "-" + str(end_pos)
                    in_sequence = False
                else:
                    if ps != "":
                        ps += ","
                    ps += str(current_start)
                    in_sequence = False
    if in_sequence:
        if ps != "":
            ps += ","
        end_pos = len(positions)
        if with_alignment:
            end_pos = end_pos + KMER_SIZE - 1
        ps += str(current_start) + "-" + str(end_pos)
    return ps


def tsv_rows(query_name: str, size_in_kmer: int, loc_start: int, loc_end: int, hits: list, entries: dict,
             position_hits: dict, align: bool, extract_positions: bool, annotations: bool, is_protein: bool) -> str:
    """QueryResultHandler, search.go:505-606: the TSV rows of ONE query.  hits: [{"Key", "Kmatch", "Alignment"}]
    (Alignment: dict with the AlignmentResult fields, float32 Identity as numpy.float32); entries: id ->
    {"EntryId", "Length"}; no feature columns (dbStats.Features empty)."""
    import numpy as np

    def go_f2(v):  # fmt.Sprintf("%.2f", v)
        v = float(v)
        if v != v:
            return "NaN"
        if v in (float("inf"), float("-inf")):
            return "+Inf" if v > 0 else "-Inf"
        return "%.2f" % v

    def go_e(v):  # fmt.Sprintf("%e", v)
        v = float(v)
        if v != v:
            return "NaN"
        if v in (float("inf"), float("-inf")):
            return "+Inf" if v > 0 else "-Inf"
        return "%e" % v

    out = ""
    if align:
        # sort.Slice by BitScore descending (search.go:491-493); a stable sort is one of its possible outcomes
        hits = sorted(hits, key=lambda h: -h["Alignment"]["BitScore"] if h["Alignment"]["BitScore"] == h["Alignment"]["BitScore"] else 0.0)
    for h in hits:
        output = ""
        output += query_name.split(" ")[0]
        output += "\t"
        output += entries[h["Key"]]["EntryId"]
        output += "\t"
        if not align:
            pos_string = ""
            output += go_f2(np.float32(h["Kmatch"]) / np.float32(size_in_kmer) * np.float32(100.00))
            output += "\t"
            output += str(size_in_kmer)
            output += "\t"
            output += str(int(h["Kmatch"]))
            output += "\t"
            if extract_positions:
                pos_string = format_positions_to_string(position_hits[h["Key"]], False)
                output += "%d" % pos_string.count(",")
            else:
                output += "N/A"
            output += "\t"
            output += str(loc_start)
            output += "\t"
            output += str(loc_end)
            output += "\t"
            output += "1"
            output += "\t"
            if annotations:
                output += "%d" % entries[h["Key"]]["Length"]
            else:
                output += "N/A"
            if extract_positions:
                output += "\t"
                output += pos_string
        else:
            a = h["Alignment"]
            output += go_f2(a["Identity"])
            output += "\t"
            output += "%d" % a["Length"]
            output += "\t"
            output += "%d" % a["Mismatches"]
            output += "\t"
            output += "%d" % a["GapOpenings"]
            output += "\t"
            if not is_protein:
                output += str(loc_start)
                output += "\t"
                output += str(loc_end)
                output += "\t"
            else:
                output += "%d" % a["QueryStart"]
                output += "\t"
                output += "%d" % a["QueryEnd"]
                output += "\t"
            output += "%d" % a["SubjectStart"]
            output += "\t"
            output += "%d" % a["SubjectEnd"]
            output += "\t"
            output += go_e(a["EValue"])
            output += "\t"
            output += go_f2(a["BitScore"])
            if extract_positions:
                output += "\t"
                output += format_positions_to_string(position_hits[h["Key"]], True)
        output += "\n"
        out += output
    return out
