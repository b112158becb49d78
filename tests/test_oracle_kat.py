"""CPU tests: the oracle against every golden vector / fixture we hold for the path
(SURVEY.md §8c — the reference ships no tests; vectors are parsed from / derived by hand
from its source, see tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as o


@pytest.fixture(scope="module")
def kat(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kat.json")))


def test_alphabet_matches_reference_source(golden_dir):
    alpha = json.load(open(os.path.join(golden_dir, "aa_alphabet.json")))["alphabet"]
    assert alpha == "ACDEFGHIKLMNPQRSTUVWY"  # k_store.go:41
    for j, a in enumerate(alpha):
        # single code in the low 5 bits; pair code (a,a) = 22 + 21 j + j in the top 9 bits
        key = o.encode_kmer((a * 7).encode())
        assert key & 0x1F == j
        assert key >> 23 == 22 + 22 * j


def test_encode_known_answers(kat):
    for kmer, expect in kat["encode"].items():
        assert o.encode_kmer(kmer.encode()) == expect, kmer


def test_encode_closed_form_and_decode():
    alpha = "ACDEFGHIKLMNPQRSTUVWY"
    rng = np.random.default_rng(7)
    for _ in range(500):
        idx = rng.integers(0, 21, 7)
        kmer = "".join(alpha[i] for i in idx)
        a, b, c, d, e, f, g = (int(x) for x in idx)
        expect = ((22 + 21 * a + b) << 23) | ((22 + 21 * c + d) << 14) | ((22 + 21 * e + f) << 5) | g
        assert o.encode_kmer(kmer.encode()) == expect
        assert o.decode_kmer(expect) == kmer.encode()


def test_unknown_residue_quirks():
    # Go map miss -> 0: pair with an unknown letter -> 0; unknown last letter aliases 'A'
    assert o.encode_kmer(b"AAAAAAX") == o.encode_kmer(b"AAAAAAA")
    assert o.encode_kmer(b"AAAAAA*") == o.encode_kmer(b"AAAAAAA")
    assert o.encode_kmer(b"aAAAAAA") == o.encode_kmer(b"XAAAAAA") == 0x000582C0
    assert o.encode_kmer(b"BZJOX*a") == 0


def test_size_in_kmer(kat):
    for c in kat["size_in_kmer"]:
        seq = b"A" * (c["len"] - 1) + (b"*" if c["star"] else b"A")
        assert o.size_in_kmer(seq) == c["expect"]


def test_fasta_id_quirk(kat):
    assert o.fasta_ids(3).tolist() == kat["fasta_ids_3"]
    assert o.fasta_ids(1).tolist() == [1]
    ids = o.fasta_ids(10)
    assert ids.tolist() == [2, 3, 4, 5, 6, 7, 8, 9, 10, 10]  # id 1 unused, last two share N


def test_orfs_known_answer(kat):
    r = o.get_orfs(kat["orfs_dna"].encode())
    assert len(r.seqs) == len(kat["orfs"])
    for i, e in enumerate(kat["orfs"]):
        assert r.seqs[i].decode() == e["seq"]
        assert bool(r.plus[i]) == e["plus"]
        assert (int(r.start[i]), int(r.end[i])) == (e["start"], e["end"])
        assert r.alts[i] == e["alts"]
        if "size_in_kmer" in e:
            assert o.size_in_kmer(r.seqs[i]) == e["size_in_kmer"]
    # case-insensitive (GetORFs lower-cases, dna.go:68)
    r2 = o.get_orfs(kat["orfs_dna"].upper().encode())
    assert r2.seqs == r.seqs


def test_gcode_table_11_matches_reference_source(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "gcode11.json")))
    assert len(g) == 64
    for codon, e in g.items():
        # one codon repeated 21 times in frame +1 -> reveals AA / start / stop of the codon
        r = o.get_orfs((codon * 21).encode())
        plus1 = [i for i in range(len(r.seqs)) if r.plus[i] and r.start[i] in (1,)]
        if e["stop"]:
            # every codon is a stop: cds "*" has length 1 < 21 -> no ORF in frame +1
            assert not any(s == b"*" * 21 for s in r.seqs)
            continue
        assert plus1, codon
        i = plus1[0]
        assert r.seqs[i] == (e["aa"] * 21).encode(), codon
        assert r.alts[i] == (list(range(21)) if e["start"] else []), codon


def test_orf_unknown_codon_contributes_no_residue():
    # 'n' codons add nothing to the cds but advance the position (dna.go:106,123-125)
    dna = b"atg" + b"gct" * 10 + b"nnn" + b"gct" * 10 + b"taa"
    r = o.get_orfs(dna)
    i = [k for k in range(len(r.seqs)) if r.plus[k] and r.start[k] == 1][0]
    assert r.seqs[i] == b"M" + b"A" * 20 + b"*"
    assert int(r.end[i]) == len(dna)


def test_filter_known_answers(kat):
    for c in kat["filter"]:
        assert o.filter_count(c["kmatch"], c["size"], o.opts(max_results=c["max_results"])) == c["keep"]
    assert o.filter_count([], 100) == 0
    assert o.filter_count([9, 9], 100) == 0
    assert o.filter_count([50] * 30, 100) == 10


def test_scores_known_answers(kat):
    s = kat["scores"]
    assert o.bitscore(50, s["lambda"], s["K"]) == pytest.approx(s["raw50_bits"], rel=1e-12)
    assert o.bitscore(100, s["lambda"], s["K"]) == pytest.approx(s["raw100_bits"], rel=1e-12)
    ev = o.evalue(350, 3500000, o.bitscore(100, s["lambda"], s["K"]))
    assert ev == pytest.approx(s["evalue_raw100_q350_n3500000"], rel=1e-9)


def test_matrix_params_match_reference_source(golden_dir):
    m = json.load(open(os.path.join(golden_dir, "matrix_scores.json")))
    assert m["aa_pos_order"] == "-ABCDEFGHIJKLMNPQRSTVWXYZ*"
    p = m["params"]["blosum62_11_1"]
    assert (p["lambda"], p["K"], p["gap_open"], p["gap_extend"]) == (0.267, 0.041, 11, 1)
    prm = o.aln_params(1000)
    assert (prm.lambda_, prm.K, prm.gap_open_opt, prm.gap_extend_opt) == (0.267, 0.041, 11, 1)


def test_blosum62_properties():
    b = o.blosum62()
    order = "-ABCDEFGHIJKLMNPQRSTVWXYZ*"
    assert (b == b.T).all()
    assert (b[0] == 0).all()
    ix = {c: i for i, c in enumerate(order)}
    assert b[ix["W"], ix["W"]] == 11 and b[ix["C"], ix["C"]] == 9 and b[ix["A"], ix["A"]] == 4
    assert b[ix["A"], ix["R"]] == -1 and b[ix["*"], ix["*"]] == 1 and b[ix["A"], ix["*"]] == -4
    assert b[ix["I"], ix["L"]] == 2 and b[ix["D"], ix["E"]] == 2 and b[ix["W"], ix["F"]] == 1
    assert int(np.trace(b[1:25, 1:25])) == 4 + 4 + 9 + 6 + 5 + 6 + 6 + 8 + 4 + 3 + 5 + 4 + 5 + 6 + 7 + 5 + 5 + 4 + 5 + 4 + 11 - 1 + 7 + 4


def test_align_identical_and_gap():
    prm = o.aln_params(3500000)
    q = b"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQFEVVHSLAKWKR"
    r = o.align(q, q, prm)
    b = o.blosum62()
    order = "-ABCDEFGHIJKLMNPQRSTVWXYZ*"
    self_score = sum(int(b[order.index(chr(c)), order.index(chr(c))]) for c in q)
    assert r.raw == r.dp_score == self_score
    assert r.identity == 100.0 and r.mismatches == 0 and r.gap_openings == 0
    assert (r.query_start, r.query_end, r.subject_start, r.subject_end) == (1, len(q), 1, len(q))
    assert r.length == len(q)
    assert r.bitscore == pytest.approx(o.bitscore(self_score), rel=1e-12)
    # one deletion in the subject: a single gap segment of score -11 (free extension in the DP,
    # kaamer re-adds (gapLen-1)*GapExtend afterwards, align.go:127-131)
    s = q[:30] + q[36:]
    r2, a, bb = o.align(q, s, prm, want_strings=True)
    assert r2.gap_openings == 1
    assert r2.dp_score == self_score - sum(int(b[order.index(chr(c)), order.index(chr(c))]) for c in q[30:36]) - 11
    assert r2.raw == r2.dp_score - 5 * 1
    assert bb.count(b"-") == 6 and a.count(b"-") == 0 and r2.length == len(q)


def test_align_illegal_letter_and_selenocysteine():
    prm = o.aln_params(1000)
    r = o.align(b"MKTAYIAKQRO", b"MKTAYIAKQR", prm)  # 'O' is not in alphabet.Protein -> biogo error, ignored
    assert r.illegal == 1 and r.length == 0 and r.raw == 0 and np.isnan(r.identity)
    assert (r.query_start, r.query_end) == (1, 0)
    r = o.align(b"MKTAYIAKQRU", b"MKTAYIAKQRU", prm)  # U -> '*' (align.go:54-55); '*'/'*' scores 1
    assert r.illegal == 0 and r.length == 11


def test_format_positions():
    # search.go:694-742, literal: a single matching position p (0-based) prints "p+1-p+2"
    assert o.format_positions([0, 0, 1, 1, 1, 0, 0]) == "3-6"
    assert o.format_positions([1, 1, 1]) == "1-3"
    assert o.format_positions([1, 0, 1, 1]) == "1-2,3-4"
    assert o.format_positions([0, 1, 1, 0], with_alignment=True) == "2-10"
    assert o.format_positions([]) == ""


def test_index_semantics_and_search_small():
    # three records, FASTA ids 2,3,3 -> the last two proteins share id 3 (their k-mer sets union)
    seqs = [b"MKTAYIAKQRQISFVKSHFSRQ", b"MKTAYIAKQRQISFVKAAAAAA", b"GGGGGGGGGGMKTAYIAKQ"]
    res, off = o.pack(seqs)
    ids = o.fasta_ids(3)
    idx = o.Index.build(res, off, ids)
    assert (idx.n_proteins, idx.n_aa, idx.n_kmers) == (3, 22 + 22 + 19, 16 + 16 + 13)
    assert (np.diff(idx.keys.astype(np.int64)) > 0).all()
    k = o.encode_kmer(b"MKTAYIA")
    i = int(np.searchsorted(idx.keys, k))
    assert idx.keys[i] == k
    assert idx.postings[int(idx.offsets[i]):int(idx.offsets[i + 1])].tolist() == [3, 2]  # descending, unique
    # GGGGGGG occurs 4x in record 3 -> one posting (set semantics)
    g = o.encode_kmer(b"GGGGGGG")
    i = int(np.searchsorted(idx.keys, g))
    assert idx.postings[int(idx.offsets[i]):int(idx.offsets[i + 1])].tolist() == [3]
    q, qo = o.pack([seqs[0], b"MKTAYI", b"GGGGGGGGGGGGGGGGGGGGGGGGGGGGGG*"])
    r = o.search_proteins(idx, q, qo, o.opts(min_kmatch=1, min_kratio=0.0))
    assert r.size_in_kmer.tolist() == [16, 0, 24]
    assert r.hits(0) == [(2, 16), (3, 10)]  # 10 shared 7-mers of the common 16-residue prefix
    assert r.hits(1) == []          # SizeInKmer < 7: skipped
    assert r.hits(2) == [(3, 24)]   # query positions count with multiplicity
