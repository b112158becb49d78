"""The C++ oracle (what the GPU kernels are compared with) against a literal Python transliteration
of the Go source (tests/go_transliteration.py), on small random inputs: index contents, protein
search with positions, GetORFs, translated search incl. SetBestStartCodon.  Two independent
restatements of the reference must agree bit for bit — this is the strongest pin the oracle can get
while the reference itself cannot be built here (no Go toolchain; the reference ships no tests)."""
import numpy as np
import pytest

from kaamer_b200 import makedb, synth
from oracle import oracle as o
from tests import go_transliteration as go

AA = "ACDEFGHIKLMNPQRSTVWY"


def _random_db(rng, n):
    """records with the quirks the FASTA front end must honour: families, lower case, unknown letters,
    ', partial' names, records shorter than a k-mer, wrapped lines"""
    recs = []
    founders = []
    for i in range(n):
        L = int(rng.integers(3, 160))
        if founders and rng.random() < 0.35:
            s = list(founders[int(rng.integers(0, len(founders)))])
            for _ in range(int(rng.integers(0, max(1, len(s) // 8)))):
                s[int(rng.integers(0, len(s)))] = AA[int(rng.integers(0, 20))]
            s = "".join(s)
        else:
            s = "".join(AA[j] for j in rng.integers(0, 20, L))
            if L > 30:
                founders.append(s)
        if rng.random() < 0.1:
            s = s.lower()
        if rng.random() < 0.1 and len(s) > 10:
            p = int(rng.integers(0, len(s)))
            s = s[:p] + "XBZU*"[int(rng.integers(0, 5))] + s[p + 1:]
        name = f"protein {i}" + (", partial" if rng.random() < 0.08 else "")
        recs.append((f">sp|S{i:05d}|SYN_{i} {name}", s))
    text = ""
    for h, s in recs:
        text += h + "\n"
        for b in range(0, len(s), 60):
            text += s[b:b + 60] + "\n"
    return text, recs


def _oracle_index(tmp_path, text):
    p = tmp_path / "db.fasta"
    p.write_text(text)
    _, _, res, off, ids = makedb.read_fasta(str(p))
    return o.Index.build(res, off, ids, 2)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_index_contents(tmp_path, seed):
    rng = np.random.default_rng(seed)
    text, _ = _random_db(rng, 80)
    gi = go.make_index(text)
    idx = _oracle_index(tmp_path, text)
    stats = gi.pop("__stats__")
    assert (idx.n_proteins, idx.n_aa) == (stats["proteins"], stats["aa"])
    keys = sorted(gi.keys())
    assert idx.keys.tolist() == keys  # ascending u32 == badger's big-endian byte order
    for i, k in enumerate(keys):
        assert idx.postings[int(idx.offsets[i]):int(idx.offsets[i + 1])].tolist() == gi[k], hex(k)
    for kmer in ("MKTAYIA", "AAAAAAX", "AXAAAAA", "yyyyyyy", "ACDEFG*"):
        assert o.encode_kmer(kmer.encode()) == go.encode_kmer(kmer)


OPTS = [dict(), dict(min_kmatch=1, min_kratio=0.0, max_results=1000), dict(min_kmatch=3, min_kratio=0.3, max_results=2),
        dict(min_kmatch=1, min_kratio=0.0, max_results=0), dict(min_kmatch=0, min_kratio=0.9, max_results=5)]


@pytest.mark.parametrize("seed", [4, 5])
def test_protein_search(tmp_path, seed):
    rng = np.random.default_rng(seed)
    text, recs = _random_db(rng, 120)
    gi = go.make_index(text)
    gi.pop("__stats__")
    idx = _oracle_index(tmp_path, text)
    queries = []
    for _ in range(60):
        s = list(recs[int(rng.integers(0, len(recs)))][1].upper())
        for _ in range(int(rng.integers(0, 1 + len(s) // 10))):
            s[int(rng.integers(0, len(s)))] = AA[int(rng.integers(0, 20))]
        queries.append("".join(s))
    queries += ["", "MKT", "A" * 12, "A" * 13, "A" * 40 + "*", recs[0][1].upper() * 3, "MKTAYIAKQRQISFVKSHFSRQX" * 2,
                recs[1][1].lower()]  # the last FASTA query is not upper-cased by the reference reader
    q, qo = o.pack([s.encode() for s in queries])
    for kw in OPTS:
        ora = o.search_proteins(idx, q, qo, o.opts(want_positions=True, **kw), 2)
        ref = go.protein_search(gi, queries, extract_positions=True,
                                **{**dict(min_kmatch=10, min_kratio=0.05, max_results=10), **kw})
        assert ora.n_rows == len(queries)
        h = 0
        for i, (size, hits, pos) in enumerate(ref):
            assert int(ora.size_in_kmer[i]) == size, (i, queries[i])
            assert ora.hits(i) == [(k, v) for k, v in hits], (i, kw)
            for k, _ in hits:
                assert ora.positions(h).tolist() == [int(b) for b in pos[k]], (i, k)
                h += 1
        assert h == len(ora.subject)


@pytest.mark.parametrize("seed", [6, 7])
def test_orfs_and_translated_search(tmp_path, seed):
    rng = np.random.default_rng(seed)
    res, off = synth.protein_db(60, config_index=1)
    names = [f"sp|S{i}|X protein {i}" for i in range(len(off) - 1)]
    p = tmp_path / "db.fasta"
    synth.write_fasta(str(p), names, res, off)
    text = p.read_text()
    gi = go.make_index(text)
    gi.pop("__stats__")
    idx = _oracle_index(tmp_path, text)
    nt, noff = synth.nucleotide_contigs(res, off, 2, 9000, config_index=2 + seed)
    contigs = [nt[int(noff[i]):int(noff[i + 1])].tobytes().decode() for i in range(len(noff) - 1)]
    acgtn = "acgtn"
    contigs += ["", "at", "atgaaa", "".join(acgtn[j] for j in rng.integers(0, 5, 700)),
                "".join("ACGT"[j] for j in rng.integers(0, 4, 1500)), "gct" * 100, "ttg" + "gca" * 40 + "tagc"]
    # GetORFs alone
    for dna in contigs:
        a = o.get_orfs(dna.encode())
        b = go.get_orfs(dna)
        assert len(a.seqs) == len(b)
        for j, x in enumerate(b):
            assert a.seqs[j].decode() == x["seq"]
            assert (int(a.start[j]), int(a.end[j]), bool(a.plus[j]), list(a.alts[j])) == \
                   (x["start"], x["end"], x["plus"], x["alts"]), (dna[:30], j)
    # translated search
    cn, co = o.pack([c.encode() for c in contigs])
    for kw in (dict(), dict(min_kmatch=1, min_kratio=0.0, max_results=50), dict(min_kmatch=4, min_kratio=0.2, max_results=1)):
        ora = o.search_nucleotide(idx, cn, co, o.opts(**kw), 2)
        rows = go.nucleotide_search(gi, contigs, **{**dict(min_kmatch=10, min_kratio=0.05, max_results=10), **kw})
        assert ora.n_rows == len(rows), kw
        h = 0
        for i, r in enumerate(rows):
            assert (int(ora.row_contig[i]), int(ora.row_start[i]), int(ora.row_end[i]), bool(ora.row_plus[i])) == \
                   (r["contig"], r["start"], r["end"], r["plus"]), (i, kw)
            assert int(ora.size_in_kmer[i]) == r["size"]
            b, e = int(ora.row_seq_off[i]), int(ora.row_seq_off[i + 1])
            assert ora.row_seq[b:e].tobytes().decode() == r["seq"]
            assert ora.hits(i) == r["hits"]
            for k, _ in r["hits"]:
                assert ora.positions(h).tolist() == [int(x) for x in r["pos"][k]]
                h += 1
        if not kw:
            assert len(rows) > 5  # the default options do find the planted genes


@pytest.mark.parametrize("seed", [8, 9])
def test_alignment_postprocessing(seed):
    """Everything align.Align computes itself from the biogo alignment (align.go:72-157: identity and
    similarity in float32, mismatches, the `score == -GapOpen` gap rule, raw score, bitscore, e-value,
    1-based coordinates) — the oracle's fields against the transliteration applied to the oracle's own
    gapped strings.  (The DP that produces the strings is biogo's: parity unpinned)"""
    rng = np.random.default_rng(seed)
    m = o.blosum62()
    n_aa = 3_500_000
    prm = o.aln_params(n_aa)
    n_gapped = 0
    for _ in range(120):
        L = int(rng.integers(20, 260))
        s = [AA[j] for j in rng.integers(0, 20, L)]
        q = list(s)
        for _ in range(int(rng.integers(0, 1 + L // 6))):  # substitutions
            q[int(rng.integers(0, len(q)))] = AA[int(rng.integers(0, 20))]
        for _ in range(int(rng.integers(0, 4))):  # insertions / deletions of 1..12 residues
            p = int(rng.integers(0, len(q)))
            n = int(rng.integers(1, 13))
            if rng.random() < 0.5:
                del q[p:p + n]
            else:
                q[p:p] = [AA[j] for j in rng.integers(0, 20, n)]
        if len(q) < 8:
            continue
        q = "".join(q) + ("*" if rng.random() < 0.3 else "")
        s = "".join(s)
        out, a, b = o.align(q.encode(), s.encode(), prm, want_strings=True)
        if out.length == 0:
            continue
        a, b = a.decode(), b.decode()
        assert len(a) == len(b) == out.length
        segs = go.segments_from_strings(a, b, out.query_start - 1, out.subject_start - 1, m)
        assert len(segs) == out.n_segments and sum(x[0] for x in segs) == out.dp_score
        ref = go.align_postprocess(a, b, segs, len(q), n_aa, m)
        n_gapped += ref["gap_openings"] > 0
        for k in ("length", "mismatches", "gap_openings", "raw", "query_start", "query_end", "subject_start", "subject_end"):
            assert getattr(out, k) == ref[k], (k, q, s)
        assert np.float32(out.identity) == np.float32(ref["identity"])
        assert np.float32(out.similarity) == np.float32(ref["similarity"])
        assert abs(out.bitscore - ref["bitscore"]) <= 1e-12 * max(1.0, abs(ref["bitscore"]))
        assert abs(out.evalue - ref["evalue"]) <= 1e-9 * abs(ref["evalue"])
    assert n_gapped > 10


def test_tsv_database(tmp_path):
    """`kaamer-db -make -f tsv`: ids 0..n-1 over the accepted rows, sequences indexed as written"""
    rng = np.random.default_rng(12)
    rows = ["Organism\tEntryID\tSEQUENCE\tProteinName"]
    for i in range(70):
        L = int(rng.integers(3, 90))
        s = "".join(AA[j] for j in rng.integers(0, 20, L))
        if rng.random() < 0.15:
            s = s.lower()  # TSV sequences are not upper-cased: every k-mer holds unknown letters
        entry = "" if rng.random() < 0.1 else f"E{i}"
        rows.append(f"org{i}\t{entry}\t{s}\tname {i}")
    rows.insert(5, "short\tE_short\tMKT\tn")
    rows.insert(9, "only two\tcolumns")
    text = "\n".join(rows) + "\n"
    p = tmp_path / "db.tsv"
    p.write_text(text)
    entries, res, off, ids = makedb.read_tsv(str(p))
    gi = go.make_index_tsv(text)
    stats = gi.pop("__stats__")
    assert entries == stats["entries"] and ids.tolist() == list(range(len(entries)))
    idx = o.Index.build(res, off, ids, 2)
    assert (idx.n_proteins, idx.n_aa) == (stats["proteins"], stats["aa"])
    keys = sorted(gi.keys())
    assert idx.keys.tolist() == keys
    for i, k in enumerate(keys):
        assert idx.postings[int(idx.offsets[i]):int(idx.offsets[i + 1])].tolist() == gi[k]


def _gotoh_local_optimum(q: str, s: str, m, gap_open=-11) -> int:
    """Independent three-state local alignment (match / gap-in-query / gap-in-subject), the model the oracle
    documents for biogo's SWAffine as kaamer configures it: the opening costs GapOpen, each gapped residue
    costs the matrix's gap row (0 in the only reading under which kaamer's `score == -GapOpen` test works).
    Returns the optimum only — it does not depend on tie-breaking."""
    NEG = -10 ** 9
    idx = [go.AA_POS_IN_MATRIX[c] for c in q], [go.AA_POS_IN_MATRIX[c] for c in s]
    n, k = len(q), len(s)
    M = [[0] * (k + 1) for _ in range(n + 1)]
    U = [[NEG] * (k + 1) for _ in range(n + 1)]  # gap run consuming query residues
    L = [[NEG] * (k + 1) for _ in range(n + 1)]  # gap run consuming subject residues
    best = 0
    for i in range(1, n + 1):
        for j in range(1, k + 1):
            U[i][j] = max(M[i - 1][j] + gap_open, U[i - 1][j], L[i - 1][j] + gap_open)
            L[i][j] = max(M[i][j - 1] + gap_open, L[i][j - 1], U[i][j - 1] + gap_open)
            d = max(M[i - 1][j - 1], U[i - 1][j - 1], L[i - 1][j - 1], 0)
            M[i][j] = max(0, d + int(m[idx[0][i - 1]][idx[1][j - 1]]))
            if M[i][j] > best:
                best = M[i][j]
    return best


def test_alignment_optimum_against_an_independent_dp():
    rng = np.random.default_rng(21)
    m = o.blosum62()
    prm = o.aln_params(1_000_000)
    checked = 0
    for _ in range(60):
        L = int(rng.integers(8, 70))
        s = [AA[j] for j in rng.integers(0, 20, L)]
        q = list(s)
        for _ in range(int(rng.integers(0, 1 + L // 5))):
            q[int(rng.integers(0, len(q)))] = AA[int(rng.integers(0, 20))]
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(0, len(q)))
            n = int(rng.integers(1, 9))
            if rng.random() < 0.5:
                del q[p:p + n]
            else:
                q[p:p] = [AA[j] for j in rng.integers(0, 20, n)]
        if len(q) < 7:
            continue
        q, s = "".join(q), "".join(s)
        out = o.align(q.encode(), s.encode(), prm)
        assert out.dp_score == _gotoh_local_optimum(q, s, m), (q, s)
        checked += 1
    # unrelated sequences too
    for _ in range(20):
        q = "".join(AA[j] for j in rng.integers(0, 20, int(rng.integers(7, 60))))
        s = "".join(AA[j] for j in rng.integers(0, 20, int(rng.integers(7, 60))))
        assert o.align(q.encode(), s.encode(), prm).dp_score == _gotoh_local_optimum(q, s, m)
    assert checked > 40
