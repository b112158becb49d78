"""GPU parity tests (through the C ABI) of six-frame translation / ORF extraction, translated
nucleotide search (incl. SetBestStartCodon) and position extraction vs the CPU oracle."""
import json
import os

import numpy as np
import pytest

from tests.helpers import assert_same_hits, assert_same_orfs, assert_same_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_small(small_db):
    from kaamer_b200 import GpuIndex

    g = GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], device=0)
    yield g
    g.close()


def _oracle_orfs(contigs):
    from oracle import oracle as o

    return [o.get_orfs(c) for c in contigs]


def test_orfs_known_answer(gpu_small, golden_dir):
    """the hand-derived vector of SURVEY §8c (atg + gct*20 + taa), lower and upper case"""
    from oracle import oracle as o

    kat = json.load(open(os.path.join(golden_dir, "kat.json")))
    for dna in (kat["orfs_dna"].encode(), kat["orfs_dna"].upper().encode()):
        nt, off = o.pack([dna])
        t = gpu_small.get_orfs(nt, off)
        assert len(t) == len(kat["orfs"])
        for i, e in enumerate(kat["orfs"]):
            assert t.sequence(i).decode() == e["seq"]
            assert (int(t.start[i]), int(t.end[i]), bool(t.plus[i])) == (e["start"], e["end"], e["plus"])
            assert t.starts_alternative(i) == e["alts"]


def test_orfs_edge_contigs(gpu_small):
    """empty / tiny contigs, every length mod 3, n's and other letters, no stop at all, stops
    only, start codons everywhere; many contigs per batch."""
    from oracle import oracle as o

    rng = np.random.default_rng(7)
    acgt = np.frombuffer(b"acgt", np.uint8)

    def rnd(n, alphabet=acgt):
        return alphabet[rng.integers(0, len(alphabet), n)].tobytes()

    contigs = [b"", b"a", b"at", b"atg", b"atga", b"atgaa"]
    contigs += [rnd(n) for n in (62, 63, 64, 65, 66, 67, 200, 1000, 1001, 1002, 5000)]
    contigs += [b"gct" * 400, b"gct" * 400 + b"g", b"atg" * 300, b"taa" * 100, b"ttg" + b"gca" * 50 + b"tag" + b"c"]
    contigs += [rnd(3000, np.frombuffer(b"acgtn", np.uint8)), rnd(3000, np.frombuffer(b"ACGTacgtNRYK-", np.uint8))]
    contigs += [rnd(2000, np.frombuffer(b"acg", np.uint8)),          # no 't': no stop codon in any plus frame
                rnd(2000, np.frombuffer(b"gca", np.uint8)) + b"n" * 7 + rnd(500)]
    contigs += [rnd(int(n)) for n in rng.integers(0, 400, 200)]      # reads-like: many short contigs
    nt, off = o.pack(contigs)
    t = gpu_small.get_orfs(nt, off)
    assert_same_orfs(t, _oracle_orfs(contigs))


def test_orfs_stop_dense_contig_takes_the_dense_fallback(gpu_small):
    """more than 12.5 % of the codons are stops: the compacted end list overflows and the
    one-thread-per-codon kernel must produce the same ORFs"""
    from oracle import oracle as o

    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"acgt", np.uint8)
    contigs = [b"taa" * 4000, b"taa" * 700 + acgt[rng.integers(0, 4, 900)].tobytes() + b"tag" * 500 + b"a",
               b"ttaa" * 3000]
    nt, off = o.pack(contigs)
    t = gpu_small.get_orfs(nt, off)
    assert len(t) > 3
    assert_same_orfs(t, _oracle_orfs(contigs))


def test_orfs_synthetic_contigs(small_db, gpu_small):
    from kaamer_b200 import synth

    nt, off = synth.nucleotide_contigs(small_db["res"], small_db["off"], 3, 150_000, config_index=2)
    contigs = [nt[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]
    t = gpu_small.get_orfs(nt, off)
    assert len(t) > 1000
    assert_same_orfs(t, _oracle_orfs(contigs))


@pytest.mark.parametrize("opts", [
    dict(),
    dict(min_kmatch=1, min_kratio=0.0, max_results=10),
    dict(min_kmatch=1, min_kratio=0.0, max_results=100),
    dict(min_kmatch=5, min_kratio=0.4, max_results=3),
    dict(min_kmatch=0, min_kratio=0.0, max_results=1),
])
def test_nucleotide_search_parity(small_db, gpu_small, opts):
    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    nt, off = synth.nucleotide_contigs(small_db["res"], small_db["off"], 2, 120_000, config_index=2)
    ora = o.search_nucleotide(small_db["idx"], nt, off, o.opts(**opts), 4)
    r = gpu_small.search_nucleotide(nt, off, SearchOptions(max_results=opts.get("max_results", 10),
                                                           min_kmatch=opts.get("min_kmatch", 10),
                                                           min_kratio=opts.get("min_kratio", 0.05)))
    assert ora.n_rows > 50
    assert_same_rows(r, ora, str(opts))
    assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments


def test_nucleotide_start_codon_correction_is_exercised(small_db, gpu_small):
    """Genes preceded by an in-frame upstream start: SetBestStartCodon must trim (dna.go:252-268)."""
    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    res, offp = small_db["res"], small_db["off"]
    rng = np.random.default_rng(11)
    lut = {a: cs for a, cs in synth._CODONS.items()}
    parts = []
    n_genes = 40
    for g in range(n_genes):
        p = int(rng.integers(0, len(offp) - 1))
        aa = res[int(offp[p]):int(offp[p + 1])].tobytes().decode()
        # upstream: stop, then a start codon, a few random non-stop codons, then the gene (its own
        # start is an alternative start) -> first alternative precedes the first matching k-mer
        lead = "TAA" + "TTG" + "".join(lut[x][0] for x in "ARNDCQEGH"[: int(rng.integers(1, 9))])
        gene = lead + "ATG" + "".join(lut[x][int(rng.integers(0, len(lut[x])))] for x in aa) + "TGA"
        pad = "".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(10, 90))))
        s = pad + gene
        if rng.random() < 0.5:
            s = s.encode().translate(synth._COMP)[::-1].decode()
        parts.append(s)
    contig = "".join(parts).encode()
    nt, off = o.pack([contig, contig[5:50_000], contig[::-1]])
    for opts in (dict(), dict(min_kmatch=1, min_kratio=0.0, max_results=20)):
        ora = o.search_nucleotide(small_db["idx"], nt, off, o.opts(**opts), 4)
        r = gpu_small.search_nucleotide(nt, off, SearchOptions(max_results=opts.get("max_results", 10),
                                                               min_kmatch=opts.get("min_kmatch", 10),
                                                               min_kratio=opts.get("min_kratio", 0.05)))
        assert_same_rows(r, ora, f"startcodon {opts}")
    # the trim really happened for many rows: a trimmed row's sequence is shorter than its ORF
    t = gpu_small.get_orfs(nt, off)
    orf_len = {(int(c), int(e), int(p)): int(t.seq_off[i + 1] - t.seq_off[i])
               for i, (c, e, p) in enumerate(zip(t.contig, t.end, t.plus))}
    trimmed = 0
    for i in range(r.n_rows):
        key = (int(r.row_contig[i]), int(r.row_end[i]), int(r.row_plus[i]))
        if int(r.row_seq_off[i + 1] - r.row_seq_off[i]) < orf_len[key]:
            trimmed += 1
    assert trimmed >= n_genes // 2


def test_nucleotide_empty_and_no_hit_batches(gpu_small):
    from kaamer_b200 import SearchOptions
    from oracle import oracle as o

    r = gpu_small.search_nucleotide(np.zeros(0, np.uint8), np.zeros(1, np.uint64), SearchOptions())
    assert r.n_rows == 0 and len(r.subject) == 0
    nt, off = o.pack([b"", b"acgtacgtacgt", b"gct" * 200])
    r = gpu_small.search_nucleotide(nt, off, SearchOptions())
    assert r.n_rows == 0 and len(r.subject) == 0 and r.n_lookups > 0


@pytest.mark.parametrize("opts", [dict(), dict(min_kmatch=1, min_kratio=0.0, max_results=50)])
def test_protein_positions_parity(small_db, gpu_small, opts):
    """ExtractPositions for protein queries (search.go:416,442-452)."""
    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 300, config_index=1, stream=9)
    extra = [b"", b"MKT", small_db["res"][:13].tobytes(), small_db["res"][:700].tobytes() * 4, b"A" * 40]
    seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)] + extra
    q, qo = o.pack(seqs)
    ora = o.search_proteins(small_db["idx"], q, qo, o.opts(want_positions=True, **opts), 4)
    r = gpu_small.search_proteins(q, qo, SearchOptions(max_results=opts.get("max_results", 10),
                                                       min_kmatch=opts.get("min_kmatch", 10),
                                                       min_kratio=opts.get("min_kratio", 0.05),
                                                       extract_positions=True))
    assert_same_hits(r, ora, f"positions {opts}")
    np.testing.assert_array_equal(r.pos_off.astype(np.int64), ora.pos_off.astype(np.int64))
    np.testing.assert_array_equal(r.pos, ora.pos)
    assert int(r.pos.sum()) > 0


def test_fastq_file_through_the_native_reader(small_db, gpu_small, tmp_path):
    """FASTQ file -> kaamer_host_read_fastq -> kaamer_gpu_search_nucleotide (the FastqSearch path,
    search_fastq.go:78-140): rows, hits, locations and positions equal the oracle on the reads the
    transliterated Go reader extracts from the same file."""
    from kaamer_b200 import SearchOptions, readers, synth
    from oracle import oracle as o
    from tests import go_transliteration as go

    nt, off = synth.nucleotide_contigs(small_db["res"], small_db["off"], 1, 60_000, config_index=2)
    rng = np.random.default_rng(17)
    p = str(tmp_path / "reads.fastq")
    with open(p, "wb") as f:
        for i in range(400):
            s = int(rng.integers(0, len(nt) - 400))
            L = int(rng.integers(60, 400))
            read = nt[s:s + L].tobytes()
            if i % 7 == 0:
                read = read.upper()
            qual = b"I" * L if i % 11 else b"@" + b"I" * (L - 1)  # a quality line that starts with '@'
            f.write(b"@read%d sampled read\n" % i + read + b"\n+\n" + qual + b"\n")
    b = readers.read_fastq(p)
    ref = go.get_queries_fastq(p)
    assert len(b) == len(ref) == 400 and b.names == [r[0] for r in ref]
    rq, rqo = o.pack([r[1].encode() for r in ref])
    np.testing.assert_array_equal(b.residues, rq)
    ora = o.search_nucleotide(small_db["idx"], rq, rqo, o.opts(), 4)
    r = gpu_small.search_nucleotide(b.residues, b.seq_off, SearchOptions())
    assert ora.n_rows > 50
    assert_same_rows(r, ora, "fastq file")


def test_other_genetic_codes_behind_the_explicit_call(golden_dir):
    """the reference always translates with table 11 (dna.go:106); the other tables of
    pkg/search/gcode.go (extracted into tests/golden/gcodes.json) are a kaamer_gpu_set_genetic_code call away:
    GetORFs under every table == the oracle under the same table, and the default is table 11 again after
    a reset"""
    import json
    import os

    from kaamer_b200 import GpuIndex
    from oracle import oracle as o
    from tests.helpers import assert_same_orfs

    tables = json.load(open(os.path.join(golden_dir, "gcodes.json")))
    assert tables["gcodeBacteria"] == tables["gcode_11"]
    rng = np.random.default_rng(8)
    contigs = [bytes(np.frombuffer(b"acgt", np.uint8)[rng.integers(0, 4, n)]) for n in (3000, 1501, 40, 7000)]
    contigs.append(b"atg" + b"gct" * 30 + b"tga" + b"ata" + b"aaa" * 25 + b"aga" + b"ttgacgtnnacg")
    nt, no = o.pack(contigs)
    res, off = o.pack([b"MKTAYIAKQRQISFVKSHFSRQ"])
    with GpuIndex.build(res, off, np.array([1], np.uint32), keep_proteins=False) as g:
        base = g.get_orfs(nt, no)
        seen = set()
        try:
            for name, t in sorted(tables.items()):
                aas, mask = t["aas"].encode(), int(t["start_mask"])
                o.set_genetic_code(aas, mask)
                g.set_genetic_code(aas, mask)
                got = g.get_orfs(nt, no)
                assert_same_orfs(got, [o.get_orfs(c) for c in contigs])
                seen.add((len(got), got.seq.tobytes()))
            assert len(seen) > 3  # the tables really differ (stops / starts move the ORFs)
        finally:
            o.set_genetic_code(None)
        g.set_genetic_code(None)
        again = g.get_orfs(nt, no)
        assert len(again) == len(base) and again.seq.tobytes() == base.seq.tobytes()
        assert_same_orfs(again, [o.get_orfs(c) for c in contigs])
