"""The one input/output pair the reference itself holds for the hot path: the worked example of
docs/client.md:117-179 (a 270-aa query, BLAN1_KLEPN, searched against a database that contains it):
SizeInKmer 264, Location 1..270, one hit with Kmatch 264, PositionHits = 264 x true, `-pos` column "1-264".
The database here = that protein among synthetic decoys (the hit set of a query only depends on the
posting lists of its own k-mers)."""
import numpy as np
import pytest

QUERY = (b"MELPNIMHPVAKLSTALAAALMLSGCMPGEIRPTIGQQMETGDQRFGDLVFRQLAPNVWQHTSYLDMPGFGAVASNGLIVRDGGRVLVVDTAWTDDQTAQILNWIKQEINLPVA"
         b"LAVVTHAHQDKMGGMDALHAAGIATYANALSNQLAPQEGMVAAQHSLTFAANGWVEPATAPNFGPLKVFYPGPGHTSDNITVGIDGTDIAFGGCLIKDSKAKSLGNLGDADTEHY"
         b"AASARAFGAAFPKASMIVMSHSAPDSRAAITHTARMADKLR")


def _db():
    from kaamer_b200 import synth

    res, off = synth.protein_db(400, config_index=1)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes() for i in range(400)]
    seqs.insert(137, QUERY)
    return seqs


def _check(r, fmt_positions, hit_id):
    assert len(QUERY) == 270
    assert int(r.size_in_kmer[0]) == 264                       # "SizeInKmer": 264
    assert r.hits(0) == [(hit_id, 264)]                        # "Hits": [{"Key": ..., "Kmatch": 264}], -m 1
    pos = r.pos[int(r.pos_off[0]):int(r.pos_off[1])]
    assert len(pos) == 264 and pos.all()                       # "PositionHits": 264 x true
    assert fmt_positions(pos) == "1-264"                       # QueryHit.Positions column


def test_oracle_reproduces_the_documented_example():
    from oracle import oracle as o

    seqs = _db()
    res, off = o.pack(seqs)
    ids = np.arange(len(seqs), dtype=np.uint32)
    idx = o.Index.build(res, off, ids, 2)
    q, qo = o.pack([QUERY])
    r = o.search_proteins(idx, q, qo, o.opts(max_results=1, want_positions=True), 1)
    _check(r, lambda p: o.format_positions(p, False), 137)


@pytest.mark.gpu
def test_gpu_reproduces_the_documented_example():
    from kaamer_b200 import GpuIndex, SearchOptions, format_tsv
    from oracle import oracle as o

    seqs = _db()
    res, off = o.pack(seqs)
    ids = np.arange(len(seqs), dtype=np.uint32)
    q, qo = o.pack([QUERY])
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        g.set_annotations(["BLAN1_KLEPN" if i == 137 else f"DECOY_{i}" for i in range(len(seqs))],
                          [len(s) for s in seqs])
        r = g.search_proteins(q, qo, SearchOptions(max_results=1, extract_positions=True))
        _check(r, lambda p: o.format_positions(p, False), 137)
        row = format_tsv(r, ["query"], seq_off=qo, with_positions=True, with_annotations=True, index=g).decode()
        # kaamer -search -t prot -m 1 -fmt tsv -ann -pos: QueryName, hit id, %KMatch, QueryKSize, KMatch, ranges, QStart, QEnd
        f = row.rstrip("\n").split("\t")
        assert f[0] == "query" and f[1] == "BLAN1_KLEPN" and f[2] == "100.00" and f[3] == "264" and f[4] == "264"
        assert f[6] == "1" and f[7] == "270" and f[9] == "270" and f[10] == "1-264"
        # align=true on the same pair: identical sequences align end to end without gaps
        a, text = g.align(q, qo, [0], [137], want_text=True)
        assert int(a["length"][0]) == 270 and int(a["mismatches"][0]) == 0 and int(a["gap_openings"][0]) == 0
        assert (int(a["query_start"][0]), int(a["query_end"][0]), int(a["subject_start"][0]), int(a["subject_end"][0])) == (1, 270, 1, 270)
        assert float(a["identity"][0]) == 100.0
        assert text[0] == QUERY + b"\n" + QUERY + b"\n" + QUERY
