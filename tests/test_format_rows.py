"""The batched row formatter (kaamer_host_format_tsv) against the transliterated Go writer
(QueryResultHandler, pkg/search/search.go:505-606): byte-identical TSV rows with and without alignment,
`-pos`, `-ann`.  The CPU tests drive the formatter with hand-built results (no device involved); the GPU
tests format real search / alignment results, with the EntryId / Length table stored in the `.kidx` file."""
import numpy as np
import pytest

from tests import go_transliteration as gt


def _fake_result(rng, nq, with_pos):
    from kaamer_b200 import SearchResult

    sizes = rng.integers(7, 400, nq).astype(np.int32)
    nh = rng.integers(0, 4, nq)
    hit_off = np.zeros(nq + 1, np.uint64)
    hit_off[1:] = np.cumsum(nh)
    n = int(hit_off[-1])
    subject = rng.integers(0, 5000, n).astype(np.uint32)
    kmatch = np.concatenate([np.sort(rng.integers(1, sizes[i] + 1, nh[i]))[::-1] for i in range(nq)]).astype(np.uint32) \
        if n else np.zeros(0, np.uint32)
    r = SearchResult(hit_off=hit_off, subject=subject, kmatch=kmatch, size_in_kmer=sizes, n_lookups=0, n_increments=0)
    if with_pos:
        rows = np.repeat(np.arange(nq), nh)
        plen = sizes[rows]
        po = np.zeros(n + 1, np.uint64)
        po[1:] = np.cumsum(plen)
        pos = (rng.random(int(po[-1])) < 0.4).astype(np.uint8)
        # runs that reach the end of the query, empty rows, single positions
        for k in range(n):
            b, e = int(po[k]), int(po[k + 1])
            if k % 5 == 0:
                pos[e - 3:e] = 1
            if k % 7 == 0:
                pos[b:e] = 0
            if k % 11 == 0:
                pos[b:e] = 0
                pos[b + 2] = 1
        r.pos_off, r.pos = po, pos
    return r


def _fake_aln(rng, n):
    from kaamer_b200.gpu import ALN_DTYPE

    a = np.zeros(n, dtype=ALN_DTYPE)
    a["identity"] = (rng.random(n) * 100).astype(np.float32)
    a["length"] = rng.integers(0, 900, n)
    a["mismatches"] = rng.integers(0, 300, n)
    a["gap_openings"] = rng.integers(0, 9, n)
    a["query_start"] = rng.integers(1, 50, n)
    a["query_end"] = rng.integers(50, 900, n)
    a["subject_start"] = rng.integers(1, 50, n)
    a["subject_end"] = rng.integers(50, 900, n)
    a["bitscore"] = np.round(rng.random(n) * 700, 3)
    a["evalue"] = 10.0 ** (-rng.random(n) * 180)
    if n > 6:
        a["identity"][1] = np.nan  # empty alignment: 0 / 0 in float32 (align.go:99)
        a["bitscore"][2] = a["bitscore"][3]  # a BitScore tie keeps the Kmatch order
        a["evalue"][4] = 0.0
        a["evalue"][5] = 1.5e300
        a["bitscore"][6] = 0.005
    return a


def _expected(r, names, seq_off, aln, with_pos, with_ann, entries):
    out = ""
    for i in range(r.n_rows):
        b, e = int(r.hit_off[i]), int(r.hit_off[i + 1])
        hits, ph = [], {}
        for k in range(b, e):
            h = {"Key": (int(r.subject[k]), k), "Kmatch": int(r.kmatch[k]), "Alignment": None}
            if aln is not None:
                x = aln[k]
                h["Alignment"] = {"Identity": np.float32(x["identity"]), "Length": int(x["length"]),
                                  "Mismatches": int(x["mismatches"]), "GapOpenings": int(x["gap_openings"]),
                                  "QueryStart": int(x["query_start"]), "QueryEnd": int(x["query_end"]),
                                  "SubjectStart": int(x["subject_start"]), "SubjectEnd": int(x["subject_end"]),
                                  "EValue": float(x["evalue"]), "BitScore": float(x["bitscore"])}
            if with_pos:
                ph[h["Key"]] = [bool(v) for v in r.pos[int(r.pos_off[k]):int(r.pos_off[k + 1])]]
            hits.append(h)
        ent = {h["Key"]: entries(h["Key"][0]) for h in hits}
        out += gt.tsv_rows(names[i], int(r.size_in_kmer[i]), 1, int(seq_off[i + 1] - seq_off[i]), hits, ent, ph,
                           aln is not None, with_pos, with_ann, True)
    return out.encode()


@pytest.mark.parametrize("with_aln", [False, True])
@pytest.mark.parametrize("with_pos", [False, True])
def test_rows_equal_the_transliterated_writer(with_aln, with_pos):
    from kaamer_b200 import format_tsv

    rng = np.random.default_rng(3 + 2 * with_aln + with_pos)
    nq = 300
    r = _fake_result(rng, nq, with_pos)
    names = [f"sp|Q{i:05d}|NAME_{i} some description {i}" if i % 3 else f"query{i}" for i in range(nq)]
    qlen = rng.integers(20, 900, nq)
    seq_off = np.zeros(nq + 1, np.uint64)
    seq_off[1:] = np.cumsum(qlen)
    aln = _fake_aln(rng, len(r.subject)) if with_aln else None
    got = format_tsv(r, names, aln=aln, seq_off=seq_off, is_protein=True, with_positions=with_pos)
    exp = _expected(r, names, seq_off, aln, with_pos, False, lambda pid: {"EntryId": str(pid), "Length": 0})
    assert got == exp
    assert got.count(b"\n") == len(r.subject) and len(got) > 1000


def test_arguments_are_checked():
    from kaamer_b200 import KaamerGpuError, format_tsv

    rng = np.random.default_rng(1)
    r = _fake_result(rng, 20, False)
    names = [f"q{i}" for i in range(20)]
    so = np.arange(21, dtype=np.uint64) * 30
    with pytest.raises(KaamerGpuError):
        format_tsv(r, names, seq_off=so, with_positions=True)      # the hits carry no positions
    with pytest.raises(KaamerGpuError):
        format_tsv(r, names, seq_off=so, with_annotations=True)    # no table without a handle
    assert format_tsv(_fake_result(rng, 0, False), [], seq_off=np.zeros(1, np.uint64)) == b""


@pytest.mark.gpu
def test_real_results_with_annotations_through_a_kidx_file(small_db, tmp_path):
    """search + align on the GPU, EntryId / Length table saved in the `.kidx` file, rows formatted natively"""
    from kaamer_b200 import GpuIndex, SearchOptions, format_tsv, synth

    res, off, ids = small_db["res"], small_db["off"], small_db["ids"]
    max_id = int(ids.max())
    entry = [""] * (max_id + 1)
    length = np.zeros(max_id + 1, np.int32)
    for i, pid in enumerate(ids.tolist()):  # protein_store[id] = last record written with the id
        entry[pid] = f"sp|S{i + 1:06d}|SYN_{i + 1}"
        length[pid] = int(off[i + 1] - off[i])
    q, qo, _ = synth.protein_queries(res, off, 150, config_index=1, stream=33)
    names = [f"query_{i} sampled" for i in range(150)]
    path = str(tmp_path / "ann.kidx")
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        g.set_annotations(entry, length)
        g.save(path)
    with GpuIndex.open(path) as g:
        r = g.search_proteins(q, qo, SearchOptions(extract_positions=True))
        nh = np.diff(r.hit_off.astype(np.int64))
        pq = np.repeat(np.arange(len(nh), dtype=np.uint32), nh)
        aln = g.align(q, qo, pq, r.subject)
        ent = lambda pid: {"EntryId": entry[pid], "Length": int(length[pid])}  # noqa: E731
        for with_aln in (False, True):
            for with_pos in (False, True):
                got = format_tsv(r, names, aln=aln if with_aln else None, seq_off=qo, with_positions=with_pos,
                                 with_annotations=True, index=g)
                exp = _expected(r, names, qo, aln if with_aln else None, with_pos, True, ent)
                assert got == exp, (with_aln, with_pos)
    assert len(r.subject) > 150


@pytest.mark.gpu
def test_nucleotide_rows_use_the_orf_location(small_db):
    from kaamer_b200 import GpuIndex, SearchOptions, format_tsv, synth

    res, off, ids = small_db["res"], small_db["off"], small_db["ids"]
    nt, no = synth.nucleotide_contigs(res, off, 2, 30_000, config_index=2)
    with GpuIndex.build(res, off, ids, keep_proteins=False) as g:
        r = g.search_nucleotide(nt, no, SearchOptions())
    names = ["contig_1 first", "contig_2 second"]
    got = format_tsv(r, names, is_protein=False, with_positions=True).decode().splitlines()
    assert len(got) == len(r.subject) > 10
    k = 0
    for i in range(r.n_rows):
        for j in range(int(r.hit_off[i]), int(r.hit_off[i + 1])):
            f = got[k].split("\t")
            assert f[0] == names[int(r.row_contig[i])].split(" ")[0]
            assert (int(f[6]), int(f[7])) == (int(r.row_start[i]), int(r.row_end[i])) and f[8] == "1"
            ph = [bool(v) for v in r.pos[int(r.pos_off[j]):int(r.pos_off[j + 1])]]
            assert f[10] == gt.format_positions_to_string(ph, False)
            k += 1
