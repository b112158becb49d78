"""GPU parity tests (through the C ABI) of the protein search path vs the CPU oracle."""
import numpy as np
import pytest

from tests.helpers import assert_same_hits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_small(small_db):
    from kaamer_b200 import GpuIndex

    g = GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], device=0)
    yield g
    g.close()


def test_device_build_equals_oracle_index(small_db, gpu_small):
    k, f, p = gpu_small.index_arrays()
    idx = small_db["idx"]
    np.testing.assert_array_equal(k, idx.keys)
    np.testing.assert_array_equal(f, idx.offsets)
    np.testing.assert_array_equal(p, idx.postings)
    st = gpu_small.dbstats()
    assert (st["NumberOfProteins"], st["NumberOfAA"], st["NumberOfKmers"]) == (idx.n_proteins, idx.n_aa, idx.n_kmers)


def test_open_view_equals_build(small_db):
    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from oracle import oracle as o

    idx = small_db["idx"]
    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 300, config_index=1, stream=5)
    ora = o.search_proteins(idx, q, qo, o.opts(), 4)
    with GpuIndex.from_arrays(idx.keys, idx.offsets, idx.postings, idx.n_proteins, idx.n_aa, idx.n_kmers) as g:
        r = g.search_proteins(q, qo, SearchOptions())
    assert_same_hits(r, ora, "open_view")


def test_kidx_roundtrip(small_db, gpu_small, tmp_path):
    from kaamer_b200 import GpuIndex

    path = str(tmp_path / "small.kidx")
    gpu_small.save(path)
    with GpuIndex.open(path) as g2:
        a, b = gpu_small.index_arrays(), g2.index_arrays()
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
        assert g2.dbstats() == gpu_small.dbstats()


@pytest.mark.parametrize("opts", [
    dict(),                                                   # reference defaults 10 / 0.05 / 10
    dict(min_kmatch=1, min_kratio=0.0, max_results=10),       # everything passes: top-10 of all subjects
    dict(min_kmatch=1, min_kratio=0.0, max_results=1000),     # slow path: > 64 candidates, bitonic sort
    dict(min_kmatch=1, min_kratio=0.0, max_results=3),        # radix select through large tie groups
    dict(min_kmatch=30, min_kratio=0.5, max_results=2),
    dict(min_kmatch=0, min_kratio=0.31, max_results=5),
    dict(max_results=0),
])
def test_search_parity_options(small_db, gpu_small, opts):
    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 500, config_index=1, stream=3)
    ora = o.search_proteins(small_db["idx"], q, qo, o.opts(**opts), 4)
    r = gpu_small.search_proteins(q, qo, SearchOptions(max_results=opts.get("max_results", 10),
                                                       min_kmatch=opts.get("min_kmatch", 10),
                                                       min_kratio=opts.get("min_kratio", 0.05)))
    assert_same_hits(r, ora, str(opts))
    assert r.n_lookups == ora.n_lookups
    assert r.n_increments == ora.n_increments


def test_edge_cases(small_db, gpu_small):
    """empty / short / ragged queries, unknown residues, trailing '*', lower case, long queries."""
    from kaamer_b200 import SearchOptions
    from oracle import oracle as o

    res, off = small_db["res"], small_db["off"]
    rec = lambda i: res[int(off[i]):int(off[i + 1])].tobytes()
    long_q = b"".join(rec(i) for i in range(40))           # > 2048 k-mers: class G
    mid_q = b"".join(rec(i) for i in range(100, 104))      # class M
    seqs = [
        b"",                      # empty
        b"MKT",                   # shorter than k
        rec(5)[:12],              # SizeInKmer 6 < 7: skipped (search_protein.go:74-76)
        rec(5)[:13],              # SizeInKmer 7: searched
        rec(6)[:14] + b"*",       # trailing '*': SizeInKmer - 1
        rec(7),
        rec(8).lower(),           # lower case: every residue unknown -> key 0 (no upper-casing on the device)
        rec(9)[:50] + b"XBZJOU*" + rec(9)[57:],   # unknown letters inside
        rec(10) + b"*",
        mid_q,
        long_q,
        b"A" * 600,               # low complexity: one k-mer repeated
        rec(11) * 3,              # repeats: positions counted with multiplicity
    ]
    q, qo = o.pack(seqs)
    for opts in (dict(), dict(min_kmatch=1, min_kratio=0.0, max_results=50)):
        ora = o.search_proteins(small_db["idx"], q, qo, o.opts(**opts), 2)
        r = gpu_small.search_proteins(q, qo, SearchOptions(max_results=opts.get("max_results", 10),
                                                           min_kmatch=opts.get("min_kmatch", 10),
                                                           min_kratio=opts.get("min_kratio", 0.05)))
        assert_same_hits(r, ora, f"edge {opts}")
        assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments
    assert r.size_in_kmer[:5].tolist() == [-6, -3, 6, 7, 8]


def test_empty_batch_and_empty_index(gpu_small):
    from kaamer_b200 import GpuIndex, SearchOptions

    r = gpu_small.search_proteins(np.zeros(0, np.uint8), np.zeros(1, np.uint64), SearchOptions())
    assert r.n_rows == 0 and len(r.subject) == 0
    with GpuIndex.build(np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros(0, np.uint32)) as g:
        q = np.frombuffer(b"MKTAYIAKQRQISFVKSHFSRQ", np.uint8)
        r = g.search_proteins(q, np.array([0, len(q)], np.uint64), SearchOptions())
        assert r.n_rows == 1 and len(r.subject) == 0 and r.size_in_kmer[0] == 16


def test_dense_collision_quirk_subjects_with_unknown_letters():
    """DB proteins with X/B/Z/U: unknown letters collapse to code 0 exactly as EncodeKmer does."""
    from kaamer_b200 import GpuIndex, SearchOptions
    from oracle import oracle as o

    seqs = [b"MKTAYIAKQRXISFVKSHFSRQLEERLGLIEV", b"MKTAYIAKQRBISFVKSHFSRQLEERLGLIEV", b"UUUUUUUUUUUUUUUUUUUU",
            b"AAAAAAAAAAAAAAAAAAAAAAAA"]
    res, off = o.pack(seqs)
    ids = np.array([5, 9, 11, 12], np.uint32)
    idx = o.Index.build(res, off, ids)
    qs = [b"MKTAYIAKQRZISFVKSHFSRQLEERLGLIEV", b"UUUUUUUUUUUUUUUUUUUUUU", b"XAXAXAXAXAXAXAXAXAXAXA"]
    q, qo = o.pack(qs)
    ora = o.search_proteins(idx, q, qo, o.opts(min_kmatch=1, min_kratio=0.0))
    with GpuIndex.build(res, off, ids) as g:
        k, f, p = g.index_arrays()
        np.testing.assert_array_equal(k, idx.keys)
        np.testing.assert_array_equal(p, idx.postings)
        r = g.search_proteins(q, qo, SearchOptions(min_kmatch=1, min_kratio=0.0))
    assert_same_hits(r, ora, "unknown letters")
    assert len(r.subject) > 0


def test_medium_scale_properties():
    """Size-independent checks at a size the oracle does not index: queries that are exact DB
    records must rank themselves first with Kmatch == SizeInKmer (every k-mer of the record is
    in its own posting list), and results are invariant under batch permutation."""
    from kaamer_b200 import GpuIndex, SearchOptions, synth

    res, off = synth.protein_db(60000, config_index=3)
    ids = np.arange(len(off) - 1, dtype=np.uint32) + 1
    q, qo, pick = synth.protein_queries(res, off, 5000, config_index=3, sub_rate=0.0)
    with GpuIndex.build(res, off, ids) as g:
        r = g.search_proteins(q, qo, SearchOptions(max_results=5))
        top = r.hit_off[:-1].astype(np.int64)
        assert (np.diff(r.hit_off.astype(np.int64)) >= 1).all()
        assert (r.kmatch[top] == r.size_in_kmer).all()
        # the query's own id is among the hits tied at the maximum (family members may tie)
        for i in range(0, 5000, 97):
            hits = r.hits(i)
            assert (int(ids[pick[i]]), int(r.size_in_kmer[i])) in hits
        # permutation invariance
        perm = np.random.default_rng(1).permutation(5000)
        lens = np.diff(qo.astype(np.int64))[perm]
        qo2 = np.zeros(5001, np.uint64)
        qo2[1:] = np.cumsum(lens)
        q2 = np.concatenate([q[int(qo[j]):int(qo[j + 1])] for j in perm])
        r2 = g.search_proteins(q2, qo2, SearchOptions(max_results=5))
        for n, j in enumerate(perm[:300]):
            assert r2.hits(n) == r.hits(int(j))
        assert r2.n_lookups == r.n_lookups and r2.n_increments == r.n_increments


def test_pinned_host_buffers_are_read_in_place(small_db, gpu_small):
    """Page-locked caller buffers take the zero-copy path of kaamer_gpu_search_proteins (kernels
    read the residues over PCIe); results must equal the staged (pageable) path and the oracle."""
    import torch

    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 3000, config_index=1, stream=17)
    seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)]
    seqs += [small_db["res"][:6000].tobytes(), small_db["res"][:50_000].tobytes(), b"", b"MKT*"]  # class G, > smem stage
    q, qo = o.pack(seqs)
    ora = o.search_proteins(small_db["idx"], q, qo, o.opts(), 4)
    hq = torch.from_numpy(q).pin_memory()
    ho = torch.from_numpy(qo.astype(np.int64)).pin_memory()
    r = gpu_small.search_proteins_ptr(hq.data_ptr(), ho.data_ptr(), len(qo) - 1, SearchOptions())
    assert_same_hits(r, ora, "pinned / zero-copy")
    assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments
    r2 = gpu_small.search_proteins(q, qo, SearchOptions())
    assert_same_hits(r2, ora, "pageable / staged")


def test_handle_is_thread_safe(small_db, gpu_small):
    """Several host threads share one handle (calls are serialised inside the library, as the
    goroutines of one kaamer request share the read-only stores, api/server.go:65)."""
    import threading

    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    batches = []
    for t in range(4):
        q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 400, config_index=1, stream=40 + t)
        batches.append((q, qo, o.search_proteins(small_db["idx"], q, qo, o.opts(), 2)))
    errors = []

    def work(t):
        try:
            q, qo, ora = batches[t]
            for _ in range(5):
                r = gpu_small.search_proteins(q, qo, SearchOptions())
                assert_same_hits(r, ora, f"thread {t}")
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors


def test_c1_fasta_pipeline(tmp_path):
    """BASELINE.json configs[0]: `kaamer-db -make` on a synthetic 10 k-protein FASTA, then 1 k protein
    queries with the default options — FASTA -> records/ids (host) -> device index -> search, against
    the oracle's makedb + indexdb + search on the same file."""
    from kaamer_b200 import GpuIndex, SearchOptions, makedb, synth
    from oracle import oracle as o

    res, off = synth.protein_db(10_000, config_index=1)
    names = [f"sp|S{i:06d}|SYN_{i} synthetic protein {i}" for i in range(1, 10_001)]
    fa = str(tmp_path / "c1.fa")
    synth.write_fasta(fa, names, res, off)
    _, _, r_res, r_off, r_ids = makedb.read_fasta(fa)
    q, qo, pick = synth.protein_queries(res, off, 1000, config_index=1, stream=1)
    idx = o.Index.build(r_res, r_off, r_ids, 8)
    ora = o.search_proteins(idx, q, qo, o.opts(), 8)
    with GpuIndex.build(r_res, r_off, r_ids, keep_proteins=True) as g:
        k, f, p = g.index_arrays()
        np.testing.assert_array_equal(k, idx.keys)
        np.testing.assert_array_equal(f, idx.offsets)
        np.testing.assert_array_equal(p, idx.postings)
        assert g.dbstats() == {"NumberOfProteins": idx.n_proteins, "NumberOfAA": idx.n_aa, "NumberOfKmers": idx.n_kmers}
        r = g.search_proteins(q, qo, SearchOptions())
        assert_same_hits(r, ora, "C1")
        assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments
        # the source record of every query is its best hit (ids follow the FASTA id quirk)
        has = np.diff(r.hit_off.astype(np.int64)) > 0
        assert has.mean() > 0.99
        top = r.subject[r.hit_off[:-1].astype(np.int64)[has]]
        assert (top == r_ids[pick][has]).mean() > 0.98
        # .kidx round trip of the C1 index
        path = str(tmp_path / "c1.kidx")
        g.save(path)
    with GpuIndex.open(path) as g2:
        r2 = g2.search_proteins(q, qo, SearchOptions())
        assert_same_hits(r2, ora, "C1 from .kidx")


@pytest.mark.parametrize("n_sharing", [700, 3300, 6000])
def test_many_candidates_per_query_walk_down_the_size_classes(n_sharing):
    """n proteins share one 600-residue core and the options let every subject through: the warp
    class overflows into the CTA class (64 candidates), the CTA class into the global-memory class
    (3040 candidates / 4096 histogram slots) — results stay bit-exact and nothing is counted twice."""
    from kaamer_b200 import GpuIndex, SearchOptions
    from oracle import oracle as o

    rng = np.random.default_rng(5)
    aa = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)
    core = aa[rng.integers(0, 20, 600)].tobytes()
    prots = [core + aa[rng.integers(0, 20, 20)].tobytes() for _ in range(n_sharing)]
    prots += [aa[rng.integers(0, 20, 300)].tobytes() for _ in range(200)]
    res, off = o.pack(prots)
    ids = np.arange(1, len(prots) + 1, dtype=np.uint32)
    idx = o.Index.build(res, off, ids, 4)
    qs = [core, core[:400], core[100:], core[:300] + prots[-1][:250], prots[-2]]
    q, qo = o.pack(qs)
    with GpuIndex.build(res, off, ids, keep_proteins=False) as g:
        for opts in (SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=10_000), SearchOptions(),
                     SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=10_000, extract_positions=True)):
            ora = o.search_proteins(idx, q, qo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results,
                                                       want_positions=opts.extract_positions), 4)
            r = g.search_proteins(q, qo, opts)
            assert_same_hits(r, ora, f"{n_sharing} sharing, {opts}")
            assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments
            if opts.extract_positions:
                np.testing.assert_array_equal(r.pos, ora.pos)


@pytest.mark.parametrize("limits", ["0,0", "40,120", "100,2048", "0,2048"])
def test_size_class_limits_do_not_change_results(small_db, gpu_small, limits, monkeypatch):
    """The class limits follow the database density (search.cu class_limits); whatever they are, every
    query must come out identical: force them so that ordinary queries run in classes M and G, with
    positions and in nucleotide mode."""
    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    from tests.helpers import assert_same_rows

    monkeypatch.setenv("KAAMER_CLASS_LIMITS", limits)
    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 300, config_index=1, stream=33)
    seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)]
    seqs += [b"", b"MKT", small_db["res"][:700].tobytes(), small_db["res"][:3000].tobytes()]
    q, qo = o.pack(seqs)
    for opts in (SearchOptions(), SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=100, extract_positions=True)):
        ora = o.search_proteins(small_db["idx"], q, qo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results,
                                                               want_positions=opts.extract_positions), 4)
        r = gpu_small.search_proteins(q, qo, opts)
        assert_same_hits(r, ora, f"limits {limits} {opts}")
        assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments
        if opts.extract_positions:
            np.testing.assert_array_equal(r.pos, ora.pos)
    nt, off = synth.nucleotide_contigs(small_db["res"], small_db["off"], 1, 60_000, config_index=2)
    ora = o.search_nucleotide(small_db["idx"], nt, off, o.opts(), 4)
    assert_same_rows(gpu_small.search_nucleotide(nt, off, SearchOptions()), ora, f"limits {limits} nucleotide")


def test_fasta_file_to_hits_through_the_native_reader(small_db, gpu_small, tmp_path):
    """query FASTA -> kaamer_host_read_fasta (pinned batch) -> kaamer_gpu_search_proteins, against the
    oracle on the reader transliteration's view of the same file (last record not upper-cased)"""
    from kaamer_b200 import SearchOptions, readers, synth
    from oracle import oracle as o
    from tests import go_transliteration as go

    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 50, config_index=1, stream=41)
    p = str(tmp_path / "queries.fasta")
    synth.write_fasta(p, [f"q{i} query {i}" for i in range(50)], q, qo)
    with open(p, "ab") as f:
        f.write(b">lower case last record\n" + small_db["res"][:300].tobytes().lower() + b"\n")
    b = readers.read_fasta(p, pinned=True)
    ref = go.get_queries_fasta(p)
    assert len(b) == 51 and [r[0] for r in ref] == b.names
    rq, rqo = o.pack([r[1].encode("latin-1") for r in ref])
    np.testing.assert_array_equal(b.residues, rq)
    ora = o.search_proteins(small_db["idx"], rq, rqo, o.opts(), 4)
    r = gpu_small.search_proteins(b.residues, b.seq_off, SearchOptions())
    assert_same_hits(r, ora, "fasta file")
    np.testing.assert_array_equal(r.size_in_kmer, b.size_in_kmer)
    assert r.hits(50) == []  # lower-case residues are unknown letters: the last record finds nothing


def test_submit_wait_pipeline_equals_the_blocking_call(small_db, gpu_small):
    """kaamer_gpu_search_proteins_submit / _wait: two batches in flight, results identical to the blocking call
    and to the oracle; a third submit is refused; pageable and pinned inputs; refused batch kinds"""
    import torch

    from kaamer_b200 import KaamerGpuError, SearchOptions, synth
    from oracle import oracle as o

    opts = SearchOptions()
    batches = []
    for b in range(5):
        q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 300 + 50 * b, config_index=1, stream=40 + b)
        if b % 2:  # page-locked inputs on every second batch
            q = torch.from_numpy(q).pin_memory().numpy()
            qo = torch.from_numpy(qo.astype(np.int64)).pin_memory().numpy().view(np.uint64)
        batches.append((q, np.ascontiguousarray(qo, dtype=np.uint64)))
    expected = [o.search_proteins(small_db["idx"], q, qo, o.opts(), 4) for q, qo in batches]
    t = gpu_small.submit_proteins_ptr(batches[0][0].ctypes.data, batches[0][1].ctypes.data, len(batches[0][1]) - 1, opts)
    for b in range(1, 5):
        t2 = gpu_small.submit_proteins_ptr(batches[b][0].ctypes.data, batches[b][1].ctypes.data, len(batches[b][1]) - 1, opts)
        if b == 1:
            with pytest.raises(KaamerGpuError):  # two in flight already
                gpu_small.submit_proteins_ptr(batches[2][0].ctypes.data, batches[2][1].ctypes.data, 10, opts)
        r = gpu_small.wait_proteins(t)
        assert_same_hits(r, expected[b - 1], f"pipelined batch {b - 1}")
        assert r.n_lookups == expected[b - 1].n_lookups and r.n_increments == expected[b - 1].n_increments
        t = t2
    assert_same_hits(gpu_small.wait_proteins(t), expected[4], "pipelined batch 4")
    with pytest.raises(KaamerGpuError):
        gpu_small.wait_proteins(t)  # nothing in flight in that slot any more
    q, qo = batches[0]
    with pytest.raises(KaamerGpuError):  # positions cannot be bounded beforehand: blocking call only
        gpu_small.submit_proteins_ptr(q.ctypes.data, qo.ctypes.data, len(qo) - 1, SearchOptions(extract_positions=True))
    # the blocking call (submit + wait inside) after the pipeline, and a MaxResults the pool must grow for
    assert_same_hits(gpu_small.search_proteins(q, qo, opts), expected[0], "blocking after pipeline")
    big = SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=100)
    ora = o.search_proteins(small_db["idx"], q, qo, o.opts(min_kmatch=1, min_kratio=0.0, max_results=100), 4)
    tk = gpu_small.submit_proteins_ptr(q.ctypes.data, qo.ctypes.data, len(qo) - 1, big)
    assert_same_hits(gpu_small.wait_proteins(tk), ora, "pool overflow inside a pipelined batch")
