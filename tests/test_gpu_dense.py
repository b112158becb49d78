"""GPU parity tests of the dense-database path: the on-device synthetic generator, the streaming index
builder and the class-D search kernel (search_dense.cuh), all through the C ABI, against the CPU oracle."""
import numpy as np
import pytest

from tests.helpers import assert_same_hits, assert_same_rows

pytestmark = pytest.mark.gpu
SEED = 20261023


def dense_db(n=1500, letters=b"ACDE", seed=7, mean_len=300):
    """a few letters only: 4^7 = 16 k possible k-mers, so every k-mer has dozens of postings (the C4
    regime at a size the oracle searches in seconds); 20 % family members give real hits"""
    rng = np.random.default_rng(seed)
    seqs = []
    for i in range(n):
        if i > 10 and rng.random() < 0.2:
            s = bytearray(seqs[int(rng.integers(0, i))])
            for p in np.flatnonzero(rng.random(len(s)) < 0.1):
                s[p] = letters[int(rng.integers(0, len(letters)))]
            seqs.append(bytes(s))
        else:
            ln = int(rng.integers(mean_len // 3, mean_len * 2))
            seqs.append(bytes(np.frombuffer(letters, np.uint8)[rng.integers(0, len(letters), ln)]))
    return seqs


@pytest.fixture(scope="module")
def dense_setup():
    from kaamer_b200 import GpuIndex
    from oracle import oracle as o

    seqs = dense_db()
    res, off = o.pack(seqs)
    ids = np.arange(len(seqs), dtype=np.uint32)
    idx = o.Index.build(res, off, ids, 4)
    g = GpuIndex.build(res, off, ids, device=0)
    rng = np.random.default_rng(11)
    qs = []
    for i in range(160):
        s = bytearray(seqs[int(rng.integers(0, len(seqs)))])
        for p in np.flatnonzero(rng.random(len(s)) < 0.1):
            s[p] = b"ACDE"[int(rng.integers(0, 4))]
        qs.append(bytes(s))
    qs += [b"ACDEACDEACDEACDEACDE", b"A" * 400, b"ACD" * 700, seqs[3] * 4, b"WWWWWWWWWWWWWWWW", b"ACDEACD", b""]
    yield {"seqs": seqs, "idx": idx, "g": g, "q": o.pack(qs)}
    g.close()


@pytest.mark.parametrize("design", ["1", "12", "11"])  # KAAMER_DENSE: class D, and its second / first designs (kept for A/B)
@pytest.mark.parametrize("opts", [
    dict(),
    dict(min_kmatch=3, min_kratio=0.0, max_results=10),     # thr = 1, one streaming warp
    dict(min_kmatch=5, min_kratio=0.0, max_results=1000),   # two streaming warps, > 64 candidates
    dict(min_kmatch=9, min_kratio=0.0, max_results=4),      # four streaming warps, thr = 1
    dict(min_kmatch=17, min_kratio=0.0, max_results=4),     # four warps x two lists at a time
    dict(min_kmatch=33, min_kratio=0.0, max_results=50),    # four warps x four lists at a time
    dict(min_kmatch=1, min_kratio=0.0, max_results=7),      # kmin < 3: class G
    dict(min_kmatch=40, min_kratio=0.3, max_results=3),
    dict(max_results=0),
])
def test_class_d_parity_on_a_dense_database(dense_setup, opts, design, monkeypatch):
    from kaamer_b200 import SearchOptions
    from oracle import oracle as o

    monkeypatch.setenv("KAAMER_DENSE", design)
    q, qo = dense_setup["q"]
    ora = o.search_proteins(dense_setup["idx"], q, qo, o.opts(**opts), 4)
    r = dense_setup["g"].search_proteins(q, qo, SearchOptions(max_results=opts.get("max_results", 10),
                                                                 min_kmatch=opts.get("min_kmatch", 10),
                                                                 min_kratio=opts.get("min_kratio", 0.05)))
    assert_same_hits(r, ora, str(opts))
    assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments
    assert r.n_increments > 20 * r.n_lookups  # the regime: dozens of postings per lookup


@pytest.mark.parametrize("design,var,val", [("11", "KAAMER_D_MAPKB", "1"), ("11", "KAAMER_D_MAPKB", "64"),
                                            ("12", "KAAMER_E_MAPW", "64,64"), ("12", "KAAMER_E_MAPW", "128,512"),
                                            ("12", "KAAMER_E_MAPW", "2048,4096"), ("1", "KAAMER_F_MAPW", "64,64,64"),
                                            ("1", "KAAMER_F_MAPW", "128,512,256"), ("1", "KAAMER_F_MAPW", "4096,2048,2048")])
def test_class_d_map_sizes(dense_setup, design, var, val, monkeypatch):
    """tiny maps (everything collides: pushes overflow H, class G takes the query) to large ones"""
    from kaamer_b200 import SearchOptions
    from oracle import oracle as o

    mapkb = f"{var}={val}"
    monkeypatch.setenv("KAAMER_DENSE", design)
    monkeypatch.setenv(var, val)
    q, qo = dense_setup["q"]
    ora = o.search_proteins(dense_setup["idx"], q, qo, o.opts(), 4)
    r = dense_setup["g"].search_proteins(q, qo, SearchOptions())
    assert_same_hits(r, ora, f"map {mapkb} KB")
    assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments


def test_class_d_equals_the_other_classes_on_a_sparse_database(small_db, monkeypatch):
    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from oracle import oracle as o

    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 600, config_index=1, stream=9)
    ora = o.search_proteins(small_db["idx"], q, qo, o.opts(), 4)
    with GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], device=0) as g:
        monkeypatch.setenv("KAAMER_DENSE", "1")
        r = g.search_proteins(q, qo, SearchOptions())
        rp = g.search_proteins(q, qo, SearchOptions(extract_positions=True))
    assert_same_hits(r, ora, "dense kernel on a sparse database")
    assert_same_hits(rp, ora, "dense kernel, positions")
    assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments


def test_class_d_nucleotide_mode(dense_setup, monkeypatch):
    """ORF queries through class D: the MinKMatch gate and the tied-best-hit flag of SetBestStartCodon"""
    from kaamer_b200 import SearchOptions
    from oracle import oracle as o

    monkeypatch.setenv("KAAMER_DENSE", "1")
    rng = np.random.default_rng(5)
    codon = {"A": "gct", "C": "tgt", "D": "gat", "E": "gaa"}
    parts = []
    for i in range(12):
        prot = dense_setup["seqs"][int(rng.integers(0, 400))][:200].decode()
        parts.append("atg" + "".join(codon[c] for c in prot) + "taa" + "acgt" * int(rng.integers(3, 12)))
    nt, no = o.pack(["".join(parts).encode(), "".join(parts[::-1]).encode()])
    ora = o.search_nucleotide(dense_setup["idx"], nt, no, o.opts(), 4)
    r = dense_setup["g"].search_nucleotide(nt, no, SearchOptions())
    assert ora.n_rows > 5
    assert_same_rows(r, ora, "class D, nucleotide")


def test_synth_generator_matches_the_cpu_twin():
    import torch

    from kaamer_b200.synthdb import SynthDB
    from oracle import oracle as o

    db = SynthDB(5000, seed=SEED)
    res, off = db.records(100, 700)
    res, off = res.cpu().numpy(), off.cpu().numpy()
    for k in (0, 1, 5, 333, 699):
        assert res[off[k]:off[k + 1]].tobytes() == o.synth_record(SEED, 100 + k), k
    q, qo = db.queries(3, 300, first=40)
    q, qo = q.cpu().numpy(), qo.cpu().numpy()
    for k in (0, 7, 299):
        assert q[qo[k]:qo[k + 1]].tobytes() == o.synth_query(SEED, 5000, 40 + k, 3)[0], k
    torch.cuda.synchronize()


@pytest.mark.parametrize("n_passes,shard", [(1, (0, 0)), (5, (0, 0)), (3, (300_000_000, 1_200_000_000))])
def test_streaming_builder_equals_the_one_shot_build(n_passes, shard, monkeypatch):
    """same records through kaamer_gpu_builder_* (device-generated chunks, several key-range passes) and
    through kaamer_gpu_build_shard: identical tables and postings, judged by the entries every k-mer of a
    query batch finds (hits, counts, lookups, increments) and by KStats"""
    from kaamer_b200 import GpuIndex, SearchOptions
    from kaamer_b200.synthdb import SynthDB
    from oracle import oracle as o

    n = 6000
    db = SynthDB(n, seed=SEED)
    res, off = db.records(0, n)
    res, off = res.cpu().numpy(), off.cpu().numpy().astype(np.uint64)
    ids = np.arange(n, dtype=np.uint32)
    q, qo = db.queries(0, 400)
    q, qo = q.cpu().numpy(), qo.cpu().numpy().astype(np.uint64)
    opts = SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=50)
    whole = shard == (0, 0)
    with db.build_index(shard=shard, n_passes=n_passes, chunk=1700) as gs:
        st = gs.dbstats()
        if whole:
            with GpuIndex.build(res, off, ids, keep_proteins=False, device=0) as g1:
                assert st == g1.dbstats()
                a, b = gs.search_proteins(q, qo, opts), g1.search_proteins(q, qo, opts)
            ora_idx = o.Index.build(res, off, ids, 4)
            ora = o.search_proteins(ora_idx, q, qo, o.opts(min_kmatch=1, min_kratio=0.0, max_results=50), 4)
            assert_same_hits(a, ora, "streaming build vs oracle")
            assert_same_hits(b, ora, "one-shot build vs oracle")
            assert a.n_increments == b.n_increments == ora.n_increments
        else:
            # a key-range shard: attach it together with the two complementary one-shot shards
            lo, hi = shard
            D = 442 ** 3 * 21
            with GpuIndex.build(res, off, ids, keep_proteins=False, device=0, shard=(0, lo)) as ga, \
                    GpuIndex.build(res, off, ids, keep_proteins=False, device=0, shard=(hi, D)) as gc:
                hs = [bytes(x.export_shard()) for x in (ga, gs, gc)]
                gs.attach_shards(hs)
                a = gs.search_proteins(q, qo, opts)
                gs.detach_shards()
            ora_idx = o.Index.build(res, off, ids, 4)
            ora = o.search_proteins(ora_idx, q, qo, o.opts(min_kmatch=1, min_kratio=0.0, max_results=50), 4)
            assert_same_hits(a, ora, "streamed shard between two one-shot shards")
            assert a.n_increments == ora.n_increments


def test_sampled_queries_against_the_restricted_oracle_index(monkeypatch):
    """the C4 parity device at a size that runs in seconds: full streamed index on the GPU, oracle on the
    index restricted to the sampled queries' k-mers (built by streaming the CPU twin)"""
    from kaamer_b200 import SearchOptions
    from kaamer_b200.synthdb import SynthDB
    from oracle import oracle as o

    n = 120_000
    db = SynthDB(n, seed=SEED)
    q, qo = db.queries(2, 2000)
    qh, qoh = q.cpu().numpy(), qo.cpu().numpy().astype(np.uint64)
    sample = list(range(0, 2000, 50))
    seqs = [qh[int(qoh[j]):int(qoh[j + 1])].tobytes() for j in sample]
    ridx = o.synth_restricted_index(SEED, n, seqs, 8)
    sq, sqo = o.pack(seqs)
    ora = o.search_proteins(ridx, sq, sqo, o.opts(), 4)
    for dense in ("0", "1"):
        monkeypatch.setenv("KAAMER_DENSE", dense)
        with db.build_index(n_passes=3, chunk=50_000) as g:
            st = g.dbstats()
            assert (st["NumberOfProteins"], st["NumberOfAA"], st["NumberOfKmers"]) == (ridx.n_proteins, ridx.n_aa, ridx.n_kmers)
            r = g.search_proteins(qh, qoh, SearchOptions())
        for i, j in enumerate(sample):
            assert r.hits(j) == [(s, int(k)) for s, k in ora.hits(i)], (dense, j)
            assert int(r.size_in_kmer[j]) == int(ora.size_in_kmer[i])
    assert len(ora.subject) >= len(sample)


@pytest.mark.parametrize("variant", ["streamed_again", "unsorted_view"])
def test_class_d_second_pass_by_streaming(dense_setup, variant, monkeypatch):
    """the final candidates are normally verified by binary search in the (descending) posting lists; an
    index view whose lists are in another order, or the A/B hook, takes the streaming second pass instead"""
    from kaamer_b200 import GpuIndex, SearchOptions
    from oracle import oracle as o

    monkeypatch.setenv("KAAMER_DENSE", "1")
    q, qo = dense_setup["q"]
    idx = dense_setup["idx"]
    ora = o.search_proteins(idx, q, qo, o.opts(), 4)
    if variant == "streamed_again":
        monkeypatch.setenv("KAAMER_D_NO_BSEARCH", "1")
        r = dense_setup["g"].search_proteins(q, qo, SearchOptions())
    else:
        post = idx.postings.copy()
        rng = np.random.default_rng(2)
        for k in range(len(idx.keys)):  # ascending instead of descending, or shuffled
            b, e = int(idx.offsets[k]), int(idx.offsets[k + 1])
            post[b:e] = post[b:e][::-1] if k % 2 else rng.permutation(post[b:e])
        with GpuIndex.from_arrays(idx.keys, idx.offsets, post, idx.n_proteins, idx.n_aa, idx.n_kmers) as g:
            r = g.search_proteins(q, qo, SearchOptions())
    assert_same_hits(r, ora, variant)
    assert r.n_increments == ora.n_increments
