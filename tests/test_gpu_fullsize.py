"""Full-size (BASELINE.json config C3: 570 k-protein DB, 100 k queries) checks through
size-independent properties, plus exact parity with the CPU oracle on a random sample of the
batch searched against the FULL database."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    from kaamer_b200 import GpuIndex, synth
    from kaamer_b200.makedb import fasta_protein_ids

    res, off = synth.protein_db(570_000, config_index=3)
    ids = fasta_protein_ids(len(off) - 1)
    q, qo, pick = synth.protein_queries(res, off, 100_000, config_index=3, stream=100)
    g = GpuIndex.build(res, off, ids, keep_proteins=False)
    yield dict(res=res, off=off, ids=ids, q=q, qo=qo, pick=pick, g=g)
    g.close()


def test_c3_properties(c3):
    from kaamer_b200 import SearchOptions

    g, q, qo = c3["g"], c3["q"], c3["qo"]
    r = g.search_proteins(q, qo, SearchOptions())
    nq = len(qo) - 1
    K = np.diff(qo.astype(np.int64)) - 6
    np.testing.assert_array_equal(r.size_in_kmer, K)
    assert r.n_lookups == int(K[K >= 7].sum())
    nh = np.diff(r.hit_off.astype(np.int64))
    assert nh.max() <= 10 and len(r.subject) == int(r.hit_off[-1])
    # FilterResults: every hit has Kmatch >= 10 and Kmatch / SizeInKmer >= 0.05 (fp64 as the reference)
    kq = np.repeat(K, nh).astype(np.float64)
    km = r.kmatch.astype(np.int64)
    assert (km >= 10).all() and not ((km.astype(np.float64) / kq) < 0.05).any() and (km <= kq).all()
    # ranking: Kmatch descending, subject id ascending inside ties, no duplicate subject per query
    row = np.repeat(np.arange(nq), nh)
    same = row[1:] == row[:-1]
    assert (km[1:][same] <= km[:-1][same]).all()
    tie = same & (km[1:] == km[:-1])
    assert (r.subject[1:][tie] > r.subject[:-1][tie]).all()
    # queries are DB records with 10 % substitutions: the source record is the top hit (or tied with it)
    src = c3["ids"][c3["pick"]]
    first = r.hit_off[:-1].astype(np.int64)
    has = nh > 0
    assert has.mean() > 0.999
    top_km = np.zeros(nq, np.int64)
    top_km[has] = km[first[has]]
    src_is_hit = np.zeros(nq, bool)
    hit_is_src = r.subject == np.repeat(src, nh)
    src_is_hit[row[hit_is_src]] = True
    assert src_is_hit[has].mean() > 0.999
    # increments: every hit's count is part of the increment total
    assert r.n_increments >= int(km.sum())


def test_c3_pinned_zero_copy_equals_staged(c3):
    import torch

    from kaamer_b200 import SearchOptions

    g, q, qo = c3["g"], c3["q"], c3["qo"]
    a = g.search_proteins(q, qo, SearchOptions())
    hq, ho = torch.from_numpy(q).pin_memory(), torch.from_numpy(qo.astype(np.int64)).pin_memory()
    b = g.search_proteins_ptr(hq.data_ptr(), ho.data_ptr(), len(qo) - 1, SearchOptions())
    for x, y in ((a.hit_off, b.hit_off), (a.subject, b.subject), (a.kmatch, b.kmatch), (a.size_in_kmer, b.size_in_kmer)):
        np.testing.assert_array_equal(x, y)
    assert (a.n_lookups, a.n_increments) == (b.n_lookups, b.n_increments)


def test_c3_sample_equals_oracle_on_full_db(c3):
    from kaamer_b200 import SearchOptions
    from oracle import oracle as o
    from tests.helpers import assert_same_hits

    idx = o.Index.build(c3["res"], c3["off"], c3["ids"], 16)
    rng = np.random.default_rng(2)
    sel = np.sort(rng.choice(100_000, 3000, replace=False))
    qo, q = c3["qo"], c3["q"]
    seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in sel]
    sq, sqo = o.pack(seqs)
    for opts in (dict(), dict(min_kmatch=1, min_kratio=0.0, max_results=25)):
        ora = o.search_proteins(idx, sq, sqo, o.opts(**opts), 16)
        r = c3["g"].search_proteins(sq, sqo, SearchOptions(max_results=opts.get("max_results", 10),
                                                         min_kmatch=opts.get("min_kmatch", 10),
                                                         min_kratio=opts.get("min_kratio", 0.05)))
        assert_same_hits(r, ora, f"C3 sample {opts}")
        assert (r.n_lookups, r.n_increments) == (ora.n_lookups, ora.n_increments)
