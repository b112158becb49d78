"""GPU parity tests (through the C ABI) of the Smith-Waterman re-alignment stage vs the CPU
oracle: integer fields bit-exact (DP optimum, raw score, alignment length, mismatches, gap
openings, coordinates), float32 identity/similarity and float64 bitscore/e-value within 1e-6
relative (BASELINE.json north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-6  # stated tolerance for %identity, bitscore, e-value
INT_FIELDS = ["length", "mismatches", "gap_openings", "raw", "query_start", "query_end", "subject_start",
              "subject_end", "dp_score"]


def _check(gpu_rows, pairs, queries, subjects, prm, what=""):
    from oracle import oracle as o

    for k, (qi, sid) in enumerate(pairs):
        exp = o.align(queries[qi], subjects[sid], prm)
        g = gpu_rows[k]
        for f in INT_FIELDS:
            assert int(g[f]) == int(getattr(exp, f)), f"{what} pair {k} (q{qi}, s{sid}): {f} {g[f]} != {getattr(exp, f)}"
        assert int(g["status"]) == int(exp.illegal), f"{what} pair {k}: status"
        for f in ("identity", "similarity", "bitscore", "evalue"):
            a, b = float(g[f]), float(getattr(exp, f))
            if np.isnan(b):
                assert np.isnan(a), f"{what} pair {k}: {f} {a} vs NaN"
            elif np.isinf(b):
                assert a == b, f"{what} pair {k}: {f} {a} != {b}"
            else:
                assert abs(a - b) <= RTOL * abs(b), f"{what} pair {k}: {f} {a} != {b}"


def _subjects_by_id(res, off, ids):
    """protein_store[id] = last accepted record written with that id."""
    out = {}
    for i, pid in enumerate(ids.tolist()):
        s = res[int(off[i]):int(off[i + 1])].tobytes()
        if len(s) >= 7:
            out[pid] = s
    return out


def test_align_top_hits_of_a_search(small_db):
    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from oracle import oracle as o

    res, off, ids = small_db["res"], small_db["off"], small_db["ids"]
    subjects = _subjects_by_id(res, off, ids)
    q, qo, _ = synth.protein_queries(res, off, 120, config_index=1, stream=21)
    queries = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)]
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        r = g.search_proteins(q, qo, SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=4))
        pairs = [(i, int(s)) for i in range(r.n_rows) for s in r.subject[int(r.hit_off[i]):int(r.hit_off[i + 1])]]
        assert len(pairs) > 200
        n_aa = g.dbstats()["NumberOfAA"]
        out = g.align(q, qo, [p[0] for p in pairs], [p[1] for p in pairs])
        _check(out, pairs, queries, subjects, o.aln_params(n_aa), "top hits")
        # a different GapOpen option only changes the `score == -GapOpen` test (align.go:127)
        out = g.align(q, qo, [p[0] for p in pairs[:60]], [p[1] for p in pairs[:60]], gap_open=10, gap_extend=2,
                      lambda_=0.3, K=0.1, number_of_aa=12345678)
        _check(out, pairs[:60], queries, subjects, o.aln_params(12345678, 0.3, 0.1, 10, 2), "opts")
        assert (out["gap_openings"] == 0).all()


def test_align_edge_cases():
    from kaamer_b200 import GpuIndex
    from oracle import oracle as o

    rng = np.random.default_rng(5)
    aa = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)

    def rnd(n):
        return aa[rng.integers(0, 20, n)].tobytes()

    def mutate(s, rate, indel=0.02):
        out = bytearray()
        for c in s:
            u = rng.random()
            if u < indel:
                continue
            if u < 2 * indel:
                out += rnd(int(rng.integers(1, 6)))
            out.append(c if rng.random() > rate else int(aa[rng.integers(0, 20)]))
        return bytes(out)

    base = rnd(300)
    long_a = rnd(1400)
    subj_list = [
        base, mutate(base, 0.2), rnd(7), b"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ", b"AAAAAAAAAAAAAAAAAAAAAAAA",
        b"MKTUUIAKQRQISFuKSHFSRQ*", b"MKTAYIAKQROISFVKSHFSRQ", b"mktayiakqrqisfvkshfsrq", b"MKT-YIAKQ-QISFVKSHFSRQ",
        long_a, mutate(long_a, 0.3, 0.01), rnd(129), rnd(257), rnd(600), b"WWWWWWWCCCCCCCWWWWWWW",
        b"BZXJBZXJBZXJ*BZXJ", rnd(5000),
    ]
    ids = np.arange(3, 3 + len(subj_list), dtype=np.uint32)
    res, off = o.pack(subj_list)
    subjects = {int(i): s for i, s in zip(ids, subj_list)}
    queries = [
        base, mutate(base, 0.1), b"", b"A", b"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ", b"MKTAYIAKQRQISFVKSHFSRQ1",
        b"mktayiakqrqisfvkshfsrq", b"MKTUYIAKQRQISFVKSHFSRQ", b"MKT-YIAKQRQISFVKSHFSRQ", long_a[100:1300],
        mutate(long_a, 0.15, 0.03), rnd(40), b"WWWWWWWWWWWWWW", b"BZXJBZXJBZXJ*BZXJ", rnd(2500), b"O" * 30,
    ]
    q, qo = o.pack(queries)
    pairs = [(i, int(s)) for i in range(len(queries)) for s in ids]
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        out = g.align(q, qo, [p[0] for p in pairs], [p[1] for p in pairs], number_of_aa=3_500_000)
        _check(out, pairs, queries, subjects, o.aln_params(3_500_000), "edge")
        assert out["status"].sum() > 0 and (out["gap_openings"] > 0).sum() > 3
        # empty pair list, unknown subject id, index without proteins
        assert len(g.align(q, qo, [], [])) == 0
        from kaamer_b200 import KaamerGpuError
        with pytest.raises(KaamerGpuError):
            g.align(q, qo, [0], [10_000])
    with GpuIndex.build(res, off, ids, keep_proteins=False) as g:
        with pytest.raises(KaamerGpuError):
            g.align(q, qo, [0], [3])


def _edge_inputs():
    """the pair set of test_align_edge_cases (mutated relatives with indels, illegal letters, U / - / *,
    lower case, multi-block subjects, a 5000-aa subject)"""
    from oracle import oracle as o

    rng = np.random.default_rng(5)
    aa = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)

    def rnd(n):
        return aa[rng.integers(0, 20, n)].tobytes()

    def mutate(s, rate, indel=0.02):
        out = bytearray()
        for c in s:
            u = rng.random()
            if u < indel:
                continue
            if u < 2 * indel:
                out += rnd(int(rng.integers(1, 6)))
            out.append(c if rng.random() > rate else int(aa[rng.integers(0, 20)]))
        return bytes(out)

    base = rnd(300)
    long_a = rnd(1400)
    subj_list = [base, mutate(base, 0.2), rnd(7), b"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ", b"AAAAAAAAAAAAAAAAAAAAAAAA",
                 b"MKTUUIAKQRQISFuKSHFSRQ*", b"MKTAYIAKQROISFVKSHFSRQ", b"mktayiakqrqisfvkshfsrq", b"MKT-YIAKQ-QISFVKSHFSRQ",
                 long_a, mutate(long_a, 0.3, 0.01), rnd(129), rnd(257), rnd(600), b"WWWWWWWCCCCCCCWWWWWWW",
                 b"BZXJBZXJBZXJ*BZXJ", rnd(3000)]
    ids = np.arange(3, 3 + len(subj_list), dtype=np.uint32)
    res, off = o.pack(subj_list)
    subjects = {int(i): s for i, s in zip(ids, subj_list)}
    queries = [base, mutate(base, 0.1), b"", b"A", b"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ", b"MKTAYIAKQRQISFVKSHFSRQ1",
               b"mktayiakqrqisfvkshfsrq", b"MKTUYIAKQRQISFVKSHFSRQ", b"MKT-YIAKQRQISFVKSHFSRQ", long_a[100:1300],
               mutate(long_a, 0.15, 0.03), rnd(40), b"WWWWWWWWWWWWWW", b"BZXJBZXJBZXJ*BZXJ", rnd(2500), b"O" * 30]
    q, qo = o.pack(queries)
    pairs = [(i, int(s)) for i in range(len(queries)) for s in ids]
    return res, off, ids, subjects, queries, q, qo, pairs


def test_aln_string_equals_the_reference_layout():
    """AlignmentResult.AlnString (align.go:69-103): gapped query, match line (letter / '+' / ' '), gapped
    subject, joined by newlines — for every pair, including empty alignments ("\\n\\n") and illegal letters"""
    from kaamer_b200 import GpuIndex
    from oracle import oracle as o

    res, off, ids, subjects, queries, q, qo, pairs = _edge_inputs()
    prm = o.aln_params(3_500_000)
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        out, texts = g.align(q, qo, [p[0] for p in pairs], [p[1] for p in pairs], number_of_aa=3_500_000, want_text=True)
        _check(out, pairs, queries, subjects, prm, "with text")
        plain = g.align(q, qo, [p[0] for p in pairs], [p[1] for p in pairs], number_of_aa=3_500_000)
        assert out.tobytes() == plain.tobytes() or np.array_equal(out["raw"], plain["raw"])
        assert g.align(q, qo, [], [], want_text=True)[1] == []
    n_gapped = 0
    for k, (qi, sid) in enumerate(pairs):
        exp = o.aln_string(queries[qi], subjects[sid], prm)
        assert texts[k] == exp, f"pair {k} (q{qi}, s{sid}):\n{texts[k].decode()}\n!=\n{exp.decode()}"
        assert len(texts[k]) == 3 * int(out["length"][k]) + 2
        n_gapped += b"-" in exp
    assert n_gapped > 10 and any(t == b"\n\n" for t in texts)


@pytest.mark.parametrize("model", ["gap_row_minus4", "gap_row_ragged", "pam_like", "zero_gap_other_matrix", "default_again"])
def test_alignment_model_is_a_runtime_parameter(model):
    """whichever gap row biogo's BLOSUM62 carries, and §8f-4's other matrices, are a kaamer_gpu_set_align_model
    call away: the general cell update against the oracle under the same model, bit-exact"""
    from kaamer_b200 import GpuIndex
    from oracle import oracle as o

    res, off, ids, subjects, queries, q, qo, pairs = _edge_inputs()
    pairs = [p for p in pairs if len(queries[p[0]]) * len(subjects[p[1]]) < 2_500_000][:150] + pairs[-17:]
    prm = o.aln_params(3_500_000)
    m0, open0 = GpuIndex.default_align_model()
    assert np.array_equal(m0.astype(np.int32), o.blosum62()) and open0 == -11
    m = m0.astype(np.int32).copy()
    gap_open = -11
    if model == "gap_row_minus4":
        m[0, 1:] = -4
        m[1:, 0] = -4
    elif model == "gap_row_ragged":
        rng = np.random.default_rng(3)
        m[0, 1:] = -rng.integers(0, 4, 25)
        m[1:, 0] = -rng.integers(0, 4, 25)
        gap_open = -7
    elif model == "pam_like":
        rng = np.random.default_rng(4)
        sym = rng.integers(-6, 7, (26, 26))
        m = np.triu(sym) + np.triu(sym, 1).T
        m[np.arange(1, 26), np.arange(1, 26)] = rng.integers(3, 12, 25)
        m[0, :] = -1
        m[:, 0] = -1
        m[0, 0] = 0
        gap_open = -9
    elif model == "zero_gap_other_matrix":
        # another matrix and gap open WITH a zero gap row: the packed int16x2 kernel takes it (its score bound
        # follows the largest entry)
        rng = np.random.default_rng(6)
        sym = rng.integers(-6, 7, (26, 26))
        m = np.triu(sym) + np.triu(sym, 1).T
        m[np.arange(1, 26), np.arange(1, 26)] = rng.integers(5, 18, 25)
        m[0, :] = 0
        m[:, 0] = 0
        gap_open = -7
    try:
        o.set_align_model(m if model != "default_again" else None, gap_open)
        with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
            if model == "default_again":
                g.set_align_model(m, -3)
                g.set_align_model(None)
            else:
                g.set_align_model(m, gap_open)
            out, texts = g.align(q, qo, [p[0] for p in pairs], [p[1] for p in pairs], number_of_aa=3_500_000, want_text=True)
            # packed jobs exactly when the model has a zero gap row
            assert (g.align_last_plan()[2] > 0) == (model in ("zero_gap_other_matrix", "default_again")), g.align_last_plan()
            _check(out, pairs, queries, subjects, prm, model)
            for k, (qi, sid) in enumerate(pairs):
                assert texts[k] == o.aln_string(queries[qi], subjects[sid], prm), (model, k)
            if model == "gap_row_minus4":
                # biogo-with-a-gap-row consequence: no segment scores exactly -GapOpen, so kaamer's test
                # (align.go:127) never fires
                assert (out["gap_openings"] == 0).all() and (out["length"] > 0).any()
    finally:
        o.set_align_model(None, 0)


def _with_env(**kv):
    import contextlib
    import os

    @contextlib.contextmanager
    def cm():
        old = {k: os.environ.get(k) for k in kv}
        try:
            for k, v in kv.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = str(v)
            yield
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    return cm()


def test_packed_int16x2_jobs_equal_the_32bit_kernels_and_the_oracle(small_db):
    """k_sw_affine_pk (two pairs per warp, int16x2 lanes, DPX) is the default for pairs it can hold; every column
    width and the 32-bit kernels must give the same AlignmentResult and AlnString for every pair, and those
    equal the oracle's"""
    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from oracle import oracle as o

    res, off, ids = small_db["res"], small_db["off"], small_db["ids"]
    subjects = _subjects_by_id(res, off, ids)
    q, qo, _ = synth.protein_queries(res, off, 300, config_index=1, stream=33)
    queries = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)]
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        r = g.search_proteins(q, qo, SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=6))
        pairs = [(i, int(s)) for i in range(r.n_rows) for s in r.subject[int(r.hit_off[i]):int(r.hit_off[i + 1])]]
        assert len(pairs) > 400
        pq, ps = [p[0] for p in pairs], [p[1] for p in pairs]
        n_aa = g.dbstats()["NumberOfAA"]
        with _with_env(KAAMER_ALIGN_PACKED=0):
            ref, ref_text = g.align(q, qo, pq, ps, number_of_aa=n_aa, want_text=True)
            assert g.align_last_plan()[2] == 0
        _check(ref[:250], pairs[:250], queries, subjects, o.aln_params(n_aa), "32-bit kernels")
        assert (ref["gap_openings"] > 0).sum() > 5 and (ref["length"] > 50).sum() > 100
        for maxcw in (None, 4, 8):
            with _with_env(KAAMER_ALIGN_PACKED=None, KAAMER_ALIGN_PK_MAXCW=maxcw):
                out, text = g.align(q, qo, pq, ps, number_of_aa=n_aa, want_text=True)
                big, single, jobs = g.align_last_plan()
            assert jobs > len(pairs) // 4, (maxcw, big, single, jobs)
            assert big + single + 2 * jobs == len(pairs)
            for f in ref.dtype.names:
                same = (out[f] == ref[f]) | ((out[f] != out[f]) & (ref[f] != ref[f]))  # NaN == NaN
                assert same.all(), (maxcw, f, int(np.flatnonzero(~same)[0]), out[f][~same][:3], ref[f][~same][:3])
            assert text == ref_text, maxcw
        # an odd number of pairs, two pairs, one pair (never packed)
        for n in (7, 2, 1):
            out = g.align(q, qo, pq[:n], ps[:n], number_of_aa=n_aa)
            assert out.tobytes() == ref[:n].tobytes() or all((out[f] == ref[f][:n]).all() for f in INT_FIELDS)
        assert g.align_last_plan()[2] == 0


def test_packed_jobs_on_the_edge_cases():
    """the edge pair set (illegal letters, U / - / *, lower case, multi-block and long subjects, saturating W runs)
    with the packed kernel taking pairs of up to 8 M cells: equal to the 32-bit kernels pair by pair"""
    from kaamer_b200 import GpuIndex

    res, off, ids, subjects, queries, q, qo, pairs = _edge_inputs()
    pq, ps = [p[0] for p in pairs], [p[1] for p in pairs]
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        with _with_env(KAAMER_ALIGN_PACKED=0):
            ref, ref_text = g.align(q, qo, pq, ps, number_of_aa=3_500_000, want_text=True)
        for cells in (None, 8 << 20):
            for maxcw in (4, 8):
                with _with_env(KAAMER_ALIGN_PK_CELLS=cells, KAAMER_ALIGN_PK_MAXCW=maxcw):
                    out, text = g.align(q, qo, pq, ps, number_of_aa=3_500_000, want_text=True)
                    plan = g.align_last_plan()
                assert plan[2] > 20, plan
                for f in ref.dtype.names:
                    same = (out[f] == ref[f]) | ((out[f] != out[f]) & (ref[f] != ref[f]))
                    assert same.all(), (cells, maxcw, f, int(np.flatnonzero(~same)[0]))
                assert text == ref_text
