"""Vectors produced by the REFERENCE ITSELF (go/cmd/kaamer-golden, run on a box with a Go toolchain):
tests/golden/ref_vectors.json.  While that file is absent — this image has no Go — the module is
skipped and parity stays "unpinned" (DESIGN.md §5); once a maintainer commits it, the oracle is checked
against every vector here and the GPU tests keep checking the CUDA path against the oracle."""
import json
import os

import numpy as np
import pytest

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.json")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/ref_vectors.json not generated "
                                "(needs the Go toolchain: go run ./cmd/kaamer-golden in a kaamer checkout)")


@pytest.fixture(scope="module")
def ref():
    return json.load(open(PATH))


def test_blosum62_and_its_gap_row(ref):
    """THE open question of §8c: biogo's gap row.  If it is not zero, the default alignment model must
    change (kaamer_gpu_set_align_model / default_align_model) — this test says so loudly."""
    from oracle import oracle as o

    m = np.array(ref["Blosum62"], dtype=np.int32)
    assert m.shape == (26, 26)
    np.testing.assert_array_equal(m[1:, 1:], o.blosum62()[1:, 1:])
    assert (m[0, :] == 0).all() and (m[:, 0] == 0).all(), \
        f"biogo's BLOSUM62 gap row is {m[0].tolist()}: make it the default model (align.cu default_align_model, oracle Blosum62)"


def test_encode_kmer(ref):
    from oracle import oracle as o

    for kmer, key in ref["EncodeKmer"].items():
        assert o.encode_kmer(kmer.encode()) == key, kmer


def test_orfs(ref):
    from oracle import oracle as o

    for v in ref["ORFs"]:
        got = o.get_orfs(v["DNA"].encode())
        exp = v["ORFs"] or []
        assert len(got.seqs) == len(exp)
        # sort.Slice is unstable: compare as multisets of (Sequence, Start, End, PlusStrand, alternatives)
        a = sorted((got.seqs[i].decode(), int(got.start[i]), int(got.end[i]), bool(got.plus[i]), tuple(got.alts[i]))
                   for i in range(len(exp)))
        b = sorted((e["Sequence"], e["Location"]["StartPosition"], e["Location"]["EndPosition"], e["Location"]["PlusStrand"],
                    tuple(e["Location"]["StartsAlternative"] or [])) for e in exp)
        assert a == b


def test_alignments(ref):
    from oracle import oracle as o

    for v in ref["Alignments"]:
        prm = o.aln_params(v["NumberOfAA"])
        got = o.align(v["Query"].encode(), v["Subject"].encode(), prm)
        r = v["Result"]
        assert (got.raw, got.length, got.mismatches, got.gap_openings) == (r["Raw"], r["Length"], r["Mismatches"], r["GapOpenings"])
        assert (got.query_start, got.query_end, got.subject_start, got.subject_end) == \
               (r["QueryStart"], r["QueryEnd"], r["SubjectStart"], r["SubjectEnd"])
        assert abs(got.bitscore - r["BitScore"]) <= 1e-6 * abs(r["BitScore"])
        assert o.aln_string(v["Query"].encode(), v["Subject"].encode(), prm).decode() == r["AlnString"]


def test_positions(ref):
    from oracle import oracle as o

    cases = {"run_to_end": [0, 1, 1, 0, 1, 1, 1], "single": [0, 1, 0, 0], "all": [1, 1, 1, 1]}
    for name, pos in cases.items():
        assert o.format_positions(pos, False) == ref["Positions"][name]
        assert o.format_positions(pos, True) == ref["Positions"][name + "_aln"]
