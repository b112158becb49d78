"""The C ABI driven from compiled host code (examples/kaamer_search.cpp, built by the library's Makefile):
.kidx file + query FASTA in, the reference's TSV lines out — compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "kaamer_b200", "_build", "kaamer_search")


def test_cpp_host_prints_the_reference_tsv(small_db, tmp_path):
    from kaamer_b200 import GpuIndex, synth
    from oracle import oracle as o
    from tests import go_transliteration as go

    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "kaamer_b200", "csrc"), "-s"])
    kidx = str(tmp_path / "db.kidx")
    with GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], keep_proteins=False) as g:
        g.save(kidx)
    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 120, config_index=1, stream=51)
    fa = str(tmp_path / "q.fasta")
    synth.write_fasta(fa, [f"query{i} some description" for i in range(120)], q, qo)
    for args, kw in (([], dict()), (["50", "1", "0.0"], dict(max_results=50, min_kmatch=1, min_kratio=0.0))):
        out = subprocess.run([BIN, kidx, fa] + args, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        recs = go.get_queries_fasta(fa)
        rq, rqo = o.pack([r[1].encode("latin-1") for r in recs])
        ora = o.search_proteins(small_db["idx"], rq, rqo, o.opts(**kw), 4)
        expect = []
        for i, (name, seq, size) in enumerate(recs):
            for subject, kmatch in ora.hits(i):
                ident = np.float32(kmatch) / np.float32(size) * np.float32(100.0)
                expect.append(f"{name.split(' ')[0]}\t{subject}\t{float(ident):.2f}\t{size}\t{kmatch}\tN/A\t1\t{len(seq)}\t1\tN/A")
        assert out.stdout.splitlines() == expect
        assert len(expect) > 100
    # --pos: ExtractPositions (search.go:521-527,541-544)
    out = subprocess.run([BIN, kidx, fa, "--pos"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    ora = o.search_proteins(small_db["idx"], rq, rqo, o.opts(want_positions=True), 4)
    expect, h = [], 0
    for i, (name, seq, size) in enumerate(recs):
        for subject, kmatch in ora.hits(i):
            ident = np.float32(kmatch) / np.float32(size) * np.float32(100.0)
            ps = go.format_positions_to_string([bool(x) for x in ora.positions(h)], False)
            h += 1
            expect.append(f"{name.split(' ')[0]}\t{subject}\t{float(ident):.2f}\t{size}\t{kmatch}\t{ps.count(',')}\t1\t{len(seq)}"
                          f"\t1\tN/A\t{ps}")
    assert out.stdout.splitlines() == expect
