"""Small end-to-end case for `compute-sanitizer --tool memcheck`: every search size class, positions,
translated search, alignment, mode S steps and mode P (owner probes) on a 1500-protein database,
checked against the CPU oracle.  Run:  compute-sanitizer --tool memcheck python profiles/tools/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from kaamer_b200 import GpuIndex, SearchOptions, synth  # noqa: E402
from kaamer_b200.peer import attach_all  # noqa: E402
from kaamer_b200.sharded import make_fences  # noqa: E402
from oracle import oracle as o  # noqa: E402

res, off = synth.protein_db(1500, config_index=1)
ids = o.fasta_ids(len(off) - 1)
idx = o.Index.build(res, off, ids, 4)
q, qo, _ = synth.protein_queries(res, off, 200, config_index=1, stream=3)
seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)]
seqs += [b"", b"MKT", res[:700].tobytes(), res[:3000].tobytes(), res[2000:12000].tobytes()]
q, qo = o.pack(seqs)
full = os.environ.get("SANITIZE_PEER", "1") == "1"
with GpuIndex.build(res, off, ids, device=0) as g:
    for opts in (SearchOptions(), SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=200, extract_positions=True)):
        ora = o.search_proteins(idx, q, qo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results,
                                                   want_positions=opts.extract_positions), 4)
        r = g.search_proteins(q, qo, opts)
        assert np.array_equal(r.subject, ora.subject) and np.array_equal(r.kmatch.astype(np.int64), ora.kmatch)
    nt, noff = synth.nucleotide_contigs(res, off, 1, 40_000, config_index=2)
    ora = o.search_nucleotide(idx, nt, noff, o.opts(), 4)
    r = g.search_nucleotide(nt, noff, SearchOptions())
    assert np.array_equal(r.subject, ora.subject) and np.array_equal(r.pos, ora.pos)
    pq = np.repeat(np.arange(20, dtype=np.uint32), 2)
    ps = np.array([int(ids[i % len(ids)]) for i in range(40)], dtype=np.uint32)
    aln = g.align(q, qo, pq, ps)
    assert len(aln) == 40
print("single index ok", flush=True)
if full:
    fences = make_fences(idx.keys, idx.offsets, 2)
    hs = [GpuIndex.build(res, off, ids, keep_proteins=False, shard=(int(fences[r]), int(fences[r + 1]))) for r in range(2)]
    attach_all(hs)
    ora = o.search_proteins(idx, q, qo, o.opts(), 4)
    for g in hs:
        r = g.search_proteins(q, qo, SearchOptions())
        assert np.array_equal(r.subject, ora.subject)
    for g in hs:
        g.detach_shards()
    for g in hs:
        g.close()
    print("mode P ok", flush=True)
print("sanitize case ok")
