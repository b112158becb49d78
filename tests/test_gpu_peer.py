"""GPU parity tests of mode P (peer-mapped key-range shards) through the C ABI.
On one GPU the G shards are G shard-range handles on the same device that attach each other by
pointer (the kernels instantiated for peer access run, the transport is local HBM); with >= 2 GPUs
the same search runs under torchrun with one process per GPU, CUDA IPC mappings and NVLink loads."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import assert_same_hits, assert_same_rows

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _peer_handles(small_db, G, dev=0, presence_filter=True, replicate_table=False, replicate_postings=False):
    from kaamer_b200 import GpuIndex
    from kaamer_b200.peer import attach_all
    from kaamer_b200.sharded import make_fences

    idx = small_db["idx"]
    fences = make_fences(idx.keys, idx.offsets, G)
    hs = [GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], keep_proteins=False, device=dev,
                         shard=(int(fences[r]), int(fences[r + 1]))) for r in range(G)]
    attach_all(hs, presence_filter, replicate_table, replicate_postings)
    return hs


@pytest.mark.parametrize("G,presence", [(1, True), (2, True), (2, False), (3, True), (8, True), (8, False),
                                        (2, "replicate"), (3, "replicate"), (2, "replicate_all"), (8, "replicate_all")])
def test_peer_protein_search_parity(small_db, G, presence):
    """every shard handle answers the whole batch exactly like the single index (all size classes:
    short, 700-residue, 3000-residue and 12000-residue queries)"""
    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    idx = small_db["idx"]
    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 400, config_index=1, stream=12)
    seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)]
    seqs += [b"", b"MKT", small_db["res"][:12].tobytes(), b"A" * 700, small_db["res"][:3000].tobytes(),
             small_db["res"][5000:17000].tobytes()]
    q, qo = o.pack(seqs)
    # "replicate_all": built sharded, searched replicated — table and postings of every shard copied at attach, the
    # entries rewritten to offsets into the copy (flat view: the non-PEER search kernels run on the replica)
    hs = _peer_handles(small_db, G, presence_filter=presence is True, replicate_table=presence == "replicate",
                       replicate_postings=presence == "replicate_all")
    try:
        for opts in (SearchOptions(), SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=100),
                     SearchOptions(min_kmatch=4, min_kratio=0.3, max_results=2)):
            ora = o.search_proteins(idx, q, qo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results), 4)
            for r in sorted({0, G - 1}):
                res = hs[r].search_proteins(q, qo, opts)
                assert_same_hits(res, ora, f"G={G} handle {r} {opts}")
                assert res.n_lookups == ora.n_lookups and res.n_increments == ora.n_increments
    finally:
        for g in hs:
            g.detach_shards()
        for g in hs:
            g.close()


@pytest.mark.parametrize("replicate", [False, True, "all"])
def test_peer_positions_and_nucleotide_parity(small_db, replicate):
    """PositionHits, translated search and SetBestStartCodon read the table too (finish.cu)"""
    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    idx = small_db["idx"]
    hs = _peer_handles(small_db, 3, replicate_table=replicate is True, replicate_postings=replicate == "all")
    try:
        q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 200, config_index=1, stream=9)
        ora = o.search_proteins(idx, q, qo, o.opts(want_positions=True), 4)
        r = hs[1].search_proteins(q, qo, SearchOptions(extract_positions=True))
        assert_same_hits(r, ora, "peer positions")
        np.testing.assert_array_equal(r.pos_off.astype(np.int64), ora.pos_off.astype(np.int64))
        np.testing.assert_array_equal(r.pos, ora.pos)
        nt, off = synth.nucleotide_contigs(small_db["res"], small_db["off"], 2, 120_000, config_index=2)
        for kw in (dict(), dict(min_kmatch=1, min_kratio=0.0, max_results=100)):
            ora = o.search_nucleotide(idx, nt, off, o.opts(**kw), 4)
            r = hs[2].search_nucleotide(nt, off, SearchOptions(max_results=kw.get("max_results", 10),
                                                               min_kmatch=kw.get("min_kmatch", 10),
                                                               min_kratio=kw.get("min_kratio", 0.05)))
            assert ora.n_rows > 50
            assert_same_rows(r, ora, f"peer nucleotide {kw}")
            assert r.n_lookups == ora.n_lookups and r.n_increments == ora.n_increments
    finally:
        for g in hs:
            g.detach_shards()
        for g in hs:
            g.close()


def test_shard_handle_alone_refuses_to_search_and_attach_checks_the_tiling(small_db):
    from kaamer_b200 import GpuIndex, KaamerGpuError, SearchOptions
    from kaamer_b200.sharded import dense_space

    half = dense_space() // 2
    q = np.frombuffer(b"MKTAYIAKQRQISFVKSHFSRQ", np.uint8)
    qo = np.array([0, len(q)], np.uint64)
    with GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], keep_proteins=False, shard=(0, half)) as a, \
            GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], keep_proteins=False,
                           shard=(half + 5, dense_space())) as b:
        with pytest.raises(KaamerGpuError) as ei:
            a.search_proteins(q, qo, SearchOptions())
        assert "shard" in str(ei.value)
        ea, eb = a.export_shard(), b.export_shard()
        assert ea.table_fd >= 0 and ea.postings_fd >= 0 and ea.table_bytes % (2 << 20) == 0  # shareable, 2 MiB pages
        try:
            with pytest.raises(KaamerGpuError) as ei:
                a.attach_shards([ea, eb])  # codes [half, half+5) are not covered
            assert "tile" in str(ei.value)
            with pytest.raises(KaamerGpuError):
                a.attach_shards([ea])  # does not reach the end of the key space
        finally:
            for e in (ea, eb):
                os.close(e.table_fd)
                os.close(e.postings_fd)
        with pytest.raises(KaamerGpuError):
            a.search_proteins(q, qo, SearchOptions())  # a failed attach leaves the handle detached


@pytest.mark.parametrize("mode", ["probe-owner", "replicate"])
def test_peer_ipc_two_gpus(mode):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "run_peer_nccl.py")]
    if mode == "replicate":
        cmd.append("replicate")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "peer ipc ok" in out.stdout


def test_open_shard_from_kidx_file(small_db, tmp_path):
    """the deployment path of a Go server: one `.kidx` file, one key-range handle per GPU
    (kaamer_gpu_kidx_fences + kaamer_gpu_open_shard), attached to each other"""
    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from kaamer_b200.peer import attach_all
    from kaamer_b200.sharded import shard_arrays
    from oracle import oracle as o

    idx = small_db["idx"]
    p = str(tmp_path / "db.kidx")
    with GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], keep_proteins=True) as full:
        full.save(p)
    fences = GpuIndex.kidx_fences(p, 3)
    hs = [GpuIndex.open_shard(p, (int(fences[r]), int(fences[r + 1]))) for r in range(3)]
    try:
        for r, g in enumerate(hs):
            k, fo, ps = g.index_arrays()
            ek, efo, eps = shard_arrays(idx.keys, idx.offsets, idx.postings, int(fences[r]), int(fences[r + 1]))
            np.testing.assert_array_equal(k, ek)
            np.testing.assert_array_equal(fo, efo)
            np.testing.assert_array_equal(ps, eps)
            assert g.dbstats()["NumberOfAA"] == idx.n_aa
        attach_all(hs)
        q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 300, config_index=1, stream=21)
        ora = o.search_proteins(idx, q, qo, o.opts(), 4)
        for g in hs:
            assert_same_hits(g.search_proteins(q, qo, SearchOptions()), ora, "open_shard")
    finally:
        for g in hs:
            g.detach_shards()
        for g in hs:
            g.close()
