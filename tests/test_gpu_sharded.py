"""GPU parity tests of the key-range sharded search (mode S): the real CUDA steps behind the C
ABI, driven by the real ShardedSearch coroutine.  On one GPU the G shards are G shard-range
indices on the same device run in lockstep; with >= 2 GPUs the same search also runs under
torchrun with NCCL all-to-alls."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.test_sharded_cpu import _check_rank

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_lockstep(idx, q, qo, G, opts, dev=0):
    from kaamer_b200 import GpuIndex
    from kaamer_b200.sharded import (CudaShardBackend, ShardedSearch, make_fences, shard_arrays, simulate_lockstep,
                                     split_queries)

    fences = make_fences(idx.keys, idx.offsets, G)
    parts = split_queries(qo, G)

    class FakeComm:
        world = G
        rank = 0

    handles, searchers, inputs = [], [], []
    for r in range(G):
        k, fo, p = shard_arrays(idx.keys, idx.offsets, idx.postings, int(fences[r]), int(fences[r + 1]))
        g = GpuIndex.from_arrays(k, fo, p, shard=(int(fences[r]), int(fences[r + 1])), device=dev)
        handles.append(g)
        searchers.append(ShardedSearch(CudaShardBackend(g), fences, FakeComm()))
        b, e = parts[r]
        inputs.append((torch.from_numpy(q[int(qo[b]):int(qo[e])].copy()).cuda(dev),
                       torch.from_numpy((qo[b:e + 1] - qo[b]).astype(np.int64)).cuda(dev), e - b))
    try:
        results = simulate_lockstep(searchers, inputs, opts)
        torch.cuda.synchronize()
        for r in results:
            r.n_hits, r.hit_base, r.size_in_kmer, r.pool = (x.cpu() for x in (r.n_hits, r.hit_base, r.size_in_kmer, r.pool))
    finally:
        for g in handles:
            g.close()
    return results, parts


@pytest.mark.parametrize("G", [1, 2, 3])
def test_sharded_lockstep_parity(small_db, G):
    from kaamer_b200 import SearchOptions, synth
    from oracle import oracle as o

    idx = small_db["idx"]
    q, qo, _ = synth.protein_queries(small_db["res"], small_db["off"], 400, config_index=1, stream=12)
    seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)]
    seqs += [b"", b"MKT", small_db["res"][:12].tobytes(), b"A" * 700, small_db["res"][:3000].tobytes()]
    q, qo = o.pack(seqs)
    for opts in (SearchOptions(), SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=100),
                 SearchOptions(min_kmatch=4, min_kratio=0.3, max_results=2)):
        ora = o.search_proteins(idx, q, qo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results), 4)
        results, parts = _run_lockstep(idx, q, qo, G, opts)
        for r in range(G):
            _check_rank(results[r], ora, *parts[r])
        assert sum(x.n_lookups for x in results) == ora.n_lookups
        assert sum(x.n_increments for x in results) == ora.n_increments


def test_sharded_large_subject_sets_use_the_global_histograms():
    """20 000 proteins sharing the same k-mers: segments and merges outgrow shared memory."""
    from kaamer_b200 import SearchOptions
    from oracle import oracle as o

    rng = np.random.default_rng(3)
    aa = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)
    core = aa[rng.integers(0, 20, 40)].tobytes()
    prots = [core[int(rng.integers(0, 10)):] + aa[rng.integers(0, 20, 12)].tobytes() for _ in range(20_000)]
    res, off = o.pack(prots)
    ids = np.arange(1, len(prots) + 1, dtype=np.uint32)
    idx = o.Index.build(res, off, ids, 4)
    qs = [core, core[5:] + b"WWWWWWWWWW", prots[7], aa[rng.integers(0, 20, 200)].tobytes(), core * 3]
    q, qo = o.pack(qs)
    for G in (1, 2):
        for opts in (SearchOptions(max_results=10), SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=30_000)):
            ora = o.search_proteins(idx, q, qo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results), 4)
            results, parts = _run_lockstep(idx, q, qo, G, opts)
            for r in range(G):
                _check_rank(results[r], ora, *parts[r])
            assert sum(x.n_increments for x in results) == ora.n_increments


def test_sharded_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "run_sharded_nccl.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "sharded nccl ok" in out.stdout


def test_build_shard_equals_slice_of_full_index(small_db):
    """kaamer_gpu_build_shard: every rank builds its own key range; the result is the slice of the
    full index (oracle) for that range, and sample-based fences balance the posting mass."""
    from kaamer_b200 import GpuIndex
    from kaamer_b200.sharded import fences_from_sample, shard_arrays

    idx = small_db["idx"]
    fences = fences_from_sample(small_db["res"], small_db["off"], 3, sample_records=500)
    assert fences[0] == 0 and len(fences) == 4
    masses = []
    for r in range(3):
        lo, hi = int(fences[r]), int(fences[r + 1])
        with GpuIndex.build(small_db["res"], small_db["off"], small_db["ids"], keep_proteins=False, shard=(lo, hi)) as g:
            k, fo, p = g.index_arrays()
            st = g.dbstats()
        ek, efo, ep = shard_arrays(idx.keys, idx.offsets, idx.postings, lo, hi)
        np.testing.assert_array_equal(k, ek)
        np.testing.assert_array_equal(fo, efo)
        np.testing.assert_array_equal(p, ep)
        assert (st["NumberOfProteins"], st["NumberOfAA"], st["NumberOfKmers"]) == (idx.n_proteins, idx.n_aa, idx.n_kmers)
        masses.append(len(k) + len(p))
    assert max(masses) < 1.25 * sum(masses) / 3
