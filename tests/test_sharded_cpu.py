"""CPU tests of the multi-GPU HOST logic: shard planning, query splitting and the
sharded-search driver's two all-to-alls — in one process (lockstep simulation) and under
torch.distributed with the gloo backend, world_size 2.  The device steps are replaced by the
test-only numpy/oracle stand-in of tests/cpu_shard_backend.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kaamer_b200 import SearchOptions, synth
from kaamer_b200.sharded import (ShardedSearch, TorchComm, dense_from_keys, dense_space, make_fences, shard_arrays,
                                 simulate_lockstep, split_queries)
from oracle import oracle as o
from tests.cpu_shard_backend import CpuShardBackend


def _db(n=300):
    res, off = synth.protein_db(n, config_index=1)
    ids = o.fasta_ids(len(off) - 1)
    return res, off, ids, o.Index.build(res, off, ids, 2)


def _check_rank(result, ora, qb, qe):
    hit_off, subject, kmatch = result.to_csr()
    b, e = int(ora.hit_off[qb]), int(ora.hit_off[qe])
    np.testing.assert_array_equal(hit_off.astype(np.int64), ora.hit_off[qb:qe + 1].astype(np.int64) - b)
    np.testing.assert_array_equal(subject, ora.subject[b:e])
    np.testing.assert_array_equal(kmatch.astype(np.int64), ora.kmatch[b:e])
    np.testing.assert_array_equal(result.size_in_kmer.numpy(), ora.size_in_kmer[qb:qe])


def test_dense_codes_and_fences():
    res, off, ids, idx = _db()
    d = dense_from_keys(idx.keys)
    assert (np.diff(d.astype(np.int64)) > 0).all(), "dense order == key order"
    assert int(d.max()) < dense_space() == 442 ** 3 * 21
    for G in (1, 2, 3, 8):
        f = make_fences(idx.keys, idx.offsets, G)
        assert f[0] == 0 and f[-1] == dense_space() and (np.diff(f.astype(np.int64)) >= 0).all()
        tot_k = tot_p = 0
        masses = []
        for s in range(G):
            k, fo, p = shard_arrays(idx.keys, idx.offsets, idx.postings, int(f[s]), int(f[s + 1]))
            assert len(fo) == len(k) + 1 and int(fo[-1]) == len(p)
            tot_k += len(k)
            tot_p += len(p)
            masses.append(len(k) + len(p))
        assert tot_k == len(idx.keys) and tot_p == len(idx.postings)
        if G > 1:
            assert max(masses) < 1.2 * (sum(masses) / G) + 10


def test_split_queries_balances_residues():
    off = np.concatenate([[0], np.cumsum(np.random.default_rng(1).integers(20, 900, 1000))]).astype(np.uint64)
    for world in (1, 2, 4, 8):
        parts = split_queries(off, world)
        assert parts[0][0] == 0 and parts[-1][1] == 1000
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        mass = [int(off[e] - off[b]) for b, e in parts]
        assert max(mass) - min(mass) <= 2 * 900
    assert split_queries(np.zeros(1, np.uint64), 4) == [(0, 0)] * 4


@pytest.mark.parametrize("G", [1, 2, 3])
def test_sharded_driver_lockstep_matches_oracle(G):
    res, off, ids, idx = _db()
    q, qo, _ = synth.protein_queries(res, off, 36, config_index=1, stream=4)
    seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)] + [b"", b"MKTAYIAKQRQI", b"A" * 30]
    q, qo = o.pack(seqs)
    for opts in (SearchOptions(), SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=5)):
        ora = o.search_proteins(idx, q, qo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results), 2)
        fences = make_fences(idx.keys, idx.offsets, G)
        parts = split_queries(qo, G)

        class FakeComm:
            world = G
            rank = 0

        searchers, inputs = [], []
        for r in range(G):
            be = CpuShardBackend(idx.keys, idx.offsets, idx.postings, fences[r], fences[r + 1])
            searchers.append(ShardedSearch(be, fences, FakeComm()))
            b, e = parts[r]
            inputs.append((torch.from_numpy(q[int(qo[b]):int(qo[e])].copy()),
                           torch.from_numpy((qo[b:e + 1] - qo[b]).astype(np.int64)), e - b))
        results = simulate_lockstep(searchers, inputs, opts)
        for r in range(G):
            _check_rank(results[r], ora, *parts[r])
        assert sum(x.n_lookups for x in results) == ora.n_lookups
        assert sum(x.n_increments for x in results) == ora.n_increments


def _gloo_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res, off, ids, idx = _db()
        q, qo, _ = synth.protein_queries(res, off, 30, config_index=1, stream=6)
        opts = SearchOptions(min_kmatch=3, min_kratio=0.01, max_results=7)
        ora = o.search_proteins(idx, q, qo, o.opts(3, 0.01, 7), 1)
        fences = make_fences(idx.keys, idx.offsets, world)
        b, e = split_queries(qo, world)[rank]
        be = CpuShardBackend(idx.keys, idx.offsets, idx.postings, fences[rank], fences[rank + 1])
        s = ShardedSearch(be, fences, TorchComm())
        r = s.search(torch.from_numpy(q[int(qo[b]):int(qo[e])].copy()),
                     torch.from_numpy((qo[b:e + 1] - qo[b]).astype(np.int64)), e - b, opts)
        _check_rank(r, ora, b, e)
        tot = torch.tensor([r.n_lookups, r.n_increments], dtype=torch.int64)
        dist.all_reduce(tot)
        assert tot.tolist() == [ora.n_lookups, ora.n_increments]
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_driver_gloo_world2():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
