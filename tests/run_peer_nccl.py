"""torchrun worker: mode P, one rank per GPU, shards mapped across processes with CUDA IPC and
probed through NVLink; every rank checks its query slice against the CPU oracle.
Launched by tests/test_gpu_peer.py::test_peer_ipc_two_gpus (needs >= 2 GPUs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kaamer_b200 import SearchOptions, synth  # noqa: E402
from kaamer_b200.peer import build_distributed  # noqa: E402
from kaamer_b200.sharded import make_fences, split_queries  # noqa: E402
from oracle import oracle as o  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res, off = synth.protein_db(3000, config_index=1)
    ids = o.fasta_ids(len(off) - 1)
    idx = o.Index.build(res, off, ids, 4)
    q, qo, _ = synth.protein_queries(res, off, 2000, config_index=1, stream=30)
    seqs = [q[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(len(qo) - 1)]
    seqs += [res[:3000].tobytes(), res[5000:17000].tobytes()] * world  # classes M and G on every rank
    q, qo = o.pack(seqs)
    fences = make_fences(idx.keys, idx.offsets, world)
    b, e = split_queries(qo, world)[rank]
    mq, mqo = q[int(qo[b]):int(qo[e])].copy(), (qo[b:e + 1] - qo[b]).astype(np.uint64)
    replicate = len(sys.argv) > 1 and sys.argv[1] == "replicate"
    g = build_distributed(res, off, ids, fences, lr, replicate_table=replicate)
    try:
        for opts in (SearchOptions(), SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=50),
                     SearchOptions(extract_positions=True)):
            ora = o.search_proteins(idx, mq, mqo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results,
                                                         want_positions=opts.extract_positions), 4)
            r = g.search_proteins(mq, mqo, opts)
            assert r.n_rows == ora.n_rows
            np.testing.assert_array_equal(r.size_in_kmer, ora.size_in_kmer)
            np.testing.assert_array_equal(r.hit_off.astype(np.int64), ora.hit_off.astype(np.int64))
            np.testing.assert_array_equal(r.subject, ora.subject)
            np.testing.assert_array_equal(r.kmatch.astype(np.int64), ora.kmatch.astype(np.int64))
            assert (r.n_lookups, r.n_increments) == (ora.n_lookups, ora.n_increments)
            if opts.extract_positions:
                np.testing.assert_array_equal(r.pos, ora.pos)
        nt, coff = synth.nucleotide_contigs(res, off, 1, 60_000, config_index=2 + rank)
        ora = o.search_nucleotide(idx, nt, coff, o.opts(), 4)
        r = g.search_nucleotide(nt, coff, SearchOptions())
        np.testing.assert_array_equal(r.subject, ora.subject)
        np.testing.assert_array_equal(r.row_start, ora.row_start)
        np.testing.assert_array_equal(r.pos, ora.pos)
    finally:
        dist.barrier()  # every rank is done with the remote shards before anyone frees its own
        g.detach_shards()
        dist.barrier()
        g.close()
    if rank == 0:
        print("peer ipc ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
