"""torchrun worker: mode S over NCCL, one rank per GPU, checked against the CPU oracle.
Launched by tests/test_gpu_sharded.py::test_sharded_nccl_two_gpus (needs >= 2 GPUs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kaamer_b200 import GpuIndex, SearchOptions, synth  # noqa: E402
from kaamer_b200.sharded import (CudaShardBackend, ShardedSearch, TorchComm, make_fences, shard_arrays,  # noqa: E402
                                 split_queries)
from oracle import oracle as o  # noqa: E402
from tests.test_sharded_cpu import _check_rank  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res, off = synth.protein_db(3000, config_index=1)
    ids = o.fasta_ids(len(off) - 1)
    idx = o.Index.build(res, off, ids, 4)
    q, qo, _ = synth.protein_queries(res, off, 2000, config_index=1, stream=30)
    fences = make_fences(idx.keys, idx.offsets, world)
    k, fo, p = shard_arrays(idx.keys, idx.offsets, idx.postings, int(fences[rank]), int(fences[rank + 1]))
    b, e = split_queries(qo, world)[rank]
    with GpuIndex.from_arrays(k, fo, p, shard=(int(fences[rank]), int(fences[rank + 1])), device=lr) as g:
        s = ShardedSearch(CudaShardBackend(g), fences, TorchComm())
        d_res = torch.from_numpy(q[int(qo[b]):int(qo[e])].copy()).cuda(lr)
        d_off = torch.from_numpy((qo[b:e + 1] - qo[b]).astype(np.int64)).cuda(lr)
        for opts in (SearchOptions(), SearchOptions(min_kmatch=1, min_kratio=0.0, max_results=50)):
            ora = o.search_proteins(idx, q, qo, o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results), 4)
            r = s.search(d_res, d_off, e - b, opts)
            torch.cuda.synchronize()
            r.n_hits, r.hit_base, r.size_in_kmer, r.pool = (x.cpu() for x in (r.n_hits, r.hit_base, r.size_in_kmer, r.pool))
            _check_rank(r, ora, b, e)
            tot = torch.tensor([r.n_lookups, r.n_increments], dtype=torch.int64, device="cuda")
            dist.all_reduce(tot)
            assert tot.tolist() == [ora.n_lookups, ora.n_increments], (tot.tolist(), ora.n_lookups, ora.n_increments)
    dist.barrier()
    if rank == 0:
        print("sharded nccl ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
