"""Mode P — peer-mapped key-range shards: the sharded index of one NVSwitch domain
behaves as ONE index.

Every rank (one process per GPU) builds or loads the key range `[fences[r], fences[r+1])` of the
dense 7-mer code space, exports it (`kaamer_gpu_shard_export`: CUDA IPC handles of the table and
the postings), gathers the exports of all ranks over `torch.distributed` and attaches them
(`kaamer_gpu_attach_shards`).  From then on the ordinary search entry points of the handle see
the whole key space: the search kernels resolve the owner shard of each query k-mer and read the
8-byte table entry — and the posting list, when there is one — from that GPU's HBM through
NVLink with plain loads.  Queries stay on their home GPU; the counts are accumulated there in
shared memory exactly as in the single-GPU path, so there is no all-to-all, no partial-count
exchange and no merge step (compare mode S in `sharded.py`, which moves the k-mers to the data).

The reference has no counterpart (one process, one badger store, pkg/search/search.go:414-440);
results are identical to the single-index search because each k-mer lookup is answered by the
one shard that owns its key.
"""
from __future__ import annotations

import numpy as np
import torch.distributed as dist

from .gpu import GpuIndex
from .sharded import fences_from_sample  # noqa: F401  (re-exported: same fences as mode S)


def attach_all(indices: "list[GpuIndex]") -> None:
    """Single process driving several shards (one Go server process with several GPUs, or the
    single-GPU tests): every handle attaches the exports of all of them."""
    handles = [g.export_shard() for g in indices]
    for g in indices:
        g.attach_shards(handles)


def attach_distributed(index: GpuIndex, group=None) -> int:
    """One process per GPU: all-gather the shard exports and attach them.  Returns the number of
    shards.  The gather is control plane (176 bytes per rank, once per index load)."""
    mine = index.export_shard()
    world = dist.get_world_size(group)
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    index.attach_shards(handles)
    dist.barrier(group)  # nobody searches before every rank has mapped every shard
    return world


def build_distributed(residues: np.ndarray, seq_off: np.ndarray, ids: np.ndarray, fences: np.ndarray, device: int,
                      group=None, keep_proteins: bool = False) -> GpuIndex:
    """Every rank builds its own key range on its GPU from the full record set
    (`kaamer_gpu_build_shard`) and maps the ranges of the others."""
    rank = dist.get_rank(group)
    g = GpuIndex.build(residues, seq_off, ids, keep_proteins=keep_proteins, device=device,
                       shard=(int(fences[rank]), int(fences[rank + 1])))
    attach_distributed(g, group)
    return g
