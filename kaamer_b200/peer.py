"""Mode P — peer-mapped key-range shards: the sharded index of one NVSwitch domain
behaves as ONE index.

Every rank (one process per GPU) builds or loads the key range `[fences[r], fences[r+1])` of the
dense 7-mer code space in shareable device memory, exports it (`kaamer_gpu_shard_export`: two
POSIX file descriptors of CUDA VMM allocations), hands the descriptors to the other ranks over
Unix-domain sockets (SCM_RIGHTS) and attaches the shards of all ranks
(`kaamer_gpu_attach_shards`).  From then on the ordinary search entry points of the handle see
the whole key space: the search kernels resolve the owner shard of each query k-mer and read the
8-byte table entry — and the posting list, when there is one — from that GPU's HBM through
NVLink with plain loads.  Queries stay on their home GPU; the counts are accumulated there in
shared memory exactly as in the single-GPU path, so there is no all-to-all, no partial-count
exchange and no merge step (compare mode S in `sharded.py`, which moves the k-mers to the data).

With `replicate_table=True` every rank also copies the table ranges of all shards into its own HBM
at attach time (the direct-address table is 14.5 GB whatever the database size; the posting lists
are what grows): the first probe of every lookup is local, NVLink carries posting lists only.

The reference has no counterpart (one process, one badger store, pkg/search/search.go:414-440);
results are identical to the single-index search because each k-mer lookup is answered by the
one shard that owns its key.
"""
from __future__ import annotations

import os
import socket

import numpy as np
import torch.distributed as dist

from . import _lib
from .gpu import GpuIndex
from .sharded import fences_from_sample  # noqa: F401  (re-exported: same fences as mode S)

_round = 0


def _close_fds(sh) -> None:
    for name in ("table_fd", "postings_fd"):
        fd = getattr(sh, name)
        if fd >= 0:
            os.close(fd)
            setattr(sh, name, -1)


def attach_all(indices: "list[GpuIndex]", presence_filter: bool = True, replicate_table: bool = False,
               replicate_postings: bool = False) -> None:
    """Single process driving several shards (one Go server process with several GPUs, or the
    single-GPU tests): every handle attaches the exports of all of them (by pointer)."""
    handles = [g.export_shard() for g in indices]
    try:
        for g in indices:
            g.attach_shards(handles, presence_filter, replicate_table, replicate_postings)
    finally:
        for sh in handles:
            _close_fds(sh)


def exchange_fds(mine: "_lib.ShardHandle", rank: int, world: int, barrier, tag: str) -> "list[tuple[int, int]]":
    """Every rank sends the two descriptors of its shard to every other rank of the node over
    abstract Unix-domain sockets (SCM_RIGHTS).  Returns, per rank, the descriptor numbers valid in
    THIS process ((-1, -1) for the own rank)."""
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(f"\0kaamer-peer-{tag}-{rank}")
    srv.listen(world)
    barrier()  # every listener exists
    got = [(-1, -1)] * world
    try:
        for r in range(world):
            if r == rank:
                for _ in range(world - 1):
                    conn, _ = srv.accept()
                    with conn:
                        socket.send_fds(conn, [b"kaamer"], [mine.table_fd, mine.postings_fd])
                        conn.recv(1)  # the receiver holds its copies before this side may close
            else:
                with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                    c.connect(f"\0kaamer-peer-{tag}-{r}")
                    _, fds, _, _ = socket.recv_fds(c, 16, 2)
                    if len(fds) != 2:
                        raise RuntimeError(f"rank {r} sent {len(fds)} descriptors instead of 2")
                    got[r] = (fds[0], fds[1])
                    c.send(b"k")
            barrier()
    finally:
        srv.close()
    return got


def attach_distributed(index: GpuIndex, group=None, presence_filter: bool = True, replicate_table: bool = False,
                       replicate_postings: bool = False) -> int:
    """One process per GPU on one node: gather the shard exports, pass the descriptors, attach.
    Control plane only (72 bytes + 2 descriptors per rank, once per index load).  Returns the
    number of shards."""
    global _round
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = index.export_shard()
    blobs = [None] * world
    dist.all_gather_object(blobs, bytes(mine), group=group)
    _round += 1
    tag = f"{os.environ.get('MASTER_PORT', '0')}-{_round}"
    fds = exchange_fds(mine, rank, world, lambda: dist.barrier(group), tag)
    handles = []
    for r in range(world):
        sh = _lib.ShardHandle.from_buffer_copy(blobs[r])
        if r == rank:
            sh = mine
        else:
            sh.table_fd, sh.postings_fd = fds[r]
        handles.append(sh)
    try:
        index.attach_shards(handles, presence_filter, replicate_table, replicate_postings)
    finally:
        for sh in handles:
            _close_fds(sh)
    dist.barrier(group)  # nobody searches before every rank has mapped every shard
    return world


def build_distributed(residues: np.ndarray, seq_off: np.ndarray, ids: np.ndarray, fences: np.ndarray, device: int,
                      group=None, keep_proteins: bool = False, presence_filter: bool = True,
                      replicate_table: bool = False) -> GpuIndex:
    """Every rank builds its own key range on its GPU from the full record set
    (`kaamer_gpu_build_shard`) and maps the ranges of the others."""
    rank = dist.get_rank(group)
    g = GpuIndex.build(residues, seq_off, ids, keep_proteins=keep_proteins, device=device,
                       shard=(int(fences[rank]), int(fences[rank + 1])))
    attach_distributed(g, group, presence_filter, replicate_table)
    return g
