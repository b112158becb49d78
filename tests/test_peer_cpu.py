"""CPU tests of the host side of mode P (kaamer_b200/peer.py): the descriptor exchange between the
ranks of one node (SCM_RIGHTS over Unix-domain sockets) under gloo, world size 2 and 3.  The
descriptors here are ordinary files standing for the shareable CUDA allocations."""
import os
import socket
import tempfile

import torch.distributed as dist
import torch.multiprocessing as mp

from kaamer_b200 import _lib
from kaamer_b200.peer import exchange_fds


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        files = []
        mine = _lib.ShardHandle()
        for what in ("table", "postings"):
            f = tempfile.TemporaryFile()
            f.write(f"{what} of rank {rank}".encode())
            f.flush()
            files.append(f)
        mine.table_fd, mine.postings_fd = files[0].fileno(), files[1].fileno()
        got = exchange_fds(mine, rank, world, dist.barrier, f"test-{port}")
        assert got[rank] == (-1, -1)
        for r in range(world):
            if r == rank:
                continue
            for fd, what in zip(got[r], ("table", "postings")):
                assert fd >= 0 and fd not in (mine.table_fd, mine.postings_fd)
                assert os.pread(fd, 100, 0) == f"{what} of rank {r}".encode()
                os.close(fd)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def _run(world):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}


def test_descriptor_exchange_gloo_world2():
    _run(2)


def test_descriptor_exchange_gloo_world3():
    _run(3)


def _write_kidx(path, keys, offsets, postings, n_proteins=0, n_aa=0, n_kmers=0):
    """the `.kidx` container of DESIGN.md §1 (no protein table)"""
    import struct

    import numpy as np

    def section(f, a):
        b = np.ascontiguousarray(a).tobytes()
        f.write(b)
        f.write(b"\0" * ((-len(b)) % 64))

    with open(path, "wb") as f:
        hd = struct.pack("<8sII5QIIQ", b"KIDX0001", 1, 7, len(keys), len(postings), n_proteins, n_aa, n_kmers,
                         int(postings.max()) if len(postings) else 0, 0, 0)
        f.write(hd + b"\0" * (128 - len(hd)))
        section(f, keys.astype(np.uint32))
        section(f, offsets.astype(np.uint64))
        section(f, postings.astype(np.uint32))


def test_kidx_fences_match_the_python_planner(tmp_path):
    """kaamer_gpu_kidx_fences (what the Go host calls) == sharded.make_fences, file I/O only"""
    import numpy as np

    from kaamer_b200 import GpuIndex, synth
    from kaamer_b200.sharded import dense_space, make_fences
    from oracle import oracle as o

    res, off = synth.protein_db(300, config_index=1)
    idx = o.Index.build(res, off, o.fasta_ids(len(off) - 1), 2)
    p = str(tmp_path / "small.kidx")
    _write_kidx(p, idx.keys, idx.offsets, idx.postings, idx.n_proteins, idx.n_aa, idx.n_kmers)
    for n in (1, 2, 3, 8):
        f = GpuIndex.kidx_fences(p, n)
        np.testing.assert_array_equal(f, make_fences(idx.keys, idx.offsets, n))
        assert f[0] == 0 and f[-1] == dense_space() and np.all(np.diff(f.astype(np.int64)) >= 0)
    _write_kidx(p, idx.keys[:0], idx.offsets[:1], idx.postings[:0])
    np.testing.assert_array_equal(GpuIndex.kidx_fences(p, 4), make_fences(idx.keys[:0], idx.offsets[:1], 4))


class _FakeIndex:
    """stands for a GpuIndex holding one key-range shard: exports two descriptors (temporary files) and
    records what attach_shards receives"""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.files = []
        self.attached = None

    def export_shard(self):
        sh = _lib.ShardHandle()
        sh.shard_lo, sh.shard_hi = self.rank * 100, (self.rank + 1) * 100
        sh.pid = os.getpid()
        fds = []
        for what in ("table", "postings"):
            f = tempfile.TemporaryFile()
            f.write(f"{what} of rank {self.rank}".encode())
            f.flush()
            fds.append(os.dup(f.fileno()))  # the driver owns (and closes) the exported descriptors
            f.close()
        sh.table_fd, sh.postings_fd = fds
        return sh

    def attach_shards(self, handles, presence_filter=True, replicate_table=False, replicate_postings=False):
        seen = []
        for r, sh in enumerate(handles):
            assert (sh.shard_lo, sh.shard_hi) == (r * 100, (r + 1) * 100)
            assert os.pread(sh.table_fd, 100, 0) == f"table of rank {r}".encode()
            assert os.pread(sh.postings_fd, 100, 0) == f"postings of rank {r}".encode()
            seen += [sh.table_fd, sh.postings_fd]
        self.attached = (len(handles), presence_filter, replicate_table, seen)


def _attach_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kaamer_b200.peer import attach_distributed

        g = _FakeIndex(rank, world)
        n = attach_distributed(g, presence_filter=False, replicate_table=True)
        assert n == world and g.attached[:3] == (world, False, True)
        for fd in g.attached[3]:  # every descriptor (own exports and received ones) is closed afterwards
            try:
                os.fstat(fd)
                raise AssertionError(f"descriptor {fd} left open")
            except OSError:
                pass
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_attach_distributed_orchestration_gloo_world3():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_attach_worker, args=(3, port, ret), nprocs=3, join=True)
    assert dict(ret) == {0: "ok", 1: "ok", 2: "ok"}
