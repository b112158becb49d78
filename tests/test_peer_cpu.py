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
