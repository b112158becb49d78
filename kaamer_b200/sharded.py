"""Multi-GPU drivers of the search path: one process per GPU,
`torch.distributed` for the plumbing (NCCL over NVLink on the GPUs; gloo in the CPU tests).

Mode R — replicated index: `split_queries` cuts a batch into contiguous per-rank slices of
equal residue mass; every rank searches its slice against its own full index; no collective
on the data path.

Mode S — key-range sharded index: `make_fences` partitions the dense 7-mer code space into
`world` contiguous ranges of equal posting mass; `ShardedSearch.search` then runs
  1. route:  every query k-mer's dense code, bucketed by owner shard      (home rank)
  2. all-to-all #1: codes + per-query counts to the owners
  3. partial counts per (query, shard) on the owner                        (owner rank)
  4. all-to-all #2: partial (subject, count) lists back to the home rank
  5. merge + FilterResults + top-N                                          (home rank)
The reference has no counterpart (single process); the decomposition is exact because
Kmatch[q, s] is a sum over the query's k-mers (pkg/search/search.go:431-436).

The device work is behind a small backend interface (`CudaShardBackend` = the C ABI of
libkaamer_gpu.so; there is no CPU implementation in this package).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check
from .gpu import GpuIndex, SearchOptions

PAIR_RADIX = 442  # dense code radix of a pair code (csrc/internal.cuh)


# ---- host-side planning -------------------------------------------------------------------------
def dense_space() -> int:
    return PAIR_RADIX ** 3 * 21


def dense_from_keys(keys: np.ndarray) -> np.ndarray:
    """EncodeKmer keys (pkg/kvstore/k_store.go:100-110) -> dense codes (csrc/internal.cuh)."""
    k = keys.astype(np.uint64)
    p0, p1, p2, s = (k >> 23) & 0x1FF, (k >> 14) & 0x1FF, (k >> 5) & 0x1FF, k & 0x1F
    f = lambda p: np.where(p == 0, 0, p - 21)  # noqa: E731
    return ((f(p0) * PAIR_RADIX + f(p1)) * PAIR_RADIX + f(p2)) * 21 + s


def make_fences(keys: np.ndarray, offsets: np.ndarray, n_shards: int) -> np.ndarray:
    """Contiguous dense-code ranges balanced by posting mass (+1 per key for the probe itself).
    Returns u64[n_shards+1] with fences[0] = 0 and fences[-1] = dense_space()."""
    fences = np.zeros(n_shards + 1, dtype=np.uint64)
    fences[-1] = dense_space()
    if len(keys) == 0 or n_shards == 1:
        for s in range(1, n_shards):
            fences[s] = dense_space() * s // n_shards
        return fences
    d = dense_from_keys(keys)  # ascending with the keys
    mass = np.diff(offsets.astype(np.int64)) + 1
    cum = np.cumsum(mass)
    for s in range(1, n_shards):
        i = int(np.searchsorted(cum, cum[-1] * s / n_shards))
        fences[s] = d[min(i, len(d) - 1)]
    return np.maximum.accumulate(fences)


def fences_from_sample(residues: np.ndarray, seq_off: np.ndarray, n_shards: int, device: int = 0,
                       sample_records: int = 50_000) -> np.ndarray:
    """Fences without the full index: the posting mass of an evenly spaced sample of the records
    (indexed on the GPU) stands for the whole database.  Every rank computes the same fences."""
    n = len(seq_off) - 1
    if n_shards == 1 or n == 0:
        return make_fences(np.zeros(0, np.uint32), np.zeros(1, np.uint64), n_shards)
    step = max(1, n // sample_records)
    pick = np.arange(0, n, step)
    lens = (seq_off[pick + 1] - seq_off[pick]).astype(np.int64)
    so = np.zeros(len(pick) + 1, dtype=np.uint64)
    so[1:] = np.cumsum(lens)
    pos = np.arange(int(so[-1]), dtype=np.int64) - np.repeat(so[:-1].astype(np.int64), lens)
    sres = residues[np.repeat(seq_off[pick].astype(np.int64), lens) + pos]
    with GpuIndex.build(sres, so, np.arange(1, len(pick) + 1, dtype=np.uint32), keep_proteins=False,
                        device=device) as g:
        keys, offsets, _ = g.index_arrays()
    return make_fences(keys, offsets, n_shards)


def shard_arrays(keys, offsets, postings, lo: int, hi: int):
    """The part of a flat index (keys, offsets, postings) whose dense codes lie in [lo, hi)."""
    d = dense_from_keys(keys)
    a, b = int(np.searchsorted(d, lo, side="left")), int(np.searchsorted(d, hi, side="left"))
    p0, p1 = int(offsets[a]), int(offsets[b])
    return keys[a:b].copy(), (offsets[a:b + 1] - offsets[a]).astype(np.uint64), postings[p0:p1].copy()


def split_queries(seq_off: np.ndarray, world: int):
    """Mode R: contiguous query ranges of (nearly) equal residue mass -> [(begin, end)] * world."""
    nq = len(seq_off) - 1
    total = int(seq_off[-1]) if nq else 0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(seq_off, total * r // world, side="left")) if nq else 0)
    cuts.append(nq)
    cuts = np.maximum.accumulate(np.minimum(cuts, nq))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


# ---- collectives --------------------------------------------------------------------------------
class TorchComm:
    """torch.distributed wrapper (nccl on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather_ints(self, values, device) -> list:
        t = torch.tensor(list(values), dtype=torch.int64, device=device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [o.tolist() for o in out]

    def all_to_all(self, inp: torch.Tensor, in_splits, out_splits) -> torch.Tensor:
        out = torch.empty(int(sum(out_splits)), dtype=inp.dtype, device=inp.device)
        dist.all_to_all_single(out, inp, output_split_sizes=[int(x) for x in out_splits],
                               input_split_sizes=[int(x) for x in in_splits], group=self.group)
        return out


class SingleComm:
    """world size 1 (the exchange steps degenerate to copies)."""

    rank, world = 0, 1

    def all_gather_ints(self, values, device):
        return [list(values)]

    def all_to_all(self, inp, in_splits, out_splits):
        return inp.clone()


# ---- device backend -----------------------------------------------------------------------------
class CudaShardBackend:
    """The four device steps through the C ABI (include/kaamer_gpu.h, csrc/shard.cu)."""

    def __init__(self, index: GpuIndex):
        self.ix = index
        self.device = torch.device("cuda", index.device)
        self._L = _lib.lib()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def route_count(self, d_res, d_off, nq, fences, n_shards):
        counts = torch.zeros(n_shards * nq, dtype=torch.int32, device=self.device)
        size = torch.zeros(nq, dtype=torch.int32, device=self.device)
        if nq == 0:
            return counts, size
        f = (C.c_uint64 * (n_shards + 1))(*[int(x) for x in fences])
        check(self._L.kaamer_gpu_shard_route(self.ix._h, C.c_void_p(d_res.data_ptr()), C.c_void_p(d_off.data_ptr()), nq,
                                             f, n_shards, C.c_void_p(counts.data_ptr()), None, None,
                                             C.c_void_p(size.data_ptr()), self._stream()))
        return counts, size

    def route_fill(self, d_res, d_off, nq, fences, n_shards, counts, offsets, total):
        codes = torch.empty(max(int(total), 1), dtype=torch.int32, device=self.device)
        if nq == 0 or int(total) == 0:
            return codes[:0]
        f = (C.c_uint64 * (n_shards + 1))(*[int(x) for x in fences])
        check(self._L.kaamer_gpu_shard_route(self.ix._h, C.c_void_p(d_res.data_ptr()), C.c_void_p(d_off.data_ptr()), nq,
                                             f, n_shards, C.c_void_p(counts.data_ptr()), C.c_void_p(offsets.data_ptr()),
                                             C.c_void_p(codes.data_ptr()), None, self._stream()))
        return codes[:int(total)]

    def shard_count(self, codes, seg_off, nseg):
        """-> (part_n i32[nseg], parts i64[sum] in segment order, n_lookups, n_increments)"""
        part_n = torch.zeros(max(nseg, 1), dtype=torch.int32, device=self.device)
        part_base = torch.zeros(max(nseg, 1), dtype=torch.int64, device=self.device)
        counters = torch.zeros(16, dtype=torch.int64, device=self.device)
        cap = max(4096, 2 * int(codes.numel()))
        for attempt in range(4):
            pool = torch.empty(cap, dtype=torch.int64, device=self.device)
            check(self._L.kaamer_gpu_shard_count(self.ix._h, C.c_void_p(codes.data_ptr()), C.c_void_p(seg_off.data_ptr()),
                                                 nseg, C.c_void_p(part_n.data_ptr()), C.c_void_p(part_base.data_ptr()),
                                                 C.c_void_p(pool.data_ptr()), cap, C.c_void_p(counters.data_ptr()),
                                                 self._stream()))
            c = counters.tolist()
            if c[3] & 2:
                raise _lib.KaamerGpuError(-6, "a segment matched more distinct subjects than the global histogram holds")
            if not (c[3] & 1):
                break
            cap = int(c[0]) + 4096
        else:
            raise _lib.KaamerGpuError(-6, "partial pool overflow")
        part_n = part_n[:nseg]
        part_off = torch.zeros(nseg + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(part_n.to(torch.int64), 0, out=part_off[1:])
        total = int(part_off[-1].item()) if nseg else 0
        parts = torch.empty(max(total, 1), dtype=torch.int64, device=self.device)
        check(self._L.kaamer_gpu_shard_gather(self.ix._h, C.c_void_p(part_n.data_ptr()), C.c_void_p(part_base.data_ptr()),
                                              C.c_void_p(part_off.data_ptr()), C.c_void_p(pool.data_ptr()), nseg,
                                              C.c_void_p(parts.data_ptr()), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # `pool` must outlive the gather
        return part_n, parts[:total], int(c[1]), int(c[2])

    def merge(self, parts, part_off, n_shards, nq, size_in_kmer, opts: SearchOptions):
        n_hits = torch.zeros(max(nq, 1), dtype=torch.int32, device=self.device)
        hit_base = torch.zeros(max(nq, 1), dtype=torch.int32, device=self.device)
        counters = torch.zeros(16, dtype=torch.int64, device=self.device)
        per_q = max(1, min(opts.max_results, 16))
        cap = nq * per_q + 4096
        o = opts.c()
        if parts.numel() == 0:
            parts = torch.zeros(1, dtype=torch.int64, device=self.device)
        for attempt in range(4):
            pool = torch.zeros(cap, dtype=torch.int64, device=self.device)
            dr = _lib.DevResult(n_hits.data_ptr(), hit_base.data_ptr(), size_in_kmer.data_ptr(), pool.data_ptr(), cap,
                                counters.data_ptr())
            check(self._L.kaamer_gpu_shard_merge(self.ix._h, C.c_void_p(parts.data_ptr()), C.c_void_p(part_off.data_ptr()),
                                                 n_shards, nq, C.c_void_p(size_in_kmer.data_ptr()), C.byref(o),
                                                 C.byref(dr), self._stream()))
            c = counters.tolist()
            if c[3] & 2:
                raise _lib.KaamerGpuError(-6, "a query matched more distinct subjects than the global histogram holds")
            if not (c[3] & 1):
                break
            cap = int(c[0]) + 4096
        else:
            raise _lib.KaamerGpuError(-6, "hit pool overflow")
        return n_hits[:nq], hit_base[:nq], pool


@dataclass
class ShardedResult:
    """Per home rank: same layout as the device-resident single-GPU result."""

    n_hits: torch.Tensor        # i32[nq]
    hit_base: torch.Tensor      # i32[nq]
    size_in_kmer: torch.Tensor  # i32[nq]
    pool: torch.Tensor          # i64: subject | kmatch << 32, rank order per query
    n_lookups: int              # lookups done by THIS rank as an owner shard
    n_increments: int
    a2a_bytes: int              # bytes this rank sent through the two all-to-alls

    def to_csr(self):
        """-> (hit_off u64[nq+1], subject u32[], kmatch u32[]) on the host"""
        n = self.n_hits.cpu().numpy().astype(np.int64)
        base = self.hit_base.cpu().numpy().astype(np.int64)
        pool = self.pool.cpu().numpy()
        hit_off = np.zeros(len(n) + 1, dtype=np.uint64)
        hit_off[1:] = np.cumsum(n)
        idx = np.repeat(base, n) + (np.arange(int(n.sum())) - np.repeat(hit_off[:-1].astype(np.int64), n))
        v = pool[idx].astype(np.uint64) if len(idx) else np.zeros(0, np.uint64)
        return hit_off, (v & np.uint64(0xFFFFFFFF)).astype(np.uint32), (v >> np.uint64(32)).astype(np.uint32)


class ShardedSearch:
    """Mode S search over `comm.world` key-range shards; this rank owns shard `comm.rank`."""

    def __init__(self, backend, fences, comm):
        self.be = backend
        self.fences = np.asarray(fences, dtype=np.uint64)
        self.comm = comm
        self.G = comm.world
        assert len(self.fences) == self.G + 1

    def search(self, d_res: torch.Tensor, d_off: torch.Tensor, nq: int, opts: SearchOptions) -> ShardedResult:
        """SPMD entry point: every rank calls it with its own slice of the queries."""
        gen = self.steps(d_res, d_off, nq, opts)
        msg = next(gen)
        while True:
            try:
                if msg[0] == "gather":
                    msg = gen.send(self.comm.all_gather_ints(msg[1], d_res.device))
                else:
                    msg = gen.send(self.comm.all_to_all(msg[1], msg[2], msg[3]))
            except StopIteration as fin:
                return fin.value

    def steps(self, d_res, d_off, nq, opts):
        """The search as a coroutine that yields at every collective:
        ("gather", ints) -> list of every rank's ints;  ("a2a", tensor, in_splits, out_splits) ->
        received tensor.  `search` drives it with torch.distributed, `simulate_lockstep` drives
        several shards inside one process (single-GPU tests of the multi-shard logic)."""
        G, be = self.G, self.be
        dev = d_res.device

        def csum0(x, n):
            out = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            if n:
                out[1:] = torch.cumsum(x.to(torch.int64), 0)
            return out

        # 1. route (home)
        counts, size_in_kmer = be.route_count(d_res, d_off, nq, self.fences, G)
        offsets = csum0(counts, G * nq)
        cut = offsets[torch.arange(0, G + 1, device=dev) * nq].tolist()  # host sync: G+1 values
        send_codes = [cut[s + 1] - cut[s] for s in range(G)]
        codes = be.route_fill(d_res, d_off, nq, self.fences, G, counts, offsets, cut[-1])
        # 2. all-to-all #1: per-query counts, then codes
        nq_all = [x[0] for x in (yield ("gather", [nq]))]
        recv_counts = yield ("a2a", counts, [nq] * G, nq_all)
        src_cut = np.concatenate([[0], np.cumsum(nq_all)]).astype(np.int64)
        nseg = int(src_cut[-1])
        seg_off = csum0(recv_counts, nseg)
        src_idx = torch.as_tensor(src_cut, device=dev)
        rc = seg_off[src_idx].tolist()
        recv_codes = yield ("a2a", codes, send_codes, [rc[r + 1] - rc[r] for r in range(G)])
        # 3. partial counts (owner)
        part_n, parts, n_lookups, n_incr = be.shard_count(recv_codes, seg_off, nseg)
        # 4. all-to-all #2: partial list lengths, then the lists, back to the home ranks
        back_n = yield ("a2a", part_n, nq_all, [nq] * G)  # [G][nq]: from shard s, my nq queries
        pc = csum0(part_n, nseg)[src_idx].tolist()
        back_cs = csum0(back_n, G * nq)
        bc = back_cs[torch.arange(0, G + 1, device=dev) * nq].tolist()
        back_parts = yield ("a2a", parts, [pc[r + 1] - pc[r] for r in range(G)], [bc[s + 1] - bc[s] for s in range(G)])
        # 5. merge (home): entries of (shard s, query q) = back_parts[back_cs[s*nq+q] : back_cs[s*nq+q+1]]
        part_off = torch.empty(G * (nq + 1), dtype=torch.int64, device=dev)
        po = part_off.view(G, nq + 1)
        po[:, :nq] = back_cs[:G * nq].view(G, nq)
        po[:, nq] = back_cs[torch.arange(1, G + 1, device=dev) * nq]
        n_hits, hit_base, pool = be.merge(back_parts, part_off, G, nq, size_in_kmer, opts)
        a2a = 4 * (counts.numel() + codes.numel() + part_n.numel()) + 8 * parts.numel()
        return ShardedResult(n_hits, hit_base, size_in_kmer, pool, n_lookups, n_incr, int(a2a))


def simulate_lockstep(searchers, inputs, opts: SearchOptions):
    """Run G ShardedSearch instances (one per shard, any backends) inside ONE process, performing
    the collectives by slicing.  inputs[r] = (d_res, d_off, nq) of "rank" r."""
    G = len(searchers)
    gens = [s.steps(*inputs[r], opts) for r, s in enumerate(searchers)]
    msgs = [next(g) for g in gens]
    results = [None] * G
    while any(r is None for r in results):
        kind = msgs[0][0]
        assert all(m[0] == kind for m in msgs)
        if kind == "gather":
            replies = [[list(m[1]) for m in msgs]] * G
        else:
            replies = []
            for dst in range(G):
                pieces = []
                for src in range(G):
                    _, t, ins, _ = msgs[src]
                    b = int(sum(ins[:dst]))
                    pieces.append(t[b:b + int(ins[dst])])
                    assert int(ins[dst]) == int(msgs[dst][3][src]), "split size mismatch"
                replies.append(torch.cat(pieces) if pieces else msgs[dst][1][:0])
        nxt = []
        for r in range(G):
            try:
                nxt.append(gens[r].send(replies[r]))
            except StopIteration as fin:
                results[r] = fin.value
                nxt.append(None)
        msgs = nxt
    return results
