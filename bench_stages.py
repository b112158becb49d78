#!/usr/bin/env python
"""bench_stages.py — per-stage measurements of the other SURVEY §8 rows (bench.py measures the
headline C3 protein search).  One JSON line per workload, same conventions as bench.py: CUDA-event
kernel times from the library's own timers, algorithmic bytes / cell updates for the roofline, the
CPU restatement (oracle/) timed beside it on a bounded sample.

    python bench_stages.py --workload c2        translated search: 5 Mb contigs vs the 10k-protein DB
    python bench_stages.py --workload c5        Smith-Waterman re-alignment of the hits of a C3 batch
    torchrun ... bench_stages.py --workload sharded   mode S (key-range shards + NCCL all-to-all)
    torchrun ... bench_stages.py --workload peer      mode P (key-range shards peer-mapped, NVLink probes)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def c2(a):
    print(json.dumps(c2_line(a.contigs, a.steps, a.warmup)))


def c2_line(contigs, steps, warmup):
    import argparse as _ap

    import torch

    a = _ap.Namespace(contigs=contigs, steps=steps, warmup=warmup)

    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from kaamer_b200.makedb import fasta_protein_ids
    from oracle import oracle as o

    res, off = synth.protein_db(10_000, config_index=1)
    ids = fasta_protein_ids(len(off) - 1)
    nt, noff = synth.nucleotide_contigs(res, off, a.contigs, 5_000_000, config_index=2)
    nt = torch.from_numpy(nt).pin_memory().numpy()  # page-locked caller buffer: H2D by DMA at PCIe speed
    threads = os.cpu_count() or 1
    with GpuIndex.build(res, off, ids, keep_proteins=False) as g:
        opts = SearchOptions()
        for _ in range(a.warmup):
            r = g.search_nucleotide(nt, noff, opts)
        g.profile_enable(True)
        g.profile_read(reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            r = g.search_nucleotide(nt, noff, opts)
        dt = (time.perf_counter() - t0) / a.steps
        prof = g.profile_read(reset=True)
        host = {k: v / a.steps for k, v in g.profile_host_read().items()}
        t0 = time.perf_counter()
        for _ in range(a.steps):
            t = g.get_orfs(nt, noff)
        dt_orf = (time.perf_counter() - t0) / a.steps
    k_ms = [x / a.steps for x in prof["kernel_ms"]]
    n_nt = int(noff[-1])
    tr_ms = k_ms[7]  # k_translate6 + k_orf_ends (+ sort, write)
    # CPU restatement on one contig
    idx = o.Index.build(res, off, ids, threads)
    one = nt[:int(noff[1])]
    t0 = time.perf_counter()
    ro = o.search_nucleotide(idx, one, noff[:2], o.opts(), threads)
    cpu_s = time.perf_counter() - t0
    line = {
        "workload": f"C2 translated search: {a.contigs} x 5 Mb synthetic contigs (~88 % coding) vs 10 000-protein DB, default options",
        "metric": "query residues/sec", "unit": "nt/s", "value": n_nt / dt, "ms_per_step": 1e3 * dt,
        "e2e": {"value": n_nt / dt, "unit": "nt/s", "note": "kaamer_gpu_search_nucleotide on pinned host buffers, H2D + ORFs + search + positions/start-codon + D2H"},
        "orfs": int(len(t)), "rows": int(r.n_rows), "hits": int(len(r.subject)), "orf_kmer_lookups": int(r.n_lookups),
        "get_orfs_ms_host_call": 1e3 * dt_orf,
        "stage_ms": {"translate_orf_kernels": tr_ms, "search_W": k_ms[0], "search_M": k_ms[1], "search_G": k_ms[2]},
        "host_phase_ms": host,
        "roofline": {"bound": "hbm", "kernel": "k_translate6+k_orf_ends+k_orf_write (incl. cub sort/scan)",
                     "algorithmic_bytes_per_nt": 3.0, "achieved": 3.0 * n_nt / (tr_ms * 1e-3) / 1e9 if tr_ms else None,
                     "peak": peaks(), "unit": "GB/s", "frac": (3.0 * n_nt / (tr_ms * 1e-3) / 1e9) / peaks() if tr_ms else None},
        "cpu_baseline": {"value": int(noff[1]) / cpu_s, "unit": "nt/s", "cores": threads, "kind": "port",
                         "sample": "one 5 Mb contig, CPU restatement (oracle/): GetORFs serial per contig as in the reference, ORF searches on all threads",
                         "rows": int(ro.n_rows)},
    }
    return line


def reads(a):
    """FASTQ-like batch: many short reads through the nucleotide path (search_fastq.go:78-140 is the
    per-ORF loop of the nucleotide search applied to every read)."""
    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from kaamer_b200.makedb import fasta_protein_ids
    from oracle import oracle as o

    res, off = synth.protein_db(10_000, config_index=1)
    ids = fasta_protein_ids(len(off) - 1)
    nt, noff = synth.nucleotide_contigs(res, off, 2, 5_000_000, config_index=2)
    rng = np.random.default_rng(7)
    n_reads, rl = a.reads, 150
    start = rng.integers(0, len(nt) - rl, n_reads)
    import torch

    rd = nt[(start[:, None] + np.arange(rl)[None, :]).reshape(-1)]
    rd = torch.from_numpy(rd).pin_memory().numpy()  # page-locked caller buffer
    roff = (np.arange(n_reads + 1, dtype=np.uint64) * rl)
    threads = os.cpu_count() or 1
    with GpuIndex.build(res, off, ids, keep_proteins=False) as g:
        opts = SearchOptions()
        for _ in range(a.warmup):
            r = g.search_nucleotide(rd, roff, opts)
        g.profile_enable(True)
        g.profile_read(reset=True)
        g.profile_host_read()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            r = g.search_nucleotide(rd, roff, opts)
        dt = (time.perf_counter() - t0) / a.steps
        prof = g.profile_read(reset=True)
        host = {k: v / a.steps for k, v in g.profile_host_read().items()}
        k_ms = [x / a.steps for x in prof["kernel_ms"]]
    idx = o.Index.build(res, off, ids, threads)
    ns = min(n_reads, 50_000)
    t0 = time.perf_counter()
    ro = o.search_nucleotide(idx, rd[:ns * rl], roff[:ns + 1], o.opts(), threads)
    cpu_s = time.perf_counter() - t0
    # parity on the CPU sample: the first rows of the GPU result are the rows of the first `ns` reads
    k = int(np.searchsorted(r.row_contig, ns))
    ok = (k == ro.n_rows and np.array_equal(r.subject[:int(r.hit_off[k])], ro.subject)
          and np.array_equal(r.row_start[:k], ro.row_start))
    print(json.dumps({
        "workload": f"reads: {n_reads} x {rl} nt sampled from 2 x 5 Mb synthetic contigs vs 10 000-protein DB, default options",
        "metric": "query residues/sec", "unit": "nt/s", "value": n_reads * rl / dt, "ms_per_step": 1e3 * dt,
        "rows": int(r.n_rows), "hits": int(len(r.subject)), "orf_kmer_lookups": int(r.n_lookups),
        "stage_ms": {"translate_orf_kernels": k_ms[7], "search_W": k_ms[0], "search_M": k_ms[1], "search_G": k_ms[2]},
        "host_phase_ms": host,
        "parity_sample": {"reads": ns, "rows_hits_locations_equal_oracle": bool(ok)},
        "cpu_baseline": {"value": ns * rl / cpu_s, "unit": "nt/s", "cores": threads, "kind": "port",
                         "sample": f"{ns} reads, CPU restatement (oracle/): reads serial, ORF searches of one read on all threads"}}))


def c5(a):
    import torch

    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from kaamer_b200.makedb import fasta_protein_ids
    from oracle import oracle as o

    res, off = synth.protein_db(a.db_proteins, config_index=3)
    ids = fasta_protein_ids(len(off) - 1)
    q, qo, _ = synth.protein_queries(res, off, a.queries, config_index=3, stream=100)
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        print(json.dumps(c5_on(g, res, off, ids, q, qo, a.steps, a.warmup, a.cpu_pairs)))


def c5_on(g, res, off, ids, q, qo, steps, warmup, cpu_pairs):
    """C5 stage on an index that holds the protein table: all (query, hit) pairs of one query batch"""
    import argparse as _ap

    import torch

    from kaamer_b200 import SearchOptions
    from oracle import oracle as o

    a = _ap.Namespace(steps=steps, warmup=warmup, cpu_pairs=cpu_pairs, queries=len(qo) - 1, db_proteins=len(ids))
    threads = os.cpu_count() or 1
    if True:
        r = g.search_proteins(q, qo, SearchOptions())
        nh = np.diff(r.hit_off.astype(np.int64))
        pq = np.repeat(np.arange(len(nh), dtype=np.uint32), nh)
        ps = r.subject
        n_aa = g.dbstats()["NumberOfAA"]
        qlen = np.diff(qo.astype(np.int64))
        # subject lengths by id: protein table semantics of the builder (last record with the id)
        slen_by_id = np.zeros(int(ids.max()) + 1, np.int64)
        slen_by_id[ids] = np.diff(off.astype(np.int64))
        cells = int((qlen[pq] * slen_by_id[ps]).sum())
        # page-locked caller buffers: queries in, one 64-byte result row per pair out
        from kaamer_b200.gpu import ALN_DTYPE
        q = torch.from_numpy(np.ascontiguousarray(q)).pin_memory().numpy()
        out_buf = torch.empty(len(pq) * ALN_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(ALN_DTYPE)
        for _ in range(a.warmup):
            out = g.align(q, qo, pq, ps, number_of_aa=n_aa, out=out_buf)
        g.profile_enable(True)
        g.profile_read(reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            out = g.align(q, qo, pq, ps, number_of_aa=n_aa, out=out_buf)
        dt = (time.perf_counter() - t0) / a.steps
        prof = g.profile_read(reset=True)
        plan = g.align_last_plan()
    k_ms = prof["kernel_ms"][3] / a.steps
    # CPU restatement on a bounded sample of the same pairs
    rng = np.random.default_rng(1)
    sample = rng.choice(len(pq), size=min(a.cpu_pairs, len(pq)), replace=False)
    subj = {}
    for i in range(len(ids)):
        subj[int(ids[i])] = (int(off[i]), int(off[i + 1]))
    prm = o.aln_params(n_aa)
    import concurrent.futures as cf

    def one(k):
        qi, sid = int(pq[k]), int(ps[k])
        b, e = subj[sid]
        x = o.align(q[int(qo[qi]):int(qo[qi + 1])].tobytes(), res[b:e].tobytes(), prm)
        return int(x.dp_score), int(x.raw)

    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(threads) as ex:  # ko_align releases the GIL inside ctypes
        ref = list(ex.map(one, sample.tolist()))
    cpu_s = time.perf_counter() - t0
    cpu_cells = int((qlen[pq[sample]] * slen_by_id[ps[sample]]).sum())
    ok = all(int(out["dp_score"][k]) == ref[i][0] and int(out["raw"][k]) == ref[i][1] for i, k in enumerate(sample.tolist()))
    line = {
        "workload": f"C5 re-alignment: all {len(pq)} (query, hit) pairs of one C3 batch ({a.queries} queries, {a.db_proteins}-protein DB), BLOSUM62 11/1",
        "metric": "cell updates/sec", "unit": "GCUPS", "value": cells / (k_ms * 1e-3) / 1e9 if k_ms else None,
        "pairs": int(len(pq)), "cells": cells, "kernel_ms": k_ms,
        "plan": {"long_pairs_one_cta_each": plan[0], "single_pairs_one_warp_each": plan[1],
                 "packed_jobs_two_pairs_per_warp_int16x2_dpx": plan[2]},
        "e2e": {"value": cells / dt / 1e9, "unit": "GCUPS", "ms_per_step": 1e3 * dt,
                "note": "kaamer_gpu_align on page-locked host buffers: pair schedule on the host, H2D queries + schedule, chunked kernels, D2H of 64 B per pair"},
        "roofline": {"bound": "integer ALU / shared memory (no dense contraction, no HBM roofline; SURVEY §8d)",
                     "achieved": cells / (k_ms * 1e-3) / 1e9 if k_ms else None, "unit": "GCUPS",
                     "traceback_bytes_per_cell": 1.0},
        "parity_spot_check": {"pairs": int(len(sample)), "dp_score_and_raw_equal_oracle": bool(ok)},
        "cpu_baseline": {"value": cpu_cells / cpu_s / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
                         "sample": f"{len(sample)} of the pairs, CPU restatement (oracle/) of align.Align on {threads} threads"},
    }
    return line


def sharded(a):
    import torch
    import torch.distributed as dist

    from kaamer_b200 import GpuIndex, SearchOptions, synth
    from kaamer_b200.makedb import fasta_protein_ids
    from kaamer_b200.sharded import CudaShardBackend, ShardedSearch, SingleComm, TorchComm, fences_from_sample

    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res, off = synth.protein_db(a.db_proteins, config_index=3)
    ids = fasta_protein_ids(len(off) - 1)
    # every rank builds its own key range on its GPU from the full record set; the fences come
    # from the posting mass of a sample of the records (identical on every rank)
    fences = fences_from_sample(res, off, world, device=lr)
    t_build = time.perf_counter()
    nq = a.queries  # per rank (weak scaling, as bench.py)
    q, qo, _ = synth.protein_queries(res, off, nq, config_index=3, stream=100 + rank)
    dev = torch.device("cuda", lr)
    d_res = torch.from_numpy(q).to(dev)
    d_off = torch.from_numpy(qo.astype(np.int64)).to(dev)
    opts = SearchOptions()
    with GpuIndex.build(res, off, ids, keep_proteins=False, device=lr,
                        shard=(int(fences[rank]), int(fences[rank + 1]))) as g:
        t_build = time.perf_counter() - t_build
        s = ShardedSearch(CudaShardBackend(g), fences, TorchComm() if world > 1 else SingleComm())
        for _ in range(a.warmup):
            r = s.search(d_res, d_off, nq, opts)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            r = s.search(d_res, d_off, nq, opts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        tot = torch.tensor([float(qo[-1]), float(r.n_lookups), float(r.a2a_bytes), float(int(r.n_hits.sum().item()))],
                           dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        if rank == 0:
            line = {"workload": f"mode S: {a.db_proteins}-protein DB key-range sharded over {world} GPU(s), {nq} queries per rank",
                    "metric": "query residues/sec", "unit": "residues/s", "n_gpus": world, "scaling": "weak",
                    "value": tot[0].item() / (t.item() * 1e-3), "ms_per_step": t.item(),
                    "kmer_lookups_per_sec": tot[1].item() / (t.item() * 1e-3),
                    "all_to_all_bytes_per_step": tot[2].item(), "hits": tot[3].item(),
                    "db_residues": int(off[-1]), "shard_build_s_rank0": t_build,
                    "note": "device-resident queries; includes route, 2 x (counts + payload) all-to-all over NCCL, partial counts, merge; "
                            "host syncs for the split sizes are inside the timed region"}
            print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def peer(a):
    """mode P: key-range shards mapped into every rank (CUDA IPC), probed through NVLink by the
    ordinary search kernels; same database, fences and per-rank query batches as `sharded`."""
    import torch
    import torch.distributed as dist

    from kaamer_b200 import SearchOptions, synth
    from kaamer_b200 import GpuIndex
    from kaamer_b200.makedb import fasta_protein_ids
    from kaamer_b200.peer import attach_all, attach_distributed
    from kaamer_b200.sharded import fences_from_sample

    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res, off = synth.protein_db(a.db_proteins, config_index=3)
    ids = fasta_protein_ids(len(off) - 1)
    fences = fences_from_sample(res, off, world, device=lr)
    nq = a.queries  # per rank (weak scaling)
    q, qo, _ = synth.protein_queries(res, off, nq, config_index=3, stream=100 + rank)
    dev = torch.device("cuda", lr)
    d_res = torch.from_numpy(q).to(dev)
    d_off = torch.from_numpy(qo.astype(np.int64)).to(dev)
    opts = SearchOptions()
    t_build = time.perf_counter()
    g = GpuIndex.build(res, off, ids, keep_proteins=False, device=lr, shard=(int(fences[rank]), int(fences[rank + 1])))
    t_build = time.perf_counter() - t_build
    if world > 1:
        attach_distributed(g, presence_filter=not a.no_presence, replicate_table=a.replicate_table)
    else:
        attach_all([g])
    pool_cap = nq * 16 + 4096
    d_nhits = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_base = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_size = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_pool = torch.zeros(pool_cap, dtype=torch.int64, device=dev)
    d_cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        g.search_proteins_device(d_res.data_ptr(), d_off.data_ptr(), nq, opts, d_nhits.data_ptr(), d_base.data_ptr(),
                                 d_size.data_ptr(), d_pool.data_ptr(), pool_cap, d_cnt.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        step()
    barrier()
    g.profile_enable(True)
    g.profile_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    prof = g.profile_read(reset=True)
    c = d_cnt.cpu().numpy().astype(np.uint64)
    assert int(c[3]) == 0, "status flags set"
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(qo[-1]), float(c[1]), float(c[2]), float(int(d_nhits.sum().item()))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        lookups = tot[1].item()
        pbar = tot[2].item() / max(1.0, lookups)
        remote = (world - 1) / world
        line = {"workload": f"mode P: {a.db_proteins}-protein DB key-range sharded over {world} GPU(s), shards peer-mapped (CUDA VMM) "
                            f"and probed through NVLink, {nq} queries per rank",
                "metric": "query residues/sec", "unit": "residues/s", "n_gpus": world, "scaling": "weak",
                "value": tot[0].item() / (t.item() * 1e-3), "ms_per_step": t.item(),
                "kmer_lookups_per_sec": lookups / (t.item() * 1e-3),
                "postings_per_lookup": pbar, "hits": tot[3].item(), "db_residues": int(off[-1]),
                "remote_owner_fraction": remote, "presence_filter": (not a.no_presence) and world > 1 and not a.replicate_table,
                "replicated_table": bool(a.replicate_table and world > 1),
                "kernel_ms_rank0": {n: prof["kernel_ms"][i] / max(1, prof["kernel_launches"][i]) for i, n in enumerate(["W", "M", "G"])},
                "shard_build_s_rank0": t_build,
                "note": "device-resident queries; no data-path collective: remote table entries and posting lists are "
                        "read by the search kernels with peer loads (sector-granular NVLink reads)"}
        print(json.dumps(line))
    barrier()
    g.detach_shards()
    barrier()
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["c2", "c5", "sharded", "peer", "reads"])
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--contigs", type=int, default=2)
    ap.add_argument("--db-proteins", type=int, default=570_000)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--cpu-pairs", type=int, default=2000)
    ap.add_argument("--no-presence", action="store_true", help="mode P without the local presence filter")
    ap.add_argument("--replicate-table", action="store_true",
                    help="mode P with the 14.5 GB table replicated on every GPU (postings stay sharded)")
    a = ap.parse_args()
    {"c2": c2, "c5": c5, "sharded": sharded, "peer": peer, "reads": reads}[a.workload](a)
