#!/usr/bin/env python
"""bench.py — kaamer search hot path on B200: query residues/s and k-mer lookups/s.

    python bench.py --gpus N --steps K --warmup W             (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   (CPU restatement, host cores)

A "step" is one pass of the hot path over one batch of synthetic protein queries.  The headline line is
workload C3 of BASELINE.md (Swiss-Prot-scale synthetic DB, 570 k proteins / ~198 M aa, 100 k queries per
batch, reference default options, no alignment): the largest configuration the CPU reference arm can
also hold.  N > 1 runs under torchrun, one rank per GPU, index replicated, one independent query batch
per rank and step (mode R: no collective on the data path).

Extra keys of the same JSON line (rank 0):
  c4        (N = 1) the UniRef90-bacteria-scale configuration C4 — 50 M synthetic proteins generated and
            indexed on the device (streaming builder), whole index resident on one GPU — with its own
            value / roofline / parity sample against the CPU oracle's restricted index;
  sharded   (N > 1) the same database built as key-range shards, one per rank, then searched (a) with the
            posting lists left sharded and read through NVLink by the search kernels (mode P, table
            replicated) and (b) after every rank copied all shards (built sharded, searched replicated);
  stages    (N = 1) translated search (C2) and Smith-Waterman re-alignment (C5) stage lines;
  reference_toolchain   probe of `go version` / module cache.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "query residues/sec"
UNIT = "residues/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--db-proteins", type=int, default=570_000)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--batches", type=int, default=4, help="distinct query batches rotated over the steps")
    ap.add_argument("--ref-queries", type=int, default=16384, help="queries per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c4-proteins", type=int, default=50_000_000, help="C4 database size (0: skip c4 / sharded)")
    ap.add_argument("--c4-sample", type=int, default=64, help="C4 queries checked against the restricted oracle index")
    ap.add_argument("--no-stages", action="store_true", help="skip the C2 / C5 stage lines")
    ap.add_argument("--sustain-s", type=float, default=1.5, help="length of the sustained loop (clock sampling)")
    return ap.parse_args()


def workload_name(a):
    return (f"C3 Swiss-Prot-scale synthetic protein search: {a.db_proteins} DB proteins, "
            f"{a.queries} protein queries/batch, MaxResults 10, MinKMatch 10, MinKRatio 0.05, no alignment")


def make_db(a):
    from kaamer_b200 import synth
    from kaamer_b200.makedb import fasta_protein_ids as fasta_ids  # ids of `kaamer-db -make -f fasta`

    res, off = synth.protein_db(a.db_proteins, config_index=3)
    ids = fasta_ids(len(off) - 1)
    return res, off, ids


def make_queries(a, res, off, rank, n_batches, nq):
    from kaamer_b200 import synth

    out = []
    for b in range(n_batches):
        q, qo, _ = synth.protein_queries(res, off, nq, config_index=3, stream=100 + rank * 64 + b)
        out.append((q, qo))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash() -> str:
    """sha256 (16 hex) of the search kernel sources: profiles/traffic.json carries the hash of the sources
    its ncu capture was taken from, a changed kernel makes the stored DRAM traffic stale (-> null)."""
    h = hashlib.sha256()
    for f in ("search.cu", "search_common.cuh", "search_dense2.cuh", "search_dense3.cuh", "internal.cuh"):
        with open(os.path.join(ROOT, "kaamer_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def traffic_from_profiles(kernel: str):
    """-> (dram bytes per launch | None, note)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
    except Exception:
        return None, "no profiles/traffic.json"
    e = (t.get("kernels") or {}).get(kernel)
    if not e:
        return None, f"no ncu capture recorded for {kernel}"
    if e.get("source_sha16") != kernel_source_hash():
        return None, f"stale: the capture ({e.get('capture')}) was taken from other kernel sources"
    return e.get("dram_bytes_per_launch"), f"ncu dram__bytes_read+write, {e.get('capture')}"


def toolchain_probe():
    """is there a Go toolchain (and a module cache) to time the real kaamer binary?"""
    go = shutil.which("go")
    out = {"go": go, "go_version": None, "gomodcache": None, "reference_tree": os.path.isdir("/root/reference")}
    if go:
        try:
            out["go_version"] = subprocess.run([go, "version"], capture_output=True, text=True, timeout=20).stdout.strip()
            mc = subprocess.run([go, "env", "GOMODCACHE"], capture_output=True, text=True, timeout=20).stdout.strip()
            out["gomodcache"] = {"path": mc, "has_badger": os.path.isdir(os.path.join(mc, "github.com", "dgraph-io"))}
        except Exception as e:  # noqa: BLE001
            out["error"] = str(e)
    out["verdict"] = ("Go toolchain present: build kaamer from the reference tree and time it (go/cmd/kaamer-golden)"
                      if go else "no Go toolchain on this box: the reference arm is the CPU restatement (oracle/)")
    return out


def cpu_arm(a, res, off, ids, batches, steps, warmup, threads):
    """The reference's algorithm (CPU restatement in oracle/) on host cores."""
    from oracle import oracle as o

    t0 = time.time()
    idx = o.Index.build(res, off, ids, threads)
    build_s = time.time() - t0
    nq = min(a.ref_queries, a.queries)
    times, residues, lookups = [], 0, 0
    for s in range(warmup + steps):
        q, qo = batches[s % len(batches)]
        qq, qqo = q[:int(qo[nq])], qo[:nq + 1]
        t = time.time()
        r = o.search_proteins(idx, qq, qqo, o.opts(), threads)
        dt = time.time() - t
        if s >= warmup:
            times.append(dt)
            residues += int(qqo[-1])
            lookups += r.n_lookups
    total = sum(times)
    return {"residues_per_s": residues / total, "lookups_per_s": lookups / total, "ms_per_step": 1e3 * total / len(times),
            "build_s": build_s, "nq": nq}


# ---------------------------------------------------------------------------------------------------
# C4: device-generated 50 M-protein database
# ---------------------------------------------------------------------------------------------------
def c4_measure(g, db, nq, steps, warmup, dev, n_batches=2, batch0=0):
    """device-resident search of C4 query batches on handle g: (ms per batch, counters, kernel times, result of batch 0)"""
    import torch

    from kaamer_b200 import SearchOptions

    d_batches = [db.queries(batch0 + b, nq) for b in range(n_batches)]
    pool_cap = nq * 16 + 4096
    d_nhits = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_base = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_size = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_pool = torch.zeros(pool_cap, dtype=torch.int64, device=dev)
    d_cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    opts = SearchOptions()

    def step(s):
        dq, dqo = d_batches[s % len(d_batches)]
        g.search_proteins_device(dq.data_ptr(), dqo.data_ptr(), nq, opts, d_nhits.data_ptr(), d_base.data_ptr(),
                                 d_size.data_ptr(), d_pool.data_ptr(), pool_cap, d_cnt.data_ptr(), stream.cuda_stream)

    for s in range(max(1, warmup)):
        step(s)
    torch.cuda.synchronize()
    g.profile_enable(True)
    g.profile_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(steps):
        step(s)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    prof = g.profile_read(reset=True)
    g.profile_enable(False)
    step(0)
    torch.cuda.synchronize()
    c = d_cnt.cpu().numpy().astype(np.uint64)
    residues = float(np.mean([int(b[1][-1].item()) for b in d_batches]))
    result0 = (d_batches[0][0].cpu().numpy(), d_batches[0][1].cpu().numpy().astype(np.uint64), d_nhits.cpu().numpy(),
               d_base.cpu().numpy(), d_pool.cpu().numpy().astype(np.uint64))
    return ms, c, prof, residues, result0


def c4_parity_sample(seed, n_proteins, result0, n_sample, threads):
    """sampled queries of batch 0 against the oracle on the restricted index (streams the CPU twin of the generator)"""
    from oracle import oracle as o

    qh, qoh, nh, hb, pool = result0
    nq = len(nh)
    sample = list(range(0, nq, max(1, nq // n_sample)))[:n_sample]
    seqs = [qh[int(qoh[j]):int(qoh[j + 1])].tobytes() for j in sample]
    t1 = time.time()
    ridx = o.synth_restricted_index(seed, n_proteins, seqs, threads)
    sq, sqo = o.pack(seqs)
    ora = o.search_proteins(ridx, sq, sqo, o.opts(), min(threads, 8))
    bad = 0
    for i, j in enumerate(sample):
        mine = [(int(v & 0xFFFFFFFF), int(v >> 32)) for v in pool[int(hb[j]):int(hb[j]) + int(nh[j])]]
        if mine != [(int(s), int(k)) for s, k in ora.hits(i)]:
            bad += 1
    return {"queries": len(sample), "mismatches": bad, "oracle_hits": int(len(ora.subject)),
            "oracle_s": time.time() - t1, "threads": threads,
            "kstats": [int(ridx.n_proteins), int(ridx.n_aa), int(ridx.n_kmers)],
            "how": "hits, Kmatch and rank of the sampled queries vs the CPU oracle on the index restricted to their "
                   "k-mers (built by streaming the generator's CPU twin over all records)"}


def c4_roofline(ms, c, prof, peak):
    lookups, incr = float(c[1]), float(c[2])
    pbar = incr / max(1.0, lookups)
    # algorithmic bytes per lookup: 16 B + 4 B per posting (+ 1 B residue); the counts stay in
    # shared memory, so the 16 B per increment of a sort/merge are not claimed
    abytes = 16.0 + 4.0 * pbar + 1.0
    k_small = prof["kernel_ms"][6] / max(1, prof["kernel_launches"][6])
    k_large = prof["kernel_ms"][1] / max(1, prof["kernel_launches"][1])
    achieved = lookups * abytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "kernel": "k_search_f (class D, search_dense3.cuh: launches by query length — 4 / 8 / 16 warps per CTA — on "
                      "two streams, plus the hand-off launch)",
            "kernel_ms_per_launch": {"short_queries": k_small, "long_queries_side_stream": k_large},
            "note": "the launches overlap (two streams): `achieved` is over the whole step, not a single launch",
            "postings_per_lookup": pbar, "algorithmic_bytes_per_lookup": abytes,
            "lookups_per_s": lookups / (ms * 1e-3), **c4_traffic()}


def c4_traffic():
    """DRAM bytes per batch of the class D launches (all query-length launches of one step), from the ncu capture
    recorded in profiles/traffic.json; null when the kernel sources changed since."""
    tot, notes = 0.0, []
    for cls in (4, 5, 6, 7):
        t, note = traffic_from_profiles(f"k_search_f<CLS{cls}>")
        if t is None:
            return {"traffic": None, "traffic_note": note}
        tot += t
        notes.append(note)
    return {"traffic": tot, "traffic_note": "sum over the launches by query length; " + notes[0]}


def c4_block(a, local_rank, dev, peak, threads):
    import torch

    from kaamer_b200.synthdb import SEED_C4, SynthDB

    db = SynthDB(a.c4_proteins, seed=SEED_C4, device=local_rank)
    t0 = time.time()
    n_aa, n_kmers = db.totals()
    g = db.build_index()
    torch.cuda.synchronize()
    build_s = time.time() - t0
    free, total = torch.cuda.mem_get_info()
    steps = max(3, min(a.steps, 10))
    ms, c, prof, residues, result0 = c4_measure(g, db, a.queries, steps, a.warmup, dev)
    out = {"workload": f"C4 UniRef90-bacteria-scale synthetic protein search: {a.c4_proteins} DB proteins generated and "
                       f"indexed on the device, {a.queries} protein queries/batch, reference default options",
           "value": residues / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "kmer_lookups_per_sec": float(c[1]) / (ms * 1e-3),
           "db_residues": n_aa, "db_kmers": n_kmers, "index_build_s": build_s, "hbm_used_gb": (total - free) / 1e9,
           "status_flags": int(c[3]), "hits_per_batch": int(result0[2].sum()),
           "lookups_by_class_W_M_G_D": [int(x) for x in c[4:8]],
           "roofline": c4_roofline(ms, c, prof, peak)}
    # the same batches through the C ABI with pinned HOST buffers, two batches in flight (what `e2e` is for C3)
    try:
        from kaamer_b200 import SearchOptions

        hb = []
        for bi in range(2):
            dq, dqo = db.queries(bi, a.queries)
            hb.append((dq.cpu().pin_memory(), dqo.to(torch.int64).cpu().pin_memory()))
        opts = SearchOptions()

        def submit(i):
            hq, ho = hb[i % 2]
            return g.submit_proteins_ptr(hq.data_ptr(), ho.data_ptr(), a.queries, opts)

        tk = submit(0)
        for i in range(1, 4):  # warm the pinned result blocks
            nx = submit(i)
            g.wait_proteins(tk)
            tk = nx
        g.wait_proteins(tk)
        n_e2e = max(3, min(a.steps, 8))
        passes = []
        for _ in range(3):  # three timed passes of n_e2e steps; the line reports the median pass and lists all
            t1 = time.perf_counter()
            tk = submit(0)
            res_e2e = hits_e2e = 0
            for i in range(1, n_e2e + 1):
                nx = submit(i) if i < n_e2e else None
                r = g.wait_proteins(tk)
                res_e2e += int(hb[(i - 1) % 2][1][-1])
                hits_e2e += len(r.subject)
                tk = nx
            passes.append(time.perf_counter() - t1)
        e2e_s = sorted(passes)[1]
        out["e2e"] = {"value": res_e2e / e2e_s, "unit": UNIT, "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e,
                      "ms_per_step_of_each_pass": [x / n_e2e * 1e3 for x in passes],
                      "h2d_bytes_per_step": int(hb[0][0].numel() + hb[0][1].numel() * 8), "hits_per_batch": hits_e2e // n_e2e,
                      "how": "kaamer_gpu_search_proteins_submit / _wait on pinned HOST buffers, two batches in flight"}
        del hb
    except Exception as e:  # noqa: BLE001
        out["e2e"] = {"error": str(e)}
    if a.c4_sample > 0:
        out["parity_sample"] = c4_parity_sample(SEED_C4, a.c4_proteins, result0, a.c4_sample, threads)
        st = g.dbstats()
        out["parity_sample"]["kstats_equal"] = [st["NumberOfProteins"], st["NumberOfAA"], st["NumberOfKmers"]] == \
            out["parity_sample"]["kstats"]
    g.close()
    del db
    torch.cuda.empty_cache()
    return out


def sharded_block(a, rank, local_rank, world, dev, peak, threads):
    """C4 database as key-range shards, one per rank (every rank sorts its own range of the key space)."""
    import torch
    import torch.distributed as dist

    from kaamer_b200 import peer
    from kaamer_b200.synthdb import SEED_C4, SynthDB, composition_fences

    db = SynthDB(a.c4_proteins, seed=SEED_C4, device=local_rank)
    fences = composition_fences(world)
    t0 = time.time()
    n_aa, n_kmers = db.totals()
    g = db.build_index(shard=(int(fences[rank]), int(fences[rank + 1])), shareable=True)
    torch.cuda.synchronize()
    dist.barrier()
    build_s = time.time() - t0
    steps = max(3, min(a.steps, 10))
    out = {"workload": f"C4: {a.c4_proteins} synthetic proteins, index built as {world} key-range shards (one per GPU, "
                       f"fences of equal expected k-mer mass), {a.queries} queries per rank and step",
           "db_residues": n_aa, "db_kmers": n_kmers, "shard_build_s": build_s, "modes": {}}
    first_result = None
    for mode, kw in (("P_postings_sharded_table_replicated", dict(replicate_table=True)),
                     ("R_built_sharded_searched_replicated", dict(replicate_table=True, replicate_postings=True))):
        t1 = time.time()
        peer.attach_distributed(g, presence_filter=False, **kw)
        torch.cuda.synchronize()
        attach_s = time.time() - t1
        dist.barrier()
        ms, c, prof, residues, result0 = c4_measure(g, db, a.queries, steps, a.warmup, dev, batch0=16 * rank)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        tot = torch.tensor([residues, float(c[1]), float(c[2]), float(c[3])], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        ms_max = float(t.item())
        res_all, lk_all, inc_all, flags = (float(x) for x in tot.tolist())
        pbar = inc_all / max(1.0, lk_all)
        abytes = 16.0 + 4.0 * pbar + 1.0
        achieved = lk_all * abytes / (ms_max * 1e-3) / 1e9
        free, total = torch.cuda.mem_get_info()
        m = {"value": res_all / (ms_max * 1e-3), "unit": UNIT, "ms_per_step_max_over_ranks": ms_max,
             "kmer_lookups_per_sec": lk_all / (ms_max * 1e-3), "attach_s": attach_s, "status_flags": int(flags),
             "hbm_used_gb_rank0": (total - free) / 1e9,
             "roofline": {"bound": "hbm" if "replicated" in mode and "sharded_searched" in mode else "nvlink+hbm",
                          "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                          "postings_per_lookup": pbar, "algorithmic_bytes_per_lookup": abytes,
                          "note": "whole-job algorithmic bytes over N x the measured single-GPU HBM peak"},
             "nvlink_bytes_per_step_per_rank": (0.0 if kw.get("replicate_postings") else
                                                4.0 * inc_all / world * (world - 1) / world),
             "collectives_on_the_data_path": 0}
        if rank == 0 and a.c4_sample > 0 and mode.startswith("P_"):
            m["parity_sample"] = c4_parity_sample(SEED_C4, a.c4_proteins, result0, a.c4_sample, threads)
            first_result = result0
        elif rank == 0 and mode.startswith("R_") and first_result is not None:
            # the same batch as in mode P (whose sample is checked against the oracle): every query of the batch must
            # come back with the same hits, Kmatch and rank
            _, _, nh0, hb0, pool0 = first_result
            _, _, nh1, hb1, pool1 = result0
            same = bool(np.array_equal(nh0, nh1))
            bad = 0
            if same:
                for j in np.flatnonzero(nh0):
                    if not np.array_equal(pool0[int(hb0[j]):int(hb0[j]) + int(nh0[j])], pool1[int(hb1[j]):int(hb1[j]) + int(nh1[j])]):
                        bad += 1
            m["parity_whole_batch_vs_mode_P"] = {"queries": int(len(nh0)), "hits": int(nh1.sum()), "n_hits_equal": same,
                                                 "queries_with_different_hits": bad}
        dist.barrier()
        out["modes"][mode] = m
        g.detach_shards()
        dist.barrier()
    out["mode_S_note"] = ("the NCCL all-to-all layout (mode S) is retired at this density: it moves 8 B per partial "
                          "(query, subject) count = twice the bytes of the posting lists mode P reads in place "
                          "(DESIGN.md §7, profiles/r1_stage_sharded_*.json)")
    g.close()
    return out


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1

    if a.impl == "reference":
        # Under torchrun only rank 0 works; the other ranks exit 0 without work.
        if rank != 0:
            return
        res, off, ids = make_db(a)
        batches = make_queries(a, res, off, 0, min(a.batches, 2), a.queries)
        r = cpu_arm(a, res, off, ids, batches, a.steps, a.warmup, threads)
        sample = (f"{r['nq']} of the {a.queries} queries of a batch per step, full {a.db_proteins}-protein DB, "
                  f"{threads} threads; CPU restatement of the Go/badger path (oracle/: sorted in-RAM index, certainly "
                  f"faster than badger), not the Go binary (no Go toolchain)")
        line = {"impl": "reference", "metric": METRIC, "value": r["residues_per_s"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "kmer_lookups_per_sec": r["lookups_per_s"],
                "config": {"workload": workload_name(a), "sample_queries_per_step": r["nq"],
                           "sample": f"every step searches {r['nq']} of the {a.queries} queries of a batch (bounded sample)"},
                "cpu_baseline": {"value": r["residues_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": r["residues_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "reference_toolchain": toolchain_probe()}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from kaamer_b200 import GpuIndex, SearchOptions

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    # ranks spread over the host's cores (8 ranks on the default affinity contend for the same NUMA node)
    try:
        ncpu = os.cpu_count() or 1
        if world > 1 and ncpu >= 2 * world:
            per = ncpu // world
            os.sched_setaffinity(0, set(range(local_rank * per, (local_rank + 1) * per)))
    except Exception:
        pass

    res, off, ids = make_db(a)
    t0 = time.time()
    g = GpuIndex.build(res, off, ids, keep_proteins=(world == 1 and not a.no_stages), device=local_rank)
    build_s = time.time() - t0
    batches = make_queries(a, res, off, rank, a.batches, a.queries)
    opts = SearchOptions()
    nq = a.queries

    # ---- device-resident inputs (value) ---------------------------------------------------
    d_batches = [(torch.from_numpy(q).to(dev), torch.from_numpy(qo.astype(np.int64)).to(dev)) for q, qo in batches]
    pool_cap = nq * 16 + 4096
    d_nhits = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_base = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_size = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_pool = torch.zeros(pool_cap, dtype=torch.int64, device=dev)
    d_cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step_device(s):
        dq, dqo = d_batches[s % len(d_batches)]
        g.search_proteins_device(dq.data_ptr(), dqo.data_ptr(), nq, opts, d_nhits.data_ptr(), d_base.data_ptr(),
                                 d_size.data_ptr(), d_pool.data_ptr(), pool_cap, d_cnt.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(a.warmup):
        step_device(s)
    barrier()
    counters0 = d_cnt.cpu().numpy().astype(np.uint64)
    assert int(counters0[3]) == 0, "status flags set (pool/hash overflow)"
    g.profile_enable(True)
    g.profile_read(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    residues = lookups = increments = 0
    cls_lookups = np.zeros(4, np.uint64)
    cls_incr = np.zeros(4, np.uint64)
    barrier()
    e0.record(stream)
    for s in range(a.steps):
        step_device(s)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = g.profile_read(reset=True)
    g.profile_enable(False)
    # sustained loop: the K timed steps above last ~25 ms; the same step repeated for >= sustain-s seconds shows
    # the rate (and the clocks, sampled every 50 ms) under sustained load
    n_sus = max(a.steps, int(a.sustain_s * 1e3 / max(1e-3, ms_total / a.steps)))
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record(stream)
    for s in range(n_sus):
        step_device(s)
    s1.record(stream)
    barrier()
    ms_sus = s0.elapsed_time(s1)
    clocks = sampler.stop() if rank == 0 else None
    # work done per step (from the library's own counters; one readback per distinct batch)
    per_batch = []
    for b in range(len(d_batches)):
        step_device(b)
        torch.cuda.synchronize()
        c = d_cnt.cpu().numpy().astype(np.uint64)
        per_batch.append(c)
    for s in range(a.steps):
        c = per_batch[s % len(per_batch)]
        residues += int(batches[s % len(batches)][1][-1])
        lookups += int(c[1])
        increments += int(c[2])
        cls_lookups += c[4:8]
        cls_incr += c[8:12]
    res_sus = sum(int(batches[s % len(batches)][1][-1]) for s in range(n_sus))
    t = torch.tensor([ms_total, ms_sus], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(residues), float(lookups), float(increments), float(res_sus)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, ms_sus_max = (float(x) for x in t.tolist())
    residues_all, lookups_all, incr_all, res_sus_all = (float(x) for x in tot.tolist())
    value = residues_all / (ms_max * 1e-3)

    # ---- end to end through the C ABI with pinned HOST buffers (e2e) ----------------------
    h_batches = []
    for q, qo in batches:
        hq = torch.from_numpy(q).pin_memory()
        ho = torch.from_numpy(qo.astype(np.int64)).pin_memory()
        h_batches.append((hq, ho))
    d2h = 0

    def step_host(s):
        hq, ho = h_batches[s % len(h_batches)]
        return g.search_proteins_ptr(hq.data_ptr(), ho.data_ptr(), nq, opts)

    r = None
    for s in range(max(a.warmup, len(h_batches) + 1)):
        # as in the timed loop the previous result is still alive when the next call allocates its own:
        # the library's pinned-block cache then holds both sets (a cudaHostAlloc costs milliseconds)
        r = step_host(s)
    barrier()
    g.profile_enable(False)  # the timed calls carry no profiling events; the stage times come from extra calls below
    t0 = time.perf_counter()
    e2e_res = 0
    step_s = []
    for s in range(a.steps):
        ts = time.perf_counter()
        r = step_host(s)
        step_s.append(time.perf_counter() - ts)
        e2e_res += int(batches[s % len(batches)][1][-1])
        d2h += r.hit_off.nbytes + r.subject.nbytes + r.kmatch.nbytes + r.size_in_kmer.nbytes
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the same K steps through the submit / wait pair of the C ABI, two batches in flight: what a server that
    # handles requests concurrently (the reference runs nbOfThreads queries at a time) gets per GPU
    def submit(s):
        hq, ho = h_batches[s % len(h_batches)]
        return g.submit_proteins_ptr(hq.data_ptr(), ho.data_ptr(), nq, opts)

    held = []
    tk = submit(0)
    for s in range(1, 8):  # warm the pinned-block cache with the steady-state population (three result sets alive)
        nx = submit(s)
        held.append(g.wait_proteins(tk))
        held = held[-1:]
        tk = nx
    r = g.wait_proteins(tk)
    barrier()
    t0 = time.perf_counter()
    pipe_res = pipe_hits = 0
    t_submit = t_wait = 0.0
    ticket = submit(0)
    for s in range(1, a.steps + 1):
        ta = time.perf_counter()
        nxt = submit(s) if s < a.steps else None
        tb = time.perf_counter()
        r = g.wait_proteins(ticket)
        t_submit += tb - ta
        t_wait += time.perf_counter() - tb
        pipe_res += int(batches[(s - 1) % len(batches)][1][-1])
        pipe_hits += len(r.subject)
        ticket = nxt
    torch.cuda.synchronize()
    pipe_s = time.perf_counter() - t0
    n_prof = min(a.steps, 4)
    g.profile_enable(True)
    g.profile_read(reset=True)
    g.profile_host_read()
    for s in range(n_prof):
        step_host(s)
    prof_e2e = g.profile_read(reset=True)
    host_e2e = g.profile_host_read()
    g.profile_enable(False)
    # What the host link gives each GPU while ALL ranks copy at once: the e2e step moves h2d bytes per batch, so
    # h2d / this bandwidth is its floor.  (GPUs of one box share PCIe switch uplinks and host memory.)
    link = {}
    try:
        big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        dbig = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        dbig.copy_(big, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            dbig.copy_(big, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = torch.tensor([4 * big.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
        gmin, gsum = gbs.clone(), gbs.clone()
        if world > 1:
            dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(gsum, op=dist.ReduceOp.SUM)
        link = {"h2d_gbs_per_gpu_min_over_ranks": float(gmin.item()), "h2d_gbs_all_ranks": float(gsum.item()),
                "how": f"4 x 256 MiB pinned host -> device copies on every rank at the same time ({world} ranks)"}
        del big, dbig
    except Exception as e:  # noqa: BLE001
        link = {"error": str(e)}
    te = torch.tensor([e2e_s, pipe_s], dtype=torch.float64, device=dev)
    re_ = torch.tensor([float(e2e_res), float(pipe_res)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(re_, op=dist.ReduceOp.SUM)
    e2e_blocking = float(re_[0].item()) / float(te[0].item())
    e2e_value = float(re_[1].item()) / float(te[1].item())
    h2d = int(np.mean([q.nbytes + (len(qo)) * 8 for q, qo in batches]))

    line = None
    peak, peak_src = peaks()
    if rank == 0:
        # dominant kernel = the class-W search kernel (one launch per step)
        # search classes W, M, G; cls counters: [4..7] lookups, [8..11] increments
        k_ms = prof["kernel_ms"][:3]
        k_n = prof["kernel_launches"][:3]
        dom = int(np.argmax(k_ms))
        dom_ms = k_ms[dom] / max(1, k_n[dom])
        pbar = float(cls_incr[dom]) / max(1.0, float(cls_lookups[dom]))
        lookups_per_launch = float(cls_lookups[dom]) / a.steps
        # algorithmic bytes per k-mer lookup: 16 B (4 B query key + 4 B index key + 8 B CSR bounds)
        # + 4 B per posting; counts stay in shared memory (not claimed); + 1 B per query residue (encode)
        bytes_per_lookup = 16.0 + 4.0 * pbar + 1.0
        achieved = lookups_per_launch * bytes_per_lookup / (dom_ms * 1e-3) / 1e9
        kname = ["k_search_wt<W>", "k_search_m", "k_search_g"][dom]
        traffic, traffic_note = traffic_from_profiles(kname)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "traffic": traffic, "traffic_note": traffic_note, "kernel_source_sha16": kernel_source_hash(),
                "peak_source": peak_src, "kernel": kname, "kernel_ms_per_launch": dom_ms,
                "kernel_share_of_step": k_ms[dom] / ms_total if ms_total else None,
                "note_overlap": "class G runs on a side stream underneath W and M: its event time is not additive",
                "kernel_ms_by_class": {n: k_ms[i] / max(1, k_n[i]) for i, n in enumerate(["W", "M", "G"])},
                "lookups_by_class": {n: float(cls_lookups[i]) / a.steps for i, n in enumerate(["W", "M", "G"])},
                "lookups_per_launch": lookups_per_launch, "postings_per_lookup": pbar,
                "algorithmic_bytes_per_lookup": bytes_per_lookup,
                "kernel_lookups_per_s": lookups_per_launch / (dom_ms * 1e-3),
                "random_probe_ceiling_note": "see profiles/: 8-byte random probes over the 14.5 GB table top out at ~36.7 G/s on B200"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic",
                "kmer_lookups_per_sec": lookups_all / (ms_max * 1e-3),
                "sustained": {"value": res_sus_all / (ms_sus_max * 1e-3), "unit": UNIT, "steps": n_sus,
                              "seconds": ms_sus_max * 1e-3, "ms_per_step": ms_sus_max / n_sus,
                              "note": "the same step repeated back to back; the clocks are sampled over the K timed steps and this loop"},
                "config": {"workload": workload_name(a),
                           "parallelism": (f"mode R: index replicated on {world} GPUs, {world} independent query batches per "
                                           f"step (one per GPU), no collective on the data path; per-GPU work fixed (weak scaling)")
                           if world > 1 else "single GPU",
                           "db_residues": int(off[-1]), "index_build_s": build_s,
                           "l2_policy": "inputs larger than L2: 14.5 GB direct-address table probed at random, "
                                        f"{a.batches} rotating query batches ({a.batches * h2d / 1e6:.0f} MB)"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h / a.steps),
                        "ms_per_step": 1e3 * float(te[1].item()) / a.steps,
                        "how": "kaamer_gpu_search_proteins_submit / _wait on pinned HOST buffers, two batches in flight per GPU "
                               "(submit k+1, wait k): every step moves its inputs host->device and its hits device->host",
                        "host_ms_per_step_rank0": {"in_submit": 1e3 * t_submit / a.steps, "in_wait": 1e3 * t_wait / a.steps},
                        "blocking": {"value": e2e_blocking, "unit": UNIT, "ms_per_step": 1e3 * float(te[0].item()) / a.steps,
                                     "how": "one kaamer_gpu_search_proteins call at a time (round 1's e2e definition)"},
                        "ms_per_call_rank0": {"min": 1e3 * min(step_s), "median": 1e3 * float(np.median(step_s)),
                                              "max": 1e3 * max(step_s), "argmax": int(np.argmax(step_s))},
                        "host_link": dict(link, **({"floor_ms_per_step": h2d / (link["h2d_gbs_per_gpu_min_over_ranks"] * 1e9) * 1e3}
                                                   if "h2d_gbs_per_gpu_min_over_ranks" in link else {})),
                        "stage_ms_per_step": {"h2d_copy_stream": prof_e2e["kernel_ms"][4] / n_prof,
                                              "search_kernels": sum(prof_e2e["kernel_ms"][:3]) / n_prof,
                                              "compaction_d2h": prof_e2e["kernel_ms"][5] / n_prof,
                                              "host_wall_whole_call": host_e2e["search_proteins_call"] / n_prof},
                        "note": "kaamer_gpu_search_proteins on pinned host buffers: residues are read in place over PCIe by the search kernels (zero-copy, aligned 16-byte loads), offsets copied H2D, hits compacted and copied D2H"},
                "gpu_launches": int(prof["all_launches"]),
                "roofline": roof,
                "reference_toolchain": toolchain_probe()}
        if not a.no_cpu_baseline:
            nb = min(len(batches), 1)
            rc = cpu_arm(a, res, off, ids, batches[:nb], 1, 0, threads)
            line["cpu_baseline"] = {"value": rc["residues_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
                                    "kmer_lookups_per_sec": rc["lookups_per_s"],
                                    "sample": f"{rc['nq']} queries of one batch against the full DB, {threads} threads, "
                                              "CPU restatement (oracle/) of the Go/badger path"}
    # ---- stage lines (C2 translated search, C5 re-alignment): N = 1 only ----------------------------
    if rank == 0 and world == 1 and not a.no_stages:
        try:
            import bench_stages

            line["stages"] = {"c5": bench_stages.c5_on(g, res, off, ids, batches[0][0], batches[0][1], steps=2, warmup=1,
                                                       cpu_pairs=256)}
        except Exception as e:  # noqa: BLE001
            line["stages"] = {"c5": {"error": repr(e)}}
    g.close()
    del d_batches, h_batches
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not a.no_stages:
        try:
            import bench_stages

            line["stages"]["c2"] = bench_stages.c2_line(contigs=2, steps=5, warmup=2)
        except Exception as e:  # noqa: BLE001
            line["stages"]["c2"] = {"error": repr(e)}
    # ---- C4 (N = 1) / sharded index (N > 1) ----------------------------------------------------------
    if a.c4_proteins > 0:
        if world == 1:
            try:
                line["c4"] = c4_block(a, local_rank, dev, peak, threads)
            except Exception as e:  # noqa: BLE001
                line["c4"] = {"error": repr(e)}
        else:
            try:
                blk = sharded_block(a, rank, local_rank, world, dev, peak, threads)
            except Exception as e:  # noqa: BLE001
                blk = {"error": repr(e)}
            if rank == 0:
                line["sharded"] = blk
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
