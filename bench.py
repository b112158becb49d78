#!/usr/bin/env python
"""bench.py — kaamer search hot path on B200: query residues/s and k-mer lookups/s.

    python bench.py --gpus N --steps K --warmup W             (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   (CPU restatement, host cores)

A "step" is one pass of the hot path over one batch of synthetic protein queries (workload
C3 of BASELINE.md: Swiss-Prot-scale synthetic DB, 570 k proteins / ~198 M aa, 100 k queries
per batch, reference default options, no alignment).  N>1 runs under torchrun, one rank per
GPU, index replicated, distinct query batches per rank (mode R, weak scaling, no collective
on the data path).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

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
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
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


def traffic_from_profiles():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


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
                  f"{threads} threads; CPU restatement of the Go/badger path (oracle/), not the Go binary (no Go toolchain)")
        line = {"impl": "reference", "metric": METRIC, "value": r["residues_per_s"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "kmer_lookups_per_sec": r["lookups_per_s"],
                "config": {"workload": workload_name(a)},
                "cpu_baseline": {"value": r["residues_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": r["residues_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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

    res, off, ids = make_db(a)
    t0 = time.time()
    g = GpuIndex.build(res, off, ids, keep_proteins=False, device=local_rank)
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
    clocks = sampler.stop() if rank == 0 else None
    prof = g.profile_read(reset=True)
    g.profile_enable(False)
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
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(residues), float(lookups), float(increments)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    residues_all, lookups_all, incr_all = (float(x) for x in tot.tolist())
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
    n_prof = min(a.steps, 4)
    g.profile_enable(True)
    g.profile_read(reset=True)
    for s in range(n_prof):
        step_host(s)
    prof_e2e = g.profile_read(reset=True)
    g.profile_enable(False)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    re_ = torch.tensor([float(e2e_res)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(re_, op=dist.ReduceOp.SUM)
    e2e_value = float(re_.item()) / float(te.item())
    h2d = int(np.mean([q.nbytes + (len(qo)) * 8 for q, qo in batches]))

    if rank == 0:
        peak, peak_src = peaks()
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
        tr = traffic_from_profiles()
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "traffic": (tr or {}).get("dram_bytes_per_launch") if dom == 0 else None, "peak_source": peak_src,
                "kernel": ["k_search_wt<W>", "k_search_m", "k_search_g"][dom], "kernel_ms_per_launch": dom_ms,
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
                "config": {"workload": workload_name(a), "parallelism": f"replicated index x{world}, queries split" if world > 1 else "single GPU",
                           "db_residues": int(off[-1]), "index_build_s": build_s,
                           "l2_policy": "inputs larger than L2: 14.5 GB direct-address table probed at random, "
                                        f"{a.batches} rotating query batches ({a.batches * h2d / 1e6:.0f} MB)"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h / a.steps),
                        "ms_per_step": 1e3 * float(te.item()) / a.steps,
                        "ms_per_call_rank0": {"min": 1e3 * min(step_s), "median": 1e3 * float(np.median(step_s)),
                                              "max": 1e3 * max(step_s), "argmax": int(np.argmax(step_s))},
                        "stage_ms_per_step": {"h2d_copy_stream": prof_e2e["kernel_ms"][4] / n_prof,
                                              "search_kernels": sum(prof_e2e["kernel_ms"][:3]) / n_prof,
                                              "compaction_d2h": prof_e2e["kernel_ms"][5] / n_prof},
                        "note": "kaamer_gpu_search_proteins on pinned host buffers: residues are read in place over PCIe by the search kernels (zero-copy, aligned 16-byte loads), offsets copied H2D, hits compacted and copied D2H"},
                "gpu_launches": int(prof["all_launches"]),
                "roofline": roof}
        if not a.no_cpu_baseline:
            nb = min(len(batches), 1)
            r = cpu_arm(a, res, off, ids, batches[:nb], 1, 0, threads)
            line["cpu_baseline"] = {"value": r["residues_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
                                    "kmer_lookups_per_sec": r["lookups_per_s"],
                                    "sample": f"{r['nq']} queries of one batch against the full DB, {threads} threads, "
                                              "CPU restatement (oracle/) of the Go/badger path"}
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
