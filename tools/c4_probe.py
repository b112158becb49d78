#!/usr/bin/env python
"""Single-GPU probe of the dense regime: streamed synthetic DB of --proteins records, device-resident
query batches, timing of kaamer_gpu_search_proteins_device with CUDA events, optional oracle sample.
Prints one JSON line per configuration (kept under profiles/ when worth keeping)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--proteins", type=int, default=2_000_000)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--passes", type=int, default=0)
    ap.add_argument("--sample", type=int, default=0, help="queries checked against the restricted oracle index")
    ap.add_argument("--dense", default="", help="force KAAMER_DENSE")
    ap.add_argument("--mapkb", default="")
    ap.add_argument("--chunk", type=int, default=1_000_000)
    a = ap.parse_args()
    if a.dense:
        os.environ["KAAMER_DENSE"] = a.dense
    if a.mapkb:
        os.environ["KAAMER_D_MAPKB"] = a.mapkb
    import torch

    from kaamer_b200 import SearchOptions
    from kaamer_b200.synthdb import SynthDB, SEED_C4

    dev = torch.device("cuda", 0)
    db = SynthDB(a.proteins, seed=SEED_C4)
    t0 = time.time()
    n_aa, n_kmers = db.totals()
    g = db.build_index(n_passes=a.passes or None, chunk=a.chunk, log=lambda m: print(m, file=sys.stderr, flush=True))
    torch.cuda.synchronize()
    build_s = time.time() - t0
    free, total = torch.cuda.mem_get_info()
    nq = a.queries
    d_batches = [db.queries(b, nq) for b in range(a.batches)]
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

    for s in range(2):
        step(s)
    torch.cuda.synchronize()
    c = d_cnt.cpu().numpy().astype(np.uint64)
    g.profile_enable(True)
    g.profile_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(a.steps):
        step(s)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    prof = g.profile_read(reset=True)
    g.profile_enable(False)
    residues = float(np.mean([int(b[1][-1].item()) for b in d_batches]))
    lookups, incr = float(c[1]), float(c[2])
    pbar = incr / max(1.0, lookups)
    abytes = 16.0 + 4.0 * pbar + 1.0
    line = {"proteins": a.proteins, "db_residues": n_aa, "db_kmers": n_kmers, "build_s": build_s,
            "hbm_used_gb": (total - free) / 1e9, "queries": nq, "ms_per_batch": ms,
            "residues_per_s": residues / (ms * 1e-3), "lookups_per_s": lookups / (ms * 1e-3),
            "postings_per_lookup": pbar, "status": int(c[3]), "hits": int(d_nhits.sum().item()),
            "algorithmic_bytes_per_lookup": abytes, "algorithmic_GBs": lookups * abytes / (ms * 1e-3) / 1e9,
            "frac_of_6530": lookups * abytes / (ms * 1e-3) / 1e9 / 6529.7,
            "lookups_by_class_WMGD": [int(x) for x in c[4:8]],
            "kernel_ms": {k: prof["kernel_ms"][i] / max(1, prof["kernel_launches"][i]) for k, i in
                          (("W", 0), ("M", 1), ("G", 2), ("D", 6))},
            "env": {"KAAMER_DENSE": os.environ.get("KAAMER_DENSE"), "KAAMER_D_MAPKB": os.environ.get("KAAMER_D_MAPKB")}}
    if a.sample:
        from oracle import oracle as o

        step(0)
        torch.cuda.synchronize()
        dq, dqo = d_batches[0]
        qh, qoh = dq.cpu().numpy(), dqo.cpu().numpy().astype(np.uint64)
        nh, hb, pool = d_nhits.cpu().numpy(), d_base.cpu().numpy(), d_pool.cpu().numpy().astype(np.uint64)
        sample = list(range(0, nq, max(1, nq // a.sample)))[:a.sample]
        seqs = [qh[int(qoh[j]):int(qoh[j + 1])].tobytes() for j in sample]
        t1 = time.time()
        ridx = o.synth_restricted_index(SEED_C4, a.proteins, seqs, os.cpu_count() or 1)
        sq, sqo = o.pack(seqs)
        ora = o.search_proteins(ridx, sq, sqo, o.opts(), 4)
        bad = 0
        for i, j in enumerate(sample):
            mine = [(int(v & 0xFFFFFFFF), int(v >> 32)) for v in pool[int(hb[j]):int(hb[j]) + int(nh[j])]]
            if mine != [(int(s), int(k)) for s, k in ora.hits(i)]:
                bad += 1
        st = g.dbstats()
        line["parity_sample"] = {"queries": len(sample), "mismatches": bad, "oracle_hits": int(len(ora.subject)),
                                 "oracle_s": time.time() - t1, "threads": os.cpu_count(),
                                 "kstats_equal": (st["NumberOfProteins"], st["NumberOfAA"], st["NumberOfKmers"]) ==
                                                 (ridx.n_proteins, ridx.n_aa, ridx.n_kmers)}
    print(json.dumps(line), flush=True)
    g.close()


if __name__ == "__main__":
    main()
