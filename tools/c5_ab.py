#!/usr/bin/env python
"""A/B of the Smith-Waterman stage on the C5 batch (all (query, hit) pairs of one C3 batch): the 32-bit kernels
against the packed int16x2 / DPX jobs at several column widths and packing limits.  One JSON line per setting
(bench_stages.c5_on: kernel GCUPS from the library's event timers, e2e GCUPS through kaamer_gpu_align on host
buffers, the schedule, a spot check of dp_score / raw against the oracle)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench_stages  # noqa: E402
from kaamer_b200 import GpuIndex, synth  # noqa: E402
from kaamer_b200.makedb import fasta_protein_ids  # noqa: E402

SETTINGS = [
    {"KAAMER_ALIGN_PACKED": "0"},
    {},
    {"KAAMER_ALIGN_SCRATCH_GB": "16"},
    {"KAAMER_ALIGN_PK_MAXCW": "4"},
    {"KAAMER_ALIGN_PK_CELLS": str(4 << 20)},
]


def main():
    db, nq = int(os.environ.get("C5_DB", 570_000)), int(os.environ.get("C5_QUERIES", 100_000))
    res, off = synth.protein_db(db, config_index=3)
    ids = fasta_protein_ids(len(off) - 1)
    q, qo, _ = synth.protein_queries(res, off, nq, config_index=3, stream=100)
    keys = sorted({k for s in SETTINGS for k in s})
    only = os.environ.get("C5_SETTINGS")  # e.g. "0,2": indices into SETTINGS
    todo = [SETTINGS[int(i)] for i in only.split(",")] if only else SETTINGS
    steps = int(os.environ.get("C5_STEPS", 3))
    with GpuIndex.build(res, off, ids, keep_proteins=True) as g:
        for s in todo:
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(s)
            line = bench_stages.c5_on(g, res, off, ids, q, qo, steps=steps, warmup=1, cpu_pairs=400)
            line["env"] = s
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
