#!/usr/bin/env python
"""Write profiles/traffic.json from an `ncu --set full` capture of `bench.py`:

    python profiles/tools/record_traffic.py "capture name" gpurun_out/X.ncu-rep [gpurun_out/Y.ncu-rep ...]

For every captured search kernel the FIRST launch's dram__bytes_read.sum + dram__bytes_write.sum is stored
under the name bench.py uses for it, together with the hash of the kernel sources the capture was taken from
(bench.kernel_source_hash): bench.py reports `roofline.traffic` only while that hash still matches."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def unit_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def bench_name(kernel):
    """ncu demangled name -> bench.py's kernel label"""
    k = re.sub(r"\((?:int|bool)\)", "", kernel)  # `<(int)512, ...>` and `<512, ...>` spellings
    m = re.search(r"k_search_wt<\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)", k)
    if m:
        return "k_search_wt<W>" if m.group(5) == "0" else "k_search_wt<M-as-warp>"
    m = re.search(r"(k_search_[ef])<\s*\d+,\s*\d+,\s*\d+,\s*(\d+)", k)
    if m:
        return f"{m.group(1)}<CLS{m.group(2)}>"
    for name in ("k_search_m", "k_search_g"):
        if name in k:
            return name
    return None


def main():
    capture, reps = sys.argv[1], sys.argv[2:]
    from bench import kernel_source_hash

    kernels = {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        kn = hdr.index("Kernel Name")
        ir, iw, it = (hdr.index(x) for x in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        for r in rows[2:]:
            name = bench_name(r[kn])
            if not name or name in kernels:
                continue
            kernels[name] = {
                "ncu_kernel": r[kn],
                "dram_bytes_per_launch": unit_bytes(r[ir], units[ir]) + unit_bytes(r[iw], units[iw]),
                "ncu_duration": f"{r[it]} {units[it]}",
                "source_sha16": kernel_source_hash(),
                "capture": f"{capture} ({os.path.basename(rep)})",
            }
    dst = os.path.join(ROOT, "profiles", "traffic.json")
    json.dump({"kernels": kernels,
               "how": "ncu --set full --clock-control none on bench.py (C3 kernels) and tools/c4_probe.py (class D "
                      "kernels at C4); first captured launch of each search kernel = one batch of 100 k queries; "
                      "dram__bytes_read.sum + dram__bytes_write.sum"}, open(dst, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
