#!/usr/bin/env python
"""Per-source-line totals of an ncu `--page source --csv` SASS export, joined by instruction address with
`nvdisasm --print-line-info` of the same cubin (the .ncu-rep holds SASS + counters, the cubin the line table).

    cuobjdump -xelf all kaamer_b200/_build/search.o; nvdisasm --print-line-info search.sm_100a.cubin > all.sass
    ncu -i X.ncu-rep --page source --csv --launch-skip N --launch-count 1 > k.csv
    python profiles/tools/ncu_by_line.py all.sass '<mangled kernel name>' k.csv [top]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass, kern, src = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    line_of = {}
    cur, inside = None, False
    for l in open(sass):
        if l.startswith(".text."):
            inside = l.strip() == f".text.{kern}:"
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, ii, isamp, iw = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
    base = None
    by = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[hi + 1:]:
        if len(r) <= iw or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None:
            base = a
        key = line_of.get(a - base, (("?", 0), ""))[0]
        v = [int(r[ii] or 0), int(r[isamp] or 0), int(r[iw] or 0)]
        for j in range(3):
            by[key][j] += v[j]
            tot[j] += v[j]
    print(f"total warp instructions {tot[0]:,}  samples {tot[1]:,}  shared wavefronts {tot[2]:,}")
    print(f"{'file:line':34s} {'inst %':>7s} {'samples %':>9s} {'smem wf %':>9s}")
    for key, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{key[0]}:{key[1]:<6d}".ljust(34), f"{100 * v[0] / tot[0]:7.2f} {100 * v[1] / max(1, tot[1]):9.2f} {100 * v[2] / max(1, tot[2]):9.2f}")


if __name__ == "__main__":
    main()
