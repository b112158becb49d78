"""Generates tests/golden/*.json from the reference SOURCE (read-only, /root/reference).

The reference ships no tests and cannot be built here (Go; no toolchain), so the golden
vectors are (a) tables parsed verbatim out of the Go source and (b) the known-answer
vectors derived by hand from the source in SURVEY.md §8c.  Run in the build container:
    python tests/golden/make_golden.py
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def gcode11():
    src = open(os.path.join(REF, "pkg/search/gcode.go")).read()
    body = src[src.index("var gcodeBacteria"):src.index("var gcode_1 ")]
    rows = re.findall(r'"([a-z]{3})":\s*AminoAcid\{AA:\s*"(.)",\s*Start:\s*(true|false),\s*Stop:\s*(true|false)\}', body)
    assert len(rows) == 64, len(rows)
    return {c: {"aa": a, "start": s == "true", "stop": t == "true"} for c, a, s, t in rows}


def gcodes():
    """every table of pkg/search/gcode.go as (aas in TCAG order, start-codon mask)"""
    src = open(os.path.join(REF, "pkg/search/gcode.go")).read()
    parts = re.split(r"\nvar (gcode\w+) = map\[string\]AminoAcid\{", src)
    out = {}
    order = "tcag"
    for name, body in zip(parts[1::2], parts[2::2]):
        rows = re.findall(r'"([a-z]{3})":\s*AminoAcid\{AA:\s*"(.)",\s*Start:\s*(true|false),\s*Stop:\s*(true|false)\}', body)
        assert len(rows) == 64, (name, len(rows))
        aas = ["?"] * 64
        mask = 0
        for codon, aa, start, stop in rows:
            i = order.index(codon[0]) * 16 + order.index(codon[1]) * 4 + order.index(codon[2])
            assert (aa == "*") == (stop == "true"), (name, codon)
            aas[i] = aa
            if start == "true":
                mask |= 1 << i
        out[name] = {"aas": "".join(aas), "start_mask": mask}
    return out


def aa_alphabet():
    src = open(os.path.join(REF, "pkg/kvstore/k_store.go")).read()
    m = re.search(r"aa := \[\]rune\{([^}]*)\}", src)
    return "".join(re.findall(r"'(.)'", m.group(1)))


def matrix_scores():
    src = open(os.path.join(REF, "pkg/align/matrixScores.go")).read()
    rows = re.findall(r'"(\w+)":\s*MatrixScores\{SubMatrix: matrix\.(\w+), GapOpen: (\d+), GapExtend: (\d+), '
                      r'Lambda: ([0-9.]+), K: ([0-9.]+)\}', src)
    pos = re.search(r"AAPosInMatrix = map\[rune\]int\{([^}]*)\}", src).group(1)
    order = "".join(c for c, _ in sorted(re.findall(r"'(.)': (\d+)", pos), key=lambda t: int(t[1])))
    return {"params": {k: {"matrix": m, "gap_open": int(go), "gap_extend": int(ge), "lambda": float(l), "K": float(kk)}
                       for k, m, go, ge, l, kk in rows}, "aa_pos_order": order}


def kat():
    # SURVEY.md §8c — derived by hand from the Go source
    return {
        "encode": {"AAAAAAA": 0x0B0582C0, "YYYYYYY": 0xE773B9D4, "WWWWWWW": 0xDC6E3713, "ACDEFGH": 0x0B90CDE6,
                   "MKTAYIA": 0x7859B820, "MELPNIM": 0x75B7E08A, "AXAAAAA": 0x000582C0,
                   "AAAAAAX": 0x0B0582C0, "AAAAAA*": 0x0B0582C0},
        "size_in_kmer": [{"len": 270, "star": False, "expect": 264}, {"len": 270, "star": True, "expect": 263}],
        "orfs_dna": "atg" + "gct" * 20 + "taa",
        "orfs": [
            {"seq": "W" + "L" * 20, "plus": True, "start": 2, "end": 64, "alts": list(range(1, 20))},
            {"seq": "K" + "Q" * 19 + "P", "plus": False, "start": 64, "end": 2, "alts": []},
            {"seq": "G" + "C" * 19 + "L", "plus": True, "start": 3, "end": 65, "alts": []},
            {"seq": "M" + "A" * 20 + "*", "plus": True, "start": 1, "end": 66, "alts": [0], "size_in_kmer": 15},
            {"seq": "L" + "S" * 20 + "H", "plus": False, "start": 66, "end": 1, "alts": []},
        ],
        "filter": [
            {"kmatch": [40, 12, 10, 9, 3], "size": 100, "max_results": 10, "keep": 3},
            {"kmatch": [40, 12, 10, 9, 3], "size": 100, "max_results": 2, "keep": 2},
            {"kmatch": [40, 12, 10, 9, 3], "size": 300, "max_results": 10, "keep": 1},
        ],
        "scores": {"lambda": 0.267, "K": 0.041, "raw50_bits": 23.868211075911667, "raw100_bits": 43.12818987177933,
                   "evalue_raw100_q350_n3500000": 1.274257735235062e-4},
        "fasta_ids_3": [2, 3, 3],
    }


if __name__ == "__main__":
    json.dump(gcode11(), open(os.path.join(OUT, "gcode11.json"), "w"), indent=0, sort_keys=True)
    json.dump(gcodes(), open(os.path.join(OUT, "gcodes.json"), "w"), indent=0, sort_keys=True)
    json.dump({"alphabet": aa_alphabet()}, open(os.path.join(OUT, "aa_alphabet.json"), "w"))
    json.dump(matrix_scores(), open(os.path.join(OUT, "matrix_scores.json"), "w"), indent=0, sort_keys=True)
    json.dump(kat(), open(os.path.join(OUT, "kat.json"), "w"), indent=1)
    print("golden written to", OUT)
