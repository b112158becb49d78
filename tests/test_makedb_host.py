"""Host-side front end of the device index build (kaamer_b200/makedb.py): the record-level rules of
`kaamer-db -make -f fasta` (pkg/makedb/inputFASTA.go:96-124,195-250) that decide which
(sequence, id) pairs reach the k-mer index."""
import gzip

import numpy as np

from kaamer_b200 import makedb, synth
from oracle import oracle as o

FASTA = b""">sp|P1|A_1 first protein
mktayiakqrqisfvkshfsrq
LEERLGLIEVQ
>sp|P2|A_2 ribosomal protein L1, partial
MKTAYIAKQRQISFVKSHFSRQ
>sp|P3|A_3 short
MKTAYI
>sp|P4|A_4 fourth
MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAAA
>sp|P5|A_5 last
AAAAAAAAAA
"""


def test_read_fasta_rules(tmp_path):
    p = tmp_path / "db.fa"
    p.write_bytes(FASTA)
    entry_ids, names, res, off, ids = makedb.read_fasta(str(p))
    # ", partial" records and records shorter than 7 residues are skipped (inputFASTA.go:219-228)
    assert entry_ids == ["sp|P1|A_1", "sp|P4|A_4", "sp|P5|A_5"]
    assert names[0] == "first protein"
    # sequence lines are upper-cased and concatenated (inputFASTA.go:215)
    assert res[int(off[0]):int(off[1])].tobytes() == b"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ"
    # FASTA id quirk (inputFASTA.go:96-124): record j (1-based, j < N) gets id j+1, the last one id N
    assert ids.tolist() == [2, 5, 5]
    assert ids.tolist() == [o.fasta_ids(5)[0], o.fasta_ids(5)[3], o.fasta_ids(5)[4]]
    np.testing.assert_array_equal(makedb.fasta_protein_ids(7), o.fasta_ids(7))
    # gzip input
    pz = tmp_path / "db.fa.gz"
    with gzip.open(pz, "wb") as f:
        f.write(FASTA)
    assert makedb.read_fasta(str(pz))[0] == entry_ids


def test_write_read_roundtrip(tmp_path):
    res, off = synth.protein_db(200, config_index=1)
    names = [f"sp|S{i:06d}|SYN_{i} synthetic protein {i}" for i in range(1, 201)]
    p = tmp_path / "syn.fa"
    synth.write_fasta(str(p), names, res, off)
    entry_ids, _, res2, off2, ids2 = makedb.read_fasta(str(p))
    assert len(entry_ids) == 200 and entry_ids[0] == "sp|S000001|SYN_1"
    np.testing.assert_array_equal(res2, res)
    np.testing.assert_array_equal(off2, off)
    np.testing.assert_array_equal(ids2, o.fasta_ids(200))


def test_read_fasta_crlf_line_endings(tmp_path):
    """bufio.Scanner / ScanLines (inputFASTA.go:86-96) drops the CR of a CRLF file: the records, and the
    k-mers that span line breaks, equal those of the LF file — checked against the transliterated Go loop"""
    from tests import go_transliteration as gt

    p = tmp_path / "crlf.fa"
    p.write_bytes(FASTA.replace(b"\n", b"\r\n"))
    a = makedb.read_fasta(str(p))
    p2 = tmp_path / "lf.fa"
    p2.write_bytes(FASTA)
    b = makedb.read_fasta(str(p2))
    assert a[0] == b[0] and a[1] == b[1]
    np.testing.assert_array_equal(a[2], b[2])
    np.testing.assert_array_equal(a[3], b[3])
    np.testing.assert_array_equal(a[4], b[4])
    assert b"\r" not in a[2].tobytes()
    # the transliterated reader loop + index step: identical k-mer index from both files
    lines = [l[:-1] if l.endswith("\r") else l for l in FASTA.replace(b"\n", b"\r\n").decode().split("\n")]
    assert gt.make_index("\n".join(lines)) == gt.make_index(FASTA.decode())
