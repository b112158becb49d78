"""The native query readers (csrc/reader.cu: GetQueriesFasta / GetQueriesFastq semantics) against the
literal transliteration of the Go readers (tests/go_transliteration.py), on files that exercise the
reference's quirks.  File I/O only: runs without a GPU."""
import gzip
import os

import numpy as np
import pytest

from kaamer_b200 import readers
from tests import go_transliteration as go


def _same(batch, ref):
    assert len(batch) == len(ref)
    for i, (name, seq, size) in enumerate(ref):
        assert batch.names[i] == name.encode("latin-1").decode("utf-8", "replace"), i
        assert batch.sequence(i) == seq.encode("latin-1"), i
        assert int(batch.size_in_kmer[i]) == size, i
    assert int(batch.seq_off[-1]) == sum(len(r[1]) for r in ref)


FASTA_CASES = {
    "plain": ">sp|P1|A first protein\nMKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ\nAPILSRVGDGTQDNLSGAEKAV\n>q2\nmktayiakqr\n>q3 last stays lower\nmkTayiakqrqisfvk\n",
    "crlf_blank_spaces": ">a desc\r\n  MKTAYIAKQR  \r\n\r\n\tQISFVKSHFS\t\r\n>b\r\nACDEFGHIKLMNPQRSTVWY*\r\n>c\r\nAAAAAAAAAAAAAAAA*",
    "no_final_newline_short_last": ">x some header here ok\nMKTAYIAKQRQISFVKSHFSRQ\n>y\nMKT",
    "headers_only_and_empty_records": ">only header one ........\n>second header\n>third\nMKTAYIAKQRQISFVK\n>fourth\n",
    "sequence_before_first_header": "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ\n>h\nACDEFGHIKLMNPQRS\n",
    "leading_whitespace_then_text": "\n\n   \n>late header after blank lines\nMKTAYIAKQRQISFVK\n",
    "html_like": "<html> this is not a fasta file at all, never\n>q\nMKTAYIAKQRQISFVK\n",
    "xml_like": "  <?xml version=1.0?> not a fasta file either..\n>q\nMKTAYIAKQRQISFVK\n",
    "bm_prefix": "BM this text starts like a bitmap signature\n>q\nMKTAYIAKQRQISFVK\n",
    "short_file": ">q\nMKTAYIAKQRQISFVK\n",  # < 32 bytes: the zero padding of the sniff buffer is binary
    "control_char": ">q with a control\x01char inside the first 32\nMKTAYIAKQRQISFVK\n",
    "nbsp_trim": ">q nbsp and nel around the residues....\n\xc2\xa0MKTAYIAKQRQISFVK\xc2\x85\n>r\nACDEFGHIKLMN\n",
}


@pytest.mark.parametrize("case", sorted(FASTA_CASES))
@pytest.mark.parametrize("gz", [False, True])
def test_fasta_reader(tmp_path, case, gz):
    data = FASTA_CASES[case].encode("latin-1")
    p = str(tmp_path / ("q.fa.gz" if gz else "q.fa"))
    with (gzip.open(p, "wb") if gz else open(p, "wb")) as f:
        f.write(data)
    ref = go.get_queries_fasta(p)
    _same(readers.read_fasta(p), ref)
    if case in ("html_like", "xml_like", "bm_prefix", "short_file", "control_char") and not gz:
        assert ref == []  # silently no queries (search.go:255-271)
    if case == "plain":
        assert [r[1] for r in ref][1:] == ["MKTAYIAKQR", "mkTayiakqrqisfvk"] and ref[0][2] == 55 - 7 + 1


def test_fasta_line_longer_than_the_scanner_buffer(tmp_path):
    """a line of 1 MiB or more ends the input; what was read so far is kept (search.go:273-274)"""
    for n in (1024 * 1024 - 1, 1024 * 1024):
        p = str(tmp_path / f"long{n}.fa")
        with open(p, "wb") as f:
            f.write(b">ok first record with a normal line\nMKTAYIAKQRQISFVK\n>long\n" + b"A" * n + b"\n>after\nACDEFGHIKLMNPQRS\n")
        ref = go.get_queries_fasta(p)
        _same(readers.read_fasta(p), ref)
        assert len(ref) == (3 if n < 1024 * 1024 else 1)


def test_multi_member_gzip(tmp_path):
    p = str(tmp_path / "two.fa.gz")
    with open(p, "wb") as f:
        f.write(gzip.compress(b">a first member of the gzip stream....\nMKTAYIAKQRQISFVK\n"))
        f.write(gzip.compress(b">b second member\nACDEFGHIKLMNPQRS\n"))
    ref = go.get_queries_fasta(p)
    assert [r[0] for r in ref] == ["a first member of the gzip stream....", "b second member"]
    _same(readers.read_fasta(p), ref)


FASTQ = ("@read1 first read of the file ........\nACGTNacgtn\n+\nIIIIIIIIII\n"
         "@read2\nACGTACGTACGTACGTACGTAC\n+read2\n@@@@IIIIIIIIIIIIIIIIII\n"   # quality line starting with '@'
         "@read3\nACGTXACGT\n+\nIIIIIIIII\n"                                   # not a sequence line: record dropped
         "@read4\nGGGGCCCC\n+\nACGTACGT\n"                                     # quality looks like a sequence: it wins
         "@read5\nTTTTTTTTTTTT")


@pytest.mark.parametrize("gz", [False, True])
def test_fastq_reader(tmp_path, gz):
    p = str(tmp_path / ("r.fq.gz" if gz else "r.fq"))
    with (gzip.open(p, "wb") if gz else open(p, "wb")) as f:
        f.write(FASTQ.encode())
    ref = go.get_queries_fastq(p)
    _same(readers.read_fastq(p), ref)
    assert [r[1] for r in ref] == ["ACGTNacgtn", "ACGTACGTACGTACGTACGTAC", "ACGTACGT", "TTTTTTTTTTTT"]
    assert ref[2][0] == "read4" and ref[1][2] == 22 - 7 + 1


def test_random_fasta_files(tmp_path):
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYacdxXBZ*- \t", np.uint8)
    for t in range(20):
        lines = [b">record zero has a long enough header line"]
        for _ in range(int(rng.integers(1, 40))):
            r = rng.random()
            if r < 0.2:
                lines.append(b">" + bytes(alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, 30)))]))
            elif r < 0.3:
                lines.append(b"")
            else:
                lines.append(bytes(alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, 90)))]))
        sep = b"\r\n" if t % 3 == 0 else b"\n"
        p = str(tmp_path / f"rnd{t}.fa")
        with open(p, "wb") as f:
            f.write(sep.join(lines) + (sep if t % 2 else b""))
        _same(readers.read_fasta(p), go.get_queries_fasta(p))


def test_missing_and_empty_files(tmp_path):
    from kaamer_b200 import KaamerGpuError

    with pytest.raises(KaamerGpuError):
        readers.read_fasta(str(tmp_path / "nope.fa"))
    p = str(tmp_path / "empty.fa")
    open(p, "wb").close()
    with pytest.raises(KaamerGpuError):
        readers.read_fasta(p)  # the reference exits on the failed read of the sniff buffer


def test_format_positions_native_oracle_and_transliteration_agree():
    from oracle import oracle as o

    rng = np.random.default_rng(13)
    cases = [[], [1], [0], [1, 1, 1], [0, 0, 0], [1, 0, 1, 0, 1], [0, 1, 1, 0, 0, 1], [1] * 40 + [0] + [1] * 3]
    cases += [(rng.random(int(rng.integers(1, 300))) < p).astype(np.uint8).tolist() for p in (0.05, 0.5, 0.9) for _ in range(20)]
    for pos in cases:
        for wa in (False, True):
            ref = go.format_positions_to_string([bool(x) for x in pos], wa)
            assert readers.format_positions(pos, wa) == ref
            assert o.format_positions(np.array(pos, np.uint8), wa) == ref
    assert readers.format_positions([1, 1, 0, 0, 1, 1, 1], False) == "1-3,5-7"
