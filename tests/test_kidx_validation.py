"""A `.kidx` file is untrusted input: truncated or corrupt files must come back as error codes
from the C ABI — never as a C++ exception through the boundary, never as out-of-bounds reads.  The header
and content checks run before anything touches a CUDA device, so these run on CPU."""
import ctypes as C
import struct

import numpy as np
import pytest

from kaamer_b200 import _lib
from oracle import oracle as o

HDR = struct.Struct("<8sII5QIIQ")  # magic, version, k, n_keys, n_postings, n_proteins, n_aa, n_kmers, max_id, flags, n_residues


def pad64(b):
    return b + b"\0" * ((-len(b)) % 64)


def make_kidx(keys, offsets, postings, max_id=None, flags=0, poff=None, pres=b"", n_keys=None, n_postings=None,
              n_residues=None):
    keys = np.asarray(keys, np.uint32)
    offsets = np.asarray(offsets, np.uint64)
    postings = np.asarray(postings, np.uint32)
    hd = HDR.pack(b"KIDX0001", 1, 7, len(keys) if n_keys is None else n_keys,
                  len(postings) if n_postings is None else n_postings, 3, 30, 24,
                  (int(postings.max()) if len(postings) else 0) if max_id is None else max_id, flags,
                  len(pres) if n_residues is None else n_residues)
    hd += b"\0" * (128 - len(hd))
    body = pad64(keys.tobytes()) + pad64(offsets.tobytes()) + pad64(postings.tobytes())
    if flags & 1:
        body += pad64(np.asarray(poff, np.uint64).tobytes()) + pad64(pres)
    return hd + body


def fences_rc(path):
    f = (C.c_uint64 * 3)()
    rc = _lib.lib().kaamer_gpu_kidx_fences(str(path).encode(), 2, f)
    return rc, _lib.lib().kaamer_gpu_last_error().decode()


K1, K2, K3 = (o.encode_kmer(b"AAAAAAA"), o.encode_kmer(b"ACDEFGH"), o.encode_kmer(b"YYYYYYY"))


def test_a_valid_file_passes_the_checks(tmp_path):
    p = tmp_path / "ok.kidx"
    p.write_bytes(make_kidx([K1, K2, K3], [0, 1, 3, 4], [5, 9, 2, 7]))
    rc, msg = fences_rc(p)
    assert rc == 0, msg


@pytest.mark.parametrize("name,blob,needle", [
    ("huge_n_keys", make_kidx([K1], [0, 1], [5], n_keys=1 << 60), "exceed the file size"),
    ("huge_n_postings", make_kidx([K1], [0, 1], [5], n_postings=1 << 61), "exceed the file size"),
    ("truncated", make_kidx([K1, K2, K3], [0, 1, 3, 4], [5, 9, 2, 7])[:200], "exceed the file size"),
    ("bad_magic", b"KIDX9999" + make_kidx([K1], [0, 1], [5])[8:], "not a kidx"),
    ("unsorted_keys", make_kidx([K2, K1, K3], [0, 1, 3, 4], [5, 9, 2, 7]), "strictly ascending"),
    ("duplicate_keys", make_kidx([K1, K1, K3], [0, 1, 3, 4], [5, 9, 2, 7]), "strictly ascending"),
    ("invalid_key", make_kidx([K1, 0x00000015, K3], [0, 1, 3, 4], [5, 9, 2, 7]), "EncodeKmer"),
    ("offsets_decrease", make_kidx([K1, K2, K3], [0, 3, 1, 4], [5, 9, 2, 7]), "non-decreasing"),
    ("offsets_past_end", make_kidx([K1, K2, K3], [0, 1, 3, 9], [5, 9, 2, 7]), "span"),
    ("offsets_start", make_kidx([K1, K2, K3], [1, 1, 3, 4], [5, 9, 2, 7]), "span"),
])
def test_corrupt_files_are_format_errors(tmp_path, name, blob, needle):
    p = tmp_path / f"{name}.kidx"
    p.write_bytes(blob)
    rc, msg = fences_rc(p)
    assert rc in (_lib_err("FORMAT"), _lib_err("IO")), (rc, msg)
    assert needle in msg, msg


def _lib_err(name):
    return {"IO": -3, "FORMAT": -4}[name]


def test_missing_file_is_an_io_error(tmp_path):
    rc, msg = fences_rc(tmp_path / "nope.kidx")
    assert rc == -3 and "cannot open" in msg


@pytest.mark.gpu
def test_open_rejects_bad_postings_and_protein_tables(tmp_path):
    from kaamer_b200 import GpuIndex, KaamerGpuError

    cases = {
        "posting_above_max_id": make_kidx([K1, K2], [0, 1, 2], [5, 900], max_id=9),
        "prot_off_not_monotonic": make_kidx([K1], [0, 1], [1], max_id=1, flags=1, poff=[0, 8, 4], pres=b"MKTAYIAK"),
        "prot_off_wrong_end": make_kidx([K1], [0, 1], [1], max_id=1, flags=1, poff=[0, 4, 6], pres=b"MKTAYIAK"),
    }
    for name, blob in cases.items():
        p = tmp_path / f"{name}.kidx"
        p.write_bytes(blob)
        with pytest.raises(KaamerGpuError) as ei:
            GpuIndex.open(str(p))
        assert ei.value.code == -4, name
    p = tmp_path / "ok_prot.kidx"
    p.write_bytes(make_kidx([K1], [0, 1], [1], max_id=1, flags=1, poff=[0, 0, 8], pres=b"MKTAYIAK"))
    with GpuIndex.open(str(p)) as g:
        assert g.dbstats()["NumberOfAA"] == 30
