"""ctypes binding of the CPU oracle (oracle/libkaamer_oracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, bench.py's cpu_baseline / --impl reference
legs and __graft_entry__.smoke() — never from kaamer_b200/ (the product path).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkaamer_oracle.so")


def build(force: bool = False) -> str:
    deps = [os.path.join(_HERE, "kaamer_oracle.cpp"), os.path.join(_HERE, "kaamer_oracle.h"),
            os.path.join(_HERE, "synth_oracle.cpp"), os.path.join(_HERE, "..", "include", "kaamer_synth_spec.h")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in deps
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class KoOpts(C.Structure):
    _fields_ = [
        ("min_kmatch", C.c_int64),
        ("min_kratio", C.c_double),
        ("max_results", C.c_int32),
        ("want_positions", C.c_int32),
    ]


class KoAln(C.Structure):
    _fields_ = [
        ("identity", C.c_float),
        ("similarity", C.c_float),
        ("length", C.c_int32),
        ("mismatches", C.c_int32),
        ("gap_openings", C.c_int32),
        ("raw", C.c_int32),
        ("bitscore", C.c_double),
        ("evalue", C.c_double),
        ("query_start", C.c_int32),
        ("query_end", C.c_int32),
        ("subject_start", C.c_int32),
        ("subject_end", C.c_int32),
        ("dp_score", C.c_int32),
        ("n_segments", C.c_int32),
        ("illegal", C.c_int32),
    ]


class KoAlnParams(C.Structure):
    _fields_ = [
        ("lambda_", C.c_double),
        ("K", C.c_double),
        ("gap_open_opt", C.c_int32),
        ("gap_extend_opt", C.c_int32),
        ("number_of_aa", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp, u8p, u32p, u64p, i64p, i32p = (
        C.c_void_p,
        C.POINTER(C.c_uint8),
        C.POINTER(C.c_uint32),
        C.POINTER(C.c_uint64),
        C.POINTER(C.c_int64),
        C.POINTER(C.c_int32),
    )
    L.ko_encode_kmer.restype = C.c_uint32
    L.ko_encode_kmer.argtypes = [C.c_char_p]
    L.ko_decode_kmer.argtypes = [C.c_uint32, C.c_char_p]
    L.ko_size_in_kmer.restype = C.c_int32
    L.ko_size_in_kmer.argtypes = [C.c_char_p, C.c_uint64]
    L.ko_fasta_ids.argtypes = [C.c_uint64, vp]
    L.ko_index_build.restype = vp
    L.ko_index_build.argtypes = [vp, vp, vp, C.c_uint64, C.c_int]
    L.ko_index_from_arrays.restype = vp
    L.ko_index_from_arrays.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
    L.ko_index_free.argtypes = [vp]
    for n in ("ko_index_n_keys", "ko_index_n_postings"):
        getattr(L, n).restype = C.c_uint64
        getattr(L, n).argtypes = [vp]
    L.ko_index_keys.restype = u32p
    L.ko_index_offsets.restype = u64p
    L.ko_index_postings.restype = u32p
    for n in ("ko_index_keys", "ko_index_offsets", "ko_index_postings"):
        getattr(L, n).argtypes = [vp]
    L.ko_index_stats.argtypes = [vp, u64p, u64p, u64p]
    L.ko_search_proteins.restype = vp
    L.ko_search_proteins.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(KoOpts), C.c_int]
    L.ko_search_nucleotide.restype = vp
    L.ko_search_nucleotide.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(KoOpts), C.c_int]
    L.ko_get_orfs.restype = vp
    L.ko_get_orfs.argtypes = [vp, C.c_uint64]
    L.ko_orfs_free.argtypes = [vp]
    L.ko_orfs_n.restype = C.c_uint64
    L.ko_orfs_n.argtypes = [vp]
    for n, t in (
        ("ko_orfs_seq", u8p),
        ("ko_orfs_seq_off", u64p),
        ("ko_orfs_start", i64p),
        ("ko_orfs_end", i64p),
        ("ko_orfs_plus", u8p),
        ("ko_orfs_alts", i32p),
        ("ko_orfs_alts_off", u64p),
    ):
        getattr(L, n).restype = t
        getattr(L, n).argtypes = [vp]
    L.ko_result_free.argtypes = [vp]
    for n in ("ko_result_n_rows", "ko_result_n_lookups", "ko_result_n_increments"):
        getattr(L, n).restype = C.c_uint64
        getattr(L, n).argtypes = [vp]
    for n, t in (
        ("ko_result_hit_off", u64p),
        ("ko_result_subject", u32p),
        ("ko_result_kmatch", i64p),
        ("ko_result_size_in_kmer", i32p),
        ("ko_result_pos_off", u64p),
        ("ko_result_pos", u8p),
        ("ko_result_row_contig", u32p),
        ("ko_result_row_start", i64p),
        ("ko_result_row_end", i64p),
        ("ko_result_row_plus", u8p),
        ("ko_result_row_seq", u8p),
        ("ko_result_row_seq_off", u64p),
    ):
        getattr(L, n).restype = t
        getattr(L, n).argtypes = [vp]
    L.ko_filter_count.restype = C.c_int32
    L.ko_filter_count.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(KoOpts)]
    L.ko_align.restype = C.c_int
    L.ko_align.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(KoAlnParams),
                           C.POINTER(KoAln), C.c_char_p, C.c_char_p, C.c_int32]
    L.ko_blosum62.restype = C.c_int32
    L.ko_blosum62.argtypes = [C.c_int32, C.c_int32]
    L.ko_bitscore.restype = C.c_double
    L.ko_bitscore.argtypes = [C.c_double, C.c_double, C.c_int32]
    L.ko_evalue.restype = C.c_double
    L.ko_evalue.argtypes = [C.c_int32, C.c_uint64, C.c_double]
    L.ko_format_positions.restype = C.c_int32
    L.ko_format_positions.argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
    L.ko_aln_string.restype = C.c_int32
    L.ko_aln_string.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
    L.ko_set_align_model.argtypes = [vp, C.c_int32]
    L.ko_set_align_model.restype = None
    L.ko_set_genetic_code.argtypes = [C.c_char_p, C.c_uint64]
    L.ko_set_genetic_code.restype = None
    L.ko_reset_align_model.argtypes = []
    L.ko_reset_align_model.restype = None
    L.kso_record.restype = C.c_uint32
    L.kso_record.argtypes = [C.c_uint64, C.c_uint64, vp, C.c_uint32]
    L.kso_record_meta.argtypes = [C.c_uint64, C.c_uint64, u64p, u32p, u32p]
    L.kso_query.restype = C.c_uint32
    L.kso_query.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, vp, C.c_uint32, u64p]
    L.kso_restricted_index.restype = C.c_int
    L.kso_restricted_index.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, vp, C.c_uint64, C.c_int, vp,
                                       C.POINTER(u32p), u64p]
    L.kso_free.argtypes = [vp]
    _lib = L
    return L


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def opts(min_kmatch=10, min_kratio=0.05, max_results=10, want_positions=False) -> KoOpts:
    """Reference defaults: api/server.go:194-207."""
    return KoOpts(int(min_kmatch), float(min_kratio), int(max_results), int(bool(want_positions)))


def encode_kmer(kmer: bytes) -> int:
    assert len(kmer) == 7
    return int(lib().ko_encode_kmer(kmer))


def decode_kmer(key: int) -> bytes:
    buf = C.create_string_buffer(8)
    lib().ko_decode_kmer(key, buf)
    return buf.raw[:7]


def size_in_kmer(seq: bytes) -> int:
    return int(lib().ko_size_in_kmer(seq, len(seq)))


def fasta_ids(n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint32)
    lib().ko_fasta_ids(n, _vp(out))
    return out


def pack(seqs) -> tuple[np.ndarray, np.ndarray]:
    """list of bytes -> (residues u8, offsets u64[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    res = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    return res, off


class Index:
    def __init__(self, handle):
        self._h = handle
        L = lib()
        nk = L.ko_index_n_keys(handle)
        npst = L.ko_index_n_postings(handle)
        self.keys = _np(L.ko_index_keys(handle), nk, np.uint32)
        self.offsets = _np(L.ko_index_offsets(handle), nk + 1, np.uint64)
        self.postings = _np(L.ko_index_postings(handle), npst, np.uint32)
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        L.ko_index_stats(handle, C.byref(a), C.byref(b), C.byref(c))
        self.n_proteins, self.n_aa, self.n_kmers = a.value, b.value, c.value

    @classmethod
    def build(cls, residues: np.ndarray, off: np.ndarray, ids: np.ndarray, n_threads: int = 1):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        h = lib().ko_index_build(_vp(residues), _vp(off), _vp(ids), len(ids), n_threads)
        return cls(h)

    @classmethod
    def from_arrays(cls, keys, offsets, postings, n_proteins=0, n_aa=0, n_kmers=0):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        postings = np.ascontiguousarray(postings, dtype=np.uint32)
        h = lib().ko_index_from_arrays(_vp(keys), _vp(offsets), _vp(postings), len(keys),
                                       n_proteins, n_aa, n_kmers)
        return cls(h)

    def __del__(self):
        try:
            if self._h:
                lib().ko_index_free(self._h)
                self._h = None
        except Exception:
            pass


@dataclass
class Result:
    hit_off: np.ndarray
    subject: np.ndarray
    kmatch: np.ndarray
    size_in_kmer: np.ndarray
    pos_off: np.ndarray
    pos: np.ndarray
    n_lookups: int
    n_increments: int
    row_contig: np.ndarray | None = None
    row_start: np.ndarray | None = None
    row_end: np.ndarray | None = None
    row_plus: np.ndarray | None = None
    row_seq: np.ndarray | None = None
    row_seq_off: np.ndarray | None = None

    @property
    def n_rows(self):
        return len(self.hit_off) - 1

    def hits(self, i):
        b, e = int(self.hit_off[i]), int(self.hit_off[i + 1])
        return list(zip(self.subject[b:e].tolist(), self.kmatch[b:e].tolist()))

    def positions(self, hit_index):
        b, e = int(self.pos_off[hit_index]), int(self.pos_off[hit_index + 1])
        return self.pos[b:e]


def _collect(h, nt: bool, want_pos: bool) -> Result:
    L = lib()
    n = L.ko_result_n_rows(h)
    hit_off = _np(L.ko_result_hit_off(h), n + 1, np.uint64)
    nh = int(hit_off[-1])
    r = Result(
        hit_off=hit_off,
        subject=_np(L.ko_result_subject(h), nh, np.uint32),
        kmatch=_np(L.ko_result_kmatch(h), nh, np.int64),
        size_in_kmer=_np(L.ko_result_size_in_kmer(h), n, np.int32),
        pos_off=np.zeros(1, np.uint64),
        pos=np.zeros(0, np.uint8),
        n_lookups=int(L.ko_result_n_lookups(h)),
        n_increments=int(L.ko_result_n_increments(h)),
    )
    if want_pos:
        r.pos_off = _np(L.ko_result_pos_off(h), nh + 1, np.uint64)
        r.pos = _np(L.ko_result_pos(h), int(r.pos_off[-1]), np.uint8)
    if nt:
        r.row_contig = _np(L.ko_result_row_contig(h), n, np.uint32)
        r.row_start = _np(L.ko_result_row_start(h), n, np.int64)
        r.row_end = _np(L.ko_result_row_end(h), n, np.int64)
        r.row_plus = _np(L.ko_result_row_plus(h), n, np.uint8)
        so = _np(L.ko_result_row_seq_off(h), n + 1, np.uint64)
        r.row_seq_off = so
        r.row_seq = _np(L.ko_result_row_seq(h), int(so[-1]), np.uint8)
    L.ko_result_free(h)
    return r


def search_proteins(idx: Index, residues, off, o: KoOpts | None = None, n_threads: int = 1) -> Result:
    o = o or opts()
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    h = lib().ko_search_proteins(idx._h, _vp(residues), _vp(off), len(off) - 1, C.byref(o), n_threads)
    return _collect(h, False, bool(o.want_positions))


def search_nucleotide(idx: Index, nt, off, o: KoOpts | None = None, n_threads: int = 1) -> Result:
    o = o or opts()
    nt = np.ascontiguousarray(nt, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    h = lib().ko_search_nucleotide(idx._h, _vp(nt), _vp(off), len(off) - 1, C.byref(o), n_threads)
    return _collect(h, True, True)


@dataclass
class Orfs:
    seqs: list
    start: np.ndarray
    end: np.ndarray
    plus: np.ndarray
    alts: list


def get_orfs(dna: bytes) -> Orfs:
    L = lib()
    a = np.frombuffer(dna, dtype=np.uint8).copy()
    h = L.ko_get_orfs(_vp(a), len(a))
    n = L.ko_orfs_n(h)
    so = _np(L.ko_orfs_seq_off(h), n + 1, np.uint64)
    seq = _np(L.ko_orfs_seq(h), int(so[-1]), np.uint8).tobytes()
    ao = _np(L.ko_orfs_alts_off(h), n + 1, np.uint64)
    al = _np(L.ko_orfs_alts(h), int(ao[-1]), np.int32)
    r = Orfs(
        seqs=[seq[int(so[i]):int(so[i + 1])] for i in range(n)],
        start=_np(L.ko_orfs_start(h), n, np.int64),
        end=_np(L.ko_orfs_end(h), n, np.int64),
        plus=_np(L.ko_orfs_plus(h), n, np.uint8),
        alts=[al[int(ao[i]):int(ao[i + 1])].tolist() for i in range(n)],
    )
    L.ko_orfs_free(h)
    return r


def filter_count(kmatch_sorted, size_in_kmer_, o: KoOpts | None = None) -> int:
    o = o or opts()
    a = np.ascontiguousarray(kmatch_sorted, dtype=np.int64)
    return int(lib().ko_filter_count(_vp(a), len(a), size_in_kmer_, C.byref(o)))


def aln_params(number_of_aa: int, lambda_=0.267, K=0.041, gap_open=11, gap_extend=1) -> KoAlnParams:
    """blosum62_11_1 defaults: pkg/align/matrixScores.go:59, api/server.go:204-206."""
    return KoAlnParams(lambda_, K, gap_open, gap_extend, number_of_aa)


def align(q: bytes, s: bytes, prm: KoAlnParams, want_strings: bool = False):
    out = KoAln()
    if want_strings:
        cap = len(q) + len(s) + 2
        a = C.create_string_buffer(cap)
        b = C.create_string_buffer(cap)
        lib().ko_align(q, len(q), s, len(s), C.byref(prm), C.byref(out), a, b, cap)
        return out, a.value, b.value
    lib().ko_align(q, len(q), s, len(s), C.byref(prm), C.byref(out), None, None, 0)
    return out


def aln_string(q: bytes, s: bytes, prm: KoAlnParams) -> bytes:
    """AlignmentResult.AlnString (align.go:103): query line, match line, subject line."""
    _, a, b = align(q, s, prm, want_strings=True)
    cap = 3 * len(a) + 8
    buf = C.create_string_buffer(cap)
    n = lib().ko_aln_string(a, b, len(a), buf, cap)
    assert n >= 0
    return buf.raw[:n]


def set_align_model(matrix26, gap_open: int) -> None:
    """scores in biogo order "-ABCDEFGHIJKLMNPQRSTVWXYZ*" (row/column 0 = gap cost); None resets"""
    if matrix26 is None:
        lib().ko_reset_align_model()
        return
    m = np.ascontiguousarray(matrix26, dtype=np.int8).reshape(26, 26)
    lib().ko_set_align_model(_vp(m), int(gap_open))


def set_genetic_code(aas64: bytes | None, start_mask: int = 0) -> None:
    lib().ko_set_genetic_code(aas64, int(start_mask))


def blosum62() -> np.ndarray:
    L = lib()
    return np.array([[L.ko_blosum62(i, j) for j in range(26)] for i in range(26)], dtype=np.int32)


def bitscore(raw, lambda_=0.267, K=0.041):
    return float(lib().ko_bitscore(lambda_, K, raw))


def evalue(qlen, number_of_aa, bits):
    return float(lib().ko_evalue(qlen, number_of_aa, bits))


def format_positions(pos, with_alignment=False) -> str:
    a = np.ascontiguousarray(pos, dtype=np.uint8)
    cap = 16 * (len(a) + 2)
    buf = C.create_string_buffer(cap)
    lib().ko_format_positions(_vp(a), len(a), int(with_alignment), buf, cap)
    return buf.value.decode()


# ---- CPU twin of the synthetic C4 generator (oracle/synth_oracle.cpp, include/kaamer_synth_spec.h) ----
def synth_record(seed: int, i: int) -> bytes:
    L = lib()
    n = L.kso_record(seed, i, None, 0)
    buf = np.zeros(n, np.uint8)
    L.kso_record(seed, i, _vp(buf), n)
    return buf.tobytes()


def synth_record_meta(seed: int, i: int):
    f, ln, th = C.c_uint64(), C.c_uint32(), C.c_uint32()
    lib().kso_record_meta(seed, i, C.byref(f), C.byref(ln), C.byref(th))
    return f.value, ln.value, th.value


def synth_query(seed: int, n_proteins: int, j: int, batch: int = 0):
    """-> (letters, record index the query was sampled from)"""
    L = lib()
    rec = C.c_uint64()
    n = L.kso_query(seed, n_proteins, j, batch, None, 0, C.byref(rec))
    buf = np.zeros(n, np.uint8)
    L.kso_query(seed, n_proteins, j, batch, _vp(buf), n, C.byref(rec))
    return buf.tobytes(), rec.value


def synth_restricted_index(seed: int, n_proteins: int, query_seqs, n_threads: int = 1, id_base: int = 0) -> "Index":
    """The index of the whole synthetic database RESTRICTED to the k-mers the given queries look up, built by
    streaming the generator over every record (what the C4 parity check searches the sample against)."""
    keys = sorted({encode_kmer(s[k:k + 7]) for s in query_seqs for k in range(max(0, len(s) - 6))})
    keys = np.array(keys, dtype=np.uint32)
    offsets = np.zeros(len(keys) + 1, dtype=np.uint64)
    pp = C.POINTER(C.c_uint32)()
    stats = (C.c_uint64 * 3)()
    rc = lib().kso_restricted_index(seed, n_proteins, id_base, _vp(keys), len(keys), n_threads, _vp(offsets),
                                    C.byref(pp), stats)
    assert rc == 0
    n = int(offsets[-1])
    postings = _np(pp, n, np.uint32)
    lib().kso_free(pp)
    keep = np.flatnonzero(offsets[1:] > offsets[:-1])  # a key without postings is a miss (search.go:421-423)
    lens = (offsets[1:] - offsets[:-1])[keep]
    off2 = np.zeros(len(keep) + 1, dtype=np.uint64)
    off2[1:] = np.cumsum(lens)
    return Index.from_arrays(keys[keep], off2, postings, stats[0], stats[1], stats[2])
