"""Query readers with the reference's semantics (GetQueriesFasta / GetQueriesFastq,
pkg/search/search.go:222-412), implemented in the native library (csrc/reader.cu) and returned as
the flat batches the search entry points take."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check


@dataclass
class QueryBatch:
    residues: np.ndarray      # u8, Query.Sequence concatenated
    seq_off: np.ndarray       # u64[n+1]
    names: list               # Query.Name
    size_in_kmer: np.ndarray  # i32[n], as the reader computes it

    def __len__(self):
        return len(self.names)

    def sequence(self, i: int) -> bytes:
        return self.residues[int(self.seq_off[i]):int(self.seq_off[i + 1])].tobytes()


def _collect(bp) -> QueryBatch:
    b = bp.contents
    n = int(b.n_queries)

    def arr(ptr, count, dtype):
        if count == 0:
            return np.zeros(0, dtype)
        buf = (C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=count).copy()

    seq_off = arr(b.seq_off, n + 1, np.uint64)
    name_off = arr(b.name_off, n + 1, np.uint64)
    blob = arr(b.names, int(name_off[-1]), np.uint8).tobytes()
    out = QueryBatch(arr(b.residues, int(b.n_residues), np.uint8), seq_off,
                     [blob[int(name_off[i]):int(name_off[i + 1])].decode("utf-8", "replace") for i in range(n)],
                     arr(b.size_in_kmer, n, np.int32))
    _lib.lib().kaamer_host_free_queries(bp)
    return out


def read_fasta(path: str, is_protein: bool = True, pinned: bool = False) -> QueryBatch:
    bp = C.POINTER(_lib.QueryBatch)()
    check(_lib.lib().kaamer_host_read_fasta(path.encode(), int(is_protein), int(pinned), C.byref(bp)))
    return _collect(bp)


def read_fastq(path: str, pinned: bool = False) -> QueryBatch:
    bp = C.POINTER(_lib.QueryBatch)()
    check(_lib.lib().kaamer_host_read_fastq(path.encode(), int(pinned), C.byref(bp)))
    return _collect(bp)


def format_positions(pos, with_alignment: bool = False) -> str:
    """FormatPositionsToString (pkg/search/search.go:694-742) of one PositionHits row."""
    a = np.ascontiguousarray(pos, dtype=np.uint8)
    cap = 24 * (len(a) // 2 + 2)
    buf = C.create_string_buffer(cap)
    n = _lib.lib().kaamer_host_format_positions(a.ctypes.data_as(C.c_void_p), len(a), int(with_alignment), buf, cap)
    if n < 0:
        raise ValueError("format_positions: buffer too small")
    return buf.value.decode()
