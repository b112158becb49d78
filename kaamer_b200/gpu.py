"""Host-side handle on the device-resident index (thin wrapper over the C ABI).

Mirrors the seam of the reference: `KVStoresNew(..., readOnly)` at server start
(api/server.go:65) -> `GpuIndex.open`; per request `KmerSearch` + `sortMapByValue` +
`FilterResults` (pkg/search/search.go:132-220,414-440) -> `GpuIndex.search_proteins`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check


@dataclass
class SearchOptions:
    """The SearchOptions fields that reach the hot path (pkg/search/search.go:56-71).
    Defaults are the server's (api/server.go:194-207)."""

    max_results: int = 10
    min_kmatch: int = 10
    min_kratio: float = 0.05
    extract_positions: bool = False

    def c(self) -> _lib.Opts:
        o = _lib.Opts()
        o.min_kmatch = int(self.min_kmatch)
        o.min_kratio = float(self.min_kratio)
        o.max_results = int(self.max_results)
        o.want_positions = 1 if self.extract_positions else 0
        return o


@dataclass
class SearchResult:
    hit_off: np.ndarray       # u64[n_rows+1]
    subject: np.ndarray       # u32[n_hits]   Hit.Key
    kmatch: np.ndarray        # u32[n_hits]   Hit.Kmatch
    size_in_kmer: np.ndarray  # i32[n_rows]   Query.SizeInKmer
    n_lookups: int
    n_increments: int
    pos_off: np.ndarray | None = None
    pos: np.ndarray | None = None
    row_contig: np.ndarray | None = None
    row_start: np.ndarray | None = None
    row_end: np.ndarray | None = None
    row_plus: np.ndarray | None = None
    row_seq_off: np.ndarray | None = None
    row_seq: np.ndarray | None = None

    @property
    def n_rows(self) -> int:
        return len(self.hit_off) - 1

    def hits(self, i: int):
        b, e = int(self.hit_off[i]), int(self.hit_off[i + 1])
        return list(zip(self.subject[b:e].tolist(), self.kmatch[b:e].tolist()))


@dataclass
class OrfTable:
    """ORF{Sequence, Location{StartPosition, EndPosition, PlusStrand, StartsAlternative}} rows
    (pkg/search/dna.go:34-44), contigs concatenated in batch order."""

    contig: np.ndarray
    start: np.ndarray
    end: np.ndarray
    plus: np.ndarray
    seq_off: np.ndarray
    seq: np.ndarray
    alts_off: np.ndarray
    alts: np.ndarray

    def __len__(self):
        return len(self.start)

    def sequence(self, i: int) -> bytes:
        return self.seq[int(self.seq_off[i]):int(self.seq_off[i + 1])].tobytes()

    def starts_alternative(self, i: int):
        return self.alts[int(self.alts_off[i]):int(self.alts_off[i + 1])].tolist()


# kaamer_aln (include/kaamer_gpu.h)
ALN_DTYPE = np.dtype([("identity", "<f4"), ("similarity", "<f4"), ("length", "<i4"), ("mismatches", "<i4"),
                      ("gap_openings", "<i4"), ("raw", "<i4"), ("bitscore", "<f8"), ("evalue", "<f8"),
                      ("query_start", "<i4"), ("query_end", "<i4"), ("subject_start", "<i4"),
                      ("subject_end", "<i4"), ("dp_score", "<i4"), ("status", "<i4")], align=True)


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _HitsOwner:
    """Keeps a library-owned kaamer_hits alive while numpy views of its pinned buffers exist."""

    def __init__(self, hp):
        self.hp = hp

    def __del__(self):
        try:
            if self.hp:
                _lib.lib().kaamer_gpu_free_hits(self.hp)
                self.hp = None
        except Exception:
            pass


def _view(ptr, n, dtype, owner):
    """zero-copy numpy view of a library-owned buffer (released when the last view dies)"""
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(int(n),))
    a = a.view(dtype) if a.dtype != np.dtype(dtype) else a
    v = a.view(_OwnedArray)
    v._owner = owner
    return v


class _OwnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


def format_tsv(result: "SearchResult", names, aln=None, seq_off=None, is_protein: bool = True,
               with_positions: bool = False, with_annotations: bool = False, index: "GpuIndex | None" = None) -> bytes:
    """The TSV rows of a whole batch in one native call (kaamer_host_format_tsv): byte for byte what the
    reference's QueryResultHandler sends to its writer (pkg/search/search.go:505-606).  `names`: Query.Name
    per query (protein batch) or per contig (nucleotide batch); `aln`: the structured array of
    GpuIndex.align for the hits, in hits order, or None."""
    names = [n if isinstance(n, bytes) else n.encode() for n in names]
    name_off = np.zeros(len(names) + 1, dtype=np.uint64)
    if names:
        name_off[1:] = np.cumsum([len(n) for n in names], dtype=np.uint64)
    blob = np.frombuffer(b"".join(names) + b"\0", dtype=np.uint8).copy()
    h = _lib.Hits()
    keep = []

    def ptr(a, dtype, ctype):
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(ctype))

    h.n_rows = result.n_rows
    h.n_hits = len(result.subject)
    h.hit_off = ptr(result.hit_off, np.uint64, C.c_uint64)
    h.subject_id = ptr(result.subject, np.uint32, C.c_uint32)
    h.kmatch = ptr(result.kmatch, np.uint32, C.c_uint32)
    h.size_in_kmer = ptr(result.size_in_kmer, np.int32, C.c_int32)
    if result.pos_off is not None and result.pos is not None:
        h.pos_off = ptr(result.pos_off, np.uint64, C.c_uint64)
        h.pos = ptr(result.pos if len(result.pos) else np.zeros(1, np.uint8), np.uint8, C.c_uint8)
    if result.row_start is not None:
        h.row_contig = ptr(result.row_contig, np.uint32, C.c_uint32)
        h.row_start = ptr(result.row_start, np.int64, C.c_int64)
        h.row_end = ptr(result.row_end, np.int64, C.c_int64)
    so = None
    if seq_off is not None:
        so = np.ascontiguousarray(seq_off, dtype=np.uint64)
    al = None
    if aln is not None:
        al = np.ascontiguousarray(aln)
        assert al.dtype == ALN_DTYPE and len(al) == h.n_hits
    out = C.POINTER(C.c_char)()
    n = C.c_uint64()
    check(_lib.lib().kaamer_host_format_tsv(index._h if index is not None else None, C.byref(h),
                                            _vp(al) if al is not None else None, _vp(blob), _vp(name_off),
                                            _vp(so) if so is not None else None, int(is_protein), int(with_positions),
                                            int(with_annotations), C.byref(out), C.byref(n)))
    text = C.string_at(out, n.value)
    _lib.lib().kaamer_host_free_text(out)
    return text


def _collect_hits(hp) -> SearchResult:
    h = hp.contents
    n, nh = h.n_rows, h.n_hits
    own = _HitsOwner(hp)
    r = SearchResult(
        hit_off=_view(h.hit_off, n + 1, np.uint64, own),
        subject=_view(h.subject_id, nh, np.uint32, own),
        kmatch=_view(h.kmatch, nh, np.uint32, own),
        size_in_kmer=_view(h.size_in_kmer, n, np.int32, own),
        n_lookups=int(h.n_lookups),
        n_increments=int(h.n_increments),
    )
    if h.pos_off:
        r.pos_off = _view(h.pos_off, nh + 1, np.uint64, own)
        r.pos = _view(h.pos, int(r.pos_off[-1]) if nh else 0, np.uint8, own)
    if h.row_start:
        r.row_contig = _view(h.row_contig, n, np.uint32, own)
        r.row_start = _view(h.row_start, n, np.int64, own)
        r.row_end = _view(h.row_end, n, np.int64, own)
        r.row_plus = _view(h.row_plus, n, np.uint8, own)
        r.row_seq_off = _view(h.row_seq_off, n + 1, np.uint64, own)
        r.row_seq = _view(h.row_seq, int(r.row_seq_off[-1]) if n else 0, np.uint8, own)
    return r


class GpuIndex:
    """A k-mer index resident in the HBM of one GPU."""

    def __init__(self, handle, device: int):
        self._h = handle
        self.device = device

    # ---- construction ------------------------------------------------------------------
    @classmethod
    def open(cls, kidx_path: str, device: int = 0) -> "GpuIndex":
        h = C.c_void_p()
        check(_lib.lib().kaamer_gpu_open(kidx_path.encode(), device, C.byref(h)))
        return cls(h, device)

    @classmethod
    def open_shard(cls, kidx_path: str, shard, device: int = 0) -> "GpuIndex":
        """The key range `shard = (lo, hi)` of a `.kidx` file, in shareable memory (mode P / mode S)."""
        h = C.c_void_p()
        check(_lib.lib().kaamer_gpu_open_shard(kidx_path.encode(), device, int(shard[0]), int(shard[1]), C.byref(h)))
        return cls(h, device)

    @staticmethod
    def kidx_fences(kidx_path: str, n_shards: int) -> np.ndarray:
        """Contiguous key ranges of equal posting mass for a `.kidx` file: u64[n_shards+1]."""
        f = (C.c_uint64 * (n_shards + 1))()
        check(_lib.lib().kaamer_gpu_kidx_fences(kidx_path.encode(), n_shards, f))
        return np.array(list(f), dtype=np.uint64)

    @classmethod
    def from_arrays(cls, keys, offsets, postings, n_proteins=0, n_aa=0, n_kmers=0, max_protein_id=None,
                    prot_seq_off=None, prot_residues=None, shard=(0, 0), device: int = 0) -> "GpuIndex":
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        postings = np.ascontiguousarray(postings, dtype=np.uint32)
        v = _lib.IndexView()
        v.n_keys, v.n_postings = len(keys), len(postings)
        v.keys, v.offsets, v.postings = _vp(keys), _vp(offsets), _vp(postings)
        v.n_proteins, v.n_aa, v.n_kmers = int(n_proteins), int(n_aa), int(n_kmers)
        if max_protein_id is None:
            max_protein_id = int(postings.max()) if len(postings) else 0
        v.max_protein_id = int(max_protein_id)
        keep = [keys, offsets, postings]
        if prot_seq_off is not None and prot_residues is not None:
            po = np.ascontiguousarray(prot_seq_off, dtype=np.uint64)
            pr = np.ascontiguousarray(prot_residues, dtype=np.uint8)
            assert len(po) == v.max_protein_id + 2
            v.prot_seq_off, v.prot_residues = _vp(po), _vp(pr)
            keep += [po, pr]
        v.shard_lo, v.shard_hi = int(shard[0]), int(shard[1])
        h = C.c_void_p()
        check(_lib.lib().kaamer_gpu_open_view(C.byref(v), device, C.byref(h)))
        return cls(h, device)

    @classmethod
    def build(cls, residues, seq_off, ids, keep_proteins: bool = True, device: int = 0, shard=(0, 0)) -> "GpuIndex":
        """makedb + indexdb semantics on the GPU (pkg/makedb/inputFASTA.go:195-250,
        pkg/indexdb/indexdb.go:68-150); ids[i] = protein id of record i.  `shard` = dense-code
        range [lo, hi) to keep (mode S), (0, 0) = everything."""
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        assert len(seq_off) == len(ids) + 1
        h = C.c_void_p()
        check(_lib.lib().kaamer_gpu_build_shard(_vp(residues), _vp(seq_off), _vp(ids), len(ids), int(keep_proteins),
                                                device, int(shard[0]), int(shard[1]), C.byref(h)))
        return cls(h, device)

    def close(self):
        if self._h:
            _lib.lib().kaamer_gpu_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- introspection -----------------------------------------------------------------
    def dbstats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(_lib.lib().kaamer_gpu_dbstats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"NumberOfProteins": a.value, "NumberOfAA": b.value, "NumberOfKmers": c.value}

    def index_arrays(self):
        nk, npst = C.c_uint64(), C.c_uint64()
        check(_lib.lib().kaamer_gpu_index_sizes(self._h, C.byref(nk), C.byref(npst)))
        keys = np.zeros(nk.value, np.uint32)
        offsets = np.zeros(nk.value + 1, np.uint64)
        postings = np.zeros(npst.value, np.uint32)
        check(_lib.lib().kaamer_gpu_index_copy(self._h, _vp(keys), _vp(offsets), _vp(postings)))
        return keys, offsets, postings

    def save(self, path: str):
        check(_lib.lib().kaamer_gpu_save(self._h, path.encode()))

    def set_annotations(self, entry_ids, lengths) -> None:
        """Protein.EntryId / Protein.Length by protein id (pkg/kvstore/protein.proto) for the row formatter
        (format_tsv); saved with the index.  entry_ids: list of str / bytes indexed by protein id ("" for
        unused ids), lengths: int32 per id."""
        ids = [e if isinstance(e, bytes) else e.encode() for e in entry_ids]
        off = np.zeros(len(ids) + 1, dtype=np.uint64)
        if ids:
            off[1:] = np.cumsum([len(e) for e in ids], dtype=np.uint64)
        blob = np.frombuffer(b"".join(ids) + b"\0", dtype=np.uint8).copy()
        ln = np.ascontiguousarray(lengths, dtype=np.int32)
        assert len(ln) == len(ids) and len(ids) >= 1
        check(_lib.lib().kaamer_gpu_set_annotations(self._h, _vp(blob), _vp(off), _vp(ln), len(ids) - 1))

    # ---- peer-mapped shards (mode P) -----------------------------------------------------
    def export_shard(self) -> "_lib.ShardHandle":
        """kaamer_shard_handle of this handle's key range.  `table_fd` / `postings_fd` are open
        descriptors of this process (-1 for a full index): send them with SCM_RIGHTS to other
        processes (kaamer_b200.peer does), then os.close them."""
        sh = _lib.ShardHandle()
        check(_lib.lib().kaamer_gpu_shard_export(self._h, C.byref(sh)))
        return sh

    def attach_shards(self, handles, presence_filter: bool = True, replicate_table: bool = False,
                      replicate_postings: bool = False) -> None:
        """Map the key-range shards of all ranks (this one included): afterwards the search entry
        points of this handle see the whole key space and probe remote shards through NVLink.
        `handles`: ShardHandle structs (or their bytes) whose descriptors are valid in THIS process."""
        arr = (_lib.ShardHandle * len(handles))()
        for i, b in enumerate(handles):
            b = bytes(b)
            assert len(b) == C.sizeof(_lib.ShardHandle)
            C.memmove(C.byref(arr[i]), b, len(b))
        flags = (0 if presence_filter else 1) | (2 if (replicate_table or replicate_postings) else 0) | \
            (4 if replicate_postings else 0)
        check(_lib.lib().kaamer_gpu_attach_shards(self._h, arr, len(handles), flags))

    def detach_shards(self) -> None:
        check(_lib.lib().kaamer_gpu_detach_shards(self._h))

    # ---- search ------------------------------------------------------------------------
    def search_proteins(self, residues, seq_off, opts: SearchOptions | None = None) -> SearchResult:
        opts = opts or SearchOptions()
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        return self.search_proteins_ptr(residues.ctypes.data, seq_off.ctypes.data, len(seq_off) - 1, opts)

    def search_proteins_ptr(self, res_ptr: int, off_ptr: int, nq: int, opts: SearchOptions) -> SearchResult:
        """Same call with raw host pointers (e.g. pinned buffers)."""
        o = opts.c()
        hp = C.POINTER(_lib.Hits)()
        check(_lib.lib().kaamer_gpu_search_proteins(self._h, C.c_void_p(res_ptr), C.c_void_p(off_ptr), nq,
                                                    C.byref(o), C.byref(hp)))
        return _collect_hits(hp)

    def submit_proteins_ptr(self, res_ptr: int, off_ptr: int, nq: int, opts: SearchOptions) -> int:
        """First half of the pipelined call (kaamer_gpu_search_proteins_submit): enqueue the batch, return a
        ticket.  The buffers behind the pointers must stay alive and unchanged until `wait_proteins`."""
        o = opts.c()
        t = C.c_int32(-1)
        check(_lib.lib().kaamer_gpu_search_proteins_submit(self._h, C.c_void_p(res_ptr), C.c_void_p(off_ptr), nq,
                                                           C.byref(o), C.byref(t)))
        return t.value

    def wait_proteins(self, ticket: int) -> SearchResult:
        hp = C.POINTER(_lib.Hits)()
        check(_lib.lib().kaamer_gpu_search_proteins_wait(self._h, ticket, C.byref(hp)))
        return _collect_hits(hp)

    def search_proteins_device(self, d_res: int, d_off: int, nq: int, opts: SearchOptions, n_hits: int,
                               hit_base: int, size_in_kmer: int, pool: int, pool_cap: int, counters: int,
                               stream: int = 0):
        """Device-resident, asynchronous on `stream` (all arguments are device pointers)."""
        o = opts.c()
        dr = _lib.DevResult(n_hits, hit_base, size_in_kmer, pool, pool_cap, counters)
        check(_lib.lib().kaamer_gpu_search_proteins_device(self._h, C.c_void_p(d_res), C.c_void_p(d_off), nq,
                                                           C.byref(o), C.byref(dr), C.c_void_p(stream)))

    def search_nucleotide(self, nt, contig_off, opts: SearchOptions | None = None) -> SearchResult:
        opts = opts or SearchOptions()
        nt = np.ascontiguousarray(nt, dtype=np.uint8)
        contig_off = np.ascontiguousarray(contig_off, dtype=np.uint64)
        o = opts.c()
        hp = C.POINTER(_lib.Hits)()
        check(_lib.lib().kaamer_gpu_search_nucleotide(self._h, _vp(nt), _vp(contig_off), len(contig_off) - 1,
                                                      C.byref(o), C.byref(hp)))
        return _collect_hits(hp)

    def set_genetic_code(self, aas64: bytes | None = None, start_mask: int = 0) -> None:
        """Genetic code of the translated search (explicit opt-in: the reference always uses table 11,
        pkg/search/dna.go:106).  aas64: amino-acid letters of the 64 codons in TCAG order ('*' = stop),
        start_mask: bit i = codon i starts an ORF.  None restores table 11."""
        check(_lib.lib().kaamer_gpu_set_genetic_code(self._h, aas64, int(start_mask)))

    def get_orfs(self, nt, contig_off) -> "OrfTable":
        """GetORFs (pkg/search/dna.go:65-181) of every contig of the batch, on the device."""
        nt = np.ascontiguousarray(nt, dtype=np.uint8)
        contig_off = np.ascontiguousarray(contig_off, dtype=np.uint64)
        op = C.POINTER(_lib.Orfs)()
        check(_lib.lib().kaamer_gpu_get_orfs(self._h, _vp(nt), _vp(contig_off), len(contig_off) - 1, C.byref(op)))
        o = op.contents
        n = int(o.n_orfs)
        seq_off = _arr(o.seq_off, n + 1, np.uint64)
        alts_off = _arr(o.alts_off, n + 1, np.uint64)
        t = OrfTable(
            contig=_arr(o.contig, n, np.uint32), start=_arr(o.start, n, np.int64), end=_arr(o.end, n, np.int64),
            plus=_arr(o.plus, n, np.uint8), seq_off=seq_off,
            seq=_arr(o.seq, int(seq_off[-1]) if n else 0, np.uint8), alts_off=alts_off,
            alts=_arr(o.alts, int(alts_off[-1]) if n else 0, np.int32))
        _lib.lib().kaamer_gpu_free_orfs(op)
        return t

    @staticmethod
    def default_align_model():
        """(int8[26, 26] scores in biogo order "-ABCDEFGHIJKLMNPQRSTVWXYZ*", gap_open) of the reference:
        BLOSUM62, zero gap row, -11 (pkg/align/align.go:62-65)"""
        m = _lib.AlnModel()
        check(_lib.lib().kaamer_gpu_default_align_model(C.byref(m)))
        return np.array(list(m.matrix), dtype=np.int8).reshape(26, 26), int(m.gap_open)

    @staticmethod
    def align_last_plan():
        """(long pairs, single pairs, packed int16x2 jobs) of this thread's last `align` call"""
        out = (C.c_uint32 * 3)()
        _lib.lib().kaamer_gpu_align_last_plan(out)
        return int(out[0]), int(out[1]), int(out[2])

    def set_align_model(self, matrix26=None, gap_open: int = -11) -> None:
        """Replace the DP model of `align` (explicit opt-in: the reference hard-wires BLOSUM62 / -11).
        matrix26[26, 26] in biogo order, row / column 0 = per-residue gap cost; None restores the default."""
        if matrix26 is None:
            check(_lib.lib().kaamer_gpu_set_align_model(self._h, None))
            return
        m = _lib.AlnModel()
        flat = np.ascontiguousarray(matrix26, dtype=np.int8).reshape(-1)
        assert len(flat) == 26 * 26
        for i, v in enumerate(flat.tolist()):
            m.matrix[i] = v
        m.gap_open = int(gap_open)
        check(_lib.lib().kaamer_gpu_set_align_model(self._h, C.byref(m)))

    def align(self, q_residues, q_off, pair_query, pair_subject, lambda_: float = 0.267, K: float = 0.041,
              gap_open: int = 11, gap_extend: int = 1, number_of_aa: int = 0, want_text: bool = False, out=None):
        """align.Align (pkg/align/align.go:46-161) for (query index, subject protein id) pairs;
        defaults = blosum62_11_1 (pkg/align/matrixScores.go:59).  Returns a structured array
        with the fields of AlignmentResult (align.go:25-40).  `out`: optional caller-owned result array (ALN_DTYPE,
        one row per pair; page-locked memory makes the device-to-host copy a DMA)."""
        q_residues = np.ascontiguousarray(q_residues, dtype=np.uint8)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        pq = np.ascontiguousarray(pair_query, dtype=np.uint32)
        ps = np.ascontiguousarray(pair_subject, dtype=np.uint32)
        assert len(pq) == len(ps)
        o = _lib.AlnOpts(lambda_, K, gap_open, gap_extend, number_of_aa)
        if out is None:
            out = np.zeros(len(pq), dtype=ALN_DTYPE)
        assert out.dtype == ALN_DTYPE and len(out) == len(pq) and out.flags["C_CONTIGUOUS"]
        if not want_text:
            check(_lib.lib().kaamer_gpu_align(self._h, _vp(q_residues), _vp(q_off), _vp(pq), _vp(ps), len(pq),
                                              C.byref(o), _vp(out)))
            return out
        # with AlignmentResult.AlnString (align.go:103) of every pair
        tp = C.POINTER(_lib.AlnText)()
        check(_lib.lib().kaamer_gpu_align_text(self._h, _vp(q_residues), _vp(q_off), _vp(pq), _vp(ps), len(pq),
                                               C.byref(o), _vp(out), C.byref(tp)))
        t = tp.contents
        off = _arr(t.off, len(pq) + 1, np.uint64)
        blob = C.string_at(t.text, int(off[-1])) if len(pq) else b""
        _lib.lib().kaamer_gpu_free_aln_text(tp)
        return out, [blob[int(off[i]):int(off[i + 1])] for i in range(len(pq))]

    # ---- profiling ---------------------------------------------------------------------
    def profile_enable(self, on: bool = True):
        check(_lib.lib().kaamer_gpu_profile_enable(self._h, int(on)))

    def profile_host_read(self, reset: bool = True):
        """Host wall clock per phase of the host-buffer calls (ms, summed since the last reset)."""
        ms = (C.c_double * 8)()
        check(_lib.lib().kaamer_gpu_profile_host_read(self._h, ms, int(reset)))
        names = ["translate_orfs", "count_pass", "finish_rows_d2h", "search_nucleotide_call", "search_proteins_call",
                 "finish_positions", "finish_result_buffers", "finish_assemble_d2h"]
        return {n: ms[i] for i, n in enumerate(names)}

    def profile_read(self, reset: bool = True):
        ms = (C.c_double * 8)()
        k = (C.c_uint64 * 8)()
        a = C.c_uint64()
        check(_lib.lib().kaamer_gpu_profile_read(self._h, ms, k, C.byref(a), int(reset)))
        return {"kernel_ms": list(ms), "kernel_launches": list(k), "all_launches": a.value}
