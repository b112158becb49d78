"""ctypes binding of libkaamer_gpu.so (the C ABI in include/kaamer_gpu.h).

The product path: there is no CPU fallback here.  If the shared library is missing the
import of the symbols fails loudly; if no CUDA device is present every entry point returns
KAAMER_ERR_CUDA and `KaamerGpuError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkaamer_gpu.so")

KAAMER_OK = 0


class KaamerGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libkaamer_gpu error {code}: {msg}")
        self.code = code


class IndexView(C.Structure):
    _fields_ = [
        ("n_keys", C.c_uint64),
        ("n_postings", C.c_uint64),
        ("keys", C.c_void_p),
        ("offsets", C.c_void_p),
        ("postings", C.c_void_p),
        ("n_proteins", C.c_uint64),
        ("n_aa", C.c_uint64),
        ("n_kmers", C.c_uint64),
        ("max_protein_id", C.c_uint32),
        ("_pad", C.c_uint32),
        ("prot_seq_off", C.c_void_p),
        ("prot_residues", C.c_void_p),
        ("shard_lo", C.c_uint64),
        ("shard_hi", C.c_uint64),
    ]


class Opts(C.Structure):
    _fields_ = [
        ("min_kmatch", C.c_int64),
        ("min_kratio", C.c_double),
        ("max_results", C.c_int32),
        ("want_positions", C.c_uint8),
        ("_pad", C.c_uint8 * 3),
    ]


class Hits(C.Structure):
    _fields_ = [
        ("n_rows", C.c_uint32),
        ("_pad", C.c_uint32),
        ("n_hits", C.c_uint64),
        ("hit_off", C.POINTER(C.c_uint64)),
        ("subject_id", C.POINTER(C.c_uint32)),
        ("kmatch", C.POINTER(C.c_uint32)),
        ("size_in_kmer", C.POINTER(C.c_int32)),
        ("pos_off", C.POINTER(C.c_uint64)),
        ("pos", C.POINTER(C.c_uint8)),
        ("row_contig", C.POINTER(C.c_uint32)),
        ("row_start", C.POINTER(C.c_int64)),
        ("row_end", C.POINTER(C.c_int64)),
        ("row_plus", C.POINTER(C.c_uint8)),
        ("row_seq_off", C.POINTER(C.c_uint64)),
        ("row_seq", C.POINTER(C.c_uint8)),
        ("n_lookups", C.c_uint64),
        ("n_increments", C.c_uint64),
        ("_owner", C.c_void_p),
    ]


class Orfs(C.Structure):
    _fields_ = [
        ("n_orfs", C.c_uint64),
        ("contig", C.POINTER(C.c_uint32)),
        ("start", C.POINTER(C.c_int64)),
        ("end", C.POINTER(C.c_int64)),
        ("plus", C.POINTER(C.c_uint8)),
        ("seq_off", C.POINTER(C.c_uint64)),
        ("seq", C.POINTER(C.c_uint8)),
        ("alts_off", C.POINTER(C.c_uint64)),
        ("alts", C.POINTER(C.c_int32)),
        ("_owner", C.c_void_p),
    ]


class AlnOpts(C.Structure):
    _fields_ = [
        ("lambda_", C.c_double),
        ("K", C.c_double),
        ("gap_open", C.c_int32),
        ("gap_extend", C.c_int32),
        ("number_of_aa", C.c_uint64),
    ]


class Aln(C.Structure):
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
        ("status", C.c_int32),
    ]


class AlnModel(C.Structure):
    _fields_ = [("matrix", C.c_int8 * (26 * 26)), ("gap_open", C.c_int32)]


class AlnText(C.Structure):
    _fields_ = [("off", C.POINTER(C.c_uint64)), ("text", C.POINTER(C.c_char)), ("_owner", C.c_void_p)]


class DevResult(C.Structure):
    _fields_ = [
        ("n_hits", C.c_void_p),
        ("hit_base", C.c_void_p),
        ("size_in_kmer", C.c_void_p),
        ("pool", C.c_void_p),
        ("pool_cap", C.c_uint64),
        ("counters", C.c_void_p),
    ]


class ShardHandle(C.Structure):
    """kaamer_shard_handle (plain bytes; the two descriptors travel separately, SCM_RIGHTS)."""
    _fields_ = [
        ("shard_lo", C.c_uint64),
        ("shard_hi", C.c_uint64),
        ("n_postings", C.c_uint64),
        ("table_ptr", C.c_uint64),
        ("postings_ptr", C.c_uint64),
        ("table_bytes", C.c_uint64),
        ("postings_bytes", C.c_uint64),
        ("device", C.c_int32),
        ("pid", C.c_int32),
        ("table_fd", C.c_int32),
        ("postings_fd", C.c_int32),
    ]


class SynthCfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_proteins", C.c_uint64)]


class QueryBatch(C.Structure):
    _fields_ = [
        ("n_queries", C.c_uint32),
        ("_pad", C.c_uint32),
        ("n_residues", C.c_uint64),
        ("residues", C.c_void_p),
        ("seq_off", C.c_void_p),
        ("names", C.c_void_p),
        ("name_off", C.c_void_p),
        ("size_in_kmer", C.c_void_p),
        ("_owner", C.c_void_p),
    ]


# every symbol include/kaamer_gpu.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "kaamer_gpu_open",
    "kaamer_gpu_open_view",
    "kaamer_gpu_open_shard",
    "kaamer_gpu_kidx_fences",
    "kaamer_gpu_build",
    "kaamer_gpu_build_shard",
    "kaamer_gpu_close",
    "kaamer_gpu_builder_open",
    "kaamer_gpu_builder_pass_begin",
    "kaamer_gpu_builder_add_device",
    "kaamer_gpu_builder_pass_end",
    "kaamer_gpu_builder_finish",
    "kaamer_gpu_builder_abort",
    "kaamer_synth_record_lengths",
    "kaamer_synth_record_residues",
    "kaamer_synth_query_lengths",
    "kaamer_synth_query_residues",
    "kaamer_gpu_dbstats",
    "kaamer_gpu_index_sizes",
    "kaamer_gpu_index_copy",
    "kaamer_gpu_save",
    "kaamer_gpu_search_proteins",
    "kaamer_gpu_search_proteins_submit",
    "kaamer_gpu_search_proteins_wait",
    "kaamer_gpu_search_nucleotide",
    "kaamer_gpu_free_hits",
    "kaamer_gpu_set_genetic_code",
    "kaamer_gpu_get_orfs",
    "kaamer_gpu_free_orfs",
    "kaamer_gpu_align",
    "kaamer_gpu_align_text",
    "kaamer_gpu_free_aln_text",
    "kaamer_gpu_default_align_model",
    "kaamer_gpu_set_align_model",
    "kaamer_gpu_align_last_plan",
    "kaamer_gpu_search_proteins_device",
    "kaamer_gpu_dense_space",
    "kaamer_gpu_shard_route",
    "kaamer_gpu_shard_count",
    "kaamer_gpu_shard_gather",
    "kaamer_gpu_shard_merge",
    "kaamer_gpu_shard_export",
    "kaamer_gpu_attach_shards",
    "kaamer_gpu_detach_shards",
    "kaamer_gpu_pinned_alloc",
    "kaamer_gpu_pinned_free",
    "kaamer_host_format_positions",
    "kaamer_host_read_fasta",
    "kaamer_host_read_fastq",
    "kaamer_host_free_queries",
    "kaamer_gpu_set_annotations",
    "kaamer_host_format_tsv",
    "kaamer_host_free_text",
    "kaamer_gpu_profile_enable",
    "kaamer_gpu_profile_read",
    "kaamer_gpu_profile_host_read",
    "kaamer_gpu_last_error",
    "kaamer_gpu_version",
]

_lib = None


def lib() -> C.CDLL:
    """Load libkaamer_gpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C kaamer_b200/csrc` (there is no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.kaamer_gpu_last_error.restype = C.c_char_p
    L.kaamer_gpu_version.restype = C.c_char_p
    L.kaamer_gpu_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.kaamer_gpu_open_view.argtypes = [C.POINTER(IndexView), C.c_int, C.POINTER(vp)]
    L.kaamer_gpu_open_shard.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(vp)]
    L.kaamer_gpu_kidx_fences.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint64)]
    L.kaamer_gpu_build.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(vp)]
    L.kaamer_gpu_build_shard.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(vp)]
    L.kaamer_gpu_close.argtypes = [vp]
    L.kaamer_gpu_close.restype = None
    L.kaamer_gpu_builder_open.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(vp)]
    L.kaamer_gpu_builder_pass_begin.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.kaamer_gpu_builder_add_device.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint64, vp]
    L.kaamer_gpu_builder_pass_end.argtypes = [vp, vp]
    L.kaamer_gpu_builder_finish.argtypes = [vp, C.POINTER(vp)]
    L.kaamer_gpu_builder_abort.argtypes = [vp]
    L.kaamer_gpu_builder_abort.restype = None
    scp = C.POINTER(SynthCfg)
    L.kaamer_synth_record_lengths.argtypes = [scp, C.c_uint64, C.c_uint64, vp, vp]
    L.kaamer_synth_record_residues.argtypes = [scp, C.c_uint64, C.c_uint64, vp, vp, vp]
    L.kaamer_synth_query_lengths.argtypes = [scp, C.c_uint32, C.c_uint64, C.c_uint64, vp, vp]
    L.kaamer_synth_query_residues.argtypes = [scp, C.c_uint32, C.c_uint64, C.c_uint64, vp, vp, vp]
    u64p = C.POINTER(C.c_uint64)
    L.kaamer_gpu_dbstats.argtypes = [vp, u64p, u64p, u64p]
    L.kaamer_gpu_index_sizes.argtypes = [vp, u64p, u64p]
    L.kaamer_gpu_index_copy.argtypes = [vp, vp, vp, vp]
    L.kaamer_gpu_save.argtypes = [vp, C.c_char_p]
    L.kaamer_gpu_search_proteins.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(Opts), C.POINTER(C.POINTER(Hits))]
    L.kaamer_gpu_search_proteins_submit.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(Opts), C.POINTER(C.c_int32)]
    L.kaamer_gpu_search_proteins_wait.argtypes = [vp, C.c_int32, C.POINTER(C.POINTER(Hits))]
    L.kaamer_gpu_search_nucleotide.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(Opts), C.POINTER(C.POINTER(Hits))]
    L.kaamer_gpu_free_hits.argtypes = [C.POINTER(Hits)]
    L.kaamer_gpu_free_hits.restype = None
    L.kaamer_gpu_set_genetic_code.argtypes = [vp, C.c_char_p, C.c_uint64]
    L.kaamer_gpu_get_orfs.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(C.POINTER(Orfs))]
    L.kaamer_gpu_free_orfs.argtypes = [C.POINTER(Orfs)]
    L.kaamer_gpu_free_orfs.restype = None
    L.kaamer_gpu_align.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.POINTER(AlnOpts), vp]
    L.kaamer_gpu_align_text.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.POINTER(AlnOpts), vp,
                                        C.POINTER(C.POINTER(AlnText))]
    L.kaamer_gpu_free_aln_text.argtypes = [C.POINTER(AlnText)]
    L.kaamer_gpu_free_aln_text.restype = None
    L.kaamer_gpu_default_align_model.argtypes = [C.POINTER(AlnModel)]
    L.kaamer_gpu_set_align_model.argtypes = [vp, C.POINTER(AlnModel)]
    L.kaamer_gpu_align_last_plan.argtypes = [C.POINTER(C.c_uint32)]
    L.kaamer_gpu_align_last_plan.restype = None
    L.kaamer_gpu_search_proteins_device.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(Opts), C.POINTER(DevResult), vp]
    L.kaamer_gpu_dense_space.restype = C.c_uint64
    L.kaamer_gpu_dense_space.argtypes = []
    L.kaamer_gpu_shard_route.argtypes = [vp, vp, vp, C.c_uint32, u64p, C.c_int, vp, vp, vp, vp, vp]
    L.kaamer_gpu_shard_count.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint64, vp, vp]
    L.kaamer_gpu_shard_gather.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, vp, vp]
    L.kaamer_gpu_shard_merge.argtypes = [vp, vp, vp, C.c_int, C.c_uint32, vp, C.POINTER(Opts), C.POINTER(DevResult), vp]
    L.kaamer_gpu_shard_export.argtypes = [vp, C.POINTER(ShardHandle)]
    L.kaamer_gpu_attach_shards.argtypes = [vp, C.POINTER(ShardHandle), C.c_int, C.c_int]
    L.kaamer_gpu_detach_shards.argtypes = [vp]
    L.kaamer_host_format_positions.argtypes = [vp, C.c_uint64, C.c_int, C.c_char_p, C.c_uint64]
    L.kaamer_host_format_positions.restype = C.c_int64
    L.kaamer_host_read_fasta.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.POINTER(QueryBatch))]
    L.kaamer_host_read_fastq.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.POINTER(QueryBatch))]
    L.kaamer_host_free_queries.argtypes = [C.POINTER(QueryBatch)]
    L.kaamer_host_free_queries.restype = None
    L.kaamer_gpu_set_annotations.argtypes = [vp, vp, vp, vp, C.c_uint32]
    L.kaamer_host_format_tsv.argtypes = [vp, C.POINTER(Hits), vp, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.POINTER(C.c_char)), u64p]
    L.kaamer_host_free_text.argtypes = [C.POINTER(C.c_char)]
    L.kaamer_host_free_text.restype = None
    L.kaamer_gpu_pinned_alloc.argtypes = [C.c_uint64, C.POINTER(vp)]
    L.kaamer_gpu_pinned_free.argtypes = [vp]
    L.kaamer_gpu_pinned_free.restype = None
    L.kaamer_gpu_profile_enable.argtypes = [vp, C.c_int]
    L.kaamer_gpu_profile_read.argtypes = [vp, vp, vp, u64p, C.c_int]
    L.kaamer_gpu_profile_host_read.argtypes = [vp, vp, C.c_int]
    for s in SYMBOLS:
        getattr(L, s)  # AttributeError if the .so lacks a declared symbol
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != KAAMER_OK:
        raise KaamerGpuError(rc, lib().kaamer_gpu_last_error().decode(errors="replace"))
