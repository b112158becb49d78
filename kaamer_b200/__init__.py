"""kaamer_b200 — B200-native (sm_100a) implementation of the zorino/kaamer search hot path.

Product path = libkaamer_gpu.so (hand-written CUDA behind the C ABI of include/kaamer_gpu.h).
This package holds only the host-side mirror of the reference interface and the loaders;
nothing here computes search results on the CPU.
"""
from ._lib import KaamerGpuError, LIB_PATH  # noqa: F401
from .gpu import GpuIndex, OrfTable, SearchOptions, SearchResult, format_tsv  # noqa: F401

__all__ = ["GpuIndex", "OrfTable", "SearchOptions", "SearchResult", "KaamerGpuError", "LIB_PATH", "format_tsv"]
