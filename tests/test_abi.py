"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/kaamer_gpu.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from kaamer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(_lib.LIB_PATH):
        entry.build()


def _header_symbols():
    hdr = open(os.path.join(ROOT, "include", "kaamer_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(kaamer_(?:gpu|host|synth)_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"libkaamer_gpu.so does not export {s}"
    assert sorted(_lib.SYMBOLS) == syms, "python binding list out of sync with the header"


def test_struct_layouts_match_header():
    # sizes the cgo side relies on (plain C layout, no packing pragmas)
    assert C.sizeof(_lib.Opts) == 24
    assert C.sizeof(_lib.DevResult) == 48
    assert C.sizeof(_lib.AlnOpts) == 32
    assert C.sizeof(_lib.Aln) == 64
    assert C.sizeof(_lib.IndexView) == 104
    assert C.sizeof(_lib.Hits) == 136 and C.sizeof(_lib.Orfs) == 80
    assert C.sizeof(_lib.ShardHandle) == 72
    assert C.sizeof(_lib.QueryBatch) == 64


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kaamer_b200 import GpuIndex, KaamerGpuError

    with pytest.raises(KaamerGpuError) as ei:
        GpuIndex.build(np.frombuffer(b"MKTAYIAKQR", np.uint8), np.array([0, 10], np.uint64), np.array([1], np.uint32))
    assert "no CPU fallback" in str(ei.value)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "kaamer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"
