import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count() -> int:
    """cudaGetDeviceCount through the runtime the product links (no torch import)."""
    import ctypes

    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    return 0


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without a CUDA device, so that a plain
    `pytest tests` is green here and on the B200 box alike."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product path has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def small_db():
    """C1-like synthetic DB at a size the oracle finishes in a second: 2000 proteins."""
    from kaamer_b200 import synth
    from oracle import oracle as o

    res, off = synth.protein_db(2000, config_index=1)
    ids = o.fasta_ids(len(off) - 1)
    idx = o.Index.build(res, off, ids, 4)
    return {"res": res, "off": off, "ids": ids, "idx": idx}
