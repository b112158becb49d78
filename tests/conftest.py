import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
