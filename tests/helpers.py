"""Parity comparators shared by the GPU tests."""
import numpy as np


def assert_same_hits(gpu, ora, what=""):
    """Bit-exact: SizeInKmer, hit sets, Kmatch and ranking (canonical tie-break).

    The reference's order among equal Kmatch is random (search.go:132-152); both sides use
    (Kmatch desc, id asc).  When MaxResults cuts through a tie group the reference keeps an
    arbitrary subset of it — the canonical choice (smallest ids) is what both sides implement.
    """
    assert gpu.n_rows == ora.n_rows, f"{what}: rows {gpu.n_rows} != {ora.n_rows}"
    np.testing.assert_array_equal(gpu.size_in_kmer, ora.size_in_kmer, err_msg=f"{what}: SizeInKmer")
    np.testing.assert_array_equal(gpu.hit_off.astype(np.int64), ora.hit_off.astype(np.int64), err_msg=f"{what}: hit_off")
    np.testing.assert_array_equal(gpu.subject, ora.subject, err_msg=f"{what}: subject ids")
    np.testing.assert_array_equal(gpu.kmatch.astype(np.int64), ora.kmatch.astype(np.int64), err_msg=f"{what}: Kmatch")
