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


def assert_same_orfs(table, oracle_orfs_per_contig):
    """GPU OrfTable (all contigs of a batch) vs a list of oracle Orfs (one per contig): bit-exact
    sequences, Location and StartsAlternative, in GetORFs order (dna.go:167-177; ties keep
    emission order on both sides)."""
    exp_n = sum(len(x.seqs) for x in oracle_orfs_per_contig)
    assert len(table) == exp_n, f"ORF count {len(table)} != {exp_n}"
    i = 0
    for c, x in enumerate(oracle_orfs_per_contig):
        for j in range(len(x.seqs)):
            what = f"contig {c} orf {j}"
            assert int(table.contig[i]) == c, what
            assert table.sequence(i) == x.seqs[j], what
            assert (int(table.start[i]), int(table.end[i]), int(table.plus[i])) == \
                   (int(x.start[j]), int(x.end[j]), int(x.plus[j])), what
            assert table.starts_alternative(i) == list(x.alts[j]), what
            i += 1


def assert_same_rows(gpu, ora, what=""):
    """Nucleotide search rows (surviving ORFs): hits, SizeInKmer, Location after
    SetBestStartCodon, trimmed Query.Sequence and PositionHits — all bit-exact."""
    assert_same_hits(gpu, ora, what)
    np.testing.assert_array_equal(gpu.row_contig, ora.row_contig, err_msg=f"{what}: contig")
    np.testing.assert_array_equal(gpu.row_start, ora.row_start, err_msg=f"{what}: StartPosition")
    np.testing.assert_array_equal(gpu.row_end, ora.row_end, err_msg=f"{what}: EndPosition")
    np.testing.assert_array_equal(gpu.row_plus, ora.row_plus, err_msg=f"{what}: PlusStrand")
    np.testing.assert_array_equal(gpu.row_seq_off.astype(np.int64), ora.row_seq_off.astype(np.int64), err_msg=f"{what}: seq_off")
    np.testing.assert_array_equal(gpu.row_seq, ora.row_seq, err_msg=f"{what}: Query.Sequence")
    np.testing.assert_array_equal(gpu.pos_off.astype(np.int64), ora.pos_off.astype(np.int64), err_msg=f"{what}: pos_off")
    np.testing.assert_array_equal(gpu.pos, ora.pos, err_msg=f"{what}: PositionHits")
