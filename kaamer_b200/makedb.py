"""Host-side front end of the device-index build: the record-level semantics of
`kaamer-db -make -f fasta` (pkg/makedb/inputFASTA.go) that decide WHICH (sequence, id) pairs
reach the k-mer index.  The k-mer work itself (windows, sort, unique, CSR) runs on the GPU
(`GpuIndex.build`)."""
from __future__ import annotations

import gzip

import numpy as np

KMER_SIZE = 7  # pkg/makedb/makedb.go:29-31


def fasta_protein_ids(n_records: int) -> np.ndarray:
    """Protein ids as runFASTA assigns them (inputFASTA.go:96-124): `proteinNb` is incremented
    on every '>' BEFORE the previous record is dispatched, so record j (1-based, j < N) gets id
    j+1 and the last record gets id N — records N-1 and N share id N; id 1 is unused."""
    if n_records == 0:
        return np.zeros(0, np.uint32)
    ids = np.arange(2, n_records + 2, dtype=np.uint32)
    ids[-1] = n_records
    return ids


def read_fasta(path: str):
    """-> (entry_ids, names, residues u8, seq_off u64, ids u32) of the ACCEPTED records.

    processProteinInputFASTA (inputFASTA.go:195-250): EntryId = first header token without
    '>', ProteinName = rest of the header; sequence lines upper-cased and concatenated;
    records whose name contains ", partial" or shorter than 7 residues are skipped (they keep
    their id number: ids are assigned before the skip)."""
    opener = gzip.open if path.endswith(".gz") else open
    headers, seqs = [], []
    with opener(path, "rb") as f:
        cur = None
        for raw in f:
            line = raw.rstrip(b"\n")
            if line.endswith(b"\r"):
                line = line[:-1]  # bufio.ScanLines drops one trailing CR (inputFASTA.go:86-96)
            if line[:1] == b">":
                headers.append(line)
                cur = []
                seqs.append(cur)
            elif cur is not None:
                cur.append(line.upper())
    ids_all = fasta_protein_ids(len(headers))
    entry_ids, names, keep_seqs, keep_ids = [], [], [], []
    for h, parts, pid in zip(headers, seqs, ids_all):
        toks = h.split(b" ")
        name = b" ".join(toks[1:])
        seq = b"".join(parts)
        if b", partial" in name or len(seq) < KMER_SIZE:
            continue
        entry_ids.append(toks[0][1:].decode())
        names.append(name.decode())
        keep_seqs.append(seq)
        keep_ids.append(pid)
    off = np.zeros(len(keep_seqs) + 1, dtype=np.uint64)
    if keep_seqs:
        off[1:] = np.cumsum([len(s) for s in keep_seqs], dtype=np.uint64)
    res = np.frombuffer(b"".join(keep_seqs), dtype=np.uint8).copy() if keep_seqs else np.zeros(0, np.uint8)
    return entry_ids, names, res, off, np.array(keep_ids, dtype=np.uint32)


def read_tsv(path: str):
    """-> (entry_ids, residues u8, seq_off u64, ids u32) of the ACCEPTED rows of a kaamer TSV database
    (`kaamer-db -make -f tsv`, pkg/makedb/inputTSV.go:94-142): the first line names the columns
    (`EntryID` and `Sequence`, case-insensitive, are mandatory); a row is skipped when its sequence is
    shorter than 7 residues or its entry id is empty; accepted rows get the ids 0, 1, 2 ... in file order
    (no id quirk here) and their sequences are NOT upper-cased (unlike the FASTA path)."""
    opener = gzip.open if path.endswith(".gz") else open
    entry_ids, seqs = [], []
    with opener(path, "rb") as f:
        header = f.readline().rstrip(b"\n").rstrip(b"\r").split(b"\t")
        low = [h.lower() for h in header]
        if b"entryid" not in low:
            raise ValueError("TSV file doesn't contain 'EntryID' header")
        if b"sequence" not in low:
            raise ValueError("TSV file doesn't contain 'Sequence' header")
        for raw in f:
            line = raw.rstrip(b"\n")
            if line.endswith(b"\r"):
                line = line[:-1]  # bufio.ScanLines drops the CR
            cols = line.split(b"\t")
            if len(cols) > len(header):
                raise ValueError("TSV row has more columns than the header (the reference panics)")
            entry, seq = b"", b""
            for i, c in enumerate(cols):
                if low[i] == b"entryid":
                    entry = c
                elif low[i] == b"sequence":
                    seq = c
            if len(seq) < KMER_SIZE or seq == b"" or entry == b"":
                continue
            entry_ids.append(entry.decode())
            seqs.append(seq)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    res = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return entry_ids, res, off, np.arange(len(seqs), dtype=np.uint32)
