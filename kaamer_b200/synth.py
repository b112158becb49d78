"""Seeded synthetic workloads (SURVEY.md §8d / BASELINE.md §4).

Protein DB: random sequences with UniProtKB/Swiss-Prot amino-acid composition, lengths
round(LogNormal(median 290, sigma 0.6)) clipped to [30, 5000]; 25 % of the records are
family members = copies of an earlier founder with 5-30 % point substitutions.
Protein queries: DB records with 10 % i.i.d. substitutions.  Nucleotide contigs: sampled DB
proteins back-translated with a random table-11 codon per residue + stop, random strand,
separated by 50-300 nt of uniform random DNA.  Generator: numpy Philox, seed 20261018 + config.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20261018

# UniProtKB/Swiss-Prot composition (%), 20 standard letters
_AA = "ARNDCQEGHILKMFPSTWYV"
_FREQ = np.array([8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96, 9.66, 5.84, 2.42, 3.86, 4.70,
                  6.56, 5.34, 1.08, 2.92, 6.87])
AA_LETTERS = np.frombuffer(_AA.encode(), dtype=np.uint8)
AA_PROBS = _FREQ / _FREQ.sum()


def rng_for(config_index: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=BASE_SEED + config_index, counter=[0, 0, 0, stream]))


def _random_residues(rng: np.random.Generator, n: int) -> np.ndarray:
    cdf = np.cumsum(AA_PROBS)
    cdf[-1] = 1.0
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 24
    for b in range(0, n, step):
        m = min(step, n - b)
        u = rng.random(m, dtype=np.float32)
        out[b:b + m] = AA_LETTERS[np.searchsorted(cdf, u, side="right").clip(0, 19)]
    return out


def _substitute(rng: np.random.Generator, res: np.ndarray, rate) -> None:
    """in-place i.i.d. point substitutions; `rate` scalar or per-residue array."""
    n = len(res)
    step = 1 << 24
    for b in range(0, n, step):
        m = min(step, n - b)
        r = rate if np.isscalar(rate) else rate[b:b + m]
        mask = rng.random(m, dtype=np.float32) < r
        k = int(mask.sum())
        if k:
            seg = res[b:b + m]
            seg[mask] = _random_residues(rng, k)


def protein_db(n_proteins: int, config_index: int = 1, family_frac: float = 0.25):
    """-> (residues u8, seq_off u64[n+1]); records are upper-case, 20 standard letters."""
    rng = rng_for(config_index, 0)
    lens = np.rint(np.exp(rng.normal(np.log(290.0), 0.6, n_proteins))).clip(30, 5000).astype(np.int64)
    member = rng.random(n_proteins) < family_frac
    member[0] = False
    founders = np.flatnonzero(~member)
    # each member copies a uniformly random founder that precedes it
    midx = np.flatnonzero(member)
    n_prev = np.searchsorted(founders, midx, side="left")  # founders before each member (>=1)
    pick = (rng.random(len(midx)) * n_prev).astype(np.int64)
    src = founders[pick]
    lens[midx] = lens[src]
    off = np.zeros(n_proteins + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    res = np.empty(total, dtype=np.uint8)
    # founders: fresh residues
    f_lens = lens[founders]
    f_res = _random_residues(rng, int(f_lens.sum()))
    f_off = np.zeros(len(founders) + 1, dtype=np.int64)
    f_off[1:] = np.cumsum(f_lens)
    # scatter founders to their places
    dst_start = off[:-1].astype(np.int64)
    pos_in_f = np.arange(len(f_res), dtype=np.int64) - np.repeat(f_off[:-1], f_lens)
    res[np.repeat(dst_start[founders], f_lens) + pos_in_f] = f_res
    # members: copy + 5-30 % substitutions
    if len(midx):
        m_lens = lens[midx]
        m_total = int(m_lens.sum())
        m_off = np.zeros(len(midx) + 1, dtype=np.int64)
        m_off[1:] = np.cumsum(m_lens)
        pos_in_m = np.arange(m_total, dtype=np.int64) - np.repeat(m_off[:-1], m_lens)
        founder_rank = np.searchsorted(founders, src)
        m_res = f_res[np.repeat(f_off[founder_rank], m_lens) + pos_in_m]
        rate = np.repeat(rng.uniform(0.05, 0.30, len(midx)).astype(np.float32), m_lens)
        _substitute(rng, m_res, rate)
        res[np.repeat(dst_start[midx], m_lens) + pos_in_m] = m_res
    return res, off


def protein_queries(db_res: np.ndarray, db_off: np.ndarray, n_queries: int, config_index: int = 1,
                    sub_rate: float = 0.10, stream: int = 1):
    """queries = uniformly sampled DB records with `sub_rate` i.i.d. substitutions."""
    rng = rng_for(config_index, stream)
    n = len(db_off) - 1
    pick = rng.integers(0, n, n_queries)
    lens = (db_off[pick + 1] - db_off[pick]).astype(np.int64)
    off = np.zeros(n_queries + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    pos = np.arange(total, dtype=np.int64) - np.repeat(off[:-1].astype(np.int64), lens)
    res = db_res[np.repeat(db_off[pick].astype(np.int64), lens) + pos].copy()
    _substitute(rng, res, sub_rate)
    return res, off, pick


# table 11 codons per amino acid (gcode.go:36-101)
_CODONS: dict[str, list[str]] = {}
_BASES = "TCAG"
_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
for _i, _a in enumerate(_AAS):
    _CODONS.setdefault(_a, []).append(_BASES[_i // 16] + _BASES[(_i // 4) % 4] + _BASES[_i % 4])
_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def nucleotide_contigs(db_res: np.ndarray, db_off: np.ndarray, n_contigs: int, contig_len: int,
                       config_index: int = 2, upper: bool = True):
    """-> (nt u8, contig_off u64[n+1]).  ~88 % coding density."""
    rng = rng_for(config_index, 2)
    n = len(db_off) - 1
    # codon lookup: for each aa letter up to 6 codons, padded by repetition
    lut = np.zeros((256, 6, 3), dtype=np.uint8)
    ncod = np.ones(256, dtype=np.int64)
    for a, cs in _CODONS.items():
        ncod[ord(a)] = len(cs)
        for j in range(6):
            lut[ord(a), j] = np.frombuffer(cs[j % len(cs)].encode(), dtype=np.uint8)
    stops = np.array([np.frombuffer(c.encode(), dtype=np.uint8) for c in _CODONS["*"]])
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    chunks_all, offs = [], [0]
    for _ in range(n_contigs):
        parts, cur = [], 0
        while cur < contig_len:
            gap = int(rng.integers(50, 301))
            parts.append(acgt[rng.integers(0, 4, gap)])
            cur += gap
            p = int(rng.integers(0, n))
            aa = db_res[int(db_off[p]):int(db_off[p + 1])]
            pickc = (rng.random(len(aa)) * ncod[aa]).astype(np.int64)
            gene = np.concatenate([lut[aa, pickc].reshape(-1), stops[int(rng.integers(0, 3))]])
            if rng.random() < 0.5:
                gene = np.frombuffer(gene.tobytes().translate(_COMP)[::-1], dtype=np.uint8)
            parts.append(gene)
            cur += len(gene)
        contig = np.concatenate(parts)[:contig_len]
        chunks_all.append(contig)
        offs.append(offs[-1] + len(contig))
    nt = np.concatenate(chunks_all) if chunks_all else np.zeros(0, np.uint8)
    if not upper:
        nt = np.frombuffer(nt.tobytes().lower(), dtype=np.uint8)
    return nt.copy(), np.array(offs, dtype=np.uint64)


def write_fasta(path: str, names, residues: np.ndarray, off: np.ndarray, width: int = 60):
    with open(path, "wb") as f:
        for i, name in enumerate(names):
            f.write(b">" + name.encode() + b"\n")
            s = residues[int(off[i]):int(off[i + 1])].tobytes()
            for b in range(0, len(s), width):
                f.write(s[b:b + width] + b"\n")
