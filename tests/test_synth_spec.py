"""The synthetic C4 generator's CPU twin (oracle/synth_oracle.cpp, which compiles
include/kaamer_synth_spec.h) against an independent statement of the specification in plain Python:
Philox4x32-10 (checked against the published known-answer vectors), record meta, residues, queries."""
import re
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M32 = 0xFFFFFFFF
SEED = 20261022


def philox(c, k):
    c = list(c)
    k0, k1 = k
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def tables():
    txt = open(os.path.join(ROOT, "include", "kaamer_synth_tables.h")).read()
    thr = [int(x, 16) for x in re.findall(r"0x([0-9A-F]{8})u", txt.split("KAAMER_SYNTH_AA_THR")[1].split("\n")[0])]
    body = txt.split("KAAMER_SYNTH_LEN_Q {")[1].split("}")[0].replace("\\", "")
    lenq = [int(x) for x in body.replace("\n", " ").split(",") if x.strip()]
    letters = re.search(r'KAAMER_SYNTH_LETTERS "([A-Z]+)"', txt).group(1)
    assert len(thr) == 20 and len(lenq) == 1024 and len(letters) == 20
    return thr, lenq, letters


THR, LENQ, LETTERS = tables()


def draw(seed, rec, block, batch, stream):
    return philox([rec & M32, block, ((rec >> 32) | (batch << 16)) & M32, stream], (seed & M32, seed >> 32))


def letter(u):
    return sum(1 for t in THR[:19] if u >= t)


def meta(seed, i):
    d = draw(seed, i, 0, 0, 1)
    if i > 0 and d[0] < 0x40000000:
        j = (d[2] * i) >> 32
        while j > 0 and draw(seed, j, 0, 0, 1)[0] < 0x40000000:
            j -= 1
        return j, LENQ[draw(seed, j, 0, 0, 1)[1] >> 22], 0x0CCCCCCD + ((d[3] * 0x40000000) >> 32)
    return i, LENQ[d[1] >> 22], 0


def record(seed, i):
    f, n, thr = meta(seed, i)
    out = []
    for b in range((n + 3) // 4):
        u = draw(seed, f, b, 0, 2)
        if thr:
            s = draw(seed, i, b, 0, 3)
            r = draw(seed, i, b, 0, 4)
            u = [r[k] if s[k] < thr else u[k] for k in range(4)]
        out += [LETTERS[letter(x)] for x in u]
    return "".join(out[:n]).encode()


def query(seed, n_proteins, j, batch):
    d = draw(seed, j, 0, batch, 5)
    rec = ((((d[0] << 32) | d[1]) * n_proteins) >> 64)
    base = record(seed, rec)
    out = bytearray(base)
    for b in range((len(base) + 3) // 4):
        s = draw(seed, j, b, batch, 6)
        r = draw(seed, j, b, batch, 7)
        for k in range(4):
            if b * 4 + k < len(base) and s[k] < 0x1999999A:
                out[b * 4 + k] = ord(LETTERS[letter(r[k])])
    return bytes(out), rec


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    assert philox([0, 0, 0, 0], (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox([M32] * 4, (M32, M32)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_record_meta_and_residues_match_the_twin():
    from oracle import oracle as o

    members = 0
    for i in list(range(40)) + [1000, 123456, 49_999_999, (1 << 32) + 5]:
        f, n, thr = meta(SEED, i)
        assert o.synth_record_meta(SEED, i) == (f, n, thr), i
        members += thr != 0
        if n <= 600:
            assert o.synth_record(SEED, i) == record(SEED, i), i
    assert members > 3


def test_queries_match_the_twin():
    from oracle import oracle as o

    for j, batch in [(0, 0), (1, 0), (7, 3), (99, 65535)]:
        s, rec = query(SEED, 5000, j, batch)
        s2, rec2 = o.synth_query(SEED, 5000, j, batch)
        assert rec == rec2 and s == s2
        base = o.synth_record(SEED, rec)
        diff = sum(a != b for a, b in zip(s, base))
        assert len(s) == len(base) and diff <= 0.25 * len(s) + 5


def test_composition_and_lengths():
    from oracle import oracle as o

    seqs = [o.synth_record(SEED, i) for i in range(3000)]
    lens = np.array([len(s) for s in seqs])
    assert 30 <= lens.min() and lens.max() <= 5000 and 300 < lens.mean() < 400
    cnt = np.bincount(np.frombuffer(b"".join(seqs), np.uint8), minlength=128)
    total = cnt.sum()
    assert abs(cnt[ord("L")] / total - 0.0966) < 0.004 and abs(cnt[ord("W")] / total - 0.0108) < 0.002
    assert set(np.flatnonzero(cnt)) == {ord(c) for c in LETTERS}


def test_restricted_index_equals_full_index_on_the_sampled_queries():
    """the C4 parity device: a query's hits only depend on the posting lists of its own k-mers"""
    from oracle import oracle as o

    n = 3000
    seqs = [o.synth_record(SEED, i) for i in range(n)]
    res, off = o.pack(seqs)
    full = o.Index.build(res, off, np.arange(n, dtype=np.uint32), 2)
    qs = [o.synth_query(SEED, n, j, 1)[0] for j in range(24)]
    q, qo = o.pack(qs)
    restricted = o.synth_restricted_index(SEED, n, qs, 2)
    assert (restricted.n_proteins, restricted.n_aa, restricted.n_kmers) == (full.n_proteins, full.n_aa, full.n_kmers)
    a = o.search_proteins(full, q, qo, o.opts(), 2)
    b = o.search_proteins(restricted, q, qo, o.opts(), 2)
    for f in ("hit_off", "subject", "kmatch", "size_in_kmer"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
    assert a.n_increments == b.n_increments and len(a.subject) > 24
