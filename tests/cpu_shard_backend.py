"""TEST INFRASTRUCTURE: a numpy/oracle stand-in for the four device steps of the sharded search
(kaamer_b200.sharded.CudaShardBackend), so that the multi-rank HOST logic (routing plan, split
sizes, offsets of the two all-to-alls, merge bookkeeping) can run on CPU tensors under gloo
with world_size 2.  Never imported by the product package."""
import numpy as np
import torch

from kaamer_b200.sharded import dense_from_keys
from oracle import oracle as o

_AA = b"ACDEFGHIKLMNPQRSTUVWY"  # pkg/kvstore/k_store.go:41


def _dense_codes(seq: bytes, K: int) -> np.ndarray:
    if K <= 0:
        return np.zeros(0, np.int64)
    keys = np.array([o.encode_kmer(seq[k:k + 7]) for k in range(K)], dtype=np.uint32)
    return dense_from_keys(keys).astype(np.int64)


class CpuShardBackend:
    def __init__(self, keys, offsets, postings, lo, hi):
        self.d = dense_from_keys(np.asarray(keys)).astype(np.int64)
        self.offsets = np.asarray(offsets).astype(np.int64)
        self.postings = np.asarray(postings)
        self.lo, self.hi = int(lo), int(hi)

    def _queries(self, d_res, d_off, nq):
        res = d_res.numpy().tobytes()
        off = d_off.numpy().astype(np.int64)
        return [res[off[i]:off[i + 1]] for i in range(nq)]

    def route_count(self, d_res, d_off, nq, fences, G):
        counts = np.zeros((G, nq), np.int32)
        size = np.zeros(nq, np.int32)
        self._codes = []
        for q, s in enumerate(self._queries(d_res, d_off, nq)):
            K = o.size_in_kmer(s)
            size[q] = K
            c = _dense_codes(s, K if K >= 7 else 0)
            sh = np.searchsorted(np.asarray(fences[1:-1]).astype(np.int64), c, side="right")
            self._codes.append((c, sh))
            for t in range(G):
                counts[t, q] = int((sh == t).sum())
        return torch.from_numpy(counts.reshape(-1)), torch.from_numpy(size)

    def route_fill(self, d_res, d_off, nq, fences, G, counts, offsets, total):
        out = np.zeros(int(total), np.int32)
        off = offsets.numpy()
        for t in range(G):
            for q, (c, sh) in enumerate(self._codes):
                sel = c[sh == t]
                b = int(off[t * nq + q])
                out[b:b + len(sel)] = sel
        return torch.from_numpy(out)

    def shard_count(self, codes, seg_off, nseg):
        codes = codes.numpy().astype(np.int64)
        so = seg_off.numpy()
        part_n = np.zeros(nseg, np.int32)
        parts = []
        lookups = incr = 0
        for g in range(nseg):
            cnt = {}
            for c in codes[so[g]:so[g + 1]]:
                lookups += 1
                assert self.lo <= c < self.hi, "code routed to the wrong shard"
                i = int(np.searchsorted(self.d, c))
                if i < len(self.d) and self.d[i] == c:
                    for p in self.postings[self.offsets[i]:self.offsets[i + 1]]:
                        cnt[int(p)] = cnt.get(int(p), 0) + 1
                        incr += 1
            part_n[g] = len(cnt)
            parts += [s | (n << 32) for s, n in cnt.items()]
        return torch.from_numpy(part_n), torch.tensor(parts, dtype=torch.int64), lookups, incr

    def merge(self, parts, part_off, G, nq, size_in_kmer, opts):
        parts = parts.numpy()
        po = part_off.numpy().reshape(G, nq + 1)
        size = size_in_kmer.numpy()
        n_hits = np.zeros(nq, np.int32)
        hit_base = np.zeros(nq, np.int32)
        pool = []
        ko = o.opts(opts.min_kmatch, opts.min_kratio, opts.max_results)
        for q in range(nq):
            cnt = {}
            for s in range(G):
                for v in parts[po[s, q]:po[s, q + 1]]:
                    cnt[int(v) & 0xFFFFFFFF] = cnt.get(int(v) & 0xFFFFFFFF, 0) + (int(v) >> 32)
            if size[q] < 7 or not cnt:
                continue
            hits = sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0]))
            keep = o.filter_count([h[1] for h in hits], int(size[q]), ko)
            hit_base[q] = len(pool)
            n_hits[q] = keep
            pool += [h[0] | (h[1] << 32) for h in hits[:keep]]
        return torch.from_numpy(n_hits), torch.from_numpy(hit_base), torch.tensor(pool + [0], dtype=torch.int64)
