import sys, time, collections
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..', '..'))
import numpy as np, torch
from kaamer_b200 import GpuIndex, SearchOptions, synth
from kaamer_b200.makedb import fasta_protein_ids
from kaamer_b200.sharded import CudaShardBackend, ShardedSearch, fences_from_sample, simulate_lockstep
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
res, off = synth.protein_db(570_000, config_index=3)
ids = fasta_protein_ids(len(off) - 1)
fences = fences_from_sample(res, off, G)
T = collections.defaultdict(float)
def timed(name, f):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); T[name] += time.perf_counter() - t
        return r
    return w
class FakeComm: world = G; rank = 0
searchers, inputs, handles = [], [], []
for r in range(G):
    g = GpuIndex.build(res, off, ids, keep_proteins=False, shard=(int(fences[r]), int(fences[r + 1])))
    handles.append(g)
    be = CudaShardBackend(g)
    for n in ("route_count", "route_fill", "shard_count", "merge"):
        setattr(be, n, timed(n, getattr(be, n)))
    searchers.append(ShardedSearch(be, fences, FakeComm()))
    q, qo, _ = synth.protein_queries(res, off, NQ, config_index=3, stream=100 + r)
    inputs.append((torch.from_numpy(q).cuda(), torch.from_numpy(qo.astype(np.int64)).cuda(), NQ))
opts = SearchOptions()
simulate_lockstep(searchers, inputs, opts)
T.clear()
torch.cuda.synchronize(); t0 = time.perf_counter()
R = 3
for _ in range(R):
    out = simulate_lockstep(searchers, inputs, opts)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / R
print("G", G, "NQ/rank", NQ, "total ms per lockstep round (all ranks serial)", tot * 1e3)
for k, v in T.items(): print(f"  {k:12s} {v / R * 1e3:8.2f} ms (sum over {G} ranks)")
print("  other (torch cumsum/cat/host)", (tot - sum(T.values()) / R) * 1e3)
print("parts per rank", [int(o.pool.numel()) for o in out][:2], "lookups", sum(o.n_lookups for o in out))
