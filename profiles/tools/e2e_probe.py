import ctypes as C, sys, time
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..', '..'))
import numpy as np, torch
from kaamer_b200 import GpuIndex, SearchOptions, synth, _lib
from kaamer_b200.makedb import fasta_protein_ids
res, off = synth.protein_db(570_000, config_index=3)
ids = fasta_protein_ids(len(off) - 1)
g = GpuIndex.build(res, off, ids, keep_proteins=False)
q, qo, _ = synth.protein_queries(res, off, 100_000, config_index=3, stream=100)
hq = torch.from_numpy(q).pin_memory(); ho = torch.from_numpy(qo.astype(np.int64)).pin_memory()
opts = SearchOptions(); nq = len(qo) - 1
L = _lib.lib()
def full():
    return g.search_proteins_ptr(hq.data_ptr(), ho.data_ptr(), nq, opts)
def raw():
    o = opts.c(); hp = C.POINTER(_lib.Hits)()
    L.kaamer_gpu_search_proteins(g._h, C.c_void_p(hq.data_ptr()), C.c_void_p(ho.data_ptr()), nq, C.byref(o), C.byref(hp))
    L.kaamer_gpu_free_hits(hp)
for name, f in (("full", full), ("raw", raw)):
    for _ in range(5): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(50): f()
    torch.cuda.synchronize(); print(name, "ms/call", (time.perf_counter() - t) / 50 * 1e3)
g.profile_enable(True)
for _ in range(5): raw()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(50): raw()
print("raw+profile ms/call", (time.perf_counter() - t) / 50 * 1e3, g.profile_read()["kernel_ms"][:6])
