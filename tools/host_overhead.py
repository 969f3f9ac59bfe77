import time, numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from metacov_b200 import CoverageEngine, synth
w = synth.c2(1.0)
db, _ = synth.generate_device(w, 0)
eng = CoverageEngine(w.contig_len, device=0, stream=torch.cuda.current_stream().cuda_stream)
g = w.n_contigs
tid = np.arange(g, dtype=np.int32); st = np.zeros(g, np.int32); en = w.contig_len.astype(np.int32)
for _ in range(5):
    eng.depth_sorted(db, wait=False); eng.region_stats(tid, st, en)
def t(fn, n=200):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("depth_sorted async only (us/call, GPU-bound when queue fills):", t(lambda: eng.depth_sorted(db, wait=False)))
print("region_stats only:", t(lambda: eng.region_stats(tid, st, en)))
print("full step:", t(lambda: (eng.depth_sorted(db, wait=False), eng.region_stats(tid, st, en))))
# python-side cost without GPU wait: time the call itself
t0 = time.perf_counter()
for _ in range(50): eng.depth_sorted(db, wait=False)
t1 = time.perf_counter(); torch.cuda.synchronize()
print("depth_sorted call return time us:", (t1 - t0) / 50 * 1e6)
import ctypes as C
from metacov_b200 import _capi
from metacov_b200.engine import _canon, _mem_kind
t0 = time.perf_counter()
for _ in range(1000): b = _canon(db); k = _mem_kind(b)
print("_canon+_mem_kind us:", (time.perf_counter() - t0) / 1000 * 1e6)
