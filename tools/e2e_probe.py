"""Where the e2e step goes: the transport block's H2D copy alone, the depth pass from the block without statistics, the
pipelined e2e loop, the time the host spends inside each call.  Run under gpurun."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from metacov_b200 import CoverageEngine, ReadBatch, synth
from metacov_b200.engine import pack_block
w = synth.c2(1.0)
db, _ = synth.generate_device(w, 0)
ws = torch.cuda.Stream()
eng = CoverageEngine(w.contig_len, device=0, stream=ws.cuda_stream)
g = w.n_contigs
tid = np.arange(g, dtype=np.int32); st = np.zeros(g, np.int32); en = w.contig_len.astype(np.int32)
hb = ReadBatch(*[t.cpu() for t in db])
blk = pack_block(hb, g, with_mapq=False, pinned=True)
out = {"block_bytes": int(blk[1])}
d = torch.empty(int(blk[1]), dtype=torch.uint8, device="cuda")
def wall(fn, n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
out["h2d_only_ms"] = wall(lambda: d.copy_(blk[0][:int(blk[1])], non_blocking=True))
out["depth_block_only_ms"] = wall(lambda: eng.depth_sorted_block(blk, wait=False))
def piped(k=50):
    prev = None
    for i in range(k):
        eng.depth_sorted_block(blk, wait=False)
        t = eng.region_stats_submit(tid, st, en, slot=i & 1)
        if prev is not None: eng.region_stats_collect(prev, copy=False)
        prev = t
    eng.region_stats_collect(prev, copy=False)
piped(5)
torch.cuda.synchronize(); t0 = time.perf_counter(); piped(50); out["piped_ms"] = (time.perf_counter() - t0) / 50 * 1e3
# host time inside the calls
tt = [0.0, 0.0, 0.0]
prev = None
for i in range(50):
    a = time.perf_counter(); eng.depth_sorted_block(blk, wait=False)
    b = time.perf_counter(); t = eng.region_stats_submit(tid, st, en, slot=i & 1)
    c = time.perf_counter()
    if prev is not None: eng.region_stats_collect(prev, copy=False)
    e = time.perf_counter(); prev = t
    tt[0] += b - a; tt[1] += c - b; tt[2] += e - c
eng.region_stats_collect(prev, copy=False)
out["host_ms_in_calls"] = {"depth_block": tt[0] / 50 * 1e3, "submit": tt[1] / 50 * 1e3, "collect": tt[2] / 50 * 1e3}
# device-side: kernels of a block pass
eng.profile(True)
for _ in range(10):
    eng.depth_sorted_block(blk, wait=False); eng.region_stats(tid, st, en)
kt = eng.profile_read(); eng.profile(False)
out["kernels_us"] = {k: round(v[1] / 10 * 1e3, 1) for k, v in kt.items()}
print(json.dumps(out))
