"""Where the from-the-file milliseconds go (AlignmentFile -> coverage_engine -> region_stats), per decode mode."""
import sys, time, json, os, tempfile
import numpy as np, torch
sys.path.insert(0, ".")
from metacov_b200 import AlignmentFile, synth
wb = synth.c2(0.2)
hb, isz = synth.generate_host(wb)
tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "x.bam")
synth.write_bam(path, wb, hb, isz)
rt = np.arange(wb.n_contigs, dtype=np.int32)
out = {"bam_bytes": os.path.getsize(path)}
for mode in ("host", "gpu", "gpu-stream"):
    best = None
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        af = AlignmentFile(path, decode=mode)
        t1 = time.perf_counter()
        eng = af.coverage_engine()
        eng.sync()
        t2 = time.perf_counter()
        st = eng.region_stats(rt, np.zeros_like(rt), wb.contig_len)
        t3 = time.perf_counter()
        af.close()
        t4 = time.perf_counter()
        r = {"open": t1 - t0, "coverage_engine": t2 - t1, "region_stats": t3 - t2, "close": t4 - t3, "total": t4 - t0}
        if best is None or r["total"] < best["total"]: best = r
    out[mode] = {k: round(v * 1e3, 2) for k, v in best.items()}
print(json.dumps(out))
