"""One-off randomized sweeps (B200): (1) the host streaming reader -> transport blocks -> streamed pass with random batch sizes
and small file reads against the one-shot depth; (2) region statistics of thousands of random regions (sizes log-uniform from 1
slot to whole contigs, overlapping, reaching past contig ends) against the C oracle."""
import sys, os, tempfile, json
import numpy as np
sys.path.insert(0, ".")
from metacov_b200 import AlignmentFile, CoverageEngine, synth
from oracle import cport
rng = np.random.default_rng(2024)
out = {}
# ---- (1) ----
w = synth.c2(0.02); hb, isz = synth.generate_host(w)
tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "f.bam")
synth.write_bam(path, w, hb, isz)
lengths = [int(x) for x in w.contig_len]
d, off, info = cport.depth(hb, lengths, mode="diff")
bad = 0; batches = []
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    br = int(rng.integers(500, 60000))
    os.environ["MCOV_STREAM_READ_BYTES"] = str(int(rng.integers(4096, 3_000_000)))
    with AlignmentFile(path, batch_reads=br) as af:
        eng = af.coverage_engine()
        ok = eng.pass_info()["n_pass"] == info["n_pass"]
        for c in range(len(lengths)):
            ok = ok and np.array_equal(eng.copy_depth(c), d[off[c]:off[c] + lengths[c]])
        batches.append(int(af.stream_batches))
    bad += 0 if ok else 1
    if not ok: print("STREAM MISMATCH", br, os.environ["MCOV_STREAM_READ_BYTES"], flush=True)
os.environ.pop("MCOV_STREAM_READ_BYTES", None)
out["host_stream"] = {"trials": len(batches), "mismatches": bad, "batches_min_max": [min(batches), max(batches)]}
# ---- (2) ----
KEYS = ("sum", "sumsq", "iq_sum", "min", "max", "med_lo", "med_hi")
bad = 0; n_regions = 0
with CoverageEngine(lengths) as eng:
    eng.compute_depth(hb)
    for t in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
        g = int(rng.integers(200, 4000))
        rt = rng.integers(0, len(lengths), g).astype(np.int32)
        ln = np.exp(rng.uniform(0, np.log(60000), g)).astype(np.int64)
        rs = np.array([rng.integers(0, lengths[c] + 1) for c in rt], np.int64)
        re = np.minimum(rs + np.maximum(ln, 1), np.array([lengths[c] for c in rt]) + rng.integers(0, 50, g))
        re = np.maximum(re, rs + 1)
        st = eng.region_stats(rt, rs.astype(np.int32), re.astype(np.int32))
        want = cport.region_stats(d, off, lengths, rt, rs.astype(np.int32), re.astype(np.int32))
        okk = all(np.array_equal(st[k], want[k]) for k in KEYS)
        bad += 0 if okk else 1
        n_regions += g
        if not okk: print("STATS MISMATCH trial", t, flush=True)
out["region_stats"] = {"trials": t + 1, "regions": n_regions, "mismatches": bad}
print(json.dumps(out))
