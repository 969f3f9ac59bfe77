"""One-off: the streamed GPU decoder (mcov_bam_gpu_stream_depth) with random chunk sizes against the whole-file GPU decode
of the same BAM -- every contig's depth must be identical whatever the chunk borders cut."""
import sys, os, tempfile, json
import numpy as np
sys.path.insert(0, ".")
from metacov_b200 import CoverageEngine, bamgpu, synth
n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(77)
w = synth.c2(0.04)                                       # 400 000 reads, 40 contigs
hb, isz = synth.generate_host(w)
tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "f.bam")
synth.write_bam(path, w, hb, isz)
size = os.path.getsize(path)
lengths = [int(x) for x in w.contig_len]
with CoverageEngine(lengths) as eng:
    soa = bamgpu.decode(eng, path)
    bamgpu.depth_sorted(eng, soa)
    want = [eng.copy_depth(c) for c in range(len(lengths))]
    pi0 = eng.pass_info()
bad = 0
chunks = []
for t in range(n_trials):
    chunk = int(rng.integers(1 << 17, max(size // 2, (1 << 17) + 1)))
    with CoverageEngine(lengths) as eng:
        info = bamgpu.stream_depth(eng, path, chunk_bytes=chunk)
        pi = eng.pass_info()
        ok = info["n_records"] == len(hb.tid) and pi["n_pass"] == pi0["n_pass"] and pi["aligned_bases"] == pi0["aligned_bases"]
        for c in range(len(lengths)):
            ok = ok and np.array_equal(eng.copy_depth(c), want[c])
        if not ok:
            det = {"chunk": chunk, "n_chunks": int(info["n_chunks"]), "n_records": int(info["n_records"]), "max_carry": int(info["max_carry"]),
                   "n_pass": [int(pi["n_pass"]), int(pi0["n_pass"])], "aligned": [int(pi["aligned_bases"]), int(pi0["aligned_bases"])], "contigs": []}
            for c in range(len(lengths)):
                g = eng.copy_depth(c)
                d = np.flatnonzero(g != want[c])
                if len(d):
                    det["contigs"].append({"c": c, "n_diff": int(len(d)), "first": int(d[0]), "last": int(d[-1]),
                                           "delta_minmax": [int((g[d].astype(np.int64) - want[c][d]).min()), int((g[d].astype(np.int64) - want[c][d]).max())]})
            print("MISMATCH", json.dumps(det), flush=True)
        bad += 0 if ok else 1
        chunks.append(int(info["n_chunks"]))
print(json.dumps({"file_bytes": size, "reads": int(len(hb.tid)), "trials": n_trials, "mismatches": bad,
                  "chunks_min_max": [min(chunks), max(chunks)]}))
