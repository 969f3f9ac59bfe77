"""One-off randomized sweep (B200): the streamed pass (mcov_stream_push) with random numbers of batches cut at random
places -- including cuts a few reads apart -- on C2 / C3 / C5-like read sets, against the C oracle's depth."""
import sys, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_stream import push_in_batches
from metacov_b200 import CoverageEngine, synth
from oracle import cport
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
res = {}
for wl, scale in (("c2", 0.01), ("c3", 0.001), ("c5", 0.002)):
    w = synth.WORKLOADS[wl](scale)
    b, _, reflen = synth.generate_host(w, want_reflen=True)
    n = len(b.tid)
    want, off, info = cport.depth(b, w.contig_len, mode="diff")
    bad = 0
    with CoverageEngine(w.contig_len) as eng:
        for t in range(trials):
            nb = int(rng.integers(2, 40))
            cuts = np.sort(rng.choice(np.arange(1, n), min(nb - 1, n - 1), replace=False))
            if t % 3 == 0 and len(cuts) > 3:                    # some cuts a few reads apart
                cuts[1] = cuts[0] + 1; cuts[2] = cuts[0] + 3
                cuts = np.unique(np.clip(cuts, 1, n - 1))
            cuts = np.r_[0, cuts, n]
            push_in_batches(eng, b, reflen, cuts)
            pi = eng.pass_info()
            ok = pi["n_pass"] == info["n_pass"] and pi["aligned_bases"] == info["aligned_bases"]
            for c in range(w.n_contigs):
                ok = ok and np.array_equal(eng.copy_depth(c), want[off[c]:off[c] + w.contig_len[c]])
            bad += 0 if ok else 1
            if not ok: print("MISMATCH", wl, t, cuts[:8], flush=True)
    res[wl] = {"reads": int(n), "trials": trials, "mismatches": bad}
print(json.dumps(res))
