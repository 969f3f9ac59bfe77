"""One-off randomized sweep (B200): htslib's max_depth cap replayed on the device.  Random sorted read sets whose depth
exceeds a LOW cap (max_depth 50 - 400, so that the cap fires all over small inputs): piles at one position, ramps, mixed
read lengths, several contigs -- the fused pass + k_cap_replay must equal the sequential htslib machine (oracle mode "plp")
slot by slot, and the whole-contig statistics computed from it."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
from metacov_b200 import CoverageEngine, ReadBatch
from oracle import cport
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 31)
KEYS = ("sum", "sumsq", "iq_sum", "min", "max", "med_lo", "med_hi")
bad = fired = 0
for t in range(trials):
    nc = int(rng.integers(1, 5))
    lengths = rng.integers(800, 6000, nc).astype(np.int32)
    cap = int(rng.choice([50, 120, 400]))
    tid, pos = [], []
    for c in range(nc):
        L = int(lengths[c])
        parts = [rng.integers(0, L, int(rng.integers(0, 4 * cap)))]
        for _ in range(int(rng.integers(0, 4))):                   # piles
            parts.append(np.full(int(rng.integers(cap // 2, 3 * cap)), int(rng.integers(0, L))))
        if rng.random() < 0.4:                                      # a ramp
            a = int(rng.integers(0, L // 2))
            parts.append(np.repeat(np.arange(a, min(a + 300, L)), int(rng.integers(1, 4))))
        ps = np.sort(np.concatenate(parts))
        tid += [c] * len(ps); pos += ps.tolist()
    n = len(tid)
    if n == 0:
        continue
    rl = rng.integers(1, int(rng.choice([40, 150, 1500])), n).astype(np.uint32)
    b = ReadBatch(np.array(tid, np.int32), np.array(pos, np.int32), np.zeros(n, np.uint16), np.full(n, 30, np.uint8),
                  np.arange(n + 1, dtype=np.uint32), rl << 4)
    want, off, info = cport.depth(b, lengths, filt=cport.default_filter(max_depth=cap), mode="plp")
    fired += 1 if info["dropped_by_cap"] > 0 else 0
    rt = np.arange(nc, dtype=np.int32)
    ref = cport.region_stats(want, off, lengths, rt, np.zeros_like(rt), lengths)
    with CoverageEngine(lengths, filt={"max_depth": cap}) as eng:
        eng.depth_sorted(b)
        ok = all(np.array_equal(eng.copy_depth(c), want[off[c]:off[c] + lengths[c]]) for c in range(nc))
        st = eng.region_stats(rt, np.zeros_like(rt), lengths)
        ok = ok and all(np.array_equal(st[k], ref[k]) for k in KEYS)
        # the asynchronous form: the replay must precede the consumer
        eng.depth_sorted(b, wait=False)
        st2 = eng.region_stats(rt, np.zeros_like(rt), lengths)
        ok = ok and all(np.array_equal(st2[k], ref[k]) for k in KEYS)
    bad += 0 if ok else 1
    if not ok: print("MISMATCH trial", t, "cap", cap, "n", n, flush=True)
print(json.dumps({"trials": trials, "cap_fired_in": fired, "mismatches": bad}))
