"""One-off randomized sweep (B200): mcov_kmer_hist (ByFlag-grouped k-mer histogram) with random K / NK / STEP / OFFSET / group
flags on random reads (lengths around the window, ambiguity codes, both strands) against oracle/scanstats.py."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
from metacov_b200 import CoverageEngine
from oracle import scanstats
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 8)
bad = 0
with CoverageEngine([1000]) as eng:
    for t in range(trials):
        K = int(rng.integers(1, 9)); NK = int(rng.integers(1, 7)); STEP = int(rng.integers(1, 9)); OFFSET = int(rng.integers(0, 6))
        win_bases = OFFSET + (NK - 1) * STEP + K
        n = int(rng.integers(1, 1500))
        need = OFFSET + STEP * NK
        l_seq = rng.integers(max(0, need - 5), need + 40, n).astype(np.int32)
        flag = rng.choice(np.array([0, 16, 64, 80, 128, 144, 4, 1024], np.uint16), n)
        gf = tuple(int(x) for x in rng.choice(np.array([0x10, 0x40, 0x80, 0x4, 0x400]), int(rng.integers(0, 3)), replace=False))
        seqs, win = [], np.zeros((n, (win_bases + 1) // 2), np.uint8)
        for i in range(n):
            L = int(l_seq[i])
            s = np.array([1, 2, 4, 8, 15, 3], np.uint8)[rng.choice(6, L, p=[.24, .24, .24, .24, .03, .01])]
            # what the reader hands over: the first (forward) / last (reverse) win_bases bases, padded with 15
            w = np.full(win_bases + (win_bases & 1), 15, np.uint8)
            rev = bool(flag[i] & 16)
            for j in range(win_bases):
                a = L - win_bases + j if rev else j
                if 0 <= a < L: w[j] = s[a]
            win[i] = (w[0::2] << 4) | w[1::2]
            seqs.append(s)
        got = eng.kmer_hist(flag, l_seq, win, win_bases, K, NK, STEP, OFFSET, gf)
        want = scanstats.kmer_hist(flag, seqs, K, NK, STEP, OFFSET, gf)             # (full stored-orientation reads)
        ok = np.array_equal(got, want)
        bad += 0 if ok else 1
        if not ok: print("MISMATCH", t, K, NK, STEP, OFFSET, gf, flush=True)
print(json.dumps({"trials": trials, "mismatches": bad}))
