#!/usr/bin/env python
"""Measurement of the SURVEY.md 8(f) rows that sit beside the coverage hot path -- the `metacov scan`
accumulators (ByFlag + IsizeHist, KmerHist) and `pileup.experimental` -- on C2-shaped synthetic reads
(10 M x 150 bp, 1 000 contigs).  bench.py stays the headline benchmark (per-base depth + per-contig
statistics); this script prints one JSON line per row:

  {"row": ..., "reads": n, "ms_per_call": t, "reads_per_s": r, "kernel_ms": {...},
   "algorithmic_bytes": b, "gbs": ..., "frac_of_hbm_peak": ..., "cpu_baseline": {...}}

`ms_per_call` is the public call with HOST arrays (H2D and the histogram/record read-back inside);
`kernel_ms` the CUDA-event time of the kernels alone.  The CPU figures time the oracle (C port for the
isize histogram, the numpy/Python restatements for the other two) on a bounded sample.

  python tools/bench_rows.py [--scale 1.0] [--calls 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(fn, calls):
    fn()
    t0 = time.perf_counter()
    for _ in range(calls):
        fn()
    return (time.perf_counter() - t0) / calls * 1e3


def kernel_ms(eng, fn, names):
    eng.profile(True)
    fn()
    kt = eng.profile_read()
    eng.profile(False)
    return {k: kt[k][1] / max(kt[k][0], 1) for k in names if k in kt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--calls", type=int, default=5)
    args = ap.parse_args()
    import torch
    from metacov_b200 import CoverageEngine, synth
    from oracle import cport, scanstats
    from oracle import experimental as oexp
    if not torch.cuda.is_available():
        raise SystemExit("bench_rows.py: no CUDA device")
    w = synth.c2(args.scale)
    batch, isize = synth.generate_host(w)
    n = len(batch.tid)
    pk = peak()
    eng = CoverageEngine(w.contig_len)
    rng = np.random.Generator(np.random.PCG64(5))

    # ---- row: ByFlag + IsizeHist (reference scan.pyx:380-420, 581-620) -------------------------------
    gf = (0x10, 0x40)                                        # -g Readdir,IsRead1: 4 groups
    call = lambda: eng.isize_hist(batch.flag, isize, gf, n_bins=1024)
    ms = timed(call, args.calls)
    km = kernel_ms(eng, call, ("k_isize_hist", "k_group_count"))
    t0 = time.perf_counter()
    cport.isize_hist(batch.flag, isize, gf, 1024)
    cpu_s = time.perf_counter() - t0
    alg = 6 * n
    k = km.get("k_isize_hist", 0.0)
    print(json.dumps({"row": "scan: ByFlag(2 flags) + IsizeHist", "reads": n, "ms_per_call": ms, "reads_per_s": n / ms * 1e3,
                      "kernel_ms": km, "algorithmic_bytes": alg, "gbs": alg / k / 1e6 if k else None,
                      "frac_of_hbm_peak": alg / k / 1e6 / pk if k else None, "h2d_bytes": 6 * n,
                      "cpu_baseline": {"reads_per_s": n / cpu_s, "kind": "port", "cores": 1,
                                       "sample": "all %d reads, C port oracle/coverage.c::orc_isize_hist" % n}}), flush=True)

    # ---- row: KmerHist (reference scan.pyx:491-533), defaults K=7 NK=8 STEP=7 OFFSET=0 -----------------
    K, NK, STEP, OFFSET = 7, 8, 7, 0
    win_bases = OFFSET + STEP * NK
    # nt16 nibbles A C G T (1 2 4 8) as a BAM record stores them, uniform; 0.1 % of the bases are N (15)
    base = np.array([1, 2, 4, 8], dtype=np.uint8)[rng.integers(0, 4, (n, win_bases + (win_bases & 1)), dtype=np.uint8)]
    base[rng.random(base.shape, dtype=np.float32) < 0.001] = 15
    win = ((base[:, 0::2] << 4) | base[:, 1::2]).astype(np.uint8)
    del base
    l_seq = np.full(n, 150, dtype=np.int32)
    call = lambda: eng.kmer_hist(batch.flag, l_seq, win, win_bases, K, NK, STEP, OFFSET, gf)
    ms = timed(call, args.calls)
    km = kernel_ms(eng, call, ("k_kmer_hist",))
    ns = min(n, 20000)
    seqs = []
    for i in range(ns):
        b = win[i]
        seqs.append(np.stack([b >> 4, b & 15], axis=1).reshape(-1)[:win_bases])
    t0 = time.perf_counter()
    scanstats.kmer_hist(batch.flag[:ns], seqs, K, NK, STEP, OFFSET, gf)
    cpu_s = time.perf_counter() - t0
    alg = (2 + 4 + win.shape[1]) * n
    k = km.get("k_kmer_hist", 0.0)
    print(json.dumps({"row": "scan: ByFlag(2 flags) + KmerHist(K=7,NK=8,STEP=7)", "reads": n, "ms_per_call": ms,
                      "reads_per_s": n / ms * 1e3, "kernel_ms": km, "algorithmic_bytes": alg,
                      "gbs": alg / k / 1e6 if k else None, "frac_of_hbm_peak": alg / k / 1e6 / pk if k else None,
                      "h2d_bytes": alg,
                      "cpu_baseline": {"reads_per_s": ns / cpu_s, "kind": "port", "cores": 1,
                                       "sample": "first %d reads, oracle/scanstats.py (reference-structured Python loop)" % ns}}),
          flush=True)

    # ---- row: pileup.experimental (reference pileup.py:38-173), one whole-contig region per contig -------
    eng.depth_sorted(batch)                                   # (the context needs a contig table + state; not timed)
    k_len = 7
    k_cor = oexp.synthetic_kcor(k_len)
    from metacov_b200 import pileup
    kc_val, kc_has = pileup._kcor_tables(k_cor, k_len)
    soa = {"pos": batch.pos, "flag": batch.flag, "cig_off": batch.cig_off, "cig": batch.cig}
    # mates share a name: reads 2j and 2j+1 of a contig form a pair
    name_hash = (np.arange(n, dtype=np.uint64) >> np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    kmer_code = rng.integers(0, 4 ** k_len, n, dtype=np.int32)
    g = w.n_contigs
    r_start = np.zeros(g, np.int32)
    r_end = w.contig_len.astype(np.int32)
    r_lb = w.read_start[:-1].astype(np.int64)
    r_ub = w.read_start[1:].astype(np.int64)
    call = lambda: eng.experimental_stats(soa, name_hash, kmer_code, k_len, kc_val, kc_has, r_start, r_end, r_lb, r_ub)
    ms = timed(call, max(2, args.calls // 2))
    km = kernel_ms(eng, call, ("k_exp_prep", "k_exp_entries", "k_exp_region", "k_exp_revsum"))
    print(json.dumps({"row": "pileup.experimental: %d whole-contig regions" % g, "reads": n, "ms_per_call": ms,
                      "reads_per_s": n / ms * 1e3, "regions_per_s": g / ms * 1e3, "kernel_ms": km,
                      "note": "kernel_ms excludes the CUB segmented sort of the (name, read) entries; ms_per_call includes "
                              "it, the H2D of pos/flag/CIGAR/name hashes/k-mer codes (%.0f MB) and the record read-back"
                              % ((4 + 2 + 4 + 8 + 4) * n / 1e6 + batch.cig.nbytes / 1e6),
                      "cpu_baseline": None}), flush=True)
    eng.close()

    # ---- row: `metacov scan` from the FILE: ByFlag(2 flags) over IsizeHist + KmerHist, host decode vs GPU decode -----------------
    # (GPU decode: the columns and the SEQ windows stay on the device -- mcov_isize_hist / mcov_kmer_hist_mem with
    #  MCOV_MEM_DEVICE -- and only the histograms come back)
    import tempfile
    from metacov_b200 import AlignmentFile, scan
    wb = synth.c2(min(args.scale, 0.2))
    hb, isz = synth.generate_host(wb)
    tmp = tempfile.mkdtemp(prefix="mcov_rows_")
    path = os.path.join(tmp, "rows.bam")
    synth.write_bam(path, wb, hb, isz)
    res = {}
    for mode in ("host", "gpu"):
        best, rows = None, None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with AlignmentFile(path, decode=mode) as af:
                bf = scan.ByFlag([scan.IsizeHist(), scan.KmerHist(K, NK, STEP, OFFSET)], [scan.Flags["Readdir"], scan.Flags["IsRead1"]])
                nrec = scan.scan_reads(af, None, bf)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            rows = [p.processors[1].counts.sum() for p in bf.processors]
        res[mode] = {"ms_total": 1e3 * best, "reads_per_s": nrec / best, "kmer_counts": [int(x) for x in rows]}
    assert res["host"]["kmer_counts"] == res["gpu"]["kmer_counts"]
    print(json.dumps({"row": "metacov scan from a BAM file: ByFlag(2 flags) over IsizeHist + KmerHist(K=7,NK=8,STEP=7)", "reads": int(len(hb.tid)),
                      "bam_bytes": os.path.getsize(path), "host_decode": res["host"], "gpu_decode": res["gpu"],
                      "speedup": res["host"]["ms_total"] / res["gpu"]["ms_total"],
                      "timed": "closed file on disk (page cache warm) -> histograms on the host, best of 3"}), flush=True)
    try:
        os.remove(path); os.remove(path + ".bai"); os.rmdir(tmp)
    except OSError:
        pass


if __name__ == "__main__":
    main()
