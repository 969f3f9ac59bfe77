"""Generate tests/golden/* by running the REFERENCE's own ``classic()``
(imported from /root/reference/metacov/pileup.py, unmodified) over the restated
pysam boundary (oracle/pysam_boundary.py).

Run in the build container (needs /root/reference):  python -m oracle.make_golden
The GPU box has no /root/reference, so the vectors are committed.
"""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bamio  # noqa: E402
from oracle.pysam_boundary import FakeAlignmentFile  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def ref_pileup_module():
    spec = importlib.util.spec_from_file_location("ref_metacov_pileup", os.path.join(REF, "metacov", "pileup.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def jsonable(d):
    return {k: (float(v) if isinstance(v, (float, np.floating)) else int(v)) for k, v in d.items()}


def save_soa(path, hdr, r):
    seq_off = np.concatenate(([0], np.cumsum([len(s) for s in r.seqs]))).astype(np.int64)
    seq = np.concatenate(r.seqs).astype(np.uint8) if r.seqs else np.zeros(0, np.uint8)
    np.savez_compressed(
        path, references=np.array(hdr.references), lengths=np.array(hdr.lengths, dtype=np.int32),
        tid=r.tid, pos=r.pos, flag=r.flag, mapq=r.mapq, l_seq=r.l_seq, isize=r.isize, mtid=r.mtid, mpos=r.mpos,
        cig_off=r.cig_off.astype(np.uint32), cig=r.cig, names=np.array(r.names), seq_off=seq_off, seq=seq)


def fixture():
    ref = ref_pileup_module()
    bam_path = os.path.join(REF, "tests", "data", "bbmap.sorted.bam")
    hdr, recs = bamio.read_bam(bam_path)
    bai = bamio.read_bai_stats(bam_path + ".bai")
    bam = FakeAlignmentFile(hdr, recs, bai)
    save_soa(os.path.join(GOLD, "fixture_soa.npz"), hdr, recs)
    out = {"source": "reference tests/data/bbmap.sorted.bam via reference metacov/pileup.py:classic",
           "references": list(hdr.references), "lengths": list(hdr.lengths),
           "mapped": bam.mapped, "unmapped": bam.unmapped, "n_records": len(recs),
           "bai": {"per_ref": [list(x) for x in bai[0]], "n_no_coor": bai[1]}}
    # regions: BAM header (test_pileup2) and regions.blast7 (test_pileup), cli.py:85-91 semantics
    regions = [(n, 0, l) for n, l in zip(hdr.references, hdr.lengths)]
    with open(os.path.join(REF, "tests", "data", "regions.blast7")) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            sacc, s, e = line.split("\t")[:3]
            s, e = sorted((int(s), int(e)))
            regions.append((sacc, s, e))
    # extra: odd/even lengths, single base, region reaching past the contig end
    regions += [("ref1", 10, 11), ("ref1", 10, 12), ("ref1", 100, 103), ("ref2", 500, 600), ("ref1", 0, 4)]
    out["classic"] = [{"ref": n, "start": s, "end": e, "result": jsonable(ref.classic(bam, n, s, e))}
                      for n, s, e in regions]
    depth = {}
    for n, l in zip(hdr.references, hdr.lengths):
        d = np.zeros(l, dtype=np.int32)
        for col in bam.pileup(n, 0, l):
            if 0 <= col.pos < l:
                d[col.pos] += col.n
        depth[n] = {"sum": int(d.sum()), "max": int(d.max()), "sha1_le_i32": hashlib.sha1(d.astype("<i4").tobytes()).hexdigest(),
                    "first8": d[:8].tolist(), "last4": d[-4:].tolist()}
        np.save(os.path.join(GOLD, "fixture_depth_%s.npy" % n), d)
    out["depth"] = depth
    with open(os.path.join(GOLD, "fixture_classic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    return out


def synthetic():
    """Small synthetic cases (deep stack to exercise the cap, indel CIGARs, clipped reads) run
    through the reference classic(); inputs are stored next to the outputs."""
    ref = ref_pileup_module()
    rng = np.random.Generator(np.random.PCG64(20261018))
    lengths = [1200, 37, 5000]
    names = ["ctgA", "ctgB extra words", "ctgC"]
    tid, pos, flag, mapq, cigs = [], [], [], [], []
    flags_pool = [99, 147, 83, 163, 99, 147, 83, 163, 97, 73, 69, 355, 1123, 611, 0, 16, 2048, 2064]
    for c, ln in enumerate(lengths):
        n = {0: 900, 1: 60, 2: 2500}[c]
        # ctgA/ctgB: reads may overhang the contig end (clipping); ctgC: none do, so that a region
        # reaching past its end sees zeros there in the reference too
        ps = np.sort(rng.integers(-3, ln if c < 2 else ln - 600, n))
        for p in ps:
            kind = rng.integers(0, 6)
            if kind == 0:
                cg = [(int(rng.integers(20, 151)) << 4) | 0]
            elif kind == 1:
                cg = [(int(rng.integers(1, 30)) << 4) | 4, (int(rng.integers(20, 120)) << 4) | 0]
            elif kind == 2:
                cg = [(int(rng.integers(10, 80)) << 4) | 7, (int(rng.integers(1, 9)) << 4) | 1, (int(rng.integers(10, 80)) << 4) | 8]
            elif kind == 3:
                cg = [(int(rng.integers(10, 80)) << 4) | 0, (int(rng.integers(1, 400)) << 4) | 3, (int(rng.integers(10, 80)) << 4) | 0]
            elif kind == 4:
                cg = [(5 << 4) | 5, (int(rng.integers(10, 80)) << 4) | 0, (int(rng.integers(1, 5)) << 4) | 2,
                      (int(rng.integers(10, 80)) << 4) | 0, (3 << 4) | 4]
            else:
                cg = [(int(rng.integers(1, 12)) << 4) | 1]            # insertion only: reflen 0
            f = flags_pool[int(rng.integers(0, len(flags_pool)))]
            tid.append(c); pos.append(max(int(p), 0)); flag.append(f); mapq.append(int(rng.integers(0, 61)))
            cigs.append(cg if not (f & 4) else [])
    # a few unplaced reads at the end
    for _ in range(5):
        tid.append(-1); pos.append(-1); flag.append(77); mapq.append(0); cigs.append([])
    cig_off = np.concatenate(([0], np.cumsum([len(c) for c in cigs]))).astype(np.uint32)
    cig = np.array([op for c in cigs for op in c], dtype=np.uint32)
    recs = bamio.BamRecords()
    recs.tid = np.array(tid, np.int32); recs.pos = np.array(pos, np.int32); recs.flag = np.array(flag, np.uint16)
    recs.mapq = np.array(mapq, np.uint8); recs.cig_off = cig_off.astype(np.int64); recs.cig = cig
    recs.l_seq = np.zeros(len(tid), np.int32); recs.isize = np.zeros(len(tid), np.int32)
    recs.mtid = np.full(len(tid), -1, np.int32); recs.mpos = np.full(len(tid), -1, np.int32)
    recs.names = ["s%d" % i for i in range(len(tid))]; recs.seqs = [np.zeros(0, np.uint8)] * len(tid)
    recs.reflen = bamio.cigar_reflen(recs.cig_off, recs.cig)
    hdr = bamio.BamHeader("", tuple(names), tuple(lengths))
    bam = FakeAlignmentFile(hdr, recs)
    np.savez_compressed(os.path.join(GOLD, "synth_small_soa.npz"), references=np.array(names),
                        lengths=np.array(lengths, np.int32), tid=recs.tid, pos=recs.pos, flag=recs.flag,
                        mapq=recs.mapq, cig_off=cig_off, cig=cig)
    regions = [(n, 0, l) for n, l in zip(names, lengths)]
    regions += [("ctgA", 1, 1200), ("ctgA", 300, 301), ("ctgA", 299, 907), ("ctgC", 1000, 4999), ("ctgC", 4990, 5010),
                ("ctgB extra words", 5, 30), ("ctgC", 0, 2), ("ctgC", 17, 20)]
    out = {"source": "synthetic SoA (oracle/make_golden.py:synthetic) via reference metacov/pileup.py:classic",
           "classic": [{"ref": n, "start": s, "end": e, "result": jsonable(ref.classic(bam, n, s, e))}
                       for n, s, e in regions]}
    with open(os.path.join(GOLD, "synth_small_classic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    return out


class _Fasta:
    """``pysam.FastaFile`` stand-in: fetch(ref, start, end) on in-memory sequences."""

    def __init__(self, seqs):
        self.seqs = seqs

    def fetch(self, ref, start, end):
        return self.seqs[ref][start:end]


def _exp_rows(ref_mod, bam, k_cor, k_len, fasta, regions):
    import contextlib
    import io
    rows = []
    for n, s, e, with_fa in regions:
        row = {"ref": n, "start": s, "end": e, "fasta": bool(with_fa)}
        try:
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):   # "RCOR is ZERO" prints (pileup.py:129)
                res = ref_mod.experimental(bam, k_cor, k_len, fasta if with_fa else None, n, s, e)
            # JSON has no nan / inf: stored as strings
            row["result"] = {k: (float(v) if np.isfinite(v) else repr(float(v))) for k, v in res.items()}
        except Exception as ex:                                    # the reference's own failure modes are part of the contract
            row["raises"] = type(ex).__name__
        rows.append(row)
    return rows


def experimental_fixture():
    """The reference's ``experimental()`` on the reference fixture BAM + FASTA with the synthetic
    k-mer ratios of oracle.experimental.synthetic_kcor."""
    import gzip
    from oracle.experimental import synthetic_kcor
    ref = ref_pileup_module()
    bam_path = os.path.join(REF, "tests", "data", "bbmap.sorted.bam")
    hdr, recs = bamio.read_bam(bam_path)
    bam = FakeAlignmentFile(hdr, recs, None)
    seqs, name = {}, None
    with gzip.open(os.path.join(REF, "tests", "data", "reference_1K.fa.gz"), "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                name = line[1:].split()[0]
                seqs[name] = []
            else:
                seqs[name].append(line.strip())
    seqs = {k: "".join(v) for k, v in seqs.items()}
    k_len = 7
    k_cor = synthetic_kcor(k_len)
    regions = [("ref1", 0, 425, True), ("ref2", 0, 575, True), ("ref1", 1, 425, True), ("ref2", 1, 575, False),
               ("ref2", 1, 300, True), ("ref2", 301, 575, True), ("ref1", 100, 103, False), ("ref1", 10, 11, False),
               ("ref2", 500, 600, False), ("ref1", 200, 420, True)]
    out = {"source": "reference tests/data/bbmap.sorted.bam + reference_1K.fa.gz via reference metacov/pileup.py:experimental",
           "k_len": k_len, "fasta": seqs, "rows": _exp_rows(ref, bam, k_cor, k_len, _Fasta(seqs), regions)}
    with open(os.path.join(GOLD, "fixture_experimental.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    return out


def experimental_synthetic():
    """Paired synthetic reads that reach the corners of ``experimental``: reverse reads whose
    interval starts before the region, mates on either side of a region border (negative slice
    indices), names seen three and four times, secondary / improper reads, soft and hard clips,
    N bases, aligned parts shorter than k_len."""
    from oracle.experimental import synthetic_kcor
    ref = ref_pileup_module()
    rng = np.random.Generator(np.random.PCG64(20261019))
    lengths = [900, 2500]
    names = ["pA", "pB second"]
    rows = []                                    # (tid, pos, flag, cigar ops, name, seq codes)
    nt = np.array([1, 2, 4, 8], dtype=np.uint8)  # A C G T in nt16
    serial = 0
    for c, ln in enumerate(lengths):
        for _ in range({0: 260, 1: 700}[c]):
            frag = int(rng.integers(60, 400))
            p1 = int(rng.integers(0, ln - 50))
            p2 = min(ln - 30, p1 + frag)
            rlen = int(rng.integers(4, 101))
            nm = "q%d" % serial
            serial += 1
            fwd_first = rng.random() < 0.5
            f1, f2 = (99, 147) if fwd_first else (83, 163)
            kind = rng.random()
            if kind < 0.08:
                f1, f2 = f1 & ~2, f2 & ~2                       # improper pair
            for p, f in ((p1, f1), (p2, f2)):
                k = rng.integers(0, 5)
                ops = [(rlen << 4) | 0]
                lseq = rlen
                if k == 1:
                    sc = int(rng.integers(1, 12)); ops = [(sc << 4) | 4, (rlen << 4) | 0]; lseq = rlen + sc
                elif k == 2:
                    sc = int(rng.integers(1, 9)); ops = [(3 << 4) | 5, (rlen << 4) | 7, (sc << 4) | 4]; lseq = rlen + sc
                elif k == 3 and rlen > 20:
                    a = rlen // 2; ops = [(a << 4) | 0, (2 << 4) | 2, ((rlen - a) << 4) | 0]
                seq = nt[rng.integers(0, 4, lseq)].copy()
                if rng.random() < 0.05:
                    seq[int(rng.integers(0, min(lseq, 8)))] = 15   # an N near the start
                rows.append((c, p, f, ops, nm, seq))
                if rng.random() < 0.04:                            # secondary copy of the same read
                    rows.append((c, p, f | 0x100, ops, nm, seq))
                if rng.random() < 0.03:                            # supplementary piece: same name, proper-pair flag kept
                    rows.append((c, min(ln - 20, p + 7), f | 0x800, [(12 << 4) | 0], nm, nt[rng.integers(0, 4, 12)].copy()))
    rows.sort(key=lambda r: (r[0], r[1]))
    recs = bamio.BamRecords()
    recs.tid = np.array([r[0] for r in rows], np.int32); recs.pos = np.array([r[1] for r in rows], np.int32)
    recs.flag = np.array([r[2] for r in rows], np.uint16); recs.mapq = np.full(len(rows), 30, np.uint8)
    cigs = [r[3] for r in rows]
    recs.cig_off = np.concatenate(([0], np.cumsum([len(c) for c in cigs]))).astype(np.int64)
    recs.cig = np.array([op for c in cigs for op in c], dtype=np.uint32)
    recs.names = [r[4] for r in rows]; recs.seqs = [r[5] for r in rows]
    recs.l_seq = np.array([len(s) for s in recs.seqs], np.int32); recs.isize = np.zeros(len(rows), np.int32)
    recs.mtid = np.full(len(rows), -1, np.int32); recs.mpos = np.full(len(rows), -1, np.int32)
    recs.reflen = bamio.cigar_reflen(recs.cig_off, recs.cig)
    hdr = bamio.BamHeader("", tuple(names), tuple(lengths))
    bam = FakeAlignmentFile(hdr, recs)
    save_soa(os.path.join(GOLD, "synth_pairs_soa.npz"), hdr, recs)
    fa = {n: "".join("ACGTN"[int(x)] for x in rng.choice(5, ln, p=[0.24, 0.26, 0.26, 0.23, 0.01])) for n, ln in zip(names, lengths)}
    k_len = 5
    k_cor = synthetic_kcor(k_len)
    regions = [("pA", 0, 900, True), ("pB second", 0, 2500, True), ("pA", 100, 500, False), ("pB second", 1200, 1201, False),
               ("pB second", 1000, 1400, True), ("pA", 850, 900, False), ("pA", 0, 60, False), ("pB second", 2400, 2600, False),
               ("pB second", 3, 2497, False), ("pA", 300, 302, False)]
    out = {"source": "synthetic paired SoA (oracle/make_golden.py:experimental_synthetic) via reference metacov/pileup.py:experimental",
           "k_len": k_len, "fasta": fa, "rows": _exp_rows(ref, bam, k_cor, k_len, _Fasta(fa), regions)}
    with open(os.path.join(GOLD, "synth_pairs_experimental.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    return out


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    f = fixture()
    s = synthetic()
    ef = experimental_fixture()
    es = experimental_synthetic()
    print("experimental rows:", len(ef["rows"]), len(es["rows"]))
    print("fixture regions:", len(f["classic"]), " synthetic regions:", len(s["classic"]))
    for row in f["classic"][:6]:
        print(row)
