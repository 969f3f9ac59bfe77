"""CPU restatement of ``metacov.pileup.experimental`` (oracle; test infrastructure
only -- never imported by the product package).

Follows reference metacov/pileup.py:38-173 over the SoA records of a file
(``oracle.bamio.BamRecords``) instead of pysam objects.  Pinned against the
reference's own function, imported unmodified and run over the restated
``fetch`` boundary, through the vectors ``oracle/make_golden.py`` writes to
``tests/golden/*_experimental.json`` (``tests/test_oracle.py``).

The per-region result is returned in two forms: ``sums`` (the exact integer
counts and the float sums the 13 outputs are made of -- what the GPU path
reports per region, ``mcov_exp_stats``) and ``result`` (the 13 outputs).
"""
import math

import numpy as np

from . import bamio

BAM_FPROPER_PAIR, BAM_FUNMAP, BAM_FREVERSE, BAM_FREAD1, BAM_FSECONDARY = 0x2, 0x4, 0x10, 0x40, 0x100


def synthetic_kcor(k_len):
    """Deterministic stand-in for ``load_kmerhist``'s two dicts (tests and golden vectors): every
    ACGT k-mer maps to a ratio in [0.5, 1.5), except that some keys are missing (KeyError paths,
    pileup.py:112, 125) and some ratios are 0 (pileup.py:114, 128)."""
    out = []
    for r in (0, 1):
        d = {}
        for code in range(4 ** k_len):
            h = (code * 2654435761 + r * 40503 + 12345) % 1000003
            if h % 17 == 0:
                continue
            kmer = "".join("ACGT"[(code >> (2 * (k_len - 1 - j))) & 3] for j in range(k_len))
            d[kmer] = 0.0 if h % 19 == 0 else 0.5 + (h % 1000) / 1000
        out.append(d)
    return out


def _py_slice_count(i, j, n):
    """len(a[i:j]) for a sequence of length n (negative indices wrap, pileup.py:105)."""
    return len(range(*slice(i, j).indices(n)))


def aligned_prefix(recs, i, k_len):
    """``read.query_alignment_sequence[0:k_len]`` (pileup.py:109, 123): SEQ as stored, soft clips removed."""
    ops = recs.cig[recs.cig_off[i]:recs.cig_off[i + 1]]
    seq = recs.seqs[i]
    lo, hi = 0, len(seq)
    for op in ops:
        o, ln = int(op) & 15, int(op) >> 4
        if o == 5:
            continue
        if o != 4:
            break
        lo += ln
    for op in ops[::-1]:
        o, ln = int(op) & 15, int(op) >> 4
        if o == 5:
            continue
        if o != 4:
            break
        hi -= ln
    return "".join(bamio.NT16[c] for c in seq[lo:hi][:k_len])


def region_sums(recs, tid, start, end, k_cor, k_len):
    """The read loop of pileup.py:90-146 for one region, as sums."""
    length = end - start
    flag, pos, reflen = recs.flag, recs.pos, recs.reflen
    unmapped = (flag & BAM_FUNMAP) != 0
    span = np.where(unmapped, 0, reflen)
    endpos = pos.astype(np.int64) + np.where(span > 0, span, 1)          # bam_endpos (Appendix A-7)
    fetched = np.nonzero((recs.tid == tid) & (endpos > start) & (pos < end))[0]
    secondary = improper = nreads = no_reflen = n_pairs = 0
    cov_sum = cov2_sum = 0
    covw_sum = wnf_sum = 0.0
    last_w = {}                       # rstart -> 1/rcor of the last read that started there
    waiting = {}                      # query_name -> record index of the unmatched mate
    for i in fetched:
        f = int(flag[i])
        if f & BAM_FSECONDARY:                                         # pileup.py:92-94
            secondary += 1
            continue
        if not f & BAM_FPROPER_PAIR:                                   # pileup.py:97-99
            improper += 1
            continue
        kmer = aligned_prefix(recs, i, k_len) if k_cor else None
        name = recs.names[i]
        if name in waiting:                                            # pileup.py:101-116
            j = waiting.pop(name)
            s = min(int(pos[i]), int(pos[j])) - start
            e = max(int(pos[i]), int(pos[j])) - start
            cov2_sum += _py_slice_count(s - 1, e + 1, length)
            term = 1.0
            if k_cor:
                kj = aligned_prefix(recs, j, k_len)
                da = k_cor[1 if f & BAM_FREVERSE else 0]
                db = k_cor[1 if int(flag[j]) & BAM_FREVERSE else 0]
                if kmer in da and kj in db:
                    prod = da[kmer] * db[kj]
                    term = 1.0 if prod == 0 else 1 / prod
            wnf_sum += term
            n_pairs += 1
        else:
            waiting[name] = i
        rcor = 1
        if k_cor:                                                      # pileup.py:121-130
            d = k_cor[0 if f & BAM_FREAD1 else 1]
            if kmer in d:
                rcor = d[kmer]
            if rcor == 0:
                rcor = 1
        has_len = not (f & BAM_FUNMAP) and recs.cig_off[i + 1] > recs.cig_off[i]
        if not has_len:                                                # reference_length is None -> TypeError in the reference
            no_reflen += 1
            continue
        rl = int(reflen[i]) if reflen[i] > 0 else 1
        if f & BAM_FREVERSE:                                           # pileup.py:132-137
            rend = int(pos[i]) - start
            rstart = rend - rl
        else:
            rstart = int(pos[i]) - start
            rend = rstart + rl
        covered = max(0, min(length, rend) - max(0, rstart))           # pileup.py:139-141
        cov_sum += covered
        covw_sum += covered * (1 / rcor)
        if 0 <= rstart < length:                                       # pileup.py:143-146
            last_w[rstart] = 1 / rcor
            nreads += 1
    cor_sum = 0.0
    for p in sorted(last_w):                                           # sum(cor) runs over positions
        cor_sum += last_w[p]
    return dict(covw_sum=covw_sum, cor_sum=cor_sum, wnf_sum=wnf_sum, cov_sum=cov_sum, cov2_sum=cov2_sum,
                n_starts=len(last_w), nreads=nreads, secondary=secondary, improper=improper,
                no_reflen=no_reflen, n_pairs=n_pairs)


def fasta_terms(region, k_cor, k_len):
    """gc and ecor of pileup.py:62-84 for the (upper-cased) region string."""
    length = len(region)
    gc = region.count("G") + region.count("C")
    gc = gc / (gc + region.count("A") + region.count("T"))
    insert, sd = 450, 150
    x = np.arange(0, 2 * insert + 1, dtype=np.float64)
    norm = np.exp(-((x - insert) / sd) ** 2 / 2) / (sd * math.sqrt(2 * math.pi))
    cor_fwd = np.zeros(length)
    cor_rev = np.zeros(length)
    for i in range(length - k_len):
        cor_fwd[i] = k_cor[0].get(region[i:i + k_len], 0)
    for j in range(k_len - 1, length):
        cor_rev[j] = k_cor[1].get(region[j - k_len + 1:j + 1][::-1], 0)
    revsum = np.zeros(length)
    for i in range(length):
        n = min(length - i, 2 * insert)
        revsum[i] = float(np.dot(norm[:n], cor_rev[i:i + n]))
    return gc, np.float64(np.dot(cor_fwd, revsum)) / length        # np.float64, as np.inner(...) / length is


def outputs(sums, length, gc=-1, ecor=-1):
    """pileup.py:150-173."""
    n_starts, nreads = sums["n_starts"], sums["nreads"]
    nz = length - n_starts
    nzef = nz / (length * (1 - 1 / length) ** nreads)
    allreads = sums["secondary"] + nreads + sums["improper"]
    wnf = sums["wnf_sum"]
    with np.errstate(invalid="ignore", divide="ignore"):
        cf = np.float64(sums["cor_sum"]) / np.float64(n_starts)
    return {
        "cov": sums["cov_sum"] / length,
        "covc": sums["covw_sum"] / length,
        # np.mean(...) is a np.float64 and numpy's round (scale, rint, unscale) differs from Python's on ties
        "den": round(np.float64(n_starts / length), 3),
        "denc": round(np.float64(sums["cor_sum"] / length), 3),
        "cov2": round(np.float64(sums["cov2_sum"] / length)),
        "cf": round(cf, 3),
        "ambig": round(sums["secondary"] / allreads, 3) if allreads > 0 else 0,
        "improper": round(sums["improper"] / allreads, 3) if allreads > 0 else 0,
        "nzef": round(nzef, 3),
        "gc": round(gc, 3),
        "ecor": round(ecor, 3),
        "wnf": round(wnf / length, 3),
        "cov3": round(200 * (wnf / ecor) / nzef / length, 3),
    }


def experimental(recs, references, k_cor, k_len, fasta_seq, ref, start, end):
    """Same arguments as the reference function, with the decoded records in place of ``bam`` and
    the contig's sequence (or None) in place of ``fasta``."""
    if end - start == 0:
        raise Exception("Length must be > 0")
    tid = list(references).index(ref)
    sums = region_sums(recs, tid, start, end, k_cor, k_len)
    if sums["no_reflen"]:
        raise TypeError("unsupported operand type(s) for +: 'int' and 'NoneType'")
    gc, ecor = (-1, -1)
    if fasta_seq is not None:
        gc, ecor = fasta_terms(fasta_seq[start:end].upper(), k_cor, k_len)
    return sums, outputs(sums, end - start, gc, ecor)
