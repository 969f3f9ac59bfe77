"""Drop-in for ``metacov.pileup`` on the B200 path.

``experimental(bam, k_cor, k_len, fasta, ref, start, end)`` keeps the signature and the 13 keys of
reference metacov/pileup.py:38-173; its read loop runs on the GPU (``mcov_experimental_run``).

``classic(bam, ref, start, end)`` keeps the signature, keys, value types and
rounding of reference metacov/pileup.py:9-26.  The per-base depth is computed
once per BAM on the GPU (instead of one htslib pileup per region) and every
region becomes a slice query answered from exact integer statistics
(``mcov_region_stats``); the float finishing below reproduces pileup.py:18-26.
"""
import numpy as np

from . import _capi


class DepthCapError(RuntimeError):
    """htslib's max_depth cap (pysam default 8000) would have dropped reads but the
    reads are not coordinate-sorted, so the (order-dependent) capped pileup is undefined.
    For sorted input the cap is replayed exactly on the GPU (k_cap_replay)."""


def finish_classic(st, n):
    """Seven outputs of ``classic`` from one ``mcov_region_stats`` record.

    Mirrors pileup.py:18-26: ``int()`` for min/max/med/sum, numpy-float64
    ``round(., 2)`` for std/avg/q23.  ``avg`` and
    ``q23`` are exact (integer sums below 2**53 divided once); ``std`` comes
    from exact integer moments, correctly rounded.
    """
    n = int(n)
    total = int(st["sum"])
    sumsq = int(st["sumsq"])
    var_num = n * sumsq - total * total          # exact (Python ints)
    std = np.sqrt(np.float64(var_num / (n * n))) if var_num > 0 else np.float64(0.0)
    q = n // 4
    iq_n = n - 2 * q
    return {
        "min": int(st["min"]),
        "max": int(st["max"]),
        "med": (int(st["med_lo"]) + int(st["med_hi"])) // 2,
        "std": round(std, 2),
        "avg": round(np.float64(total / n), 2),
        "q23": round(np.float64(int(st["iq_sum"]) / iq_n), 2) if iq_n > 0 else round(np.float64("nan"), 2),
        "sum": total,
    }


def finish_classic_many(st, ns):
    """``finish_classic`` for an array of records: the same seven values per region, computed with numpy over all
    regions at once (500 k contigs cost the per-record version 8 s of Python; the GPU pass behind them takes 4 ms).
    Bit-identical to ``finish_classic``: a region whose integer moments leave the range where float64 holds them
    exactly (2**53) takes the scalar path."""
    g = len(st)
    ns = np.asarray(ns, dtype=np.int64)
    if g == 0:
        return []
    total = st["sum"].astype(np.int64)
    sumsq = st["sumsq"].astype(np.int64)
    lim = float(1 << 53)
    nf, tf, sf = ns.astype(np.float64), total.astype(np.float64), sumsq.astype(np.float64)
    exact = (nf * sf < lim) & (tf * tf < lim) & (nf * nf < lim) & (ns > 0)
    var_num = ns * sumsq - total * total                       # (only looked at where `exact`)
    with np.errstate(invalid="ignore", divide="ignore"):
        std = np.where(var_num > 0, np.sqrt(var_num.astype(np.float64) / (ns * ns).astype(np.float64)), 0.0)
        avg = tf / nf
        iq_n = ns - 2 * (ns // 4)
        q23 = np.where(iq_n > 0, st["iq_sum"].astype(np.float64) / np.where(iq_n > 0, iq_n, 1).astype(np.float64), np.nan)
    exact &= st["iq_sum"].astype(np.float64) < lim
    std, avg, q23 = np.round(std, 2), np.round(avg, 2), np.round(q23, 2)
    med = (st["med_lo"].astype(np.int64) + st["med_hi"].astype(np.int64)) // 2
    mn, mx = st["min"].tolist(), st["max"].tolist()
    med, total_l = med.tolist(), total.tolist()
    out = [{"min": a, "max": b, "med": c, "std": d, "avg": e, "q23": f, "sum": t}
           for a, b, c, d, e, f, t in zip(mn, mx, med, list(std), list(avg), list(q23), total_l)]     # (list(): numpy.float64 items)
    for i in np.nonzero(~exact)[0]:
        out[i] = finish_classic(st[i], int(ns[i]))
    return out


def _engine_of(bam):
    eng = getattr(bam, "coverage_engine", None)
    if eng is None:
        raise TypeError(
            "metacov_b200.pileup needs a metacov_b200.AlignmentFile (got %r); the GPU path has no "
            "CPU fallback for foreign bam objects" % type(bam).__name__)
    return eng()


def classic_many(bam, refs, starts, ends):
    """``classic`` for many regions in one GPU pass (the loop of reference
    metacov/cli.py:85-95).  Returns a list of dicts."""
    eng = _engine_of(bam)
    tids = np.asarray([bam.get_tid(r) for r in refs], dtype=np.int32)
    starts = np.asarray(starts, dtype=np.int64)
    ends = np.asarray(ends, dtype=np.int64)
    if np.any(ends - starts <= 0):
        # np.amin of an empty vector (pileup.py:19); negative lengths fail earlier in np.zeros
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    # A region may reach past its contig: the reference's vector is end-start long whatever the
    # contig length (pileup.py:10-11) and stays 0 there; the C-ABI counts those positions as 0.
    st = eng.region_stats(tids, starts, ends)
    return finish_classic_many(st, ends - starts)


def classic(bam, ref, start, end):
    """Drop-in for ``metacov.pileup.classic`` (reference pileup.py:9)."""
    return classic_many(bam, [ref], [start], [end])[0]


def load_kmerhist(f, k_len=7):
    """Drop-in for ``metacov.pileup.load_kmerhist`` (reference pileup.py:29-35):
    k-mer -> n0 / mean(n1..) for R1 and R2 from ``metacov scan``'s CSV.

    The reference expects the columns ``Mapped`` and ``R`` and averages "all other columns", which
    only worked while pandas silently skipped the text columns; ``metacov scan`` at the reference's
    HEAD names the read-number column ``IsRead1`` / ``IsRead2`` instead.
    Both layouts are read here: the ratio is taken over the numeric columns, and the read number
    comes from ``R`` if present, else from ``IsRead1`` / ``IsRead2`` (values ``R1`` / ``R2``)."""
    import pandas as pd
    df = pd.read_csv(f)
    df = df[~((df.Mapped == "Unmapped") | (df.kmer == "N" * k_len))]
    df = df.set_index("kmer")
    num = df.select_dtypes("number")
    cor = num[num.columns[0]] / num[num.columns[1:]].mean(axis=1)
    rcol = df.R if "R" in df.columns else (df.IsRead1 if "IsRead1" in df.columns else df.IsRead2)
    return [cor[rcol == r].to_dict() for r in ("R1", "R2")]


# ---- pileup.experimental -------------------------------------------------------------------------

_NT = {"A": 0, "C": 1, "G": 2, "T": 3}


def _kcor_tables(k_cor, k_len):
    """The two dicts of ``load_kmerhist`` as dense tables (value, key-present) indexed by the
    2-bit code of the k-mer, first base most significant (mcov_bam_qas_kmer)."""
    if not k_cor:
        return None, None
    n = 4 ** k_len
    val = np.zeros((2, n), dtype=np.float64)
    has = np.zeros((2, n), dtype=np.uint8)
    for r in (0, 1):
        for kmer, v in k_cor[r].items():
            if len(kmer) != k_len:
                continue
            code = 0
            for ch in kmer:
                c = _NT.get(ch)
                if c is None:
                    code = -1
                    break
                code = code * 4 + c
            if code >= 0:
                val[r, code] = v
                has[r, code] = 1
    return val, has


def _region_kmer_cor(region, k_cor, k_len):
    """cor_fwd / cor_rev of reference pileup.py:66-77 (missing key -> 0), vectorised."""
    L = len(region)
    val, has = _kcor_tables(k_cor, k_len)
    b = np.frombuffer(region.encode("latin-1"), dtype=np.uint8)
    code = np.full(L, -1, dtype=np.int64)
    for ch, c in _NT.items():
        code[b == ord(ch)] = c
    cor_fwd = np.zeros(L)
    cor_rev = np.zeros(L)
    if L >= k_len:
        win = np.lib.stride_tricks.sliding_window_view(code, k_len)          # win[j] = codes of region[j:j+k]
        ok = (win >= 0).all(axis=1)
        pw = 4 ** np.arange(k_len - 1, -1, -1, dtype=np.int64)
        fwd_idx = np.where(ok, (win * pw).sum(axis=1), 0)
        rev_idx = np.where(ok, (win * pw[::-1]).sum(axis=1), 0)              # the same bases read backwards
        # cor_fwd[i] = k_cor[0][region[i:i+k]] for i in range(L-k)   (the last window is never looked up)
        n_f = max(L - k_len, 0)
        cor_fwd[:n_f] = np.where(ok[:n_f] & (has[0, fwd_idx[:n_f]] != 0), val[0, fwd_idx[:n_f]], 0.0)
        # cor_rev[j] = k_cor[1][reversed(region[j-k+1 .. j])] for j in [k-1, L-1]
        cor_rev[k_len - 1:] = np.where(ok & (has[1, rev_idx] != 0), val[1, rev_idx], 0.0)
    return cor_fwd, cor_rev


def _insert_model():
    """norm of reference pileup.py:54-60: N(450, 150) pdf at 0..900 (scipy.stats.norm.pdf)."""
    insert, sd = 450, 150
    x = (np.arange(0, 2 * insert + 1, dtype=np.float64) - insert) / sd
    return np.exp(-x ** 2 / 2.0) / np.sqrt(2 * np.pi) / sd, 2 * insert


def finish_experimental(st, length, gc, ecor):
    """The 13 outputs of ``experimental`` (reference pileup.py:150-173) from one
    ``mcov_exp_stats`` record; same expressions, types and rounding."""
    length = int(length)
    n_starts, nreads = int(st["n_starts"]), int(st["nreads"])
    secondary, improper = int(st["secondary"]), int(st["improper"])
    nz = length - n_starts
    nz_e = length * (1 - 1 / length) ** nreads
    nzef = nz / nz_e
    allreads = secondary + nreads + improper
    wnf = float(st["wnf_sum"])
    with np.errstate(invalid="ignore", divide="ignore"):
        cf = np.float64(st["cor_sum"]) / np.float64(n_starts)
    return {
        "cov": np.float64(int(st["cov_sum"]) / length),
        "covc": np.float64(st["covw_sum"]) / length,
        "den": round(np.float64(n_starts / length), 3),
        "denc": round(np.float64(st["cor_sum"]) / length, 3),
        "cov2": round(np.float64(int(st["cov2_sum"]) / length)),
        "cf": round(cf, 3),
        "ambig": round(secondary / allreads, 3) if allreads > 0 else 0,
        "improper": round(improper / allreads, 3) if allreads > 0 else 0,
        "nzef": round(nzef, 3),
        "gc": round(gc, 3),
        "ecor": round(ecor, 3),
        "wnf": round(wnf / length, 3),
        "cov3": round(200 * (wnf / ecor) / nzef / length, 3),
    }


def experimental_many(bam, k_cor, k_len, fasta, refs, starts, ends):
    """``experimental`` for many regions: one GPU pass per batch of regions."""
    if not hasattr(bam, "experimental_stats"):
        raise TypeError(
            "metacov_b200.pileup needs a metacov_b200.AlignmentFile (got %r); the GPU path has no "
            "CPU fallback for foreign bam objects" % type(bam).__name__)
    starts = [int(x) for x in starts]
    ends = [int(x) for x in ends]
    for s0, e0 in zip(starts, ends):
        if e0 - s0 == 0:
            raise Exception("Length must be > 0")                      # pileup.py:40-41
        if e0 - s0 < 0:
            raise ValueError("negative dimensions are not allowed")    # np.zeros(length), pileup.py:42
    kc_val, kc_has = _kcor_tables(k_cor, k_len)
    st = bam.experimental_stats(refs, starts, ends, k_len, kc_val, kc_has)
    norm, n_w = _insert_model()
    out = []
    for i, (ref, s0, e0) in enumerate(zip(refs, starts, ends)):
        length = e0 - s0
        if int(st[i]["no_reflen"]):
            # rend = rstart + read.reference_length with reference_length None (pileup.py:134-137)
            raise TypeError("unsupported operand type(s) for +: 'int' and 'NoneType'")
        if fasta:
            region = fasta.fetch(ref, s0, e0).upper()                   # pileup.py:62-65
            gc = region.count("G") + region.count("C")
            gc = gc / (gc + region.count("A") + region.count("T"))
            if not k_cor:
                raise UnboundLocalError("cannot access local variable 'ecor' where it is not associated with a value")
            cor_fwd, cor_rev = _region_kmer_cor(region, k_cor, k_len)
            # the reference indexes region with the region length; a FASTA shorter than the region fails there
            if len(region) != length:
                raise IndexError("string index out of range")
            revsum = bam.coverage_engine().exp_revsum(cor_rev, norm[:n_w])   # pileup.py:78-83 on the GPU
            ecor = np.inner(cor_fwd, revsum) / length
        else:
            gc = -1
            ecor = -1
        out.append(finish_experimental(st[i], length, gc, ecor))
    return out


def experimental(bam, k_cor, k_len, fasta, ref, start, end):
    """Drop-in for ``metacov.pileup.experimental`` (reference pileup.py:38)."""
    return experimental_many(bam, k_cor, k_len, fasta, [ref], [start], [end])[0]
