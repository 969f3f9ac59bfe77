"""Drop-in for ``metacov.pileup`` on the B200 path.

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
    return [finish_classic(st[i], int(ends[i] - starts[i])) for i in range(len(tids))]


def classic(bam, ref, start, end):
    """Drop-in for ``metacov.pileup.classic`` (reference pileup.py:9)."""
    return classic_many(bam, [ref], [start], [end])[0]


def load_kmerhist(f, k_len=7):
    """Drop-in for ``metacov.pileup.load_kmerhist`` (reference pileup.py:29-35):
    k-mer -> n0 / mean(n1..) for R1 and R2 from ``metacov scan``'s CSV."""
    import pandas as pd
    df = pd.read_csv(f)
    df = df[~((df.Mapped == "Unmapped") | (df.kmer == "N" * k_len))]
    df = df.set_index("kmer")
    cor = df[df.columns[0]] / df[df.columns[1:]].mean(axis=1)
    return [cor[df.R == r].to_dict() for r in ("R1", "R2")]
