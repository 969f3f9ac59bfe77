"""CPU restatement of ``metacov.pileup.classic`` (oracle; test infrastructure
only -- never imported by the product package).

Follows reference metacov/pileup.py:9-26 (the function) and cli.py:85-108 (the
region loop that drives it).  Two formulations of the per-base depth are kept
so that each cross-checks the other:

* ``depth_columns``  -- per-column accumulation exactly as pileup.py:11-16
  does it, fed by the streaming htslib engine restated in
  ``pysam_boundary.plp_columns`` (includes the max_depth cap);
* ``depth_diffarray`` -- +1 @ pos, -1 @ pos+reflen and a running sum; equal to
  the first whenever the cap is idle (``cap_is_idle``).
"""
import numpy as np

from . import bamio
from .pysam_boundary import PileupFilter


def depth_columns(bam, ref, start, end):
    """pileup.py:10-16: float64 zeros, ``columns[pos-start] += column.n``."""
    columns = np.zeros(end - start)
    for col in bam.pileup(ref, start, end):
        if start <= col.pos < end:
            columns[col.pos - start] += col.n
    return columns


def stats_from_columns(columns):
    """pileup.py:18-26 on a float64 vector of per-base depth.

    Types follow the reference: int() casts for min/max/med/sum, numpy
    float64 ``round(., 2)`` for std/avg/q23.
    """
    columns = np.asarray(columns, dtype=np.float64)
    n = len(columns)
    q = n // 4
    ordered = np.sort(columns)
    return {
        "min": int(columns.min()),
        "max": int(columns.max()),
        "med": int(np.median(columns)),
        "std": round(np.std(columns), 2),
        "avg": round(np.mean(columns), 2),
        "q23": round(np.mean(ordered[q:n - q]), 2),
        "sum": int(columns.sum()),
    }


def classic(bam, ref, start, end):
    """Restated ``metacov.pileup.classic(bam, ref, start, end)``."""
    return stats_from_columns(depth_columns(bam, ref, start, end))


def depth_diffarray(tid, pos, flag, mapq, cig_off, cig, lengths, filt=None):
    """Whole-file per-base depth by difference arrays.

    Returns a list of int64 arrays, one per contig (length = contig length).
    Intervals are clipped to [0, len).  Valid when the max_depth cap is idle.
    """
    filt = filt or PileupFilter()
    tid = np.asarray(tid)
    pos = np.asarray(pos).astype(np.int64)
    ok = filt.passes(flag, mapq) & (tid >= 0)
    reflen = bamio.cigar_reflen(cig_off, cig)
    out = []
    for c, ln in enumerate(lengths):
        sel = ok & (tid == c)
        s = np.clip(pos[sel], 0, ln)
        e = np.clip(pos[sel] + reflen[sel], 0, ln)
        d = np.zeros(ln + 1, dtype=np.int64)
        np.add.at(d, s, 1)
        np.add.at(d, e, -1)
        out.append(np.cumsum(d)[:ln])
    return out


def cap_is_idle(depth, starts, max_depth):
    """Sufficient and necessary no-op condition of htslib's maxcnt for a
    whole-contig pileup: for every position p, buffered reads before the last
    push at p = depth[p-1] + starts[p] - 1 < max_depth."""
    prev = np.concatenate(([0], depth[:-1]))
    return bool(np.all(prev + starts <= max_depth))


def region_rows(bam, regions):
    """cli.py:85-108: rows of the ``metacov pileup`` CSV as dicts.

    ``regions`` yields objects with ``sacc, sstart, send`` (util.Region)."""
    name2ref = {w.split()[0]: w for w in bam.references}
    for hit in regions:
        ref = name2ref[hit.sacc]
        start, end = sorted((int(hit.sstart), int(hit.send)))
        row = classic(bam, ref, start, end)
        row.update({"sacc": hit.sacc, "start": hit.sstart, "end": hit.send})
        yield row
