"""Restated pysam/htslib boundary of the coverage hot path (oracle; test
infrastructure only -- never imported by the product package).

The reference's ``classic`` (reference metacov/pileup.py:9-26) and
``experimental`` (pileup.py:38-173) consume a duck-typed ``bam`` object:
``bam.pileup(ref, start, end)`` -> objects with ``.pos`` / ``.n``
(pileup.py:13-16) and ``bam.fetch(ref, start, end)`` -> read objects
(pileup.py:90-137).  In the reference that object is ``pysam.AlignmentFile``
(cli.py:56); pysam/htslib is an un-vendored, unpinned dependency
(requirements.txt:2) that is absent here, so its published algorithm is
restated:

* read filter  = pysam ``libcalignmentfile.pyx:__advance_samtools`` with the
  defaults of ``AlignmentFile.pileup`` (stepper="samtools": flag_filter =
  UNMAP|SECONDARY|QCFAIL|DUP, ignore_orphans=True, min_mapping_quality=0,
  max_depth=8000, truncate=False);
* column engine = htslib ``sam.c`` ``bam_plp_push`` / ``bam_plp_next`` /
  ``bam_plp_auto`` -- a sequential machine, including the order-dependent
  ``maxcnt`` cap;
* ``column.n``  = number of buffered passing reads with beg <= pos < end, where
  end = pos + bam_cigar2rlen (reads with D / N at the column are counted).

PARITY UNPINNED at this boundary (no reference test pins a number here).
"""
import numpy as np

from . import bamio

BAM_FPAIRED = 0x1
BAM_FPROPER_PAIR = 0x2
BAM_FUNMAP = 0x4
BAM_FMUNMAP = 0x8
BAM_FREVERSE = 0x10
BAM_FMREVERSE = 0x20
BAM_FREAD1 = 0x40
BAM_FREAD2 = 0x80
BAM_FSECONDARY = 0x100
BAM_FQCFAIL = 0x200
BAM_FDUP = 0x400
BAM_FSUPPLEMENTARY = 0x800

DEFAULT_FLAG_FILTER = BAM_FUNMAP | BAM_FSECONDARY | BAM_FQCFAIL | BAM_FDUP  # 0x704
DEFAULT_MAX_DEPTH = 8000


class PileupFilter:
    """The implicit arguments of ``bam.pileup(ref, start, end)`` made explicit."""

    def __init__(self, flag_filter=DEFAULT_FLAG_FILTER, flag_require=0, min_mapq=0,
                 ignore_orphans=True, max_depth=DEFAULT_MAX_DEPTH):
        self.flag_filter = flag_filter
        self.flag_require = flag_require
        self.min_mapq = min_mapq
        self.ignore_orphans = ignore_orphans
        self.max_depth = max_depth

    def passes(self, flag, mapq):
        """Vectorised ``__advance_samtools`` predicate (plus bam_plp_push's own
        ``flag & BAM_FUNMAP`` drop)."""
        flag = np.asarray(flag).astype(np.int64)
        mapq = np.asarray(mapq).astype(np.int64)
        ok = (flag & self.flag_filter) == 0
        if self.flag_require:
            ok &= (flag & self.flag_require) != 0
        if self.min_mapq > 0:
            ok &= mapq >= self.min_mapq
        if self.ignore_orphans:
            ok &= ~(((flag & BAM_FPAIRED) != 0) & ((flag & BAM_FPROPER_PAIR) == 0))
        ok &= (flag & BAM_FUNMAP) == 0
        return ok


class Column:
    """Stand-in for pysam.PileupColumn: only ``pos`` and ``n`` are consumed
    (reference metacov/pileup.py:14-16)."""
    __slots__ = ("pos", "n", "tid")

    def __init__(self, tid, pos, n):
        self.tid = tid
        self.pos = pos
        self.n = n

    nsegments = property(lambda self: self.n)
    reference_pos = property(lambda self: self.pos)


def plp_columns(reads, maxcnt=DEFAULT_MAX_DEPTH):
    """htslib pileup engine as a generator of (tid, pos, n).

    ``reads`` yields (tid, pos, end) of records that already passed the
    filter, in file order.  Follows sam.c: ``bam_plp_auto`` loops
    {``bam_plp_next``; read one record; ``bam_plp_push``}.
    """
    it_tid, it_pos = 0, 0          # bam_plp_init: calloc -> 0, 0
    max_tid, max_pos = -1, -1
    buf = []                       # linked list head..tail (excluding sentinel)
    is_eof = False
    reads = iter(reads)

    def plp_next():
        # one call of bam_plp_next: returns a (tid,pos,n) with n>0 or None
        nonlocal it_tid, it_pos, buf
        if is_eof and not buf:
            return None
        while is_eof or max_tid > it_tid or (max_tid == it_tid and max_pos > it_pos):
            n_plp = 0
            keep = []
            for node in buf:
                t, b, e = node
                if t < it_tid or (t == it_tid and e <= it_pos):
                    continue                      # retire (mp_free)
                if t == it_tid and b <= it_pos:
                    n_plp += 1                    # resolve_cigar2 always succeeds
                keep.append(node)
            buf = keep
            out = (it_tid, it_pos, n_plp)
            if buf:
                h_tid, h_beg, _ = buf[0]
                if it_tid < h_tid:
                    it_tid, it_pos = h_tid, h_beg
                elif it_pos < h_beg:
                    it_pos = h_beg
                else:
                    it_pos += 1
            else:
                # head == tail: the sentinel node is zero-initialised (tid 0, beg 0)
                if it_tid < 0:
                    it_tid, it_pos = 0, 0
                elif it_pos < 0:
                    it_pos = 0
                else:
                    it_pos += 1
            if n_plp:
                return out
            if is_eof and not buf:
                break
        return None

    while True:
        col = plp_next()
        if col is not None:
            yield col
            continue
        if is_eof:
            return
        # read alignments until a column can be produced
        got = False
        for (t, p, e) in reads:
            # bam_plp_push
            if t < 0:
                continue
            # mp->cnt = 1 (sentinel) + len(buf)
            if it_tid == t and it_pos == p and (1 + len(buf)) > maxcnt:
                continue                          # max_depth cap: read dropped
            if t < max_tid or (t == max_tid and p < max_pos):
                raise ValueError("the input is not sorted")
            max_tid, max_pos = t, p
            if e > it_pos or t > it_tid:
                buf.append((t, p, e))
            col = plp_next()
            if col is not None:
                got = True
                break
        if got:
            yield col
            continue
        is_eof = True                              # bam_plp_push(iter, 0)


class Read:
    """Stand-in for pysam.AlignedSegment: the attributes ``experimental``
    touches (reference metacov/pileup.py:92-137)."""

    def __init__(self, recs, i):
        self._r = recs
        self._i = i

    @property
    def flag(self):
        return int(self._r.flag[self._i])

    is_secondary = property(lambda s: bool(s.flag & BAM_FSECONDARY))
    is_proper_pair = property(lambda s: bool(s.flag & BAM_FPROPER_PAIR))
    is_reverse = property(lambda s: bool(s.flag & BAM_FREVERSE))
    is_read1 = property(lambda s: bool(s.flag & BAM_FREAD1))
    is_unmapped = property(lambda s: bool(s.flag & BAM_FUNMAP))

    @property
    def query_name(self):
        return self._r.names[self._i]

    @property
    def reference_start(self):
        return int(self._r.pos[self._i])

    @property
    def reference_length(self):
        # pysam: None when unmapped or without CIGAR, else bam_endpos - pos
        i = self._i
        if self.flag & BAM_FUNMAP or self._r.cig_off[i + 1] == self._r.cig_off[i]:
            return None
        rl = int(self._r.reflen[i])
        return rl if rl > 0 else 1

    @property
    def query_alignment_sequence(self):
        i = self._i
        ops = self._r.cig[self._r.cig_off[i]:self._r.cig_off[i + 1]]
        seq = self._r.seqs[i]
        s, e = 0, len(seq)
        for op in ops:                       # leading S (H is skipped)
            o, l = int(op) & 15, int(op) >> 4
            if o == 5:
                continue
            if o == 4:
                s += l
            else:
                break
        for op in ops[::-1]:
            o, l = int(op) & 15, int(op) >> 4
            if o == 5:
                continue
            if o == 4:
                e -= l
            else:
                break
        return "".join(bamio.NT16[c] for c in seq[s:e])


class FakeAlignmentFile:
    """Duck-typed ``pysam.AlignmentFile`` over decoded records."""

    def __init__(self, header, recs, bai_stats=None, filt=None):
        self.header = header
        self.references = tuple(header.references)
        self.lengths = tuple(header.lengths)
        self.recs = recs
        self.filter = filt or PileupFilter()
        self._name2tid = {n: i for i, n in enumerate(self.references)}
        # bam_endpos: reflen==0 / unmapped -> pos+1
        rl = np.where((recs.flag & BAM_FUNMAP) != 0, 0, recs.reflen)
        self._endpos = recs.pos.astype(np.int64) + np.where(rl > 0, rl, 1)
        if bai_stats is not None:
            per_ref, n_no_coor = bai_stats
            self.mapped = sum(m for m, _ in per_ref)
            self.unmapped = sum(u for _, u in per_ref) + n_no_coor
        else:
            un = (recs.flag & BAM_FUNMAP) != 0
            self.mapped = int(np.sum(~un & (recs.tid >= 0)))
            self.unmapped = int(np.sum(un))

    @classmethod
    def from_bam(cls, path, filt=None):
        import os
        hdr, recs = bamio.read_bam(path)
        bai = path + ".bai"
        stats = bamio.read_bai_stats(bai) if os.path.exists(bai) else None
        return cls(hdr, recs, stats, filt)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def _region_indices(self, contig, start, stop):
        tid = self._name2tid[contig]
        start = 0 if start is None else start
        stop = self.lengths[tid] if stop is None else stop
        r = self.recs
        sel = (r.tid == tid) & (self._endpos > start) & (r.pos < stop)
        return tid, np.nonzero(sel)[0]

    def fetch(self, contig, start=None, stop=None):
        """Every record overlapping the region, unfiltered, file order
        (SURVEY.md Appendix A-7)."""
        _, idx = self._region_indices(contig, start, stop)
        for i in idx:
            yield Read(self.recs, int(i))

    def pileup(self, contig, start=None, stop=None, **kw):
        """Columns of the samtools-stepper pileup (no truncation)."""
        f = self.filter
        maxcnt = kw.get("max_depth", f.max_depth)
        tid, idx = self._region_indices(contig, start, stop)
        r = self.recs
        ok = f.passes(r.flag[idx], r.mapq[idx])
        idx = idx[ok]
        pos = r.pos[idx].astype(np.int64)
        end = pos + r.reflen[idx]

        def gen():
            for p, e in zip(pos.tolist(), end.tolist()):
                yield (tid, p, e)

        for t, p, n in plp_columns(gen(), maxcnt):
            yield Column(t, p, n)
