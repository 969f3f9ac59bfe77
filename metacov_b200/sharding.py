"""Multi-GPU layout of the coverage path: contiguous contig ranges per rank.

Reads never span contigs and a coordinate-sorted BAM stores each contig's reads
contiguously, so every rank runs the whole kernel pipeline on its own contig
range with no data-path collective; the only exchange is ONE gather of the
64-byte per-region statistics records.  The reference is
single-process (reference metacov/cli.py:85-108 loops over regions serially).
"""
import numpy as np

from . import _capi

# cost model of one contig: slots written/read (12 B per slot over the push path, 8 on the fused
# path) + per-read bytes; only the ratio matters for balancing
SLOT_COST = 12
READ_COST = 27


def partition_contigs(contig_len, reads_per_contig, n_ranks):
    """Split contigs [0, C) into n_ranks contiguous ranges balanced on
    SLOT_COST*len + READ_COST*reads.  Returns bounds[n_ranks+1] (contig indices)."""
    contig_len = np.asarray(contig_len, dtype=np.int64)
    reads_per_contig = np.asarray(reads_per_contig, dtype=np.int64)
    n_contigs = len(contig_len)
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    cost = SLOT_COST * (contig_len + 1) + READ_COST * reads_per_contig
    csum = np.concatenate(([0], np.cumsum(cost)))
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, n_ranks):
        target = total * r // n_ranks
        # first boundary whose prefix cost reaches the target, at least one past the previous
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, bounds[-1]), n_contigs)
        bounds.append(b)
    bounds.append(n_contigs)
    return np.asarray(bounds, dtype=np.int64)


def reads_per_contig_from_tid(tid, n_contigs):
    """Counting pre-pass over tid[] (the alternative to the BAI pseudo-bins)."""
    tid = np.asarray(tid)
    return np.bincount(tid[(tid >= 0) & (tid < n_contigs)], minlength=n_contigs).astype(np.int64)


def shard_read_range(read_start, bounds, rank):
    """Read index range [lo, hi) of a rank given the per-contig read prefix."""
    return int(read_start[bounds[rank]]), int(read_start[bounds[rank + 1]])


def assign_regions(region_tid, bounds):
    """Owner rank of every region (regions never span contigs)."""
    region_tid = np.asarray(region_tid, dtype=np.int64)
    return (np.searchsorted(np.asarray(bounds), region_tid, side="right") - 1).astype(np.int64)


def gather_region_stats(local_stats, owner, rank, world_size, device=None, group=None):
    """All-gather the per-region statistics records so that every rank (rank 0
    writes the CSV) holds all G records in region order.

    local_stats: structured array (REGION_STATS_DTYPE) of THIS rank's regions, in
    the order of ``np.nonzero(owner == rank)``.  One collective, fixed-size
    records padded to the largest per-rank count.  Works over NCCL (device
    tensors) and gloo (CPU tensors).
    """
    import torch
    import torch.distributed as dist
    owner = np.asarray(owner)
    counts = np.bincount(owner, minlength=world_size)
    cap = int(counts.max()) if len(counts) else 0
    rec = _capi.REGION_STATS_DTYPE.itemsize
    buf = np.zeros(cap * rec, dtype=np.uint8)
    mine = np.ascontiguousarray(local_stats).view(np.uint8).reshape(-1)
    buf[:len(mine)] = mine
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world_size * cap * rec, dtype=torch.uint8, device=t.device)
    if world_size > 1:
        dist.all_gather_into_tensor(out, t, group=group)
    else:
        out.copy_(t)
    flat = out.cpu().numpy().reshape(world_size, cap * rec)
    merged = np.zeros(len(owner), dtype=_capi.REGION_STATS_DTYPE)
    for r in range(world_size):
        idx = np.nonzero(owner == r)[0]
        if len(idx):
            merged[idx] = flat[r, :len(idx) * rec].view(_capi.REGION_STATS_DTYPE)
    return merged


class DeviceGather:
    """Device-resident gather of the per-region statistics records (N>1).

    Every rank writes its records with ``CoverageEngine.region_stats_enqueue`` into
    ``local_dev``; ONE all-gather over NCCL (no host round trip) and one D2H copy into pinned
    memory deliver all records in region order.  The region->rank ownership is fixed at
    construction so that the per-step merge is a handful of slice copies.
    """

    def __init__(self, owner, world_size, device, group=None):
        import torch
        self.world = int(world_size)
        self.group = group
        self.rec = _capi.REGION_STATS_DTYPE.itemsize
        owner = np.asarray(owner)
        self.n_regions = len(owner)
        counts = np.bincount(owner, minlength=self.world) if len(owner) else np.zeros(self.world, np.int64)
        self.cap = max(int(counts.max()) if len(counts) else 0, 1)
        self.counts = counts
        # contiguous contig ranges => regions listed in contig order are already grouped by rank
        self.monotone = bool(np.all(np.diff(owner) >= 0)) if len(owner) else True
        self.index = None if self.monotone else [np.nonzero(owner == r)[0] for r in range(self.world)]
        # two slots: step k+1 can be enqueued (records, all-gather, copy-back) before step k is waited for
        self.local = [torch.zeros(self.cap * self.rec, dtype=torch.uint8, device=device) for _ in range(2)]
        self.out = [torch.empty(self.world * self.cap * self.rec, dtype=torch.uint8, device=device) for _ in range(2)]
        self.host = [torch.empty(self.world * self.cap * self.rec, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        # the collective and the copy-back run on a stream of their own: the next pass's kernels (enqueued on
        # the caller's stream right after submit) overlap the all-gather instead of queueing behind it
        self.side = torch.cuda.Stream(device=device)
        self.local_dev, self.out_dev, self.out_host = self.local[0], self.out[0], self.host[0]
        self.merged = np.zeros(self.n_regions, dtype=_capi.REGION_STATS_DTYPE)

    def submit(self, slot=0, want_host=True):
        """Enqueue the all-gather of ``local[slot]`` (written on the current stream) and, with want_host,
        the copy of all records to pinned memory -- both on the gather's side stream; returns without
        waiting.  ``local[slot]`` may be rewritten once ``collect(slot)`` has returned."""
        import torch
        import torch.distributed as dist
        self.ready[slot].record(torch.cuda.current_stream())
        self.side.wait_event(self.ready[slot])
        with torch.cuda.stream(self.side):
            if self.world > 1:
                dist.all_gather_into_tensor(self.out[slot], self.local[slot], group=self.group)
            else:
                self.out[slot].copy_(self.local[slot])
            if want_host:
                self.host[slot].copy_(self.out[slot], non_blocking=True)
            self.done[slot].record(self.side)

    def collect(self, slot=0, want_host=True):
        """Wait for ``submit(slot)``; with want_host return the records in region order."""
        self.done[slot].synchronize()
        if not want_host:
            return None
        flat = self.host[slot].numpy().reshape(self.world, self.cap * self.rec)
        if self.monotone and np.all(self.counts == self.cap):
            return flat.reshape(-1).view(_capi.REGION_STATS_DTYPE)       # zero-copy: already in region order
        out_u8 = self.merged.view(np.uint8).reshape(-1, self.rec)
        pos = 0
        for r in range(self.world):
            k = int(self.counts[r])
            if not k:
                continue
            part = flat[r, :k * self.rec].reshape(k, self.rec)          # plain byte rows: memcpy speed
            if self.monotone:
                out_u8[pos:pos + k] = part
                pos += k
            else:
                out_u8[self.index[r]] = part
        return self.merged

    def gather(self, want_host=True):
        """All-gather the records (slot 0) and wait; with want_host (rank 0: it writes the CSV) return
        them in region order from pinned host memory."""
        self.submit(0, want_host)
        return self.collect(0, want_host)


# ---- contigs cut between ranks ("boundary-read carry") --------------------------------
#
# A contig too large for one rank's share is cut at a position p.  What the right-hand rank needs
# from the left is the depth carried across p, i.e. the reads that start before p and end after it.
# A coordinate-sorted, indexed BAM hands those over directly (they are what fetch(ref, p, p+1)
# returns), so the carry is expressed as DATA, not as a collective: the boundary reads are given to
# both sides, each side shifts positions by its piece's origin, and the kernels' own clipping
# (start < 0 -> 0, end > len -> len; the sentinel slot takes the -1) makes every piece's deltas sum
# to zero exactly as for a whole contig.  Depth per piece is then bit-identical to the slice of the
# unsplit result.  The one quantity that does need an exchange is the statistics of a REGION that
# straddles a cut: `med` and `q23` (reference pileup.py:21,24) are order statistics of the merged
# multiset, so the ranks sum exact counting histograms (one all-reduce of 32 KB per cut region)
# and the merged histogram is walked on the GPU (mcov_region_hist_enqueue / mcov_hist_stats_enqueue).

from collections import namedtuple

Piece = namedtuple("Piece", "tid p0 p1")
Piece.__doc__ = "Positions [p0, p1) of contig tid owned by one rank."


def partition_positions(contig_len, reads_per_contig, n_ranks, min_piece=4096, align=4):
    """Balanced split of the concatenated contigs into n_ranks contiguous shares where a cut MAY
    fall inside a contig (reads assumed uniform within a contig).  Returns a list (per rank) of
    lists of ``Piece``; pieces shorter than ``min_piece`` are avoided by moving the cut to the
    nearer contig border."""
    contig_len = np.asarray(contig_len, dtype=np.int64)
    reads_per_contig = np.asarray(reads_per_contig, dtype=np.int64)
    n_contigs = len(contig_len)
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    cost = (SLOT_COST * (contig_len + 1) + READ_COST * reads_per_contig).astype(np.float64)
    csum = np.concatenate(([0.0], np.cumsum(cost)))
    total = float(csum[-1])
    cuts = [(0, 0)]                                            # (contig, position): start of each rank's share
    for r in range(1, n_ranks):
        target = total * r / n_ranks
        c = int(np.searchsorted(csum, target, side="right")) - 1
        c = min(max(c, 0), n_contigs - 1) if n_contigs else 0
        p = 0
        if n_contigs:
            ln = int(contig_len[c])
            p = int((target - csum[c]) / cost[c] * ln) if cost[c] > 0 else 0
            p -= p % align
            if p < min_piece:
                p = 0
            elif ln - p < min_piece:
                c, p = c + 1, 0
        cut = (c, p)
        cuts.append(max(cut, cuts[-1]))                        # never before the previous cut
    cuts.append((n_contigs, 0))
    shares = []
    for r in range(n_ranks):
        (c0, p0), (c1, p1) = cuts[r], cuts[r + 1]
        pieces = []
        for c in range(c0, min(c1 + (1 if p1 > 0 else 0), n_contigs)):
            a = p0 if c == c0 else 0
            b = p1 if (c == c1 and p1 > 0) else int(contig_len[c])
            if b > a or (int(contig_len[c]) == 0 and c < c1):
                pieces.append(Piece(c, a, b))
        shares.append(pieces)
    return shares


def _is_torch(a):
    return hasattr(a, "data_ptr") and hasattr(a, "is_cuda")


def piece_read_range(pos_contig, piece, contig_len, reach=None):
    """Index range [lo, hi) (relative to the contig's first read) of the reads a piece needs: those
    that start inside it plus the boundary reads that may reach into it (start in [p0-reach, p0);
    reach = an upper bound of a read's reference span, None = everything before p0)."""
    n = len(pos_contig)
    if _is_torch(pos_contig):
        import torch
        ss = lambda v: int(torch.searchsorted(pos_contig, torch.tensor([v], dtype=pos_contig.dtype, device=pos_contig.device)).item())
    else:
        ss = lambda v: int(np.searchsorted(pos_contig, v, side="left"))
    lo = 0 if (piece.p0 == 0 or reach is None) else ss(piece.p0 - int(reach))
    hi = n if piece.p1 >= int(contig_len) else ss(piece.p1)
    return lo, max(hi, lo)


def localize_reads(batch, read_start, pieces, contig_len, reach=None):
    """The read batch of one rank: for every piece the reads of ``piece_read_range`` with
    tid = index of the piece in ``pieces`` and pos shifted by -p0 (boundary reads get negative
    positions; the kernels clip them at the piece's first slot).  ``batch`` holds ALL reads in
    coordinate order (numpy arrays or torch tensors); ``read_start`` is the per-contig read prefix."""
    from .engine import ReadBatch
    tor = _is_torch(batch.tid)
    if tor:
        import torch
        cat = lambda xs, like: torch.cat(xs) if xs else like[:0]
    else:
        cat = lambda xs, like: np.concatenate(xs) if xs else like[:0]
    cols = {k: [] for k in ("tid", "pos", "flag", "mapq", "off", "cig")}
    ops = 0
    for k, pc in enumerate(pieces):
        rs, re = int(read_start[pc.tid]), int(read_start[pc.tid + 1])
        lo, hi = piece_read_range(batch.pos[rs:re], pc, contig_len[pc.tid], reach)
        lo, hi = rs + lo, rs + hi
        o0, o1 = int(batch.cig_off[lo]) & 0xFFFFFFFF, int(batch.cig_off[hi]) & 0xFFFFFFFF
        if tor:
            cols["tid"].append(torch.full((hi - lo,), k, dtype=batch.tid.dtype, device=batch.tid.device))
            off = (batch.cig_off[lo:hi].to(torch.int64) & 0xFFFFFFFF) - o0 + ops
        else:
            cols["tid"].append(np.full(hi - lo, k, dtype=np.int32))
            off = batch.cig_off[lo:hi].astype(np.int64) - o0 + ops
        cols["pos"].append(batch.pos[lo:hi] - pc.p0)
        cols["flag"].append(batch.flag[lo:hi])
        cols["mapq"].append(batch.mapq[lo:hi])
        cols["off"].append(off)
        cols["cig"].append(batch.cig[o0:o1])
        ops += o1 - o0
    if ops >= 2 ** 32:
        raise ValueError("more than 2^32-1 CIGAR ops in one shard")
    if tor:
        off = torch.cat(cols["off"] + [torch.tensor([ops], dtype=torch.int64, device=batch.tid.device)])
        off = torch.where(off >= 2 ** 31, off - 2 ** 32, off).to(torch.int32)      # uint32 bit pattern in an int32 tensor
    else:
        off = np.concatenate(cols["off"] + [np.array([ops], np.int64)]).astype(np.uint32)
    return ReadBatch(cat(cols["tid"], batch.tid), cat(cols["pos"], batch.pos), cat(cols["flag"], batch.flag),
                     cat(cols["mapq"], batch.mapq), off, cat(cols["cig"], batch.cig))


class RegionSplit:
    """Regions mapped onto the pieces of every rank.

    whole[r] = (index, tid, start, end): regions that lie in ONE piece of rank r (local coordinates;
               ``index`` = position in the caller's region list) -- finished locally, 64-byte records.
    cut[r]   = (cut_index, tid, start, end): rank r's parts of the regions that straddle a cut
               (``cut_index`` = position in ``cut_regions``) -- partial histograms, merged.
    cut_regions = caller-side indices of the straddling regions.
    """

    def __init__(self, n_ranks):
        self.whole = [([], [], [], []) for _ in range(n_ranks)]
        self.cut = [([], [], [], []) for _ in range(n_ranks)]
        self.cut_regions = []

    def arrays(self, which, rank):
        idx, tid, st, en = (self.whole if which == "whole" else self.cut)[rank]
        return (np.asarray(idx, np.int64), np.asarray(tid, np.int32), np.asarray(st, np.int32), np.asarray(en, np.int32))


def split_regions(reg_tid, reg_start, reg_end, shares, contig_len):
    """Map regions [start, end) of contig tid (end may reach past the contig: those positions count
    as depth 0, reference pileup.py:10-11) onto the pieces of ``shares``."""
    plan = RegionSplit(len(shares))
    by_contig = {}
    for r, pieces in enumerate(shares):
        for k, pc in enumerate(pieces):
            by_contig.setdefault(pc.tid, []).append((pc.p0, pc.p1, r, k))
    for i, (t, s, e) in enumerate(zip(reg_tid, reg_start, reg_end)):
        t, s, e = int(t), int(s), int(e)
        ln = int(contig_len[t])
        parts = []
        for p0, p1, r, k in by_contig.get(t, []):
            last = p1 >= ln                                   # the piece holding the contig's end takes the overhang
            a, b = max(s, p0), (e if last else min(e, p1))
            if b > a or (e == s and p0 <= s and (s < p1 or last)):
                parts.append((r, k, a - p0, b - p0))
        if e == s:
            parts = parts[:1]
        if len(parts) == 1:
            r, k, a, b = parts[0]
            w = plan.whole[r]
            w[0].append(i); w[1].append(k); w[2].append(a); w[3].append(b)
        elif len(parts) > 1:
            ci = len(plan.cut_regions)
            plan.cut_regions.append(i)
            for r, k, a, b in parts:
                c = plan.cut[r]
                c[0].append(ci); c[1].append(k); c[2].append(a); c[3].append(b)
        else:
            raise ValueError("region %d (tid %d, %d-%d) lies in no piece" % (i, t, s, e))
    return plan


def sharded_region_stats(eng, plan, rank, world_size, device, group=None, breadth_n=1):
    """Statistics of all regions of ``plan`` with contigs cut between ranks: local records for the
    regions inside one piece, partial histograms -> all-reduce (NCCL) -> GPU walk for the regions
    that straddle a cut, then the one gather of records.  Returns the records in the caller's
    region order (every rank)."""
    import torch
    import torch.distributed as dist
    n_regions = sum(len(plan.whole[r][0]) for r in range(world_size)) + len(plan.cut_regions)
    idx, tid, st, en = plan.arrays("whole", rank)
    local = eng.region_stats(tid, st, en, breadth_n=breadth_n) if len(idx) else np.zeros(0, _capi.REGION_STATS_DTYPE)
    owner = np.full(n_regions, -1, dtype=np.int64)
    for r in range(world_size):
        owner[np.asarray(plan.whole[r][0], np.int64)] = r
    whole_idx = np.nonzero(owner >= 0)[0]
    # gather_region_stats expects each rank's records in the order of np.nonzero(owner == rank)
    order = np.argsort(idx, kind="stable")
    merged_whole = gather_region_stats(np.ascontiguousarray(local[order]), owner[whole_idx], rank, world_size,
                                       device=device, group=group)
    out = np.zeros(n_regions, dtype=_capi.REGION_STATS_DTYPE)
    out[whole_idx] = merged_whole
    n_cut = len(plan.cut_regions)
    if n_cut:
        hist = torch.zeros((n_cut, _capi.HIST_BINS), dtype=torch.int32, device=device)
        ci, ctid, cst, cen = plan.arrays("cut", rank)
        if len(ci):
            # parts of one region held by the same rank (two pieces of one contig cannot be) are distinct rows
            part = torch.zeros((len(ci), _capi.HIST_BINS), dtype=torch.int32, device=device)
            torch.cuda.current_stream(device).synchronize()          # `part` is zeroed (torch's stream) ...
            eng.region_hist_enqueue(ctid, cst, cen, part)
            eng.sync()                                                # ... and filled (the engine's stream)
            hist.index_add_(0, torch.from_numpy(ci).to(device), part)
        if world_size > 1:
            dist.all_reduce(hist, group=group)
        rec = torch.empty(n_cut * 64, dtype=torch.uint8, device=device)
        torch.cuda.current_stream(device).synchronize()
        eng.hist_stats_enqueue(hist, n_cut, rec, breadth_n=breadth_n)
        eng.sync()
        out[np.asarray(plan.cut_regions, np.int64)] = rec.cpu().numpy().view(_capi.REGION_STATS_DTYPE)
    return out
