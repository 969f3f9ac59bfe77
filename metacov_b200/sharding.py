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
        self.local_dev, self.out_dev, self.out_host = self.local[0], self.out[0], self.host[0]
        self.merged = np.zeros(self.n_regions, dtype=_capi.REGION_STATS_DTYPE)

    def submit(self, slot=0, want_host=True):
        """Enqueue (current stream) the all-gather of ``local[slot]`` and, with want_host, the copy of
        all records to pinned memory; returns without waiting."""
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.all_gather_into_tensor(self.out[slot], self.local[slot], group=self.group)
        else:
            self.out[slot].copy_(self.local[slot])
        if want_host:
            self.host[slot].copy_(self.out[slot], non_blocking=True)
        self.done[slot].record(torch.cuda.current_stream())

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
