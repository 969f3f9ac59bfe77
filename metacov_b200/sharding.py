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


def gather_region_stats_device(local_dev, out_dev, owner, world_size, group=None):
    """Device-resident variant: ``local_dev`` (uint8 CUDA tensor, cap*64 bytes, this rank's records
    written by ``CoverageEngine.region_stats_enqueue``) is all-gathered into ``out_dev``
    (world*cap*64 bytes) over NCCL with no host round trip; one D2H copy then delivers all
    records in region order."""
    import torch.distributed as dist
    rec = _capi.REGION_STATS_DTYPE.itemsize
    if world_size > 1:
        dist.all_gather_into_tensor(out_dev, local_dev, group=group)
    else:
        out_dev.copy_(local_dev)
    cap = local_dev.numel() // rec
    flat = out_dev.cpu().numpy().reshape(world_size, cap * rec)
    owner = np.asarray(owner)
    merged = np.zeros(len(owner), dtype=_capi.REGION_STATS_DTYPE)
    for r in range(world_size):
        idx = np.nonzero(owner == r)[0]
        if len(idx):
            merged[idx] = flat[r, :len(idx) * rec].view(_capi.REGION_STATS_DTYPE)
    return merged
