"""N>1 host logic on CPU: contig-range partition, per-rank read ranges, region ownership and the
single gather of statistics records over a 2-rank gloo group.  The per-rank compute stands in with
the C oracle (no GPU here); on the GPU box bench.py --gpus N runs the same plumbing over NCCL."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_balances_and_covers():
    from metacov_b200 import sharding
    rng = np.random.default_rng(1)
    ln = rng.integers(500, 50_000, 1000)
    rd = (ln * rng.uniform(0.01, 3.0, 1000)).astype(np.int64)
    for n in (1, 2, 3, 8):
        b = sharding.partition_contigs(ln, rd, n)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0) and len(b) == n + 1
        cost = sharding.SLOT_COST * (ln + 1) + sharding.READ_COST * rd
        per = np.array([cost[b[r]:b[r + 1]].sum() for r in range(n)])
        assert per.max() <= per.mean() * 1.15 + cost.max()
    # more ranks than contigs: empty shards are legal
    b = sharding.partition_contigs([10, 10], [1, 1], 4)
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)
    owner = sharding.assign_regions([0, 1, 1, 0], b)
    assert hasattr(sharding, "DeviceGather")
    assert all(b[o] <= t < b[o + 1] for o, t in zip(owner, [0, 1, 1, 0]))
    tid = np.array([0, 0, 2, 2, 2, -1])
    assert sharding.reads_per_contig_from_tid(tid, 3).tolist() == [2, 0, 3]


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    from metacov_b200 import ReadBatch, sharding, synth
    from oracle import cport
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    w = synth.c3(0.0004)                                   # skewed abundance, 200 small contigs
    rpc = np.diff(w.read_start)
    bounds = sharding.partition_contigs(w.contig_len, rpc, world)
    c0, c1 = int(bounds[rank]), int(bounds[rank + 1])
    r0, r1 = sharding.shard_read_range(w.read_start, bounds, rank)
    batch, _ = synth.generate_host(w, i0=r0, n=r1 - r0, tid_base=c0)       # this rank's reads only
    lengths = w.contig_len[c0:c1]
    # regions: every contig plus a few sub-ranges, owned by the rank that holds the contig
    rng = np.random.default_rng(5)
    reg_tid = np.r_[np.arange(w.n_contigs), rng.integers(0, w.n_contigs, 40)]
    reg_start = np.r_[np.zeros(w.n_contigs, np.int64), [rng.integers(0, w.contig_len[t] // 2) for t in reg_tid[w.n_contigs:]]]
    reg_end = np.r_[w.contig_len.astype(np.int64), [w.contig_len[t] for t in reg_tid[w.n_contigs:]]]
    owner = sharding.assign_regions(reg_tid, bounds)
    mine = np.nonzero(owner == rank)[0]
    d, off, _ = cport.depth(batch, lengths, mode="diff") if len(lengths) else (np.zeros(0, np.int32), np.zeros(1, np.int64), {})
    local = cport.region_stats(d, off, lengths, reg_tid[mine] - c0, reg_start[mine], reg_end[mine]) if len(mine) else \
        np.zeros(0, cport.ORC_STATS_DTYPE)
    merged = sharding.gather_region_stats(local, owner, rank, world)
    if rank == 0:
        full, _ = synth.generate_host(w)
        d, off, _ = cport.depth(full, w.contig_len, mode="diff")
        want = cport.region_stats(d, off, w.contig_len, reg_tid, reg_start, reg_end)
        ok = all(np.array_equal(merged[k], want[k]) for k in ("sum", "sumsq", "iq_sum", "min", "max", "med_lo", "med_hi"))
        with open(out_path, "w") as fh:
            fh.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


# ---- contigs cut between ranks: boundary reads on both sides, histograms merged for straddling regions ----

def _stats_from_hist(hist, breadth_n=1):
    """Records from an exact counting histogram -- the arithmetic of hist_walk (k_stats.cuh), in numpy
    (stands in for mcov_hist_stats_enqueue, which needs a GPU)."""
    from oracle import cport
    out = np.zeros(1, cport.ORC_STATS_DTYPE)
    b = np.nonzero(hist)[0]
    cn = hist[b].astype(np.int64)
    n = int(cn.sum())
    if n == 0:
        return out[0]
    csum = np.concatenate(([0], np.cumsum(cn)))
    k1, k2, m1, m2 = n // 4, n - n // 4, (n - 1) // 2, n // 2
    lo, hi = np.maximum(csum[:-1], k1), np.minimum(csum[1:], k2)
    out["sum"] = int((b * cn).sum()); out["sumsq"] = int((b * b * cn).sum())
    out["iq_sum"] = int((np.maximum(hi - lo, 0) * b).sum())
    out["n_ge1"] = int(cn[b >= 1].sum()); out["n_geN"] = int(cn[b >= breadth_n].sum())
    out["min"] = int(b[0]); out["max"] = int(b[-1])
    out["med_lo"] = int(b[np.searchsorted(csum, m1, side="right") - 1])
    out["med_hi"] = int(b[np.searchsorted(csum, m2, side="right") - 1])
    return out[0]


def _split_workload():
    from metacov_b200 import synth
    # three contigs, the big ones several times a rank's share: every cut falls inside a contig
    return synth.Workload("split", [90_000, 7_000, 120_000], [18_000, 1_400, 24_000], seed=77)


def _split_regions(w, shares):
    """Whole contigs, windows straddling every cut, a window inside one piece, one past the contig end,
    an empty one."""
    t, s, e = [], [], []
    for c in range(w.n_contigs):
        t.append(c); s.append(0); e.append(int(w.contig_len[c]))
    for pieces in shares:
        for pc in pieces:
            if pc.p0 > 0:
                t += [pc.tid, pc.tid]; s += [pc.p0 - 700, pc.p0 - 1]; e += [pc.p0 + 900, pc.p0 + 1]
                t.append(pc.tid); s.append(pc.p0); e.append(min(pc.p0 + 500, int(w.contig_len[pc.tid])))
    t += [2, 0, 1]; s += [119_000, 10, 6_900]; e += [121_500, 10, 7_300]
    return np.asarray(t, np.int32), np.asarray(s, np.int32), np.asarray(e, np.int32)


def _rank_records(w, full, shares, plan, rank):
    """One rank's local compute with the C oracle standing in for the GPU engine: depth of its pieces
    from the localized reads, records of whole regions, partial histograms of cut regions."""
    from metacov_b200 import sharding
    from oracle import cport
    pieces = shares[rank]
    lengths = np.asarray([pc.p1 - pc.p0 for pc in pieces], np.int32)
    local = sharding.localize_reads(full, w.read_start, pieces, w.contig_len, reach=160)
    d, off, _ = cport.depth(local, lengths, mode="diff") if len(pieces) else (np.zeros(0, np.int32), np.zeros(1, np.int64), {})
    idx, tid, st, en = plan.arrays("whole", rank)
    rec = cport.region_stats(d, off, lengths, tid, st, en) if len(idx) else np.zeros(0, cport.ORC_STATS_DTYPE)
    hist = np.zeros((len(plan.cut_regions), 8192), np.int64)
    ci, ctid, cst, cen = plan.arrays("cut", rank)
    for k in range(len(ci)):
        ln = int(lengths[ctid[k]])
        a, b = min(int(cst[k]), ln), min(int(cen[k]), ln)
        hist[ci[k]] += np.bincount(d[off[ctid[k]] + a:off[ctid[k]] + b], minlength=8192)
        hist[ci[k], 0] += (int(cen[k]) - int(cst[k])) - (b - a)          # beyond the contig end: depth 0
    return d, off, lengths, idx, rec, hist


def test_split_contigs_depth_and_regions_in_process():
    """3 shares of 3 contigs (cuts inside contigs 0 and 2): every piece's depth equals the slice of the
    unsplit depth bit for bit; records of all regions -- including those straddling a cut -- equal
    the single-process records."""
    from metacov_b200 import sharding, synth
    from oracle import cport
    w = _split_workload()
    full, _ = synth.generate_host(w)
    rpc = np.diff(w.read_start)
    shares = sharding.partition_positions(w.contig_len, rpc, 3)
    assert sum(len(p) for p in shares) == 5 and shares[1][0].p0 > 0 and shares[2][0].p0 > 0
    covered = {}
    for pieces in shares:
        for pc in pieces:
            covered.setdefault(pc.tid, []).append((pc.p0, pc.p1))
    for c, iv in covered.items():
        iv.sort()
        assert iv[0][0] == 0 and iv[-1][1] == w.contig_len[c] and all(iv[k][1] == iv[k + 1][0] for k in range(len(iv) - 1))
    reg = _split_regions(w, shares)
    plan = sharding.split_regions(*reg, shares, w.contig_len)
    assert len(plan.cut_regions) >= 5                                  # 2 cuts x 2 windows + whole contigs 0 and 2
    want_d, want_off, _ = cport.depth(full, w.contig_len, mode="diff")
    want = cport.region_stats(want_d, want_off, w.contig_len, *reg)
    got = np.zeros(len(reg[0]), cport.ORC_STATS_DTYPE)
    hist = np.zeros((len(plan.cut_regions), 8192), np.int64)
    for r in range(3):
        d, off, lengths, idx, rec, h = _rank_records(w, full, shares, plan, r)
        for k, pc in enumerate(shares[r]):
            assert np.array_equal(d[off[k]:off[k] + lengths[k]], want_d[want_off[pc.tid] + pc.p0:want_off[pc.tid] + pc.p1])
            assert d[off[k] + lengths[k]] == 0                         # the sentinel slot closes every piece at 0
        got[idx] = rec
        hist += h
    for k, i in enumerate(plan.cut_regions):
        got[i] = _stats_from_hist(hist[k])
    for key in ("sum", "sumsq", "iq_sum", "n_ge1", "min", "max", "med_lo", "med_hi"):
        assert np.array_equal(got[key], want[key]), key
    # a reach that is too generous only costs reads; None means "everything before the cut"
    loc_all = sharding.localize_reads(full, w.read_start, shares[1], w.contig_len, reach=None)
    loc_160 = sharding.localize_reads(full, w.read_start, shares[1], w.contig_len, reach=160)
    assert len(loc_all.tid) > len(loc_160.tid)
    l1 = np.asarray([pc.p1 - pc.p0 for pc in shares[1]], np.int32)
    assert np.array_equal(cport.depth(loc_all, l1)[0], cport.depth(loc_160, l1)[0])


def _split_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    from metacov_b200 import sharding, synth
    from oracle import cport
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    w = _split_workload()
    full, _ = synth.generate_host(w)
    shares = sharding.partition_positions(w.contig_len, np.diff(w.read_start), world)
    reg = _split_regions(w, shares)
    plan = sharding.split_regions(*reg, shares, w.contig_len)
    d, off, lengths, idx, rec, hist = _rank_records(w, full, shares, plan, rank)
    # the exchange of the real path: one all-reduce of the cut regions' histograms, one gather of records
    ht = torch.from_numpy(hist)
    dist.all_reduce(ht)
    n_regions = len(reg[0])
    owner = np.full(n_regions, -1, np.int64)
    for r in range(world):
        owner[np.asarray(plan.whole[r][0], np.int64)] = r
    whole_idx = np.nonzero(owner >= 0)[0]
    merged = sharding.gather_region_stats(np.ascontiguousarray(rec[np.argsort(idx, kind="stable")]), owner[whole_idx], rank, world)
    got = np.zeros(n_regions, cport.ORC_STATS_DTYPE)
    got[whole_idx] = merged
    for k, i in enumerate(plan.cut_regions):
        got[i] = _stats_from_hist(ht[k].numpy())
    if rank == 0:
        want_d, want_off, _ = cport.depth(full, w.contig_len, mode="diff")
        want = cport.region_stats(want_d, want_off, w.contig_len, *reg)
        ok = len(plan.cut_regions) > 0 and all(np.array_equal(got[k], want[k]) for k in
                                               ("sum", "sumsq", "iq_sum", "n_ge1", "min", "max", "med_lo", "med_hi"))
        with open(out_path, "w") as fh:
            fh.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_split_contig_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_split_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"
