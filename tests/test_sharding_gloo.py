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
