"""Contigs cut between ranks on the GPU: every piece's depth equals the slice of
the unsplit depth bit for bit (boundary reads given to both sides, clipped by the kernels), and
regions that straddle a cut are finished from merged counting histograms
(mcov_region_hist_enqueue -> sum -> mcov_hist_stats_enqueue) with records identical to the unsplit
engine's.  The N>1 plumbing (all-reduce, gather) is covered on CPU by tests/test_sharding_gloo.py;
here the ranks are emulated one after the other on cuda:0."""
import numpy as np
import pytest

from test_sharding_gloo import _split_regions, _split_workload

pytestmark = pytest.mark.gpu

KEYS = ("sum", "sumsq", "iq_sum", "n_ge1", "n_geN", "min", "max", "med_lo", "med_hi", "flags")


@pytest.mark.parametrize("resident", [False, True])
def test_split_contigs_match_unsplit_engine(resident):
    import torch
    from metacov_b200 import CoverageEngine, ReadBatch, _capi, sharding, synth
    w = _split_workload()
    if resident:
        full, _ = synth.generate_device(w, 0)
    else:
        full, _ = synth.generate_host(w)
    shares = sharding.partition_positions(w.contig_len, np.diff(w.read_start), 3)
    reg = _split_regions(w, shares)
    plan = sharding.split_regions(*reg, shares, w.contig_len)
    assert len(plan.cut_regions) >= 5
    with CoverageEngine(w.contig_len) as ref:
        ref.depth_sorted(full)
        want = ref.region_stats(*reg).copy()
        want_depth = [ref.copy_depth(c) for c in range(w.n_contigs)]
    got = np.zeros(len(reg[0]), _capi.REGION_STATS_DTYPE)
    hist = torch.zeros((len(plan.cut_regions), _capi.HIST_BINS), dtype=torch.int32, device="cuda")
    for r, pieces in enumerate(shares):
        lengths = np.asarray([pc.p1 - pc.p0 for pc in pieces], np.int32)
        local = sharding.localize_reads(full, w.read_start, pieces, w.contig_len, reach=160)
        assert (len(local.tid) > 0) and (r == 0 or int(local.pos.min()) < 0)      # boundary reads travel with the piece
        with CoverageEngine(lengths) as eng:
            eng.depth_sorted(local)
            for k, pc in enumerate(pieces):
                assert np.array_equal(eng.copy_depth(k), want_depth[pc.tid][pc.p0:pc.p1])
            idx, tid, st, en = plan.arrays("whole", r)
            if len(idx):
                got[idx] = eng.region_stats(tid, st, en)
            ci, ctid, cst, cen = plan.arrays("cut", r)
            if len(ci):
                part = torch.zeros((len(ci), _capi.HIST_BINS), dtype=torch.int32, device="cuda")
                torch.cuda.synchronize()
                eng.region_hist_enqueue(ctid, cst, cen, part)
                eng.sync()
                hist.index_add_(0, torch.from_numpy(ci).cuda(), part)       # stands in for the all-reduce
            if r == len(shares) - 1:
                rec = torch.empty(len(plan.cut_regions) * 64, dtype=torch.uint8, device="cuda")
                torch.cuda.synchronize()
                eng.hist_stats_enqueue(hist, len(plan.cut_regions), rec)
                eng.sync()
                got[np.asarray(plan.cut_regions)] = rec.cpu().numpy().view(_capi.REGION_STATS_DTYPE)
    for key in KEYS:
        assert np.array_equal(got[key], want[key]), key


def test_sharded_region_stats_single_rank_two_pieces_of_one_contig():
    """sharding.sharded_region_stats end to end (world_size 1, no process group): one rank holding a
    contig as two pieces -- whole regions, straddling regions, a region past the contig end."""
    import torch
    from metacov_b200 import CoverageEngine, sharding, synth
    w = synth.c2(0.004)                                                    # 4 contigs of 50 kb
    full, _ = synth.generate_host(w)
    P = sharding.Piece
    shares = [[P(0, 0, 50_000), P(1, 0, 20_000), P(1, 20_000, 50_000), P(2, 0, 50_000), P(3, 0, 31_000), P(3, 31_000, 50_000)]]
    t = np.asarray([0, 1, 2, 3, 1, 1, 3, 3, 1], np.int32)
    s = np.asarray([0, 0, 0, 0, 19_000, 100, 30_999, 49_000, 20_000], np.int32)
    e = np.asarray([50_000, 50_000, 50_000, 50_000, 26_000, 900, 31_001, 50_700, 20_000], np.int32)
    plan = sharding.split_regions(t, s, e, shares, w.contig_len)
    assert sorted(plan.cut_regions) == [1, 3, 4, 6]
    with CoverageEngine(w.contig_len) as ref:
        ref.depth_sorted(full)
        want = ref.region_stats(t, s, e).copy()
    lengths = np.asarray([pc.p1 - pc.p0 for pc in shares[0]], np.int32)
    local = sharding.localize_reads(full, w.read_start, shares[0], w.contig_len, reach=200)
    with CoverageEngine(lengths) as eng:
        eng.depth_sorted(local)
        got = sharding.sharded_region_stats(eng, plan, 0, 1, torch.device("cuda", 0))
    for key in KEYS:
        assert np.array_equal(got[key], want[key]), key
