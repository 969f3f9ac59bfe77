"""Randomised parity: many small read sets with adversarial shapes -- contigs of 1 bp to a few tiles, reads
hanging over both contig ends, negative positions, empty CIGARs, zero-length and non-reference ops, every
flag bit, spans longer than a tile (the far-read tables), unplaced reads, duplicates of a position by the
thousand, random filters, random (overlapping, overhanging, empty) regions -- against the C oracle, on both
depth formulations, the packed host transport, the run-length export and the insert-size histogram.  Seeds
are fixed: failures reproduce."""
import os

import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu

KEYS = ("sum", "sumsq", "iq_sum", "n_ge1", "min", "max", "med_lo", "med_hi")


def _case(seed):
    from metacov_b200 import ReadBatch
    rng = np.random.Generator(np.random.PCG64(seed))
    n_contigs = int(rng.integers(1, 12))
    kind = seed % 4
    if kind == 0:
        lengths = rng.integers(1, 40, n_contigs)                         # tiny contigs
    elif kind == 1:
        lengths = rng.integers(1, 7000, n_contigs)                       # around the 2048-slot tile size
    elif kind == 2:
        lengths = np.r_[rng.integers(4000, 30000, 1), rng.integers(1, 300, n_contigs - 1)]
    else:
        lengths = rng.integers(2040, 2060, n_contigs)                    # tile borders near contig borders
    lengths = lengths.astype(np.int32)
    n = int(rng.integers(0, 4000))
    tid = np.sort(rng.integers(0, n_contigs, n)).astype(np.int32)
    if n and rng.random() < 0.5:                                          # unplaced reads sort last
        k = int(rng.integers(0, min(n, 20) + 1))
        if k:
            tid[n - k:] = -1
    pos = np.zeros(n, np.int32)
    for c in range(-1, n_contigs):
        m = tid == c
        cnt = int(m.sum())
        if not cnt:
            continue
        ln = 100 if c < 0 else int(lengths[c])
        p = rng.integers(-60, ln + 60, cnt) if rng.random() < 0.7 else np.full(cnt, rng.integers(0, ln))   # pile-ups
        pos[m] = np.sort(p)
    flag = rng.choice(np.array([0, 16, 99, 147, 83, 163, 97, 145, 73, 4, 69, 256, 1024, 512, 2048, 1, 3], np.uint16), n)
    flag = (flag | (rng.integers(0, 4096, n).astype(np.uint16) * (rng.random(n) < 0.1))).astype(np.uint16)
    mapq = rng.integers(0, 61, n).astype(np.uint8)
    n_op = rng.integers(0, 9, n)
    n_op[rng.random(n) < 0.5] = 1
    cig_off = np.concatenate(([0], np.cumsum(n_op))).astype(np.uint32)
    tot = int(cig_off[-1])
    ops = rng.integers(0, 9, tot).astype(np.uint32)                      # M I D N S H P = X
    lens = rng.integers(0, 200, tot).astype(np.uint32)
    far = rng.random(tot) < (0.02 if kind in (1, 2) else 0.0)
    lens[far] = rng.integers(2048, 9000, int(far.sum()))                 # spans beyond a tile: far reads
    cig = (lens << 4 | ops).astype(np.uint32)
    isize = rng.integers(-700, 700, n).astype(np.int32)
    # regions: whole contigs, random windows, overhanging, empty
    rt = rng.integers(0, n_contigs, 12).astype(np.int32)
    rs = np.array([rng.integers(0, lengths[t] + 1) for t in rt], np.int32)
    re = np.array([s + rng.integers(0, lengths[t] + 40) for s, t in zip(rs, rt)], np.int32)
    rt = np.r_[np.arange(n_contigs, dtype=np.int32), rt]
    rs = np.r_[np.zeros(n_contigs, np.int32), rs]
    re = np.r_[lengths, re]
    filt = {}
    if seed % 3 == 1:
        filt = dict(flag_filter=int(rng.choice([0, 0x704, 0x4, 0xF04])), min_mapq=int(rng.integers(0, 40)),
                    ignore_orphans=int(rng.integers(0, 2)), flag_require=int(rng.choice([0, 0, 16, 64])))
    return ReadBatch(tid, pos, flag, mapq, cig_off, cig), isize, lengths, (rt, rs, re), filt


# MCOV_FUZZ_BLOCKS=60 runs 600 cases instead of 60 (a one-off of the builder: profiles/r02_fuzz_600.txt)
@pytest.mark.parametrize("block", range(int(os.environ.get("MCOV_FUZZ_BLOCKS", "6"))))
def test_random_cases_match_oracle(block):
    from metacov_b200 import CoverageEngine, McovError, _capi
    from metacov_b200.engine import pack_batch, pack_batch_delta, pack_block
    for seed in range(block * 10, block * 10 + 10):
        b, isize, lengths, (rt, rs, re), filt = _case(1000 + seed)
        of = cport.default_filter(**filt) if filt else None
        d, off, info = cport.depth(b, lengths, filt=of, mode="diff")
        want = cport.region_stats(d, off, lengths, rt, rs, re)
        nonempty = re > rs
        with CoverageEngine(lengths, filt=filt or None) as eng:
            for path in ("auto", "push", "packed", "delta", "block"):
                if path == "auto":
                    eng.compute_depth(b)
                elif path == "push":
                    eng.begin()
                    h = len(b.tid) // 2
                    o = int(b.cig_off[h])
                    from metacov_b200 import ReadBatch
                    eng.push(ReadBatch(b.tid[:h], b.pos[:h], b.flag[:h], b.mapq[:h], b.cig_off[:h + 1], b.cig[:o]))
                    eng.push(ReadBatch(b.tid[h:], b.pos[h:], b.flag[h:], b.mapq[h:], (b.cig_off[h:] - o).astype(np.uint32), b.cig[o:]))
                    eng.finalize()
                else:
                    try:
                        if path == "packed":
                            eng.depth_sorted_packed(pack_batch(b, len(lengths), with_mapq=True))
                        elif path == "block":
                            blk = pack_block(b, len(lengths), with_mapq=True)
                            u = eng.block_unpack(blk)                      # the rebuilt columns themselves
                            ok = (b.tid >= 0) & (b.tid < len(lengths))
                            assert np.array_equal(u["tid"][ok], b.tid[ok]) and np.all(u["tid"][~ok] == -1), seed
                            assert np.array_equal(u["pos"], b.pos) and np.array_equal(u["flag"], b.flag) and np.array_equal(u["mapq"], b.mapq), seed
                            assert np.array_equal(u["cig_off"], b.cig_off) and np.array_equal(u["cig"], b.cig), seed
                            eng.depth_sorted_block(blk)
                        else:
                            eng.depth_sorted_delta(pack_batch_delta(b, len(lengths), with_mapq=True))
                    except ValueError:
                        continue                                          # the batch does not qualify for this transport
                    except McovError as e:
                        assert e.code in (_capi.MCOV_ERR_UNSORTED, _capi.MCOV_ERR_RANGE), (seed, path, e)
                        continue                                          # (the auto path took the push formulation for it)
                pi = eng.pass_info()
                assert pi["n_pass"] == info["n_pass"] and pi["aligned_bases"] == info["aligned_bases"], (seed, path)
                for c in range(len(lengths)):
                    assert np.array_equal(eng.copy_depth(c), d[off[c]:off[c] + lengths[c]]), (seed, path, c)
                st = eng.region_stats(rt, rs, re)
                for k in KEYS:
                    assert np.array_equal(st[k][nonempty], want[k][nonempty]), (seed, path, k)
                assert np.all(st["sum"][~nonempty] == 0)
            runs = eng.depth_runs()
            total = int(((runs["end"] - runs["start"]).astype(np.int64) * runs["depth"]).sum())
            assert total == int(sum(int(d[off[c]:off[c] + lengths[c]].astype(np.int64).sum()) for c in range(len(lengths)))), seed
            assert np.all(runs["end"] > runs["start"])
            hist, cnt, mx = eng.isize_hist(b.flag, isize, (16, 64), n_bins=1024)
            rh, rc, rmx = cport.isize_hist(b.flag, isize, (16, 64), hist.shape[1])
            assert np.array_equal(hist, rh) and np.array_equal(cnt, rc) and mx == rmx, seed
