"""Host-side packers of the narrow transports (CPU): what the device rebuilds must be what was packed.
The device side (mcov_depth_sorted_packed / _delta) is covered by tests/test_gpu_parity.py and
tests/test_gpu_fuzz.py."""
import numpy as np
import pytest

from helpers import load_soa


def _rebuild_delta(p):
    """The arithmetic of k_delta_seed / k_delta_patch / prefix sum / k_delta_finish, in numpy."""
    n = p["n"]
    d = p["dpos"].astype(np.int64)
    d[p["exc_index"]] = p["exc_delta"]
    S = np.cumsum(d).astype(np.int64) & 0xFFFFFFFF                 # the device sums in 32 bits, wrap-around included
    crs = p["contig_read_start"]
    pos = np.empty(n, np.int64)
    for a, b in list(zip(crs[:-1], crs[1:])) + [(crs[-1], n)]:
        if b > a:
            pos[a:b] = (S[a:b] - (S[a - 1] if a > 0 else 0)) & 0xFFFFFFFF
    return pos.astype(np.uint32).view(np.int32)


def test_delta_packer_round_trip():
    from metacov_b200 import ReadBatch, synth
    from metacov_b200.engine import pack_batch, pack_batch_delta, packed_bytes
    w = synth.c2(0.01)
    b, _ = synth.generate_host(w)
    p = pack_batch_delta(b, w.n_contigs)
    assert np.array_equal(_rebuild_delta(p), b.pos) and len(p["exc_index"]) == 0
    assert np.array_equal(np.cumsum(p["n_cigar"], dtype=np.int64), b.cig_off[1:].astype(np.int64))
    assert np.array_equal(p["cig"].astype(np.uint32), b.cig)
    assert packed_bytes(p) / len(b.tid) < 7.5 < 12.0 < packed_bytes(pack_batch(b, w.n_contigs)) / len(b.tid)
    # fixture: unplaced reads at the end form one more segment
    z, fb = load_soa("fixture_soa.npz")
    pf = pack_batch_delta(fb, 2)
    assert np.array_equal(_rebuild_delta(pf), fb.pos) and pf["contig_read_start"][-1] < len(fb.tid)
    # gaps beyond 16 bits, first reads far from 0, negative differences (unsorted), negative positions
    rng = np.random.default_rng(4)
    for trial in range(20):
        n_contigs = int(rng.integers(1, 6))
        counts = rng.integers(0, 60, n_contigs)
        tid = np.repeat(np.arange(n_contigs, dtype=np.int32), counts)
        pos = rng.integers(-1000, 2_000_000_000, len(tid)).astype(np.int32)
        if trial % 2 == 0:
            pos = np.concatenate([np.sort(pos[tid == c]) for c in range(n_contigs)]).astype(np.int32) if len(tid) else pos
        n = len(tid)
        rb = ReadBatch(tid, pos, np.zeros(n, np.uint16), np.zeros(n, np.uint8), np.arange(n + 1, dtype=np.uint32),
                       np.full(n, 100 << 4, np.uint32))
        pk = pack_batch_delta(rb, n_contigs)
        assert np.array_equal(_rebuild_delta(pk), pos), trial
        if n:
            assert pk["dpos"][pk["exc_index"]].max(initial=0) == 0          # placeholders of the exceptions
    # what does not qualify
    n = 4
    long_op = ReadBatch(np.zeros(n, np.int32), np.arange(n, dtype=np.int32), np.zeros(n, np.uint16), np.zeros(n, np.uint8),
                        np.arange(n + 1, dtype=np.uint32), np.full(n, 5000 << 4, np.uint32))
    with pytest.raises(ValueError):
        pack_batch_delta(long_op, 1)
    many_ops = ReadBatch(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.uint16), np.zeros(1, np.uint8),
                         np.array([0, 300], np.uint32), np.full(300, 1 << 4, np.uint32))
    with pytest.raises(ValueError):
        pack_batch_delta(many_ops, 1)
    ungrouped = ReadBatch(np.array([1, 0], np.int32), np.zeros(2, np.int32), np.zeros(2, np.uint16), np.zeros(2, np.uint8),
                          np.arange(3, dtype=np.uint32), np.full(2, 1 << 4, np.uint32))
    with pytest.raises(ValueError):
        pack_batch_delta(ungrouped, 2)


def test_partition_positions_covers_every_base_once():
    from metacov_b200 import sharding
    rng = np.random.default_rng(12)
    for trial in range(40):
        c = int(rng.integers(1, 30))
        ln = rng.integers(1, 400_000, c)
        rd = (ln * rng.uniform(0.0, 0.5, c)).astype(np.int64)
        for n in (1, 2, 3, 8):
            shares = sharding.partition_positions(ln, rd, n)
            assert len(shares) == n
            seen = {}
            last = (-1, 0)
            for pieces in shares:
                for pc in pieces:
                    assert 0 <= pc.p0 < pc.p1 <= ln[pc.tid]
                    assert (pc.tid, pc.p0) >= last                      # contiguous, in order
                    last = (pc.tid, pc.p1)
                    seen.setdefault(pc.tid, []).append((pc.p0, pc.p1))
            for t in range(c):
                iv = sorted(seen.get(t, []))
                assert iv and iv[0][0] == 0 and iv[-1][1] == ln[t] and all(iv[k][1] == iv[k + 1][0] for k in range(len(iv) - 1)), (trial, n, t)
            # every region lands somewhere, whole or cut
            rt = rng.integers(0, c, 10)
            rs = np.array([rng.integers(0, ln[t]) for t in rt])
            re = np.array([s + rng.integers(0, ln[t] + 50) for s, t in zip(rs, rt)])
            plan = sharding.split_regions(rt, rs, re, shares, ln)
            n_whole = sum(len(plan.whole[r][0]) for r in range(n))
            assert n_whole + len(plan.cut_regions) == 10
            for r in range(n):
                ci, ctid, cst, cen = plan.arrays("cut", r)
                assert np.all(cen > cst)


def _unpack_block(buf):
    """The transport block (include/metacov_b200.h: mcov_block_hdr, version 4) decoded in numpy: what k_block_index /
    k_block_reduce / k_block_prefix / k_block_expand rebuild on the device."""
    import struct
    raw = np.asarray(buf, dtype=np.uint8)
    (magic, version, n, n_carry, n_cigar, n_exc, n_esc, n_xops, total, n_contigs, n_jt, n_dict, n_dictops, has_mapq, xop_bytes,
     last_tid, last_pos, o_crs, o_dpos, o_ei, o_ev, o_fc, o_jt, o_qi, o_qf, o_qc, o_doff, o_dops, o_xops, o_mapq, nib,
     n_dq, n_fq, o_nb, o_dq, o_fq, o_chunk) = struct.unpack_from("<IIqqqqqqqiiiiiiiiIIIIIIIIIIIIIIqqIIII", raw.tobytes()[:200])
    assert magic == 0x4256434D and version == 4 and total <= len(raw) and xop_bytes in (2, 4) and nib in (0, 1)
    view = lambda off, cnt, dt: raw[off:off + cnt * np.dtype(dt).itemsize].view(dt)
    crs = view(o_crs, n_contigs + 1, np.int64)
    if nib:
        nbv = view(o_nb, n, np.uint8)
        d, fc = (nbv & 15).astype(np.int64), (nbv >> 4).astype(np.int64)
        dq, fq = view(o_dq, n_dq, np.uint8), view(o_fq, n_fq, np.uint8)
        assert int((d == 15).sum()) == n_dq and int((fc == 15).sum()) == n_fq and np.all(dq >= 15) and np.all(fq >= 15)
        d[d == 15] = dq
        fc[fc == 15] = fq
        fc = fc.astype(np.uint8)
    else:
        d = view(o_dpos, n, np.uint8).astype(np.int64)
        fc = view(o_fc, n, np.uint8)
    assert np.all(np.diff(view(o_ei, n_exc, np.uint32).astype(np.int64)) > 0)            # ascending: the device relies on it
    d[view(o_ei, n_exc, np.uint32)] = view(o_ev, n_exc, np.int32)
    S = np.cumsum(d)
    pos = np.empty(n, np.int64)
    tid = np.full(n, -1, np.int32)
    for c, (a, b) in enumerate(list(zip(crs[:-1], crs[1:])) + [(crs[-1], n)]):
        if b > a:
            pos[a:b] = S[a:b] - (S[a - 1] if a > 0 else 0)
            tid[a:b] = c if c < n_contigs else -1
    jt = np.concatenate([view(o_jt, n_jt, np.uint32), np.zeros(256 - n_jt, np.uint32)])
    assert np.all((fc < n_jt) | (fc == 255))
    e = jt[fc]
    flag, cc = (e >> 8).astype(np.uint16), (e & 255).astype(np.int64)
    qi = view(o_qi, n_esc, np.uint32)
    assert np.array_equal(qi, np.nonzero(fc == 255)[0])                                  # ascending
    flag[qi] = view(o_qf, n_esc, np.uint16)
    cc[qi] = view(o_qc, n_esc, np.uint8)
    doff = view(o_doff, n_dict + 1, np.uint32).astype(np.int64)
    dops = view(o_dops, n_dictops, np.uint32)
    xops = view(o_xops, n_xops, np.uint16 if xop_bytes == 2 else np.uint32).astype(np.uint32)
    cig, ncig, x = [], np.zeros(n, np.int64), 0
    for i in range(n):
        if cc[i] < 128:
            cig.append(dops[doff[cc[i]]:doff[cc[i] + 1]])
        else:
            cig.append(xops[x:x + cc[i] - 128]); x += cc[i] - 128
        ncig[i] = len(cig[-1])
    assert x == n_xops and int(ncig.sum()) == n_cigar
    mapq = view(o_mapq, n, np.uint8) if has_mapq else None
    # the chunk table (one row of eight u32 per 2 048 reads): where the chunk's entries of the side lists, the explicit ops, the
    # escapes and the exceptions begin; op offset of its first read; position of the read in front of it
    ct = view(o_chunk, 8 * ((n + 2047) // 2048), np.uint32).reshape(-1, 8)
    starts = np.arange(0, n, 2048)
    if nib:
        assert np.array_equal(ct[:, 0], np.concatenate(([0], np.cumsum(nbv & 15 == 15)))[starts])
        assert np.array_equal(ct[:, 1], np.concatenate(([0], np.cumsum(nbv >> 4 == 15)))[starts])
    assert np.array_equal(ct[:, 2], np.concatenate(([0], np.cumsum(ncig)))[starts])
    assert np.array_equal(ct[:, 3], np.concatenate(([0], np.cumsum(np.where(cc >= 128, cc - 128, 0))))[starts])
    assert np.array_equal(ct[1:, 4].view(np.int32), pos[starts[1:] - 1].astype(np.int32)) and (len(ct) == 0 or ct[0, 4] == 0)
    for col, lst in ((5, qi), (6, view(o_ei, n_exc, np.uint32))):
        first = np.full(len(ct), 0xFFFFFFFF, np.uint32)
        for k in range(len(lst) - 1, -1, -1):
            first[lst[k] // 2048] = k
        assert np.array_equal(ct[:, col], first)
    return dict(n=n, n_carry=n_carry, tid=tid, pos=pos.astype(np.int64), flag=flag, ncig=ncig, mapq=mapq,
                cig=np.concatenate(cig) if cig else np.zeros(0, np.uint32), last=(last_tid, last_pos), n_exc=n_exc, n_esc=n_esc,
                xop_bytes=xop_bytes, nib=nib)


def test_block_packer_round_trip():
    """mcov_pack_block (native): the block decodes to the columns it was packed from -- short reads, the fixture with
    its unplaced tail, gaps beyond 8 bits (exceptions), negative differences, thousands of distinct (flag, CIGAR) pairs (escapes)."""
    from metacov_b200 import ReadBatch, synth
    from metacov_b200.engine import pack_block

    def check(b, n_contigs, **kw):
        buf, nb = pack_block(b, n_contigs, **kw)
        u = _unpack_block(buf[:nb])
        valid = (b.tid >= 0) & (b.tid < n_contigs)
        assert np.array_equal(u["tid"][valid], b.tid[valid]) and np.all(u["tid"][~valid] == -1)
        assert np.array_equal(u["pos"], b.pos.astype(np.int64)) and np.array_equal(u["flag"], b.flag)
        assert np.array_equal(u["ncig"], np.diff(b.cig_off.astype(np.int64))) and np.array_equal(u["cig"], b.cig)
        if kw.get("with_mapq"):
            assert np.array_equal(u["mapq"], b.mapq)
        if len(b.tid):
            assert u["last"] == (int(b.tid[-1]), int(b.pos[-1]))
        return u, nb

    w = synth.c2(0.01)
    b, _ = synth.generate_host(w)
    u, nb = check(b, w.n_contigs, with_mapq=False, n_carry=17)
    assert nb / len(b.tid) < 1.6 and u["nib"] == 1 and u["xop_bytes"] == 2                 # C2: the nibble form, u16 explicit ops
    assert u["n_carry"] == 17 and u["n_exc"] == 0 and u["n_esc"] < 0.02 * len(b.tid)
    check(b, w.n_contigs, with_mapq=True, threads=3)
    z, fb = load_soa("fixture_soa.npz")
    check(fb, 2, with_mapq=True)
    rng = np.random.default_rng(4)
    for trial in range(12):
        nc = int(rng.integers(1, 6))
        n = int(rng.integers(0, 3000))
        tid = np.sort(rng.integers(0, nc, n)).astype(np.int32)
        if n > 10 and trial % 3 == 0:
            tid[-5:] = -1
        pos = rng.integers(-100, 5_000_000 if trial % 2 else 3000, n).astype(np.int32)
        if trial % 4:
            for c in range(-1, nc):
                m = tid == c
                pos[m] = np.sort(pos[m])
        flag = rng.integers(0, 4096 if trial % 5 else 65536, n).astype(np.uint16) if trial % 2 else rng.choice(np.array([99, 147, 83, 163, 4], np.uint16), n)
        n_op = rng.integers(0, 7, n)
        off = np.concatenate(([0], np.cumsum(n_op))).astype(np.uint32)
        cig = ((rng.integers(1, 200, int(off[-1])).astype(np.uint32) << 4) | rng.integers(0, 9, int(off[-1])).astype(np.uint32))
        rb = ReadBatch(tid, pos, flag, rng.integers(0, 61, n).astype(np.uint8), off, cig)
        u, _ = check(rb, nc, with_mapq=True)
        if n > 600 and trial % 2:
            assert u["n_esc"] > 0                                # thousands of distinct flags: most pairs are escapes
    # what does not qualify
    w5 = synth.c5(0.0005)
    b5, _ = synth.generate_host(w5)
    with pytest.raises(ValueError):
        pack_block(b5, w5.n_contigs)
    with pytest.raises(ValueError):
        pack_block(ReadBatch(b.tid[::-1].copy(), b.pos, b.flag, b.mapq, b.cig_off, b.cig), w.n_contigs)   # not grouped by contig


def test_narrow_packers_refuse_64bit_offsets():
    """pack_block / pack_batch / pack_batch_delta take 64-bit offset arrays only when they fit 32 bits: a batch of 2^32 or
    more ops (config C5 at full size) is refused instead of truncated (it travels through the *_wide entry points)."""
    from metacov_b200 import ReadBatch
    from metacov_b200.engine import pack_batch, pack_batch_delta, pack_block
    n = 4
    cols = (np.zeros(n, np.int32), np.arange(n, dtype=np.int32), np.zeros(n, np.uint16), np.zeros(n, np.uint8))
    big = ReadBatch(*cols, np.array([0, 1, 2, 3, 1 << 33], np.int64), np.full(4, 100 << 4, np.uint32))
    for f in (pack_block, pack_batch, pack_batch_delta):
        with pytest.raises(ValueError):
            f(big, 1)
    ok = ReadBatch(*cols, np.array([0, 1, 2, 3, 4], np.int64), np.full(4, 100 << 4, np.uint32))
    assert pack_batch(ok, 1)["n_cig_total"] == 4 and pack_block(ok, 1)[1] > 0 and pack_batch_delta(ok, 1)["n_cig_total"] == 4
