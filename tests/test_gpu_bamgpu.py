"""BAM decode on the GPU (mcov_bam_decode_gpu: BGZF inflate + record chain + SoA columns) against the host
decoder (mcov_bam_open / mcov_bam_load) on the same files: the fixture, short reads over many BGZF
blocks, long reads whose records are larger than a 64 KiB chunk, stored (level 0) blocks, an empty file;
corrupt files fail loudly; depth from the device-resident columns equals the oracle."""
import struct
import zlib

import os

import numpy as np
import pytest

from helpers import load_soa
from oracle import bamio, cport

pytestmark = pytest.mark.gpu

COLS = ("tid", "pos", "flag", "mapq", "l_seq", "isize", "cig_off", "cig")


def _write(tmp, name, refs, lengths, b, level=6, **kw):
    path = str(tmp / name)
    old = bamio._bgzf_block
    if level != 6:
        def blk(payload):
            comp = zlib.compressobj(level, zlib.DEFLATED, -15)
            cdata = comp.compress(payload) + comp.flush()
            head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, len(cdata) + 25)
            return head + cdata + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))
        bamio._bgzf_block = blk
    try:
        bamio.write_bam(path, refs, [int(x) for x in lengths], b.tid, b.pos, b.flag, b.mapq, b.cig_off, b.cig, **kw)
    finally:
        bamio._bgzf_block = old
    return path


def _host_columns(path):
    from metacov_b200 import AlignmentFile
    with AlignmentFile(path) as bam:
        s = bam.soa()                                       # views of the reader's arrays: copy before it closes
        return {c: np.array(s[c]) for c in COLS}, list(bam.references), list(bam.lengths)


def _check_file(path, lengths, batch=None):
    from metacov_b200 import CoverageEngine, bamgpu
    want, refs, lens = _host_columns(path)
    with CoverageEngine(lengths) as eng:
        soa = bamgpu.decode(eng, path)
        assert soa.n_records == len(want["tid"]) and soa.n_cigar == len(want["cig"]) and soa.n_ref == len(refs)
        for c in COLS:
            assert np.array_equal(soa.to_host(c), np.asarray(want[c])), c
        text, hrefs = soa.header()
        assert [r for r, _ in hrefs] == [r.split()[0] for r in refs] and [l for _, l in hrefs] == lens
        if batch is not None:
            bamgpu.depth_sorted(eng, soa)
            d, off, info = cport.depth(batch, lengths, mode="diff")
            assert eng.pass_info()["n_pass"] == info["n_pass"]
            for c in range(len(lengths)):
                assert np.array_equal(eng.copy_depth(c), d[off[c]:off[c] + lengths[c]])
        return soa.n_segments


def test_fixture_and_multi_block_files(tmp_path):
    from metacov_b200 import synth
    z, b = load_soa("fixture_soa.npz")
    so = z["seq_off"]
    seqs = [z["seq"][so[i]:so[i + 1]] for i in range(len(b.tid))]
    p = _write(tmp_path, "fixture.bam", [str(x) for x in z["references"]], z["lengths"], b, isize=z["isize"],
               names=[str(x) for x in z["names"]], seqs=seqs)
    _check_file(p, z["lengths"], b)
    w = synth.c2(0.003)                                      # 30 000 reads: ~100 BGZF blocks / chunks
    hb, isz = synth.generate_host(w)
    refs = ["c%d" % c for c in range(w.n_contigs)]
    for level in (6, 1, 0):                                  # dynamic Huffman, fast, stored blocks
        p = _write(tmp_path, "c2_l%d.bam" % level, refs, w.contig_len, hb, level=level, isize=isz)
        nseg = _check_file(p, w.contig_len, hb)
        assert nseg > 20                                     # the record chain really was walked in parallel


def test_long_records_span_chunks(tmp_path):
    from metacov_b200 import synth
    w = synth.c5(0.0005)                                     # 10-50 kb reads: records of 30-150 KB
    hb, isz = synth.generate_host(w)
    p = _write(tmp_path, "c5.bam", ["c%d" % c for c in range(w.n_contigs)], w.contig_len, hb, isize=isz)
    _check_file(p, w.contig_len, hb)


def test_empty_and_corrupt_files(tmp_path):
    from metacov_b200 import CoverageEngine, McovError, ReadBatch, bamgpu
    z, b = load_soa("fixture_soa.npz")
    empty = ReadBatch(*(np.zeros(0, a.dtype) for a in (b.tid, b.pos, b.flag, b.mapq)), np.zeros(1, np.uint32), np.zeros(0, np.uint32))
    p = _write(tmp_path, "empty.bam", ["ref1", "ref2"], [425, 575], empty)
    with CoverageEngine([425, 575]) as eng:
        soa = bamgpu.decode(eng, p)
        assert soa.n_records == 0 and soa.n_ref == 2
        good = _write(tmp_path, "good.bam", [str(x) for x in z["references"]], z["lengths"], b)
        raw = bytearray(open(good, "rb").read())
        bad = bytes(raw[:200]) + bytes([raw[200] ^ 0x5A]) + bytes(raw[201:])          # one payload byte flipped
        with pytest.raises(McovError):
            bamgpu.decode(eng, bad)
        with pytest.raises(McovError):
            bamgpu.decode(eng, bytes(raw[:len(raw) // 2]))                             # truncated file
        with pytest.raises(McovError):
            bamgpu.decode(eng, b"not a bam file at all" * 10)
        assert bamgpu.decode(eng, bytes(raw)).n_records == len(b.tid)                  # the context survives errors


def test_names_and_seq_on_the_device(tmp_path):
    """mcov_bam_gpu_names_seq against the host reader on reads with soft / hard clips, both strands, odd lengths,
    ambiguity codes, empty SEQ and names of every length; then `experimental` and the k-mer histogram of a
    GPU-decoded file equal the host-decoded file's without a host reader being opened."""
    from metacov_b200 import AlignmentFile, ReadBatch, pileup
    from metacov_b200 import scan as mscan
    rng = np.random.default_rng(77)
    n = 6000
    lengths = np.array([40000, 30000], np.int32)
    tid = np.sort(rng.integers(0, 2, n)).astype(np.int32)
    pos = np.zeros(n, np.int32)
    for c in range(2):
        m = tid == c
        pos[m] = np.sort(rng.integers(0, lengths[c] - 400, m.sum()))
    flag = (rng.integers(0, 2, n) * 0x10 + rng.integers(0, 2, n) * 0x1 + rng.integers(0, 2, n) * 0x2 +
            rng.choice([0x40, 0x80], n)).astype(np.uint16)
    cig, off, seqs, names = [], [0], [], []
    for i in range(n):
        ops = []
        if rng.random() < 0.2: ops.append((5, int(rng.integers(1, 9))))
        if rng.random() < 0.4: ops.append((4, int(rng.integers(1, 30))))
        ops.append((0, int(rng.integers(1, 120))))
        if rng.random() < 0.3: ops += [(1, int(rng.integers(1, 5))), (0, int(rng.integers(1, 60)))]
        if rng.random() < 0.3: ops += [(2, int(rng.integers(1, 5))), (0, int(rng.integers(1, 60)))]
        if rng.random() < 0.4: ops.append((4, int(rng.integers(1, 30))))
        if rng.random() < 0.2: ops.append((5, int(rng.integers(1, 9))))
        qlen = sum(l for o, l in ops if o in (0, 1, 4))
        if i % 97 == 0:
            qlen = 0                                         # SEQ "*"
        cig += [(l << 4) | o for o, l in ops]
        off.append(len(cig))
        seqs.append(rng.choice(np.array([1, 2, 4, 8, 15, 1, 2, 4, 8, 1, 2, 4, 8, 3], np.uint8), qlen))
        names.append("r%d/%s" % (i // 2, "x" * int(rng.integers(0, 40))))
    b = ReadBatch(tid, pos, flag, np.full(n, 30, np.uint8), np.array(off, np.uint32), np.array(cig, np.uint32))
    p = _write(tmp_path, "seq.bam", ["c0", "c1"], lengths, b, isize=rng.integers(-600, 600, n).astype(np.int32),
               names=names, seqs=seqs)
    with AlignmentFile(p) as host, AlignmentFile(p, decode="gpu") as gpu:
        assert np.array_equal(gpu.name_hashes(), host.name_hashes())
        for k in (1, 3, 7, 15):
            assert np.array_equal(gpu.qas_kmer_codes(k), host.qas_kmer_codes(k)), k
        for w in (1, 2, 56, 57, 300):
            assert np.array_equal(gpu.seq_windows(w), host.seq_windows(w)), w
        kc = [{}, {}]
        rr = np.random.default_rng(5)
        for r in range(2):
            for code in range(4 ** 3):
                if rr.random() < 0.8:
                    kc[r]["".join("ACGT"[(code >> (2 * (2 - j))) & 3] for j in range(3))] = float(rr.uniform(0.5, 2.0))
        for ref, s0, e0 in (("c0", 0, 4000), ("c1", 1000, 9000)):
            a = pileup.experimental(host, kc, 3, None, ref, s0, e0)
            g = pileup.experimental(gpu, kc, 3, None, ref, s0, e0)
            assert a == g, (ref, s0, e0)
        rows = []
        for f in (host, gpu):
            kh = mscan.KmerHist(3, 4, 3, 1)
            mscan.scan_reads(f, None, [kh])
            rows.append(kh.counts.copy())
        assert np.array_equal(rows[0], rows[1]) and rows[0].sum() > 0
        # the whole `scan` stack on the device-resident columns (mcov_kmer_hist_mem / mcov_isize_hist with MCOV_MEM_DEVICE):
        # ByFlag over IsizeHist + KmerHist, all records and a maxreads prefix, equal the host-decoded file's
        for maxreads in (0, 37):
            res = []
            for f in (host, gpu):
                bf = mscan.ByFlag([mscan.IsizeHist(), mscan.KmerHist(3, 4, 3, 1)], [mscan.Flags["Readdir"], mscan.Flags["IsRead1"]])
                nrec = mscan.scan_reads(f, None, bf, maxreads=maxreads)
                res.append((nrec, [[list(map(str, r)) for r in bf.get_rows(i)] for i in range(2)]))
            assert res[0][0] == res[1][0] == (maxreads or n)
            assert res[0][1] == res[1][1]
        assert gpu._h is None
    with AlignmentFile(p, decode="gpu") as fresh:
        mscan.scan_reads(fresh, None, mscan.ByFlag([mscan.IsizeHist(), mscan.KmerHist(3, 4, 3, 1)], [mscan.Flags["IsRead1"]]))
        assert fresh._h is None and fresh._soa is None           # neither the host reader nor a host copy of the columns


def test_alignmentfile_and_cli_with_gpu_decode(tmp_path):
    """AlignmentFile(decode="gpu") answers exactly like the host-decoded file: header, index statistics,
    classic() of the golden regions, the pileup() protocol, the records; `metacov pileup --bam-decode gpu`
    writes the same CSV."""
    from click.testing import CliRunner
    from helpers import load_json
    from metacov_b200 import AlignmentFile, pileup
    from metacov_b200.cli import pileup as cli_pileup
    z, b = load_soa("fixture_soa.npz")
    so = z["seq_off"]
    seqs = [z["seq"][so[i]:so[i + 1]] for i in range(len(b.tid))]
    p = _write(tmp_path, "fixture.bam", [str(x) for x in z["references"]], z["lengths"], b, isize=z["isize"],
               names=[str(x) for x in z["names"]], seqs=seqs)
    gold = load_json("fixture_classic.json")
    with AlignmentFile(p) as host, AlignmentFile(p, decode="gpu") as gpu:
        assert gpu.references == host.references and gpu.lengths == host.lengths and gpu.text == host.text
        assert (gpu.mapped, gpu.unmapped) == (host.mapped, host.unmapped) == (gold["mapped"], gold["unmapped"])
        for row in gold["classic"]:
            assert pileup.classic(gpu, row["ref"], row["start"], row["end"]) == row["result"], row
        assert [(c.pos, c.n) for c in gpu.pileup("ref1", 0, 425)] == [(c.pos, c.n) for c in host.pileup("ref1", 0, 425)]
        hs, gs = host.soa(), gpu.soa()
        for c in COLS:
            assert np.array_equal(hs[c], gs[c]), c
        # read names and SEQ are read on the device too: same hashes, k-mer keys and SEQ windows as the host reader's,
        # and the GPU-decoded file never opens a host reader
        assert np.array_equal(gpu.name_hashes(), host.name_hashes())
        for k in (1, 4, 7, 15):
            assert np.array_equal(gpu.qas_kmer_codes(k), host.qas_kmer_codes(k)), k
        for w in (1, 7, 56, 101, 400):
            assert np.array_equal(gpu.seq_windows(w), host.seq_windows(w)), w
        assert gpu._h is None
        gpu.set_pileup_filter(min_mapq=20)
        host.set_pileup_filter(min_mapq=20)
        assert pileup.classic(gpu, "ref1", 0, 425) == pileup.classic(host, "ref1", 0, 425)
    outs = []
    for mode in ("host", "gpu"):
        out = tmp_path / ("cov_%s.csv" % mode)
        res = CliRunner().invoke(cli_pileup, ["-b", p, "-o", str(out), "--bam-decode", mode])
        assert res.exit_code == 0, res.output
        outs.append(out.read_text())
    assert outs[0] == outs[1] and len(outs[0].splitlines()) == 3
    with pytest.raises(OSError):
        AlignmentFile(str(tmp_path / "fixture.bam.bai"), decode="gpu")          # not a BGZF file
    # decode="auto": a file this small fits the device -> GPU decode; with an impossible footprint -> streamed GPU decode
    with AlignmentFile(p, decode="auto") as auto:
        assert auto.decode == "gpu"
    old = AlignmentFile.GPU_DECODE_FOOTPRINT
    AlignmentFile.GPU_DECODE_FOOTPRINT = 1 << 50
    try:
        with AlignmentFile(p, decode="auto") as auto:
            assert auto.decode == "gpu-stream"
            assert pileup.classic(auto, "ref1", 0, 425) == gold["classic"][0]["result"] or gold["classic"][0]["ref"] != "ref1"
    finally:
        AlignmentFile.GPU_DECODE_FOOTPRINT = old


def test_streamed_gpu_decode_matches_whole_file(tmp_path):
    """mcov_bam_gpu_stream_depth: the file in chunks through the GPU decoder into the streamed pass.  Chunks far smaller
    than the file (records and BGZF blocks cut by every chunk border, reads carried from batch to batch, long reads that
    reach across several tiles) must give the depth of the one-shot pass bit for bit; unplaced reads at the end are not
    carried; a truncated file and an unsorted file fail loudly; AlignmentFile(decode="gpu-stream") answers like the host."""
    from metacov_b200 import AlignmentFile, CoverageEngine, McovError, ReadBatch, bamgpu, pileup, synth
    w = synth.c2(0.004)                                     # 40 000 reads, 4 contigs
    hb, isz = synth.generate_host(w)
    rng = np.random.default_rng(4)
    n = len(hb.tid)
    # a few long reads (span >> a tile) and a block of unplaced reads at the end
    cig = np.array(hb.cig); off = np.array(hb.cig_off)
    longs = rng.choice(n, 200, replace=False)
    single = np.flatnonzero(np.diff(off) == 1)
    longs = np.intersect1d(longs, single)
    cig[off[longs]] = (rng.integers(3000, 9000, len(longs)).astype(np.uint32) << 4)
    nu = 500
    tid = np.r_[hb.tid, np.full(nu, -1, np.int32)]; pos = np.r_[hb.pos, np.full(nu, -1, np.int32)]
    flag = np.r_[hb.flag, np.full(nu, 4, np.uint16)]; mapq = np.r_[hb.mapq, np.zeros(nu, np.uint8)]
    off2 = np.r_[off, np.full(nu, off[-1], np.uint32)]
    b = ReadBatch(tid, pos, flag, mapq, off2, cig)
    lengths = [int(x) for x in w.contig_len]
    p = _write(tmp_path, "big.bam", ["c%d" % c for c in range(w.n_contigs)], lengths, b)
    size = (tmp_path / "big.bam").stat().st_size
    want, woff, winfo = cport.depth(b, lengths, mode="diff")
    for chunk in (1 << 17, 300_000, 1 << 30):
        with CoverageEngine(lengths) as eng:
            info = bamgpu.stream_depth(eng, p, chunk_bytes=chunk)
            assert info["n_records"] == len(tid) and info["file_bytes"] == size
            assert info["n_chunks"] == 1 if chunk > size else info["n_chunks"] >= (4 if chunk == 1 << 17 else 2)
            pi = eng.pass_info()
            assert pi["n_pass"] == winfo["n_pass"] and pi["aligned_bases"] == winfo["aligned_bases"] and pi["sorted"] == 1
            for c in range(len(lengths)):
                assert np.array_equal(eng.copy_depth(c), want[woff[c]:woff[c] + lengths[c]]), (chunk, c)
            assert info["max_carry"] < 5000                    # the unplaced tail is not dragged along
    # the long-stream arrangement (pinned host buffers, taken by files of 1 GiB and more) on the same file: a child process,
    # because the switch is read once per process
    import subprocess, sys
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "from metacov_b200 import CoverageEngine, bamgpu\n"
            "lengths = %r\n"
            "with CoverageEngine(lengths) as eng:\n"
            "    info = bamgpu.stream_depth(eng, %r, chunk_bytes=1 << 17)\n"
            "    pi = eng.pass_info()\n"
            "    print(info['n_records'], info['n_chunks'], pi['n_pass'], pi['aligned_bases'], int(eng.copy_depth(1).astype(np.int64).sum()))\n"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), lengths, p))
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MCOV_STREAM_PIN="1"), capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    got = [int(x) for x in res.stdout.split()]
    assert got[0] == len(tid) and got[1] >= 4 and got[2] == winfo["n_pass"] and got[3] == winfo["aligned_bases"]
    assert got[4] == int(want[woff[1]:woff[1] + lengths[1]].astype(np.int64).sum())
    # truncated file
    raw = open(p, "rb").read()
    (tmp_path / "cut.bam").write_bytes(raw[:len(raw) * 2 // 3])
    with CoverageEngine(lengths) as eng:
        with pytest.raises(McovError):
            bamgpu.stream_depth(eng, str(tmp_path / "cut.bam"), chunk_bytes=1 << 17)
    # unsorted file: reported by the verdict
    perm = np.arange(len(hb.tid)); perm[100], perm[30000] = perm[30000], perm[100]
    ub = ReadBatch(hb.tid[perm], hb.pos[perm], hb.flag[perm], hb.mapq[perm], np.arange(len(perm) + 1, dtype=np.uint32),
                   np.full(len(perm), 100 << 4, np.uint32))
    pu = _write(tmp_path, "unsorted.bam", ["c%d" % c for c in range(w.n_contigs)], lengths, ub)
    with CoverageEngine(lengths) as eng:
        with pytest.raises(McovError):
            bamgpu.stream_depth(eng, pu, chunk_bytes=1 << 17)
            eng.pass_info()
    # the drop-in: same classic() as the host-decoded file, and the unsorted file still gets its depth (whole-file path)
    with AlignmentFile(p) as host, AlignmentFile(p, decode="gpu-stream", gpu_chunk_bytes=1 << 17) as gs:
        assert gs.references == host.references and gs.lengths == host.lengths
        for ref, a0, e0 in (("c0", 0, lengths[0]), ("c2", 1000, 30000), ("c3", 5, 50)):
            assert pileup.classic(gs, ref, a0, e0) == pileup.classic(host, ref, a0, e0)
        assert gs.stream_batches >= 4 and gs.gpu_stream_info["n_records"] == len(tid)
    with AlignmentFile(pu) as host, AlignmentFile(pu, decode="gpu-stream", gpu_chunk_bytes=1 << 17) as gs:
        assert pileup.classic(gs, "c1", 0, lengths[1]) == pileup.classic(host, "c1", 0, lengths[1])


def test_streamed_gpu_decode_random_chunk_sizes(tmp_path):
    """The streamed GPU decoder against the whole-file GPU decode of the same 400 000-read file for a dozen chunk sizes:
    whatever the chunk borders cut -- BGZF blocks, records, a record of which fewer than 36 bytes are present (chunk size
    1 287 135 on this file: the write pass once parsed such a stub and wrote its garbage op count past the op column) --
    every contig's depth and the pass counters are identical.  `tools/stream_fuzz.py N` runs N random sizes."""
    from metacov_b200 import CoverageEngine, bamgpu, synth
    w = synth.c2(0.04)
    hb, isz = synth.generate_host(w)
    path = str(tmp_path / "f.bam")
    synth.write_bam(path, w, hb, isz)
    size = os.path.getsize(path)
    lengths = [int(x) for x in w.contig_len]
    with CoverageEngine(lengths) as eng:
        bamgpu.depth_sorted(eng, bamgpu.decode(eng, path))
        want = [eng.copy_depth(c) for c in range(len(lengths))]
        pi0 = eng.pass_info()
    rng = np.random.default_rng(77)
    sizes = [1_287_135] + [int(rng.integers(1 << 17, size // 2)) for _ in range(11)]
    for chunk in sizes:
        with CoverageEngine(lengths) as eng:
            info = bamgpu.stream_depth(eng, path, chunk_bytes=chunk)
            pi = eng.pass_info()
            assert info["n_records"] == len(hb.tid), chunk
            assert pi["n_pass"] == pi0["n_pass"] and pi["aligned_bases"] == pi0["aligned_bases"], chunk
            for c in range(len(lengths)):
                assert np.array_equal(eng.copy_depth(c), want[c]), (chunk, c)
