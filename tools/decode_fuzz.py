"""One-off randomized sweep (B200): random BAM files -- record sizes from tens of bytes to tens of kilobytes (CIGARs of 1..400
ops, names of 1..200 characters, SEQ of 0..6 000 bases), compression levels 0 / 1 / 6 / 9, unplaced tails -- decoded on the GPU
(whole file, and streamed with a random chunk size) against the host decoder: columns, names / SEQ derivatives, depth."""
import sys, os, tempfile, json, struct, zlib
import numpy as np
sys.path.insert(0, ".")
from metacov_b200 import AlignmentFile, CoverageEngine, ReadBatch, bamgpu
from oracle import bamio
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 10
tmp = tempfile.mkdtemp()
res = {"files": 0, "column_mismatches": 0, "stream_mismatches": 0, "nameseq_mismatches": 0}
for f in range(n_files):
    n = int(rng.integers(200, 25000))
    nc = int(rng.integers(1, 30))
    lengths = rng.integers(50_000, 400_000, nc)
    tid = np.sort(rng.integers(0, nc, n)).astype(np.int32)
    pos = np.zeros(n, np.int32)
    for c in range(nc):
        m = tid == c
        pos[m] = np.sort(rng.integers(0, lengths[c] - 1, int(m.sum())))
    nu = int(rng.integers(0, 200)) if f % 2 else 0
    tid = np.r_[tid, np.full(nu, -1, np.int32)]; pos = np.r_[pos, np.full(nu, -1, np.int32)]
    n = len(tid)
    big = rng.random(n) < 0.02
    n_op = np.where(big, rng.integers(50, 400, n), rng.integers(1, 6, n))
    off = np.concatenate(([0], np.cumsum(n_op))).astype(np.uint32)
    cig = ((rng.integers(1, 60, int(off[-1])).astype(np.uint32) << 4) | rng.choice(np.array([0, 1, 2, 4, 7, 8], np.uint32), int(off[-1])).astype(np.uint32))
    flag = rng.choice(np.array([99, 147, 83, 163, 0, 16, 1024 + 99, 4], np.uint16), n)
    flag[tid < 0] = 4
    mapq = rng.integers(0, 61, n).astype(np.uint8)
    names = ["r" * int(rng.integers(1, 200)) + str(i) for i in range(n)]
    seqs = [np.array([1, 2, 4, 8, 15], np.uint8)[rng.integers(0, 5, int(rng.integers(0, 6000)) if big[i] else int(rng.integers(0, 160)))] for i in range(n)]
    level = int(rng.choice([0, 1, 6, 9]))
    old = bamio._bgzf_block
    def blk(payload, level=level):
        comp = zlib.compressobj(level, zlib.DEFLATED, -15)
        cdata = comp.compress(payload) + comp.flush()
        head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, len(cdata) + 25)
        return head + cdata + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))
    bamio._bgzf_block = blk
    path = os.path.join(tmp, "f%d.bam" % f)
    try:
        bamio.write_bam(path, ["c%d" % c for c in range(nc)], [int(x) for x in lengths], tid, pos, flag, mapq, off, cig,
                        isize=rng.integers(-900, 900, n).astype(np.int32), names=names, seqs=seqs)
    finally:
        bamio._bgzf_block = old
    with AlignmentFile(path) as host, AlignmentFile(path, decode="gpu") as gpu:
        hs, gs = host.soa(), gpu.soa()
        ok = all(np.array_equal(hs[k], gs[k]) for k in hs)
        res["column_mismatches"] += 0 if ok else 1
        ok2 = np.array_equal(host.name_hashes(), gpu.name_hashes()) and np.array_equal(host.qas_kmer_codes(5), gpu.qas_kmer_codes(5)) and \
            np.array_equal(host.seq_windows(56), gpu.seq_windows(56))
        res["nameseq_mismatches"] += 0 if ok2 else 1
        want = [host.coverage_engine().copy_depth(c) for c in range(nc)]
    size = os.path.getsize(path)
    chunk = int(rng.integers(1 << 17, max((1 << 17) + 1, size)))
    with CoverageEngine([int(x) for x in lengths]) as eng:
        info = bamgpu.stream_depth(eng, path, chunk_bytes=chunk)
        ok3 = info["n_records"] == n and all(np.array_equal(eng.copy_depth(c), want[c]) for c in range(nc))
    res["stream_mismatches"] += 0 if ok3 else 1
    if not (ok and ok2 and ok3): print("MISMATCH file", f, "n", n, "level", level, "chunk", chunk, ok, ok2, ok3, flush=True)
    res["files"] += 1
print(json.dumps(res))
