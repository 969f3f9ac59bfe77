"""The device DEFLATE / CRC-32 code (csrc/inflate.cuh) compiled for the host, against zlib: every block
type (stored, fixed, dynamic), empty and 1-byte payloads, the BGZF blocks of a BAM written by the fixture
writer, and clean failure on truncated / corrupt input.  The GPU run of the same code is covered by
tests/test_gpu_bamgpu.py."""
import zlib

import numpy as np

from helpers import load_soa
from oracle import bamio


def _inflate(comp, ulen):
    """Decode with and without the circular output window (1 KiB: most matches take the far path, 8 KiB: the
    GPU's size); both must agree."""
    from metacov_b200 import _capi
    src = np.frombuffer(comp, dtype=np.uint8) if len(comp) else np.zeros(1, np.uint8)
    res = []
    for window in (0, 1024, 8192):
        dst = np.full(ulen + 8, 0xEE, dtype=np.uint8)
        if window:
            rc = _capi.lib.mcov_inflate_host_win(src.ctypes.data, len(comp), dst.ctypes.data, ulen, window)
        else:
            rc = _capi.lib.mcov_inflate_host(src.ctypes.data, len(comp), dst.ctypes.data, ulen)
        assert np.all(dst[ulen:] == 0xEE), "wrote past the end of the output"
        res.append((rc, dst[:ulen].tobytes()))
    assert all(r[0] == res[0][0] for r in res) and (res[0][0] != 0 or all(r[1] == res[0][1] for r in res))
    return res[0]


def _deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return c.compress(data) + c.flush()


def test_inflate_matches_zlib_on_every_block_type():
    from metacov_b200 import _capi
    rng = np.random.default_rng(3)
    payloads = [b"", b"A", rng.integers(0, 256, 65280, dtype=np.uint8).tobytes(),
                rng.integers(0, 4, 65280, dtype=np.uint8).tobytes(), (b"ACGT" * 20 + b"N") * 700,
                bytes(range(256)) * 200, b"\0" * 65280]
    for data in payloads:
        for level in (0, 1, 6, 9):                                  # 0: stored blocks
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                comp = _deflate(data, level, strat)
                rc, out = _inflate(comp, len(data))
                assert rc == 0 and out == data, (len(data), level, strat, rc)
        buf = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
        assert _capi.lib.mcov_crc32_host(buf.ctypes.data, len(data)) == (zlib.crc32(data) & 0xFFFFFFFF)
        for lanes in (1, 2, 7, 32):                                     # the warp's lane-sliced CRC fold
            for n in {len(data), min(len(data), 31), min(len(data), 33), min(len(data), 4097)}:
                assert _capi.lib.mcov_crc32_sliced_host(buf.ctypes.data, n, lanes) == (zlib.crc32(data[:n]) & 0xFFFFFFFF), (lanes, n)


def test_inflate_rejects_bad_streams():
    data = (b"ACGTTGCA" * 31 + b"xyz") * 200
    comp = _deflate(data, 6)
    assert _inflate(comp[:len(comp) // 2], len(data))[0] != 0          # truncated input
    assert _inflate(comp, len(data) - 1)[0] != 0                        # output one byte short
    assert _inflate(comp, len(data) + 1)[0] != 0                        # stream ends early
    assert _inflate(b"\x07" + comp[1:], len(data))[0] != 0             # block type 3
    bad = bytearray(_deflate(data, 0))
    bad[3] ^= 0xFF                                                      # stored block: LEN / NLEN mismatch
    assert _inflate(bytes(bad), len(data))[0] != 0


def test_inflate_bgzf_blocks_of_fixture_bam(tmp_path):
    z, b = load_soa("fixture_soa.npz")
    path = str(tmp_path / "f.bam")
    bamio.write_bam(path, [str(x) for x in z["references"]], z["lengths"].tolist(), b.tid, b.pos, b.flag, b.mapq, b.cig_off, b.cig)
    raw = open(path, "rb").read()
    off, n_blocks = 0, 0
    while off < len(raw):
        xlen = int.from_bytes(raw[off + 10:off + 12], "little")
        bsize = int.from_bytes(raw[off + 16:off + 18], "little")
        comp = raw[off + 12 + xlen:off + bsize + 1 - 8]
        ulen = int.from_bytes(raw[off + bsize + 1 - 4:off + bsize + 1], "little")
        rc, out = _inflate(comp, ulen)
        assert rc == 0 and out == zlib.decompress(comp, -15)
        off += bsize + 1
        n_blocks += 1
    assert n_blocks >= 2
