"""ctypes access to oracle/liboracle.so (C restatement; test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

ORC_STATS_DTYPE = np.dtype([
    ("sum", "<i8"), ("sumsq", "<u8"), ("iq_sum", "<i8"), ("n_ge1", "<i8"), ("n_geN", "<i8"),
    ("min", "<i4"), ("max", "<i4"), ("med_lo", "<i4"), ("med_hi", "<i4"), ("reserved", "<i4"), ("flags", "<i4")])


class OrcFilter(C.Structure):
    _fields_ = [("flag_filter", C.c_uint16), ("flag_require", C.c_uint16), ("min_mapq", C.c_uint8),
                ("ignore_orphans", C.c_uint8), ("count_del", C.c_uint8), ("reflen0_as_one", C.c_uint8), ("max_depth", C.c_int32)]


def default_filter(**kw):
    f = OrcFilter(0x704, 0, 0, 1, 1, 0, 8000)
    for k, v in kw.items():
        setattr(f, k, v)
    return f


def build(force=False):
    srcs = [os.path.join(HERE, "coverage.c"), os.path.join(HERE, "synthgen.c"),
            os.path.join(os.path.dirname(HERE), "include", "mcov_synth.h")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        _lib.orc_depth_diff.restype = i64
        _lib.orc_depth_diff.argtypes = [i64, vp, vp, vp, vp, vp, vp, C.POINTER(OrcFilter), i32, vp, vp, vp, vp]
        _lib.orc_depth_diff_par.restype = i64
        _lib.orc_depth_diff_par.argtypes = [i64, vp, vp, vp, vp, vp, vp, C.POINTER(OrcFilter), i32, vp, vp, vp, vp, C.c_int]
        _lib.orc_depth_plp.restype = i64
        _lib.orc_depth_plp.argtypes = [i64, vp, vp, vp, vp, vp, vp, C.POINTER(OrcFilter), i32, vp, vp, vp]
        _lib.orc_region_stats.restype = None
        _lib.orc_region_stats.argtypes = [vp, vp, vp, i64, vp, vp, vp, i32, vp, C.c_int]
        _lib.orc_isize_hist.restype = i32
        _lib.orc_isize_hist.argtypes = [i64, vp, vp, i32, vp, i32, vp, vp]
        _lib.orc_reflen_all.restype = i64
        _lib.orc_reflen_all.argtypes = [i64, vp, vp, vp]
    return _lib


def layout(lengths):
    """Slot offsets: contig c owns len+1 slots, 16-byte aligned (same rule as mcov_set_contigs)."""
    lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    sizes = (lengths.astype(np.int64) + 1 + 3) & ~3
    off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    return lengths, off


def _arrs(b):
    return (np.ascontiguousarray(b.tid, np.int32), np.ascontiguousarray(b.pos, np.int32),
            np.ascontiguousarray(b.flag, np.uint16), np.ascontiguousarray(b.mapq, np.uint8),
            np.ascontiguousarray(b.cig_off, np.uint32), np.ascontiguousarray(b.cig, np.uint32))


def depth(batch, lengths, filt=None, mode="diff", threads=1):
    """Per-base depth of every contig.  mode: 'diff' | 'par' | 'plp'.
    Returns (depth int32[off[-1]], off, info dict)."""
    L = lib()
    f = filt or default_filter()
    lengths, off = layout(lengths)
    a = _arrs(batch)
    n = len(a[0])
    d = np.empty(int(off[-1]), dtype=np.int32)
    p = [x.ctypes.data for x in a]
    aligned = C.c_int64(0)
    if mode == "diff":
        r = L.orc_depth_diff(n, *p, C.byref(f), len(lengths), lengths.ctypes.data, off.ctypes.data, d.ctypes.data,
                             C.addressof(aligned))
        info = {"n_pass": r, "aligned_bases": aligned.value}
    elif mode == "par":
        r = L.orc_depth_diff_par(n, *p, C.byref(f), len(lengths), lengths.ctypes.data, off.ctypes.data, d.ctypes.data,
                                 C.addressof(aligned), threads)
        if r < 0:
            raise ValueError("orc_depth_diff_par: unsorted input")
        info = {"n_pass": r, "aligned_bases": aligned.value}
    elif mode == "plp":
        r = L.orc_depth_plp(n, *p, C.byref(f), len(lengths), lengths.ctypes.data, off.ctypes.data, d.ctypes.data)
        if r < 0:
            raise ValueError("orc_depth_plp: unsorted input")
        info = {"dropped_by_cap": r}
    else:
        raise ValueError(mode)
    return d, off, info


def region_stats(d, off, lengths, tid, start, end, breadth_n=1, threads=1):
    L = lib()
    lengths = np.ascontiguousarray(lengths, np.int32)
    tid = np.ascontiguousarray(tid, np.int32); start = np.ascontiguousarray(start, np.int32)
    end = np.ascontiguousarray(end, np.int32)
    out = np.zeros(len(tid), dtype=ORC_STATS_DTYPE)
    L.orc_region_stats(d.ctypes.data, off.ctypes.data, lengths.ctypes.data, len(tid), tid.ctypes.data,
                       start.ctypes.data, end.ctypes.data, breadth_n, out.ctypes.data, threads)
    return out


def isize_hist(flag, isize, group_flags=(), n_bins=1024):
    L = lib()
    flag = np.ascontiguousarray(flag, np.uint16); isize = np.ascontiguousarray(isize, np.int32)
    gf = np.ascontiguousarray(group_flags, np.uint16)
    groups = 1 << len(gf)
    hist = np.zeros((groups, n_bins), np.uint32); cnt = np.zeros(groups, np.uint64)
    mx = L.orc_isize_hist(len(flag), flag.ctypes.data, isize.ctypes.data, len(gf), gf.ctypes.data if len(gf) else None,
                          n_bins, hist.ctypes.data, cnt.ctypes.data)
    return hist, cnt, mx
