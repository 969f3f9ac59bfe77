"""BASELINE.json's synthetic workloads for the CPU arms of bench.py, WITHOUT the product library: the workload
tables (contig lengths, reads per contig: SURVEY.md 8(d)) restated in numpy, the reads from the shared counter-based
generator header include/mcov_synth.h compiled into liboracle.so (oracle/synthgen.c).  Test / bench infrastructure
only; metacov_b200/synth.py is the product-side twin and tests/test_oracle.py checks that both give the same reads."""
import ctypes as C
import os
from collections import namedtuple

import numpy as np

from . import cport

Batch = namedtuple("Batch", "tid pos flag mapq cig_off cig")


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("mode", C.c_int32), ("read_len", C.c_int32), ("span_min", C.c_int32),
                ("span_max", C.c_int32), ("margin", C.c_int32), ("reserved", C.c_int32)]


class Workload:
    def __init__(self, name, contig_len, reads_per_contig, seed, mode=0, read_len=150, span_min=0, span_max=0):
        self.name = name
        self.contig_len = np.ascontiguousarray(contig_len, dtype=np.int32)
        rpc = np.asarray(reads_per_contig, dtype=np.int64)
        self.read_start = np.concatenate(([0], np.cumsum(rpc))).astype(np.int64)
        self.n_reads = int(self.read_start[-1])
        self.n_contigs = len(self.contig_len)
        margin = read_len + 5 if mode == 0 else 2 * span_max
        self.params = SynthParams(seed=seed, mode=mode, read_len=read_len, span_min=span_min, span_max=span_max,
                                  margin=margin, reserved=0)

    def describe(self):
        return "%s: %d reads, %d contigs, %d bp" % (self.name, self.n_reads, self.n_contigs,
                                                    int(self.contig_len.astype(np.int64).sum()))


def c2(scale=1.0, seed=1001):
    n = max(1, int(round(1000 * scale)))
    return Workload("C2", np.full(n, 50_000), np.full(n, 10_000), seed)


def c3(scale=1.0, seed=1003):
    n = max(1, int(round(500_000 * scale)))
    rng = np.random.Generator(np.random.PCG64(seed))
    ln = np.clip(np.exp(rng.normal(np.log(2000.0), 0.6, n)), 500, 50_000).astype(np.int64)
    w = np.exp(rng.normal(0.0, 1.5, n))
    reads = w * ln
    reads = reads / reads.sum() * int(round(200_000_000 * scale))
    reads = np.minimum(reads, 4000.0 * ln / 150.0)
    return Workload("C3", ln, np.maximum(np.floor(reads), 1).astype(np.int64), seed)


def c4(scale=1.0, seed=1004):
    n = max(1, int(round(100_000 * scale)))
    return Workload("C4", np.full(n, 50_000), np.full(n, 10_000), seed)


def c5(scale=1.0, seed=1005, span_min=10_000, span_max=50_000):
    n = max(1, int(round(20_000 * scale)))
    rng = np.random.Generator(np.random.PCG64(seed))
    ln = rng.integers(max(100_000, 2 * span_max + 1), max(300_000, 6 * span_max) + 1, n).astype(np.int64)
    return Workload("C5", ln, np.full(n, 250), seed, mode=1, span_min=span_min, span_max=span_max)


WORKLOADS = {"c2": c2, "c3": c3, "c4": c4, "c5": c5}


def generate(w, i0=0, n=None, tid_base=0, threads=None):
    """Reads [i0, i0+n) of the workload as numpy arrays -> (Batch, isize)."""
    L = cport.lib()
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.orc_synth_ncigar.restype = None
    L.orc_synth_ncigar.argtypes = [C.POINTER(SynthParams), i64, i64, vp, C.c_int]
    L.orc_synth_reads.restype = None
    L.orc_synth_reads.argtypes = [C.POINTER(SynthParams), i64, i64, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    n = w.n_reads - i0 if n is None else n
    threads = threads or (os.cpu_count() or 1)
    ncig = np.empty(n, dtype=np.uint32)
    L.orc_synth_ncigar(C.byref(w.params), i0, n, ncig.ctypes.data, threads)
    tot = int(ncig.sum(dtype=np.int64))
    if tot >= 2 ** 32:
        raise ValueError("more than 2^32-1 CIGAR ops: generate in smaller pieces")
    off = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(ncig, dtype=np.uint32, out=off[1:])
    tid = np.empty(n, np.int32); pos = np.empty(n, np.int32); flag = np.empty(n, np.uint16)
    mapq = np.empty(n, np.uint8); isize = np.empty(n, np.int32); cig = np.empty(max(tot, 1), np.uint32)
    L.orc_synth_reads(C.byref(w.params), i0, n, w.read_start.ctypes.data, w.contig_len.ctypes.data, w.n_contigs, tid_base,
                      off.ctypes.data, tid.ctypes.data, pos.ctypes.data, flag.ctypes.data, mapq.ctypes.data,
                      isize.ctypes.data, cig.ctypes.data, threads)
    return Batch(tid, pos, flag, mapq, off, cig[:tot]), isize
