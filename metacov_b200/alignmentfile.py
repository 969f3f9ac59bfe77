"""``AlignmentFile``: the subset of the pysam object protocol the coverage path
consumes, backed by the native BGZF/BAM/BAI reader (csrc/bamio.cpp) and the
CUDA depth engine.

Protocol provided (reference call sites): constructor from a path
(metacov/cli.py:56, 211); context manager (cli.py:258); ``references`` /
``lengths`` (util.py:67-68, cli.py:80); ``mapped`` / ``unmapped``
(cli.py:73-75, 214-216); ``pileup(ref, start, end)`` yielding objects with
``.pos`` / ``.n`` (pileup.py:13-16).
"""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import McovError, lib
from .engine import CoverageEngine, ReadBatch


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class PileupColumn:
    """Stand-in for pysam.PileupColumn (only ``pos`` and ``n`` are read by the
    reference, pileup.py:14-16)."""
    __slots__ = ("reference_id", "pos", "n")

    def __init__(self, tid, pos, n):
        self.reference_id = tid
        self.pos = pos
        self.n = n

    reference_pos = property(lambda self: self.pos)
    nsegments = property(lambda self: self.n)


class BamStream:
    """The streaming host reader (csrc/bamio.cpp, mcov_bam_stream_*): batches of a BAM decoded into pinned SoA
    buffers, each led by the reads the depth engine asked to see again.  Iterate with ``next_batch(resend)``."""

    def __init__(self, filename, batch_reads=1 << 21, threads=0):
        self._h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = lib.mcov_bam_stream_open(C.byref(self._h), str(filename).encode(), int(batch_reads), int(threads), err, len(err))
        if rc != 0:
            self._h = None
            raise OSError("%s: %s" % (filename, err.value.decode()))
        n = lib.mcov_bam_stream_n_ref(self._h)
        self.references = tuple(lib.mcov_bam_stream_ref_name(self._h, i).decode() for i in range(n))
        self.lengths = tuple(lib.mcov_bam_stream_ref_len(self._h, i) for i in range(n))
        self.text = lib.mcov_bam_stream_header_text(self._h).decode()

    def next_batch(self, resend=(-1, 0)):
        """(ReadBatch of pinned numpy views, n_carry, last, extra columns) or None after the last batch."""
        b = _capi.BamBatch()
        rc = lib.mcov_bam_stream_next(self._h, int(resend[0]), int(resend[1]), C.byref(b))
        if rc < 0:
            raise McovError(rc, lib.mcov_bam_stream_error(self._h).decode())
        if rc == 0:
            return None
        n = b.n
        batch = ReadBatch(_view(b.tid, n, np.int32), _view(b.pos, n, np.int32), _view(b.flag, n, np.uint16),
                          _view(b.mapq, n, np.uint8), _view(b.cig_off, n + 1, np.uint32), _view(b.cig, b.n_cigar, np.uint32))
        extra = dict(l_seq=_view(b.l_seq, n, np.int32), isize=_view(b.isize, n, np.int32), reflen=_view(b.reflen, n, np.int32))
        return batch, int(b.n_carry), bool(b.last), extra

    def next_block(self, resend=(-1, 0), with_mapq=False):
        """Like ``next_batch`` plus the batch as a transport block: (block or None, ReadBatch, n_carry, last, extra)."""
        b = _capi.BamBatch()
        blk, nb = C.c_void_p(), C.c_int64(0)
        rc = lib.mcov_bam_stream_next_block(self._h, int(resend[0]), int(resend[1]), 1 if with_mapq else 0,
                                            C.byref(blk), C.byref(nb), C.byref(b))
        if rc < 0:
            raise McovError(rc, lib.mcov_bam_stream_error(self._h).decode())
        if rc == 0:
            return None
        n = b.n
        batch = ReadBatch(_view(b.tid, n, np.int32), _view(b.pos, n, np.int32), _view(b.flag, n, np.uint16),
                          _view(b.mapq, n, np.uint8), _view(b.cig_off, n + 1, np.uint32), _view(b.cig, b.n_cigar, np.uint32))
        extra = dict(l_seq=_view(b.l_seq, n, np.int32), isize=_view(b.isize, n, np.int32), reflen=_view(b.reflen, n, np.int32))
        block = (blk.value, nb.value) if blk.value else None
        return block, batch, int(b.n_carry), bool(b.last), extra

    @property
    def n_records(self):
        return lib.mcov_bam_stream_records(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib.mcov_bam_stream_close(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


def stream_depth(engine, stream):
    """Per-base depth of a whole sorted BAM through the streamed fused path: batch k+1 is decoded on the host while
    the GPU works on batch k.  Returns the number of batches.  Raises McovError(MCOV_ERR_UNSORTED) -- at the first
    synchronising call on the engine -- when the file is not coordinate-sorted."""
    engine.stream_begin()
    resend = (-1, 0)
    k = 0
    while True:
        item = stream.next_block(resend, with_mapq=engine.filter.min_mapq > 0)
        if item is None:
            if k == 0:
                engine.stream_push(ReadBatch(*[np.zeros(0, d) for d in (np.int32, np.int32, np.uint16, np.uint8)],
                                             np.zeros(1, np.uint32), np.zeros(0, np.uint32)), 0, last=True)
            break
        block, batch, n_carry, last, _ = item
        # the transport block (one narrow host-to-device copy) when the batch qualifies, else its columns
        if block is not None and len(batch.tid):
            rt = engine.stream_push_block(block, last=last)
        else:
            rt = engine.stream_push(batch, n_carry=n_carry, last=last)
        if len(batch.tid):
            resend = rt
        k += 1
        if last:
            break
    return k


class AlignmentFile:
    def __init__(self, filename, mode="rb", device=0, decode="host", **kw):
        """decode="host": BGZF inflate and record parsing by the native host reader (csrc/bamio.cpp), streamed batch by
        batch; decode="gpu": the compressed file goes to the device and is inflated and parsed there (csrc/bam_gpu.cu) --
        the coverage path then never holds the records on the host, and the file is decoded ~10x faster than 16 host
        cores manage (tools/bench_bam.py), but the compressed image, the inflated stream and the columns must fit the
        device; decode="gpu-stream": the file is read in chunks (``gpu_chunk_bytes``, default 64 MiB) and every chunk is
        inflated, parsed and pushed into the streamed depth pass on the device (mcov_bam_gpu_stream_depth) -- any file size,
        the host only reads; accessors that need every record on the host (soa(), read names) open the host reader on
        demand; decode="auto": "gpu" when the whole file fits the device (file size x GPU_DECODE_FOOTPRINT below the free
        device memory), else "gpu-stream"."""
        if "w" in mode:
            raise ValueError("metacov_b200.AlignmentFile is read-only")
        if decode not in ("host", "gpu", "gpu-stream", "auto"):
            raise ValueError("decode must be 'host', 'gpu', 'gpu-stream' or 'auto'")
        if decode == "auto":
            decode = "gpu" if self._fits_device(filename, device) else "gpu-stream"
        self._gpu_chunk_bytes = int(kw.get("gpu_chunk_bytes", 64 << 20))
        self.gpu_stream_info = None
        self.decode = decode
        self.filename = filename
        self._gpu = None
        self._device = device
        self._soa = None
        self._engine = None
        self._filter_kw = None
        self._index_stats = None
        self._batch_reads = int(kw.get("batch_reads", 1 << 21))     # records per batch of the streamed depth pass
        self.stream_batches = 0
        self._h = None
        self._stream = None
        if decode == "gpu":
            self._open_gpu(filename, device)
            return
        # host decode: only the header is read here (streaming reader: the first blocks of the file); the records are
        # streamed batch by batch by coverage_engine(), or loaded whole by soa() for the callers that need every column
        self._stream = BamStream(filename, batch_reads=self._batch_reads)
        self.references, self.lengths, self.text = self._stream.references, self._stream.lengths, self._stream.text
        self.nreferences = len(self.references)
        self._tid = {name: i for i, name in enumerate(self.references)}

    # device bytes per byte of BAM that a GPU decode needs: the image itself, the inflated stream (BAM records deflate
    # ~4-5x), the SoA columns, and as much again for the depth arrays and scratch of the pass that follows
    GPU_DECODE_FOOTPRINT = 12

    @classmethod
    def _fits_device(cls, filename, device):
        import torch
        try:
            size = os.path.getsize(filename)
            free, _total = torch.cuda.mem_get_info(device)
        except (OSError, RuntimeError):
            return False
        return size * cls.GPU_DECODE_FOOTPRINT < free

    def _open_gpu(self, filename, device):
        from . import bamgpu
        self._h = None
        eng = CoverageEngine([1], device=device)            # the contig table follows once the header is decoded
        try:
            dsoa = bamgpu.decode(eng, os.fspath(filename))      # (the library maps the file: no host copy of the image)
            self.text, refs = dsoa.header()
        except McovError as e:
            eng.close()
            raise OSError("%s: %s" % (filename, e))
        except BaseException:
            eng.close()
            raise
        self.references = tuple(r for r, _ in refs)
        self.lengths = tuple(l for _, l in refs)
        self.nreferences = len(refs)
        self._tid = {name: i for i, name in enumerate(self.references)}
        eng.set_contigs(self.lengths)
        self._gpu = (eng, dsoa)

    def _host_handle(self):
        """The host reader's handle (a GPU-decoded file never needs it: columns, read names and SEQ all come from
        the device, bamgpu.DeviceSoA)."""
        if self._h is None:
            self._open_host(self.filename, header=False)
        return self._h

    def _open_host(self, filename, header=True):
        self._h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = lib.mcov_bam_open(C.byref(self._h), str(filename).encode(), err, len(err))
        if rc != 0:
            self._h = None
            # pysam raises OSError/ValueError for unreadable files
            raise OSError("%s: %s" % (filename, err.value.decode()))
        if not header:
            return
        n = lib.mcov_bam_n_ref(self._h)
        self.references = tuple(lib.mcov_bam_ref_name(self._h, i).decode() for i in range(n))
        self.lengths = tuple(lib.mcov_bam_ref_len(self._h, i) for i in range(n))
        self.nreferences = n
        self.text = lib.mcov_bam_header_text(self._h).decode()
        self._tid = {name: i for i, name in enumerate(self.references)}

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "_engine", None) is not None:
            self._engine.close()
            self._engine = None
        if getattr(self, "_gpu", None) is not None:
            self._gpu[0].close()                            # (the same engine, if the depth was computed)
            self._gpu = None
        if getattr(self, "_h", None):
            lib.mcov_bam_close(self._h)
            self._h = None
        if getattr(self, "_stream", None) is not None:
            self._stream.close()
            self._stream = None
        self._soa = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    # -- header / index ---------------------------------------------------
    def get_tid(self, reference):
        return self._tid.get(reference, -1)

    def get_reference_name(self, tid):
        return self.references[tid]

    def _idx(self):
        if self._index_stats is None:
            m, u = C.c_int64(0), C.c_int64(0)
            rc = lib.mcov_bai_stats(str(self.filename).encode(), C.byref(m), C.byref(u))
            if rc != 0:
                raise ValueError("mapping information not recorded in index or index not available")
            self._index_stats = (m.value, u.value)
        return self._index_stats

    @property
    def mapped(self):
        return self._idx()[0]

    @property
    def unmapped(self):
        return self._idx()[1]

    def has_index(self):
        try:
            self._idx()
            return True
        except ValueError:
            return False

    # -- records ----------------------------------------------------------
    def soa(self):
        """Every record of the file as SoA numpy views (file order; the arrays
        the reference reads field by field at scan.pyx:243-294)."""
        if self._soa is None and self._gpu is not None:
            self._soa = {c: self._gpu[1].to_host(c) for c, _ in self._gpu[1].COLS}
        if self._soa is None:
            self._host_handle()
            rc = lib.mcov_bam_load(self._h, 0)
            if rc != 0:
                raise McovError(rc, "BAM record decode failed")
            n = lib.mcov_bam_n_records(self._h)
            nc = lib.mcov_bam_n_cigar(self._h)
            self._soa = dict(
                tid=_view(lib.mcov_bam_tid(self._h), n, np.int32),
                pos=_view(lib.mcov_bam_pos(self._h), n, np.int32),
                flag=_view(lib.mcov_bam_flag(self._h), n, np.uint16),
                mapq=_view(lib.mcov_bam_mapq(self._h), n, np.uint8),
                l_seq=_view(lib.mcov_bam_lseq(self._h), n, np.int32),
                isize=_view(lib.mcov_bam_isize(self._h), n, np.int32),
                cig_off=_view(lib.mcov_bam_cig_off(self._h), n + 1, np.uint32),
                cig=_view(lib.mcov_bam_cig(self._h), nc, np.uint32),
            )
        return self._soa

    def seq_windows(self, win_bases):
        """uint8[n, (win_bases+1)//2]: per read the first (forward) / last (reverse) win_bases bases,
        nt16 two per byte -- the part of SEQ the k-mer histogram looks at."""
        if self._gpu is not None:                       # decoded on the GPU: SEQ is read there too
            return self._gpu[1].names_seq(win_bases=win_bases)["seq_win"]
        n = len(self.soa()["tid"])
        rc = lib.mcov_bam_load_seq(self._host_handle())
        if rc != 0:
            raise McovError(rc, "BAM SEQ decode failed")
        out = np.empty((n, (win_bases + 1) // 2), dtype=np.uint8)
        rc = lib.mcov_bam_seq_windows(self._host_handle(), int(win_bases), _capi.ptr(out))
        if rc != 0:
            raise McovError(rc, "mcov_bam_seq_windows failed")
        return out

    def __len__(self):
        return len(self.soa()["tid"])

    # -- pileup.experimental ------------------------------------------------
    def name_hashes(self):
        """uint64[n]: FNV-1a of every read name (the key of the reference's mate dict, pileup.py:101)."""
        if self._gpu is not None:
            return self._gpu[1].names_seq(name_hash=True)["name_hash"]
        n = len(self.soa()["tid"])
        rc = lib.mcov_bam_load_seq(self._host_handle())
        if rc != 0:
            raise McovError(rc, "BAM SEQ decode failed")
        return _view(lib.mcov_bam_name_hash(self._host_handle()), n, np.uint64)

    def qas_kmer_codes(self, k_len):
        """int32[n]: code of ``read.query_alignment_sequence[0:k_len]`` (pileup.py:109, 123), -1 = no key."""
        if self._gpu is not None:
            return self._gpu[1].names_seq(k_len=k_len)["kmer_code"]
        n = len(self.soa()["tid"])
        rc = lib.mcov_bam_load_seq(self._host_handle())
        if rc != 0:
            raise McovError(rc, "BAM SEQ decode failed")
        out = np.empty(n, dtype=np.int32)
        rc = lib.mcov_bam_qas_kmer(self._host_handle(), int(k_len), _capi.ptr(out))
        if rc != 0:
            raise McovError(rc, "mcov_bam_qas_kmer failed")
        return out

    def _fetch_index(self):
        """Per-contig record ranges and the longest reference span: what stands in for the BAI when
        a region's candidate records are looked up (``bam.fetch`` needs a sorted, indexed file)."""
        if getattr(self, "_fidx", None) is None:
            s = self.soa()
            # (contig, position) as one sortable key; unplaced reads (tid -1 -> 2^32-1) sort last
            key = (s["tid"].view(np.uint32).astype(np.int64) << 31) | s["pos"].astype(np.int64).clip(0)
            if len(key) > 1 and np.any(key[1:] < key[:-1]):
                raise ValueError("fetch called on bamfile without index")          # pysam's message
            ut = s["tid"].view(np.uint32)
            cstart = np.searchsorted(ut, np.arange(self.nreferences + 1, dtype=np.uint32), side="left")
            consumes = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)   # M D N = X
            per_op = (s["cig"] >> 4).astype(np.int64) * consumes[s["cig"] & 15]
            csum = np.concatenate(([0], np.cumsum(per_op)))
            reflen = csum[s["cig_off"][1:]] - csum[s["cig_off"][:-1]]
            self._fidx = (cstart, int(max(reflen.max(initial=0), 1)))
        return self._fidx

    def experimental_stats(self, refs, starts, ends, k_len, kc_val, kc_has):
        """``mcov_exp_stats`` records of the regions (the read loop of pileup.py:90-146 on the GPU)."""
        s = self.soa()
        cstart, max_span = self._fetch_index()
        pos = s["pos"]
        g = len(refs)
        lb = np.zeros(g, dtype=np.int64)
        ub = np.zeros(g, dtype=np.int64)
        for i, (ref, s0, e0) in enumerate(zip(refs, starts, ends)):
            tid = self._tid[ref]                                             # KeyError for unknown contigs, like pysam
            c0, c1 = int(cstart[tid]), int(cstart[tid + 1])
            lb[i] = c0 + np.searchsorted(pos[c0:c1], s0 - max_span, side="right")
            ub[i] = c0 + np.searchsorted(pos[c0:c1], e0, side="left")
            ub[i] = max(ub[i], lb[i])
        hashes = self.name_hashes()
        codes = self.qas_kmer_codes(k_len) if kc_val is not None else np.full(len(pos), -1, dtype=np.int32)
        eng = self.coverage_engine()
        out = np.zeros(g, dtype=_capi.EXP_STATS_DTYPE)
        # batches bounded by the scratch they need: (region, read) entries and last-writer slots
        i = 0
        while i < g:
            j, ent, slots = i, 0, 0
            while j < g and (j == i or (ent + ub[j] - lb[j] < (1 << 30) and slots + ends[j] - starts[j] < (1 << 29))):
                ent += int(ub[j] - lb[j]); slots += int(ends[j] - starts[j]); j += 1
            # only the reads the batch's regions can touch travel to the device: [lo, hi) of the file order (a single region
            # of a 10 M-read file is a few thousand reads, not 266 MB of columns)
            lo, hi = int(lb[i:j].min()), int(ub[i:j].max())
            hi = max(hi, lo)
            o0, o1 = int(s["cig_off"][lo]), int(s["cig_off"][hi])
            sub = {"pos": s["pos"][lo:hi], "flag": s["flag"][lo:hi], "cig": s["cig"][o0:o1],
                   "cig_off": (s["cig_off"][lo:hi + 1] - np.uint32(o0)).astype(np.uint32)}
            out[i:j] = eng.experimental_stats(sub, hashes[lo:hi], codes[lo:hi], k_len, kc_val, kc_has,
                                              np.asarray(starts[i:j], dtype=np.int32), np.asarray(ends[i:j], dtype=np.int32),
                                              lb[i:j] - lo, ub[i:j] - lo)
            i = j
        return out

    # -- coverage ---------------------------------------------------------
    def set_pileup_filter(self, **kw):
        """Override pysam's implicit pileup arguments (flag_filter, flag_require,
        min_mapq, ignore_orphans, max_depth); invalidates the cached depth."""
        self._filter_kw = kw
        if self._engine is not None:
            if self._gpu is None:
                self._engine.close()                        # (a GPU-decoded file keeps its engine: it owns the columns)
            self._engine = None

    def coverage_engine(self):
        """The per-base depth of the whole file, computed once on the GPU."""
        if self._engine is None and self._gpu is not None:
            eng, path = self._gpu_depth()
        elif self._engine is None and self._soa is None:
            # host decode, records not loaded: stream the file in batches (never held in memory); a file that is not
            # coordinate-sorted, or one where htslib's max_depth cap fires, takes the whole-file path below
            eng = CoverageEngine(self.lengths, device=self._device, filt=self._filter_kw)
            path = None
            if self.decode == "gpu-stream":
                # chunks of the file through the GPU decoder into the streamed pass; a file it cannot take (not sorted, the
                # max_depth cap fires) falls through to the host paths below, which know what to do with it
                from . import bamgpu
                try:
                    self.gpu_stream_info = bamgpu.stream_depth(eng, self.filename, chunk_bytes=self._gpu_chunk_bytes)
                    eng.pass_info()
                    path = "fused"
                    self.stream_batches = self.gpu_stream_info["n_chunks"]
                except McovError as e:
                    if e.code not in (_capi.MCOV_ERR_UNSORTED, _capi.MCOV_ERR_RANGE, _capi.MCOV_ERR_STATE):
                        eng.close()
                        raise
            if path is None and self.decode != "gpu-stream":    # (what the GPU stream refused, the host stream would refuse too)
                try:
                    st, self._stream = self._stream, None       # the reader opened for the header, not yet advanced
                    if st is None:
                        st = BamStream(self.filename, batch_reads=self._batch_reads)
                    with st:
                        self.stream_batches = stream_depth(eng, st)
                    eng.pass_info()                              # delivers the verdict of the stream
                    path = "fused"
                except McovError as e:
                    if e.code not in (_capi.MCOV_ERR_UNSORTED, _capi.MCOV_ERR_RANGE, _capi.MCOV_ERR_STATE):
                        eng.close()
                        raise
            if path is None:
                s = self.soa()
                path = eng.compute_depth(ReadBatch(s["tid"], s["pos"], s["flag"], s["mapq"], s["cig_off"], s["cig"]))
        elif self._engine is None:
            s = self.soa()
            eng = CoverageEngine(self.lengths, device=self._device, filt=self._filter_kw)
            path = eng.compute_depth(ReadBatch(s["tid"], s["pos"], s["flag"], s["mapq"], s["cig_off"], s["cig"]))
        if self._engine is None:
            info = eng.pass_info()
            md = eng.filter.max_depth
            # the fused (sorted) path replays htslib's cap exactly on the GPU; the any-order path cannot
            # (the cap is only defined for sorted input), so it refuses rather than return uncapped numbers
            if path != "fused" and md > 0 and info["cap_metric"] > md:
                if self._gpu is None:
                    eng.close()
                from .pileup import DepthCapError
                raise DepthCapError(
                    "depth[p-1]+starts[p] reaches %d > max_depth=%d: htslib's pileup would drop reads here "
                    "(order-dependent cap); raise max_depth via set_pileup_filter(max_depth=...)"
                    % (info["cap_metric"], md))
            self._engine = eng
        return self._engine

    def _gpu_depth(self):
        """Depth straight from the device-resident columns of a GPU-decoded file."""
        from . import bamgpu
        eng, dsoa = self._gpu
        if self._filter_kw:
            eng.set_filter(**self._filter_kw)
        try:
            bamgpu.depth_sorted(eng, dsoa)
            return eng, "fused"
        except McovError as e:
            if e.code not in (_capi.MCOV_ERR_UNSORTED, _capi.MCOV_ERR_RANGE):
                raise
        r = dsoa.raw
        eng.begin()
        eng._check(lib.mcov_push_reads(eng._ctx, dsoa.n_records, r.tid, r.pos, r.flag, r.mapq, r.cig_off, r.cig, _capi.MEM_DEVICE))
        eng.finalize()
        return eng, "push"

    def pileup(self, contig=None, start=None, stop=None, **kw):
        """Columns with n > 0 inside [start, stop) (pysam also emits columns
        outside the region when truncate=False; the reference discards them,
        pileup.py:14-15)."""
        tid = self._tid[contig]
        start = 0 if start is None else max(int(start), 0)
        stop = self.lengths[tid] if stop is None else min(int(stop), self.lengths[tid])
        d = self.coverage_engine().copy_depth(tid, start, stop) if stop > start else np.zeros(0, np.int32)
        for off in np.nonzero(d)[0]:
            yield PileupColumn(tid, start + int(off), int(d[off]))

    def count_coverage_depth(self, contig, start=None, stop=None):
        """Per-base depth as an int32 array (additive helper)."""
        tid = self._tid[contig]
        return self.coverage_engine().copy_depth(tid, 0 if start is None else start, stop)
