// k_prep_tma.cuh -- first kernel of the fused path, TMA-staged (the short-read hot path).
//
// Same work and same outputs as k_fused_prep (k_fused.cuh): filter + CIGAR reduce of every read,
// 4-byte records, tile borders, far-read tables.  What differs is how the reads reach the SM.
// k_fused_prep issues five 128-bit loads per thread and then a dependent load per CIGAR op; ncu
// showed it waiting on exactly those (long_scoreboard 46 % of the stall samples: 20 % on the first
// use of the SoA columns, 17 % on the first use of the CIGAR op) at 64 registers / 50 % occupancy.
// Here a PRODUCER warp stages whole chunks of the SoA -- 1024 consecutive reads: tid, pos, flag,
// mapq, offsets and the chunk's contiguous CIGAR range -- into a ring of shared-memory stages
// with cp.async.bulk (TMA) completing on mbarriers; the 8 CONSUMER warps wait on the stage's
// "full" barrier, take everything from shared memory and release the stage through its "empty"
// barrier.  The dependent load (offset -> op) disappears: the producer knows a chunk's op range
// from the two offsets at its ends (fetched one chunk ahead), so ops and columns travel together.
//
// Requirements (else the caller launches k_fused_prep): 32-bit offsets, every column 16-byte
// aligned.  A chunk whose ops do not fit the stage (long reads) leaves them in global memory.
#pragma once
#include "k_fused.cuh"

namespace mcov {

#ifndef MCOV_PREP_STAGES
#define MCOV_PREP_STAGES 3
#endif
#ifndef MCOV_PREP_CTAS
#define MCOV_PREP_CTAS 3
#endif
constexpr int kPtStages = MCOV_PREP_STAGES;
constexpr int kPtChunk = kPrepThreads * kPrepPer;          // 1024 reads per stage
constexpr int kPtCigCap = 2048;                            // ops staged per chunk (average 2 per read)
constexpr int kPtThreads = kPrepThreads + 32;              // 8 consumer warps + the producer warp
// stage layout (byte offsets; every part a multiple of 16)
constexpr int kPtOffTid = 0;                               // 4 reads before the chunk + the chunk
constexpr int kPtOffPos = kPtOffTid + 16 + 4 * kPtChunk;
constexpr int kPtOffOff = kPtOffPos + 16 + 4 * kPtChunk;   // kPtChunk + 1 offsets (+ padding)
constexpr int kPtOffFlag = kPtOffOff + 4 * kPtChunk + 16;
constexpr int kPtOffMapq = kPtOffFlag + 2 * kPtChunk;
constexpr int kPtOffCig = kPtOffMapq + kPtChunk;
constexpr int kPtStageBytes = kPtOffCig + 4 * kPtCigCap + 16;   // (+16: unpredicated op loads may look 3 ops past a read's last one)
constexpr int kPtSmemBytes = kPtStages * kPtStageBytes;
static_assert(kPtStageBytes % 16 == 0, "stage size must keep every stage 16-byte aligned");

struct PtMeta {            // what the producer tells the consumers about a stage
  int64_t chunk;           // chunk in the stage (-1: no more)
  uint32_t a0, in_smem;    // first staged op index; the ops are in the stage
  uint32_t nread, last;    // reads of the chunk; the chunk ends the batch
};

// Loads + filter + CIGAR reduction of the 4 reads of one consumer thread, from a landed stage.
// OPS_STAGED: the chunk's ops are in the stage (LDS), else they stayed in global memory.
template <bool OPS_STAGED>
__device__ __forceinline__ PrepReads prep_reduce_staged(const FusedArgs& f, const char* st, const PtMeta m, int64_t i0, int lane,
                                                        const uint32_t* s_lut) {
  const ExpandArgs& a = f.e;
  const uint32_t n_contigs = (uint32_t)a.n_contigs;
  const uint32_t minq = a.filt.min_mapq;
  const int t = threadIdx.x;
  PrepReads R;
  R.nv = min(kPrepPer, max(0, (int)m.nread - kPrepPer * t));
  const int4 t4 = *reinterpret_cast<const int4*>(st + kPtOffTid + 16 + 16 * t);
  const int4 p4 = *reinterpret_cast<const int4*>(st + kPtOffPos + 16 + 16 * t);
  const uint4 o4 = *reinterpret_cast<const uint4*>(st + kPtOffOff + 16 * t);
  const uint32_t o_end = *reinterpret_cast<const uint32_t*>(st + kPtOffOff + 16 * t + 16);
  const ushort4 f4 = *reinterpret_cast<const ushort4*>(st + kPtOffFlag + 8 * t);
  R.T[0] = t4.x; R.T[1] = t4.y; R.T[2] = t4.z; R.T[3] = t4.w;
  R.P[0] = p4.x; R.P[1] = p4.y; R.P[2] = p4.z; R.P[3] = p4.w;
  const uint32_t F[4] = {f4.x, f4.y, f4.z, f4.w};
  // The flag filter (pysam __advance_samtools + bam_plp_push's UNMAP drop) is a function of
  // the flag alone: one bit per flag value in a 4096-bit table (the 12 bits BAM defines), instead of four mask tests
  // on the ALU pipe that bounds this kernel.  A flag with higher bits set is evaluated in full; mapq only when asked.
  uint32_t fpass[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) fpass[r] = (s_lut[(F[r] >> 5) & 127u] >> (F[r] & 31u)) & 1u;
  if (((F[0] | F[1]) | (F[2] | F[3])) > 0xFFFu) {
    const uint32_t drop = (uint32_t)a.filt.flag_filter | 0x4u, req = a.filt.flag_require;
    const uint32_t req_none = req == 0 ? 1u : 0u, orph_mask = a.filt.ignore_orphans ? 3u : 0u;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      fpass[r] = (((F[r] & drop) == 0u) & (((F[r] & req) | req_none) != 0u) & ((F[r] & orph_mask) != 1u)) ? 1u : 0u;
  }
  if (minq) {
    const uchar4 q4 = *reinterpret_cast<const uchar4*>(st + kPtOffMapq + 4 * t);
    const uint32_t Q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) fpass[r] &= Q[r] >= minq ? 1u : 0u;
  }
  uint32_t O[5] = {o4.x, o4.y, o4.z, o4.w, o_end};
  if (R.nv < kPrepPer) {                         // ragged end of the batch: what lies behind it in the stage is not data
    const uint32_t o_last = R.nv == 0 ? m.a0 : R.nv == 1 ? O[1] : R.nv == 2 ? O[2] : O[3];     // = cig_off[n] for nv > 0
#pragma unroll
    for (int r = 0; r < 4; ++r) if (r >= R.nv) { R.T[r] = -1; R.P[r] = 0; O[r + 1] = o_last; if (R.nv == 0) O[r] = o_last; }
  }
  // ops of this chunk: in the stage or left in global memory
  unsigned passm = 0, coop = 0;
  uint32_t nc[4];
  uint32_t nc_max = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const bool p = (r < R.nv) & (fpass[r] != 0u) & ((uint32_t)R.T[r] < n_contigs);
    nc[r] = O[r + 1] - O[r];
    const bool c = p && nc[r] > kThreadOps;
    passm |= p ? (1u << r) : 0u;
    coop |= c ? (1u << r) : 0u;
    if (!p || c) nc[r] = 0;                                           // ops this thread reduces itself
    nc_max = max(nc_max, nc[r]);
  }
  if (OPS_STAGED) {
    // Shared-window byte address of each read's first op.  The loads are NOT predicated -- op k of a read with
    // fewer ops is the next read's op or padding inside the stage, harmless to load -- only the accumulation is,
    // which keeps a (read, op) evaluation at six instructions; rounds 2..4 run only if some lane of the warp needs them.
    const uint32_t cig_s = smem_u32(st + kPtOffCig);
    uint32_t A[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) A[r] = cig_s + ((O[r] - m.a0) << 2);
    auto lds = [](uint32_t addr) -> uint32_t { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; };
#pragma unroll
    for (int r = 0; r < 4; ++r) { const uint32_t v = cigar_ref_len_add(0u, lds(A[r])); R.reflen[r] = nc[r] > 0 ? v : 0u; }
    if (nc_max > 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r) { const uint32_t v = cigar_ref_len_add(R.reflen[r], lds(A[r] + 4)); if (nc[r] > 1) R.reflen[r] = v; }
      if (nc_max > 2) {
#pragma unroll
        for (int r = 0; r < 4; ++r) { const uint32_t v = cigar_ref_len_add(R.reflen[r], lds(A[r] + 8)); if (nc[r] > 2) R.reflen[r] = v; }
        if (nc_max > 3) {
#pragma unroll
          for (int r = 0; r < 4; ++r) { const uint32_t v = cigar_ref_len_add(R.reflen[r], lds(A[r] + 12)); if (nc[r] > 3) R.reflen[r] = v; }
        }
      }
    }
  } else {
    const uint32_t* __restrict__ g_cig = a.cig;
#pragma unroll
    for (int r = 0; r < 4; ++r) R.reflen[r] = cigar_ref_len_add(0u, nc[r] > 0 ? __ldg(g_cig + O[r]) : 0u);
#pragma unroll 1
    for (uint32_t k = 1; k < nc_max; ++k) {
#pragma unroll
      for (int r = 0; r < 4; ++r) R.reflen[r] = cigar_ref_len_add(R.reflen[r], k < nc[r] ? __ldg(g_cig + O[r] + k) : 0u);
    }
  }
  // long CIGARs: the whole warp reduces one read at a time with 128-bit loads from global memory
  while (__any_sync(0xffffffffu, coop != 0)) {
    const unsigned lanes = __ballot_sync(0xffffffffu, coop != 0);
    const int src = __ffs(lanes) - 1;
    const int r = __ffs(__shfl_sync(0xffffffffu, coop, src)) - 1;
    uint32_t ob = r == 0 ? O[0] : r == 1 ? O[1] : r == 2 ? O[2] : O[3];
    const uint32_t oe = r == 0 ? O[1] : r == 1 ? O[2] : r == 2 ? O[3] : O[4];
    uint32_t cnt = oe - ob;
    ob = __shfl_sync(0xffffffffu, ob, src);
    cnt = __shfl_sync(0xffffffffu, cnt, src);
    const unsigned long long v = warp_cigar_reflen_call(a.cig + ob, cnt, lane);
    if (lane == src) {
      const uint32_t v32 = v > 0x7fffffffull ? 0x7fffffffu : (uint32_t)v;
      if (r == 0) R.reflen[0] = v32; else if (r == 1) R.reflen[1] = v32; else if (r == 2) R.reflen[2] = v32; else R.reflen[3] = v32;
      coop &= ~(1u << r);
    }
  }
  if (a.filt.reflen0_as_one) {                     // older htslib: a read that consumes no reference still occupies `pos`
#pragma unroll
    for (int r = 0; r < 4; ++r) if (((passm >> r) & 1u) && R.reflen[r] == 0u) R.reflen[r] = 1u;
  }
  R.passm = passm;
  return R;
}

// prep_emit (k_fused.cuh) for the staged kernel.  Same outputs; the fixed per-thread part of the warp's bookkeeping is
// trimmed, because this kernel is bound by its instruction count (integer pipe): every thread takes its predecessor read
// (pvT, pvP) from the stage instead of by shuffle + lane-0 fix-up, the end-of-batch test is a per-chunk flag, and the rare
// events (tile border, end of batch, far read) share ONE warp vote.
__device__ __forceinline__ void prep_emit_staged(const FusedArgs& f, const int64_t i0, const PrepReads& R, const int32_t pvT,
                                                 const int32_t pvP, const bool is_last, const int lane, PrepWarp& W) {
  const ExpandArgs& a = f.e;
  const uint32_t n_contigs = (uint32_t)a.n_contigs;
  const int32_t* T = R.T;
  const int32_t* P = R.P;
  const uint32_t* reflen = R.reflen;
  const unsigned passm = R.passm;
  const int32_t Tw = __shfl_sync(0xffffffffu, T[0], 0);
  const bool simple = R.nv == kPrepPer && pvT == Tw && T[0] == Tw && T[1] == Tw && T[2] == Tw && T[3] == Tw &&
                      (pvP | P[0] | P[1] | P[2] | P[3]) >= 0;
  if (!(__all_sync(0xffffffffu, simple) && (uint32_t)Tw < n_contigs)) {
    const PrepAcc pa = prep_general(f, i0, R.nv, T[0], T[1], T[2], T[3], P[0], P[1], P[2], P[3], reflen[0], reflen[1],
                                    reflen[2], reflen[3], passm, lane);
    W.n_pass += pa.n_pass; W.aligned += pa.al32; W.max_span = max(W.max_span, pa.max_span); W.unsorted |= pa.unsorted;
    return;
  }
  if (Tw != W.w_tid) {                               // warp-uniform
    W.w_tid = Tw;
    const int64_t base = a.contig_off[Tw];
    W.w_len = (uint32_t)a.contig_len[Tw];
    W.w_tb = (uint32_t)(base >> kTileShift); W.w_bo = (uint32_t)base & (kTile - 1);
  }
  const uint32_t w_len = W.w_len, w_tb = W.w_tb, w_bo = W.w_bo;
  uint32_t q[4], sq[4], rc[4], sp[4];
  uint32_t al32 = 0, n_pass = 0;
  unsigned farmask = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    q[r] = min((uint32_t)P[r], w_len);
    const uint32_t e = min((uint32_t)P[r] + reflen[r], w_len);     // P >= 0 and reflen < 2^31: no wrap
    sp[r] = ((passm >> r) & 1u) ? e - q[r] : 0u;
    sq[r] = w_bo + q[r];
    rc[r] = sp[r] * (uint32_t)kTile + (sq[r] & (kTile - 1));         // {start offset in its tile, span}: one IMAD
    n_pass += sp[r] ? 1u : 0u;
    al32 += sp[r] ? reflen[r] : 0u;
  }
  // far reads (span beyond one tile) are the exception: one test on the largest span, the per-read work only when it fires
  const uint32_t sp_max = max(max(sp[0], sp[1]), max(sp[2], sp[3]));
  if (sp_max > kNearSpan) {
    uint32_t near_max = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const bool far = sp[r] > kNearSpan;
      if (far) { farmask |= 1u << r; rc[r] = (sq[r] & (kTile - 1)) | (kRecFar << kTileShift); }
      near_max = max(near_max, far ? 0u : sp[r]);
    }
    W.max_span = max(W.max_span, near_max);
  } else {
    W.max_span = max(W.max_span, sp_max);
  }
  W.aligned += al32; W.n_pass += n_pass;
  *reinterpret_cast<uint4*>(f.rec + i0) = make_uint4(rc[0], rc[1], rc[2], rc[3]);
  // sorted inside the contig <=> clamped positions never decrease (the predecessor is in the same contig: pvT == Tw)
  const uint32_t pq = min((uint32_t)pvP, w_len);
  W.unsorted |= (q[0] < pq || q[1] < q[0] || q[2] < q[1] || q[3] < q[2]) ? 1u : 0u;
  // tiles relative to the contig's first tile
  const uint32_t t3 = sq[3] >> kTileShift, ptile = (w_bo + pq) >> kTileShift;
  if (__any_sync(0xffffffffu, t3 > ptile || is_last || farmask != 0)) {
    if (__any_sync(0xffffffffu, t3 > ptile || is_last)) {
      // a tile border inside the warp's reads: the first read of every tile entered writes its index
      if (__any_sync(0xffffffffu, t3 - ptile > 4u || is_last)) {       // long gaps / end of batch: warp-cooperative
        prep_tile_boundaries(f.tile_first, i0, kPrepPer, (int64_t)(w_tb + ptile), w_tb + (sq[0] >> kTileShift),
                             w_tb + (sq[1] >> kTileShift), w_tb + (sq[2] >> kTileShift), w_tb + t3, is_last, a.n,
                             (uint32_t)f.n_tiles, lane);
      } else if (t3 > ptile) {
        uint32_t* __restrict__ tf = f.tile_first + w_tb;
        uint32_t prev = ptile;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint32_t tr = sq[r] >> kTileShift;
          for (uint32_t T2 = prev + 1; T2 <= tr; ++T2) tf[T2] = (uint32_t)(i0 + r);
          prev = max(prev, tr);
        }
      }
    }
    if (__any_sync(0xffffffffu, farmask != 0)) {
      const int64_t base = a.contig_off[Tw];
      int64_t e64[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) e64[r] = base + min((uint32_t)P[r] + reflen[r], w_len);
      prep_far(f, w_tb + (sq[0] >> kTileShift), w_tb + (sq[1] >> kTileShift), w_tb + (sq[2] >> kTileShift), w_tb + t3, e64[0],
               e64[1], e64[2], e64[3], farmask, lane);
    }
  }
}

// The producer: lane 0 of the last warp.  Chunk c = reads [c*1024, min(n, (c+1)*1024)).  A CTA's first two chunks
// are blockIdx.x and blockIdx.x + gridDim.x; after that chunks are drawn from a ticket counter (one atomic per
// chunk, issued two chunks ahead of its use): SMs do not run at the same speed, and with a fixed stride the
// slowest ones kept the kernel alive for ~2 us (ncu: max - mean of sm__cycles_active) after the others had finished.
__device__ __forceinline__ void prep_producer(const FusedArgs& f, char* smem, uint64_t* full, uint64_t* empty, PtMeta* meta,
                                              int64_t n_chunks) {
  const ExpandArgs& a = f.e;
  const int64_t n = a.n;
  const uint32_t* __restrict__ g_off = a.cig_off;
  const uint32_t cig_total = __ldg(g_off + n);
  const uint64_t pol = l2_policy_evict_first();
  const unsigned G = gridDim.x;
  int64_t c = blockIdx.x, cn = (int64_t)blockIdx.x + G;
  uint32_t ob = 0, oe = 0;
  if (c < n_chunks) { ob = __ldg(g_off + c * kPtChunk); oe = __ldg(g_off + min((c + 1) * (int64_t)kPtChunk, n)); }
  unsigned it = 0;
#pragma unroll 1
  for (;; ++it) {
    // the chunk after next (ticket) and the op range of the next chunk: fetched now, needed one iteration from now
    const int64_t cnn = 2 * (int64_t)G + atomicAdd(&a.pc->ticket, 1u);
    uint32_t nob = 0, noe = 0;
    if (cn < n_chunks) { nob = __ldg(g_off + cn * kPtChunk); noe = __ldg(g_off + min((cn + 1) * (int64_t)kPtChunk, n)); }
    const unsigned s = it % kPtStages, k = it / kPtStages;
    if (k > 0) mbar_wait_relaxed(&empty[s], (k - 1) & 1);          // the consumers have released this stage
    if (c >= n_chunks) {                                            // no more chunks: tell the consumers
      meta[s].chunk = -1; meta[s].a0 = 0; meta[s].in_smem = 0; meta[s].nread = 0; meta[s].last = 0;
      mbar_arrive(&full[s]);
      break;
    }
    char* st = smem + (size_t)s * kPtStageBytes;
    const int64_t r0 = c * kPtChunk;
    const uint32_t nread = (uint32_t)min((int64_t)kPtChunk, n - r0);
    const uint32_t n4 = nread & ~3u, n8 = nread & ~7u, n16 = nread & ~15u;   // reads covered by whole 16-byte units
    const uint32_t prev = c > 0 ? 16u : 0u;                        // the four reads before the chunk (sortedness, tile border)
    // CIGAR ops [ob, oe) -> aligned window [a0, a1); never past the last whole vector of the array
    const uint32_t a0 = ob & ~3u;
    uint32_t a1 = (oe + 3u) & ~3u;
    const bool fits = a1 - a0 <= (uint32_t)kPtCigCap;
    a1 = min(a1, cig_total & ~3u);
    const uint32_t cig_bytes = (fits && a1 > a0) ? 4u * (a1 - a0) : 0u;
    // what whole 16-byte units do not cover (only at the ragged end of the batch) is copied by hand, BEFORE
    // the arrival below publishes the stage
    for (uint32_t r = n4; r < nread; ++r) {
      reinterpret_cast<int32_t*>(st + kPtOffTid + 16)[r] = a.tid[r0 + r];
      reinterpret_cast<int32_t*>(st + kPtOffPos + 16)[r] = a.pos[r0 + r];
      reinterpret_cast<uint32_t*>(st + kPtOffOff)[r] = g_off[r0 + r];
    }
    for (uint32_t r = n8; r < nread; ++r) reinterpret_cast<uint16_t*>(st + kPtOffFlag)[r] = a.flag[r0 + r];
    for (uint32_t r = n16; r < nread; ++r) reinterpret_cast<uint8_t*>(st + kPtOffMapq)[r] = a.mapq[r0 + r];
    reinterpret_cast<uint32_t*>(st + kPtOffOff)[nread] = oe;        // offsets: one entry more than reads
    if (fits) for (uint32_t o = max(a1, a0); o < oe; ++o) reinterpret_cast<uint32_t*>(st + kPtOffCig)[o - a0] = a.cig[o];
    meta[s].chunk = c; meta[s].a0 = a0; meta[s].in_smem = fits ? 1u : 0u; meta[s].nread = nread; meta[s].last = (r0 + nread == n) ? 1u : 0u;
    const uint32_t tx = 2u * (prev + 4u * n4) + 4u * n4 + 2u * n8 + n16 + cig_bytes;
    mbar_arrive_expect_tx(&full[s], tx);
    if (prev + n4) {
      tma_load_1d_hint(st + kPtOffTid + 16 - prev, a.tid + r0 - (prev >> 2), prev + 4u * n4, &full[s], pol);
      tma_load_1d_hint(st + kPtOffPos + 16 - prev, a.pos + r0 - (prev >> 2), prev + 4u * n4, &full[s], pol);
    }
    if (n4) tma_load_1d_hint(st + kPtOffOff, g_off + r0, 4u * n4, &full[s], pol);
    if (n8) tma_load_1d_hint(st + kPtOffFlag, a.flag + r0, 2u * n8, &full[s], pol);
    if (n16) tma_load_1d_hint(st + kPtOffMapq, a.mapq + r0, n16, &full[s], pol);
    if (cig_bytes) tma_load_1d_hint(st + kPtOffCig, a.cig + a0, cig_bytes, &full[s], pol);
    c = cn; cn = cnn; ob = nob; oe = noe;
  }
}

__global__ void __launch_bounds__(kPtThreads, MCOV_PREP_CTAS)
k_fused_prep_tma(const __grid_constant__ FusedArgs f) {
  extern __shared__ __align__(128) char pt_smem[];
  __shared__ __align__(8) uint64_t s_full[kPtStages], s_empty[kPtStages];
  __shared__ PtMeta s_meta[kPtStages];
  __shared__ uint32_t s_lut[128];
  if (threadIdx.x < 128) s_lut[threadIdx.x] = f.flag_lut[threadIdx.x];
  pdl_launch_dependents();                                    // k_scan_counts may take free slots as this grid drains
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kPtStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kPrepThreads / 32); }
    mbar_fence_init();
  }
  __syncthreads();
  const int64_t n = f.e.n;
  const int64_t n_chunks = (n + kPtChunk - 1) / kPtChunk;
  if (warp == kPrepThreads / 32) {                             // producer warp
    if (lane == 0) prep_producer(f, pt_smem, s_full, s_empty, s_meta, n_chunks);
    return;
  }
  PrepWarp W = {0ull, 0u, 0u, 0u, -1, 0u, 0u, 0u};
  unsigned s = 0, ph = 0;                                       // stage of this iteration and the parity of its "full" phase
#pragma unroll 1
  for (;; s = (s + 1 == kPtStages) ? 0u : s + 1, ph ^= (s == 0) ? 1u : 0u) {
    mbar_wait(&s_full[s], ph);                                  // the chunk has landed
    const char* st = pt_smem + (size_t)s * kPtStageBytes;
    const PtMeta m = s_meta[s];
    if (m.chunk < 0) break;
    const int64_t c = m.chunk;
    const int64_t i0 = c * kPtChunk + (int64_t)threadIdx.x * kPrepPer;
    // every thread's predecessor read, straight from the stage (the four reads before the chunk lead it); the very first
    // read of the batch has none
    int32_t pvT = reinterpret_cast<const int32_t*>(st + kPtOffTid + 16)[(int)threadIdx.x * kPrepPer - 1];
    const int32_t pvP = reinterpret_cast<const int32_t*>(st + kPtOffPos + 16)[(int)threadIdx.x * kPrepPer - 1];
    if (i0 == 0) pvT = -1;
    const bool is_last = m.last != 0u && (uint32_t)(threadIdx.x * kPrepPer + kPrepPer) == m.nread;
    const PrepReads R = m.in_smem ? prep_reduce_staged<true>(f, st, m, i0, lane, s_lut) : prep_reduce_staged<false>(f, st, m, i0, lane, s_lut);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[s]);                    // this warp is done with the stage
    prep_emit_staged(f, i0, R, pvT, pvP, is_last, lane, W);
  }
  prep_flush_counters<true>(f.e.pc, W.n_pass, W.aligned, W.unsorted, W.max_span);
}

}  // namespace mcov
