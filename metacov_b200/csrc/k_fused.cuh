// k_fused.cuh -- fused depth path for coordinate-sorted reads.
//
// Same result as K0 (clear) + K1 (expand, k_expand.cuh) + K2 (scan, k_scan.cuh)
// -- the per-base `column.n` of reference metacov/pileup.py:13-16 -- but the
// slot space is written exactly once and never read, no atomic ever reaches
// L2 per read, and no CTA waits on another:
//
//   k_fused_prep   one pass over the reads, 4 consecutive reads per thread with
//                  128-bit loads: filter + CIGAR reduce, emits
//                  rec[i] = {low 32 bits of the start slot, clipped span}
//                  (span 0 = read contributes nothing).  Because the reads are
//                  sorted it also produces, with warp-aggregated updates,
//                    tile_first[T]  first read whose start slot >= T*kTile
//                    tile_agg[T]    (#starts - #ends) falling in tile T
//                  and, for reads whose span exceeds kNearSpan ("far" reads:
//                  long reads, spliced reads), the end slot list + per-tile
//                  far-end counts.
//   k_scan_inplace one small scan over [tile_agg | far counts]: the inclusive
//                  sum of tile_agg up to T-1 IS the depth entering tile T, so
//                  the tile kernel needs no look-back and no ordering.
//   k_far_scatter  bucket the far ends by tile (counting sort).
//   k_fused_tile   one CTA per tile of kTile slots: +1/-1 of the tile's reads
//                  go to SHARED-memory counters (starts and ends kept apart),
//                  the ends of near reads that started before the tile are
//                  found by walking back at most max_span slots in the sorted
//                  order, far ends come from the tile's bucket; block scan
//                  from the known carry; 128-bit streaming stores.  Keeping
//                  starts and ends apart gives htslib's max_depth no-op
//                  condition exactly: cap[p] = depth[p-1] + starts[p]
//                  = depth[p] + ends[p].
//
// HBM bytes (algorithmic): prep 15R + 4*sum(n_cigar of passing reads) + 8R;
// tile 8R + 4(L+C).
#pragma once
#include "ctx.cuh"
#include "k_expand.cuh"
#include "k_scan.cuh"

namespace mcov {

constexpr int kTile = 4096;               // slots per CTA of the tile kernel
constexpr int kTileShift = 12;
constexpr uint32_t kNearSpan = kTile;     // spans above this take the bucket path
constexpr int kFusedThreads = 256;
constexpr int kPrepThreads = 256;
constexpr int kPrepPer = 4;               // reads per thread
constexpr uint32_t kFarCapDefault = 1u << 26;

struct FusedArgs {
  ExpandArgs e;
  uint2* rec;                 // [n] {key_lo, span}
  int64_t n_slots;
  int64_t n_tiles;
  int64_t* far_end;           // [far_cap] end slots of far reads
  uint32_t far_cap;
  int32_t* tile_agg;          // [cnt_pad] (#starts - #ends) per tile -> inclusive scan in place
  uint32_t* tile_cnt;         // [cnt_pad] far ends per tile -> inclusive scan in place (follows tile_agg)
  uint32_t* tile_cursor;      // [n_tiles] scatter cursors
  uint32_t* far_sorted;       // [far_cap] in-tile offsets bucketed by tile
  int64_t* tile_first;        // [n_tiles+1]
  int32_t* depth;
  int vec_ok;                 // SoA base pointers aligned for 128-bit loads
};

struct Key { int64_t key; int64_t len_off; };

// slot key of a read: contig offset + clamped position; reads without a valid
// contig sort after every slot
__device__ __forceinline__ int64_t slot_key(const ExpandArgs& a, int32_t t, int32_t p, int64_t n_slots, int64_t& len,
                                            int64_t& base) {
  if (t < 0 || t >= a.n_contigs) { len = 0; base = n_slots; return n_slots; }
  len = a.contig_len[t];
  base = a.contig_off[t];
  int64_t q = p < 0 ? 0 : (p > len ? len : (int64_t)p);
  return base + q;
}

// fill tile_first[lo..hi] = v; long gaps are written by the whole warp
__device__ __forceinline__ void fill_tile_first(int64_t* tile_first, int64_t lo, int64_t hi, int64_t v, int lane,
                                                unsigned active) {
  // short gaps inline
  bool big = (hi - lo) >= 8;
  if (!big) for (int64_t T = lo; T <= hi; ++T) tile_first[T] = v;
  unsigned todo = __ballot_sync(active, big);
  while (todo) {
    int src = __ffs(todo) - 1;
    todo &= todo - 1;
    int64_t l2 = __shfl_sync(active, lo, src), h2 = __shfl_sync(active, hi, src), v2 = __shfl_sync(active, v, src);
    for (int64_t T = l2 + lane; T <= h2; T += 32) tile_first[T] = v2;
  }
}

__global__ void __launch_bounds__(kPrepThreads)
k_fused_prep(FusedArgs f) {
  const ExpandArgs& a = f.e;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long n_pass = 0, aligned = 0;
  int unsorted = 0;
  uint32_t max_span = 0;
  const int64_t n = a.n;
  const int64_t n_groups = (n + kPrepPer - 1) / kPrepPer;
  const int64_t g_round = (n_groups + 31) & ~(int64_t)31;     // whole warps iterate together
  const int64_t g_stride = (int64_t)gridDim.x * kPrepThreads;
  const int64_t last_tile = f.n_tiles;                        // tile_first has n_tiles+1 entries

  for (int64_t g = (int64_t)blockIdx.x * kPrepThreads + threadIdx.x; g < g_round; g += g_stride) {
    const int64_t i0 = g * kPrepPer;
    int32_t T[4], P[4];
    uint32_t F[4], Q[4], O[5];
    int nv = (int)min((int64_t)kPrepPer, max((int64_t)0, n - i0));     // valid reads of this thread
    if (nv == kPrepPer && f.vec_ok) {
      int4 t4 = *reinterpret_cast<const int4*>(a.tid + i0);
      int4 p4 = *reinterpret_cast<const int4*>(a.pos + i0);
      ushort4 f4 = *reinterpret_cast<const ushort4*>(a.flag + i0);
      uchar4 q4 = *reinterpret_cast<const uchar4*>(a.mapq + i0);
      uint4 o4 = *reinterpret_cast<const uint4*>(a.cig_off + i0);
      O[4] = a.cig_off[i0 + 4];
      T[0] = t4.x; T[1] = t4.y; T[2] = t4.z; T[3] = t4.w;
      P[0] = p4.x; P[1] = p4.y; P[2] = p4.z; P[3] = p4.w;
      F[0] = f4.x; F[1] = f4.y; F[2] = f4.z; F[3] = f4.w;
      Q[0] = q4.x; Q[1] = q4.y; Q[2] = q4.z; Q[3] = q4.w;
      O[0] = o4.x; O[1] = o4.y; O[2] = o4.z; O[3] = o4.w;
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        bool v = r < nv;
        T[r] = v ? a.tid[i0 + r] : -1;
        P[r] = v ? a.pos[i0 + r] : 0;
        F[r] = v ? (uint32_t)a.flag[i0 + r] : 0x4u;
        Q[r] = v ? (uint32_t)a.mapq[i0 + r] : 0u;
      }
      // padding reads get an empty CIGAR: their offsets all equal cig_off[n]
#pragma unroll
      for (int r = 0; r <= 4; ++r) O[r] = (i0 <= n) ? a.cig_off[min(i0 + r, n)] : 0u;
    }
    // previous read's (tid,pos): from the neighbouring lane, lane 0 reloads
    int32_t pt = __shfl_up_sync(0xffffffffu, T[3], 1), pp = __shfl_up_sync(0xffffffffu, P[3], 1);
    if (lane == 0) {
      if (i0 > 0 && i0 - 1 < n) { pt = a.tid[i0 - 1]; pp = a.pos[i0 - 1]; }
      else { pt = INT_MIN; pp = INT_MIN; }                        // INT_MIN: "no previous read"
    }

    bool pass[4];
    unsigned long long reflen[4];
    bool coop[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      pass[r] = r < nv && read_passes(F[r], Q[r], a.filt) && T[r] >= 0 && T[r] < a.n_contigs;
      uint32_t nc = O[r + 1] - O[r];
      coop[r] = pass[r] && nc > kThreadOps;
      reflen[r] = 0;
    }
    // first op of every short CIGAR: four independent loads
    uint32_t op0[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) op0[r] = (pass[r] && !coop[r] && O[r + 1] > O[r]) ? __ldg(a.cig + O[r]) : 0u;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      reflen[r] = cigar_ref_len(op0[r]);
      if (pass[r] && !coop[r])
        for (uint32_t k = O[r] + 1; k < O[r + 1]; ++k) reflen[r] += cigar_ref_len(__ldg(a.cig + k));
    }
    // long CIGARs: the whole warp reduces one read at a time with 128-bit loads
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      unsigned todo = __ballot_sync(0xffffffffu, coop[r]);
      while (todo) {
        int src = __ffs(todo) - 1;
        todo &= todo - 1;
        uint32_t b0 = __shfl_sync(0xffffffffu, O[r], src), b1 = __shfl_sync(0xffffffffu, O[r + 1], src);
        unsigned long long v = warp_cigar_reflen(a.cig, b0, b1, lane, a.cig_aligned16 != 0);
        if (lane == src) reflen[r] = v;
      }
    }

    // slot keys, clipped intervals, records
    int64_t key[4], endk[4];
    uint32_t span[4];
    int64_t c_len = 0, c_base = 0; int32_t c_tid = INT_MIN;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (T[r] != c_tid) { c_tid = T[r]; (void)slot_key(a, T[r], 0, f.n_slots, c_len, c_base); }
      bool valid_t = T[r] >= 0 && T[r] < a.n_contigs;
      int64_t q = P[r] < 0 ? 0 : (P[r] > c_len ? c_len : (int64_t)P[r]);
      key[r] = valid_t ? c_base + q : f.n_slots;
      span[r] = 0; endk[r] = key[r];
      if (pass[r]) {
        int64_t e = (int64_t)P[r] + (int64_t)reflen[r];
        e = e < 0 ? 0 : (e > c_len ? c_len : e);
        if (e > q) { span[r] = (uint32_t)(e - q); endk[r] = c_base + e; n_pass += 1; aligned += reflen[r]; }
      }
    }
    if (nv == kPrepPer) {
      uint4* out = reinterpret_cast<uint4*>(f.rec + i0);
      out[0] = make_uint4((uint32_t)key[0], span[0], (uint32_t)key[1], span[1]);
      out[1] = make_uint4((uint32_t)key[2], span[2], (uint32_t)key[3], span[3]);
    } else {
      for (int r = 0; r < nv; ++r) f.rec[i0 + r] = make_uint2((uint32_t)key[r], span[r]);
    }

    // sortedness + tile boundaries (tile_first)
    {
      int32_t qt = pt, qp = pp;
      int64_t plen, pbase;
      int64_t prev_tile = (pt == INT_MIN) ? -1 : (slot_key(a, pt, pp, f.n_slots, plen, pbase) >> kTileShift);
      // per-thread pending fill (at most one long gap per read; short ones inline)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        bool v = r < nv;
        if (v && qt != INT_MIN) {
          uint32_t u0 = (uint32_t)qt, u1 = (uint32_t)T[r];
          if (u1 < u0 || (u1 == u0 && P[r] < qp)) unsorted = 1;
        }
        int64_t t_cur = v ? min(key[r] >> kTileShift, last_tile) : prev_tile;
        int64_t lo = prev_tile + 1, hi = t_cur;
        bool need = v && hi >= lo;
        unsigned active = 0xffffffffu;
        fill_tile_first(f.tile_first, need ? lo : 1, need ? hi : 0, i0 + r, lane, active);
        if (v) { prev_tile = max(prev_tile, t_cur); qt = T[r]; qp = P[r]; }
      }
      // the last read closes the table
      bool is_last = nv > 0 && (i0 + nv == n);
      fill_tile_first(f.tile_first, is_last ? prev_tile + 1 : 1, is_last ? last_tile : 0, n, lane, 0xffffffffu);
    }

    // far reads: end list + per-tile far counts (warp-aggregated slot claim); near: max span
    {
      int nfar = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (span[r] > kNearSpan) ++nfar;
        else max_span = max(max_span, span[r]);
      }
      int incl = nfar;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&a.pc->n_far, (uint32_t)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint32_t idx = base + (uint32_t)(incl - nfar);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (span[r] > kNearSpan) {
            if (idx < f.far_cap) {
              f.far_end[idx] = endk[r];
              atomicAdd(f.tile_cnt + (endk[r] >> kTileShift), 1u);
              atomicAdd(f.tile_agg + (endk[r] >> kTileShift), -1);
            }
            ++idx;
          }
        }
      }
    }

    // tile_agg: +1 per start, -1 per near end, aggregated across the warp per distinct tile
    {
      int64_t pend_t[8];
      int pend_v[8];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        bool c = span[r] > 0;
        pend_t[r] = c ? (key[r] >> kTileShift) : INT64_MAX;
        pend_v[r] = 1;
        bool nearr = c && span[r] <= kNearSpan;
        pend_t[4 + r] = nearr ? (endk[r] >> kTileShift) : INT64_MAX;
        pend_v[4 + r] = -1;
      }
      for (int iter = 0; iter < 4; ++iter) {
        int64_t m = INT64_MAX;
#pragma unroll
        for (int k = 0; k < 8; ++k) m = min(m, pend_t[k]);
        // warp minimum of a 64-bit value: high then low word
        unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)((unsigned long long)m >> 32));
        unsigned lo = __reduce_min_sync(0xffffffffu, ((unsigned)((unsigned long long)m >> 32) == hi) ? (unsigned)m : 0xffffffffu);
        int64_t Tm = (int64_t)(((unsigned long long)hi << 32) | lo);
        if (Tm == INT64_MAX) break;
        int local = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (pend_t[k] == Tm) { local += pend_v[k]; pend_t[k] = INT64_MAX; }
        int v = __reduce_add_sync(0xffffffffu, local);
        if (lane == 0 && v != 0) atomicAdd(f.tile_agg + Tm, v);
      }
      // anything still pending (reads of this warp spread over many tiles): direct updates
#pragma unroll
      for (int k = 0; k < 8; ++k) if (pend_t[k] != INT64_MAX) atomicAdd(f.tile_agg + pend_t[k], pend_v[k]);
    }
  }

  // block-level reduction of the pass counters
  n_pass = warp_sum(n_pass);
  aligned = warp_sum(aligned);
  unsorted = __any_sync(0xffffffffu, unsorted);
  max_span = (uint32_t)warp_max((int)max_span);
  __shared__ unsigned long long s_np[kPrepThreads / 32], s_al[kPrepThreads / 32];
  __shared__ int s_un[kPrepThreads / 32];
  __shared__ uint32_t s_ms[kPrepThreads / 32];
  if (lane == 0) { s_np[warp] = n_pass; s_al[warp] = aligned; s_un[warp] = unsorted; s_ms[warp] = max_span; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long np = 0, al = 0; int un = 0; uint32_t ms = 0;
    for (int k = 0; k < kPrepThreads / 32; ++k) { np += s_np[k]; al += s_al[k]; un |= s_un[k]; ms = max(ms, s_ms[k]); }
    if (np) atomicAdd(&a.pc->n_pass, np);
    if (al) atomicAdd(&a.pc->aligned_bases, al);
    if (un) atomicOr(&a.pc->unsorted, 1);
    if (ms) atomicMax(&a.pc->max_span, ms);
  }
}

// bucket far ends by tile; tile_cnt holds the INCLUSIVE scan of the per-tile counts.  Because the
// scan runs over [tile_agg | tile_cnt] as one array and tile_agg sums to zero, tile_cnt's running
// sum starts from zero.
__global__ void k_far_scatter(FusedArgs f) {
  uint32_t n_far = min(f.e.pc->n_far, f.far_cap);
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_far; k += gridDim.x * blockDim.x) {
    int64_t e = f.far_end[k];
    int64_t T = e >> kTileShift;
    uint32_t base = T > 0 ? f.tile_cnt[T - 1] : 0u;
    uint32_t p = atomicAdd(f.tile_cursor + T, 1u);
    f.far_sorted[base + p] = (uint32_t)(e - (T << kTileShift));
  }
}

__global__ void __launch_bounds__(kFusedThreads, 5)
k_fused_tile(FusedArgs f) {
  __shared__ __align__(16) int s_start[kTile];
  __shared__ __align__(16) int s_end[kTile];
  __shared__ int s_warp[kFusedThreads / 32];
  __shared__ int s_warp2[kFusedThreads / 32];
  PassCounters* pc = f.e.pc;
  const int64_t tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // tile metadata: independent loads issued before the shared memory is cleared
  const int64_t r0 = f.tile_first[tile], r1 = f.tile_first[tile + 1];
  const int64_t jmin = tile > 0 ? f.tile_first[tile - 1] : 0;
  const int carry = tile > 0 ? f.tile_agg[tile - 1] : 0;           // depth entering the tile
  const uint32_t k0 = tile > 0 ? f.tile_cnt[tile - 1] : 0u, k1 = f.tile_cnt[tile];
  const uint32_t reach = pc->max_span;
  {
    int4* z0 = reinterpret_cast<int4*>(s_start);
    int4* z1 = reinterpret_cast<int4*>(s_end);
    for (int k = threadIdx.x; k < kTile / 4; k += kFusedThreads) { z0[k] = make_int4(0, 0, 0, 0); z1[k] = make_int4(0, 0, 0, 0); }
  }
  __syncthreads();
  const int64_t base = tile << kTileShift;
  const uint32_t base_lo = (uint32_t)base;

  // reads that start in this tile
  for (int64_t j = r0 + threadIdx.x; j < r1; j += kFusedThreads) {
    uint2 r = f.rec[j];
    if (r.y) {
      uint32_t local = r.x - base_lo;                  // < kTile for sorted input (tile_first)
      if (local < (uint32_t)kTile) {                   // guard: unsorted input must not corrupt smem
        atomicAdd(&s_start[local], 1);
        uint32_t el = local + r.y;
        if (r.y <= kNearSpan && el < (uint32_t)kTile) atomicAdd(&s_end[el], 1);
      }
    }
  }
  // near reads that started before the tile and end inside it: walk back while the start is within
  // max_span of the tile (sorted order => monotone distance).  reach <= kNearSpan = kTile, so every
  // candidate started in the previous tile; staying inside it keeps the 32-bit distance from wrapping.
  for (int64_t j = r0 - 1 - threadIdx.x; j >= jmin; j -= kFusedThreads) {
    uint2 r = f.rec[j];
    uint32_t d = base_lo - r.x;                        // distance behind the tile start (>= 1)
    if (d > reach) break;
    if (r.y >= d && r.y <= kNearSpan) {
      uint32_t el = r.y - d;
      if (el < (uint32_t)kTile) atomicAdd(&s_end[el], 1);
    }
  }
  // far ends bucketed for this tile
  for (uint32_t k = k0 + threadIdx.x; k < k1; k += kFusedThreads) atomicAdd(&s_end[f.far_sorted[k] & (kTile - 1)], 1);
  __syncthreads();

  // block scan of (starts - ends), warp-striped like k_scan_inplace
  const int4* vs = reinterpret_cast<const int4*>(s_start);
  const int4* ve = reinterpret_cast<const int4*>(s_end);
  int4 v[kScanVec], en[kScanVec];
  int run[kScanVec];
#pragma unroll
  for (int j = 0; j < kScanVec; ++j) {
    int idx = (warp * kScanVec + j) * 32 + lane;
    int4 st = vs[idx];
    en[j] = ve[idx];
    v[j].x = st.x - en[j].x;
    v[j].y = v[j].x + st.y - en[j].y;
    v[j].z = v[j].y + st.z - en[j].z;
    v[j].w = v[j].z + st.w - en[j].w;
    run[j] = v[j].w;
  }
  int acc = 0;
#pragma unroll
  for (int j = 0; j < kScanVec; ++j) {
    int x = run[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    int total = __shfl_sync(0xffffffffu, x, 31);
    run[j] = x - run[j] + acc;
    acc += total;
  }
  if (lane == 31) s_warp[warp] = acc;
  __syncthreads();
  int off = carry;
#pragma unroll
  for (int k = 0; k < kFusedThreads / 32; ++k) off += (k < warp) ? s_warp[k] : 0;
  int4* out = reinterpret_cast<int4*>(f.depth + base);
  const int64_t n_vec = (f.n_slots - base) >> 2;
  int mx = 0, cap = 0;
#pragma unroll
  for (int j = 0; j < kScanVec; ++j) {
    int idx = (warp * kScanVec + j) * 32 + lane;
    int o = off + run[j];
    v[j].x += o; v[j].y += o; v[j].z += o; v[j].w += o;
    mx = max(mx, max(max(v[j].x, v[j].y), max(v[j].z, v[j].w)));
    cap = max(cap, max(max(v[j].x + en[j].x, v[j].y + en[j].y), max(v[j].z + en[j].z, v[j].w + en[j].w)));
    if (idx < n_vec) st_stream_int4(out + idx, v[j]);
  }
  mx = warp_max(mx);
  cap = warp_max(cap);
  __syncthreads();
  if (lane == 0) { s_warp[warp] = mx; s_warp2[warp] = cap; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0, c2 = 0;
    for (int k = 0; k < kFusedThreads / 32; ++k) { m = max(m, s_warp[k]); c2 = max(c2, s_warp2[k]); }
    // only touch the global maxima when this tile can raise them (stale reads only cost an extra atomic)
    if (m > *((volatile int*)&pc->max_depth_seen)) atomicMax(&pc->max_depth_seen, m);
    if (c2 > *((volatile int*)&pc->cap_metric)) atomicMax(&pc->cap_metric, c2);
  }
}

}  // namespace mcov
