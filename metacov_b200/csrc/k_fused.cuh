// k_fused.cuh -- fused depth path for coordinate-sorted reads.
//
// Same result as K0 (clear) + K1 (expand, k_expand.cuh) + K2 (scan, k_scan.cuh)
// -- the per-base `column.n` of reference metacov/pileup.py:13-16 -- but the
// slot space is written exactly once and never read:
//
//   k_fused_prep   one pass over the reads: filter + CIGAR reduce (shared with
//                  K1), emits rec[i] = {low 32 bits of the start slot, clipped
//                  span}; span 0 = read contributes nothing.  Reads whose span
//                  exceeds kNearSpan ("far" reads: long reads, spliced reads)
//                  also append their end slot to a list and bump the counter
//                  of the tile their end falls in.
//   k_tile_first   tile_first[T] = first read whose start slot >= T*kTile
//                  (binary search over the sorted (tid,pos) keys).
//   k_far_scatter  bucket the far ends by tile (counting sort, order inside a
//                  tile is irrelevant).
//   k_fused_tile   one CTA per tile of kTile slots: +1/-1 of the tile's reads
//                  go to SHARED-memory counters (starts and ends kept apart),
//                  the ends of near reads that started before the tile are
//                  found by walking back at most max_span slots in the sorted
//                  order, far ends come from the tile's bucket; then a block
//                  scan + decoupled look-back carry, and the depth of the tile
//                  is written with 128-bit stores.  Keeping starts and ends
//                  apart gives htslib's max_depth no-op condition exactly:
//                  cap[p] = depth[p-1] + starts[p] = depth[p] + ends[p]
//                  (SURVEY.md Appendix A-6).
//
// HBM bytes (algorithmic): prep 15R + 4*sum(n_cigar of passing reads) + 8R;
// tile 8R + 4(L+C).
#pragma once
#include "ctx.cuh"
#include "k_expand.cuh"
#include "k_scan.cuh"

namespace mcov {

constexpr int kTile = kScanTile;          // 4096 slots per CTA
constexpr uint32_t kNearSpan = kTile;     // spans above this take the bucket path
constexpr int kFusedThreads = kScanThreads;
constexpr uint32_t kFarCapDefault = 1u << 26;

struct FusedArgs {
  ExpandArgs e;
  uint2* rec;                 // [n] {key_lo, span}
  int64_t n_slots;
  int64_t n_tiles;
  int64_t* far_end;           // [far_cap] end slots of far reads
  uint32_t far_cap;
  uint32_t* tile_cnt;         // [n_tiles(+pad)] far ends per tile -> inclusive scan in place
  uint32_t* tile_cursor;      // [n_tiles] scatter cursors
  uint32_t* far_sorted;       // [far_cap] in-tile offsets bucketed by tile
  int64_t* tile_first;        // [n_tiles+1]
  unsigned long long* status; // look-back status words
  int32_t* depth;
};

__device__ __forceinline__ int64_t read_key64(const ExpandArgs& a, int64_t i, int64_t n_slots) {
  int32_t t = a.tid[i];
  if (t < 0 || t >= a.n_contigs) return n_slots;      // sorts after every slot
  int64_t len = a.contig_len[t], p = a.pos[i];
  p = p < 0 ? 0 : (p > len ? len : p);
  return a.contig_off[t] + p;
}

__global__ void __launch_bounds__(kExpandThreads)
k_fused_prep(FusedArgs f) {
  const ExpandArgs& a = f.e;
  const int lane = threadIdx.x & 31;
  unsigned long long n_pass = 0, aligned = 0;
  int unsorted = 0;
  uint32_t max_span = 0;
  const int64_t stride = (int64_t)gridDim.x * kExpandThreads;
  const int64_t n_round = (a.n + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * kExpandThreads + threadIdx.x; i < n_round; i += stride) {
    bool in_range = i < a.n;
    int64_t s = 0, e = 0;
    unsigned long long reflen;
    bool ok = expand_one(a, i, in_range, lane, s, e, reflen);
    if (in_range) {
      uint2 r;
      if (ok) {
        uint32_t span = (uint32_t)(e - s);
        r.x = (uint32_t)s; r.y = span;
        n_pass += 1; aligned += reflen;
        if (span <= kNearSpan) max_span = max(max_span, span);
        else {
          uint32_t idx = atomicAdd(&a.pc->n_far, 1u);
          if (idx < f.far_cap) {
            f.far_end[idx] = e;
            atomicAdd(f.tile_cnt + e / kTile, 1u);
          }
        }
      } else {
        r.x = (uint32_t)read_key64(a, i, f.n_slots); r.y = 0u;
      }
      f.rec[i] = r;
      if (i > 0) {
        uint32_t u0 = (uint32_t)a.tid[i - 1], u1 = (uint32_t)a.tid[i];
        if (u1 < u0 || (u1 == u0 && a.pos[i] < a.pos[i - 1])) unsorted = 1;
      }
    }
  }
  n_pass = warp_sum(n_pass);
  aligned = warp_sum(aligned);
  unsorted = __any_sync(0xffffffffu, unsorted);
  max_span = (uint32_t)warp_max((int)max_span);
  if (lane == 0) {
    if (n_pass) atomicAdd(&a.pc->n_pass, n_pass);
    if (aligned) atomicAdd(&a.pc->aligned_bases, aligned);
    if (unsorted) atomicOr(&a.pc->unsorted, 1);
    if (max_span) atomicMax(&a.pc->max_span, max_span);
  }
}

// tile_first[T] = lower_bound over reads of key64 >= T*kTile, T in [0, n_tiles]
__global__ void k_tile_first(FusedArgs f) {
  int64_t T = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (T > f.n_tiles) return;
  if (T == f.n_tiles) { f.tile_first[T] = f.e.n; return; }
  const int64_t want = T * kTile;
  int64_t lo = 0, hi = f.e.n;                 // first index with key >= want
  while (lo < hi) {
    int64_t mid = lo + ((hi - lo) >> 1);
    if (read_key64(f.e, mid, f.n_slots) < want) lo = mid + 1; else hi = mid;
  }
  f.tile_first[T] = lo;
}

// bucket far ends by tile; tile_cnt holds the INCLUSIVE scan of the per-tile counts
__global__ void k_far_scatter(FusedArgs f) {
  uint32_t n_far = min(f.e.pc->n_far, f.far_cap);
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_far; k += gridDim.x * blockDim.x) {
    int64_t e = f.far_end[k];
    int64_t T = e / kTile;
    uint32_t base = T > 0 ? f.tile_cnt[T - 1] : 0u;
    uint32_t p = atomicAdd(f.tile_cursor + T, 1u);
    f.far_sorted[base + p] = (uint32_t)(e - T * kTile);
  }
}

__global__ void __launch_bounds__(kFusedThreads)
k_fused_tile(FusedArgs f) {
  __shared__ __align__(16) int s_start[kTile];
  __shared__ __align__(16) int s_end[kTile];
  __shared__ int64_t s_tile;
  __shared__ int s_warp[kFusedThreads / 32];
  __shared__ int s_warp2[kFusedThreads / 32];
  __shared__ int s_prefix;
  PassCounters* pc = f.e.pc;
  if (threadIdx.x == 0) s_tile = atomicAdd(&pc->ticket2, 1u);
  {
    int4* z0 = reinterpret_cast<int4*>(s_start);
    int4* z1 = reinterpret_cast<int4*>(s_end);
    for (int k = threadIdx.x; k < kTile / 4; k += kFusedThreads) { z0[k] = make_int4(0, 0, 0, 0); z1[k] = make_int4(0, 0, 0, 0); }
  }
  __syncthreads();
  const int64_t tile = s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = tile * kTile;
  const uint32_t base_lo = (uint32_t)base;
  const int64_t r0 = f.tile_first[tile], r1 = f.tile_first[tile + 1];

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
  // near reads that started before the tile and end inside it: walk back while
  // the start is within max_span of the tile (sorted order => monotone distance)
  {
    const uint32_t reach = pc->max_span;                 // written by k_fused_prep
    // reach <= kNearSpan = kTile, so every candidate started in the previous tile; staying
    // inside it also keeps the 32-bit distance below from wrapping
    const int64_t jmin = tile > 0 ? f.tile_first[tile - 1] : 0;
    for (int64_t j = r0 - 1 - threadIdx.x; j >= jmin; j -= kFusedThreads) {
      uint2 r = f.rec[j];
      uint32_t d = base_lo - r.x;                        // distance behind the tile start (>= 1)
      if (d > reach) break;
      if (r.y >= d && r.y <= kNearSpan) {
        uint32_t el = r.y - d;                           // < kTile because span <= kNearSpan = kTile and d >= 1
        if (el < (uint32_t)kTile) atomicAdd(&s_end[el], 1);
      }
    }
  }
  // far ends bucketed for this tile
  {
    uint32_t k0 = tile > 0 ? f.tile_cnt[tile - 1] : 0u, k1 = f.tile_cnt[tile];
    for (uint32_t k = k0 + threadIdx.x; k < k1; k += kFusedThreads) atomicAdd(&s_end[f.far_sorted[k] & (kTile - 1)], 1);
  }
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
  int carry = 0;
#pragma unroll
  for (int j = 0; j < kScanVec; ++j) {
    int x = run[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    int total = __shfl_sync(0xffffffffu, x, 31);
    run[j] = x - run[j] + carry;
    carry += total;
  }
  if (lane == 31) s_warp[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    int wv = (lane < kFusedThreads / 32) ? s_warp[lane] : 0;
    int x = wv;
#pragma unroll
    for (int o = 1; o < kFusedThreads / 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    int tile_sum = __shfl_sync(0xffffffffu, x, kFusedThreads / 32 - 1);
    if (lane < kFusedThreads / 32) s_warp[lane] = x - wv;
    int prefix = 0;
    if (tile > 0) {
      if (lane == 0) st_release_u64(f.status + tile, kTileAggregate | (uint32_t)tile_sum);
      prefix = scan_lookback(f.status, tile, lane);
    }
    if (lane == 0) {
      st_release_u64(f.status + tile, kTilePrefix | (uint32_t)(prefix + tile_sum));
      s_prefix = prefix;
    }
  }
  __syncthreads();
  const int off = s_prefix + s_warp[warp];
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
  if (lane == 0) { s_warp2[warp] = cap; }
  __syncthreads();
  if (lane == 0) s_warp[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0, c2 = 0;
    for (int k = 0; k < kFusedThreads / 32; ++k) { m = max(m, s_warp[k]); c2 = max(c2, s_warp2[k]); }
    if (m > 0) atomicMax(&pc->max_depth_seen, m);
    if (c2 > 0) atomicMax(&pc->cap_metric, c2);
  }
}

}  // namespace mcov
