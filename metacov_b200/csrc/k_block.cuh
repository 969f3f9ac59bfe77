// k_block.cuh -- the transport block (include/metacov_b200.h: mcov_block_hdr, version 4) widened into SoA columns on the
// device.
//
// One host-to-device copy brings the block; four launches rebuild the columns, each byte of the block touched twice
// and every column written once:
//   k_block_index   (one thread per escape / exception) where the ascending escape and exception lists enter each
//                   chunk of 2 048 reads;
//   k_block_reduce  (one CTA per chunk) the chunk's totals: CIGAR ops, explicit ops, and the position sum since the
//                   chunk's last contig start (positions are a SEGMENTED sum of the differences, one segment per
//                   contig: the contig starts come from the read prefix `crs`, found by a warp-wide 32-ary search);
//   k_block_prefix  (one CTA) exclusive prefixes of the chunk totals;
//   k_block_expand  (one CTA per chunk) tid (running maximum over the contig starts inside the chunk), pos (segmented
//                   scan with the chunk's carry), flag, mapq, op offsets, and every read's ops copied from the
//                   dictionary (shared memory) or the explicit list -- eight consecutive reads per thread, 64-bit
//                   loads of the per-read bytes, 128-bit stores of the columns.
// The per-read bytes come wide (dpos[], fc[]) or as nibbles with side lists (blk_load), as the packer chose.
// Round 2's first version took eight launches (seed, two patches, counts, three look-back scans, finish) and 335 us for
// 10 M reads.
#pragma once
#include "common.cuh"

namespace mcov {

constexpr int kBlkThreads = 256;
constexpr int kBlkPer = 8;
constexpr int kBlkChunk = kBlkThreads * kBlkPer;          // 2 048 reads per CTA
static_assert(kBlkChunk == MCOV_BLOCK_CHUNK, "the block's chunk table is per CTA of the unpack kernels");
constexpr uint32_t kBlkNone = 0xFFFFFFFFu;

struct BlockArgs {
  const char* blk;            // the block on the device
  mcov_block_hdr h;
  int64_t off_len;            // entries of cig_off[] that are written: n + 1 rounded up to a multiple of 4
  int64_t n_chunks;           // ceil(off_len / kBlkChunk)
  uint32_t* esc_first;        // [n_chunks] index of the chunk's first escape (kBlkNone: none); likewise the exceptions
  uint32_t* exc_first;
  uint4* agg;                 // [n_chunks] {position sum since the last contig start in the chunk, chunk holds a contig start, ops, explicit ops}
  uint4* pre;                 // [n_chunks] exclusive prefixes: {position carry, -, op offset, explicit-op offset}
  int32_t* tid; int32_t* pos; uint16_t* flag; uint8_t* mapq; uint32_t* cig_off; uint32_t* cig;
};

__global__ void k_block_index(const __grid_constant__ BlockArgs a) {
  const mcov_block_hdr& h = a.h;
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t* idx;
  uint32_t* first;
  if (k < h.n_esc) { idx = reinterpret_cast<const uint32_t*>(a.blk + h.off_esc_idx); first = a.esc_first; }
  else {
    k -= h.n_esc;
    if (k >= h.n_exc) return;
    idx = reinterpret_cast<const uint32_t*>(a.blk + h.off_exc_idx); first = a.exc_first;
  }
  const uint32_t i = idx[k];
  if ((int64_t)i >= h.n) return;
  const uint32_t ch = i / kBlkChunk;
  if (k == 0 || idx[k - 1] / kBlkChunk != ch) first[ch] = (uint32_t)k;
}

// largest c in [0, n_contigs] with crs[c] <= x (crs[0] = 0 <= x, crs non-decreasing): 32-ary search by one warp
__device__ __forceinline__ int32_t blk_warp_search(const int64_t* __restrict__ crs, int32_t n_contigs, int64_t x) {
  int32_t lo = 0, hi = n_contigs + 1;
  const int lane = threadIdx.x & 31;
  while (hi - lo > 1) {
    const int32_t step = (hi - lo - 1 + 31) >> 5;
    const int64_t p = (int64_t)lo + (int64_t)(lane + 1) * step;
    const bool ok = p < hi && crs[p] <= x;
    const int k = __popc(__ballot_sync(0xffffffffu, ok));
    const int64_t nhi = (int64_t)lo + (int64_t)(k + 1) * step;
    lo += k * step;
    if (nhi < hi) hi = (int32_t)nhi;
  }
  return lo;
}

struct BlkShared {
  uint32_t jt[256];           // joint table: flag << 8 | class
  uint32_t dn[128];           // ops of a dictionary entry
  uint32_t e[kBlkChunk];      // per read: flag << 8 | class (patching the escapes) / contig-start marks (expand)
  int32_t d[kBlkChunk];       // per read: position difference (patching the exceptions)
  unsigned long long w64[kBlkThreads / 32];
  int32_t w_s[kBlkThreads / 32], w_f[kBlkThreads / 32], w_m[kBlkThreads / 32];
  uint32_t w_n[kBlkThreads / 32];
  int32_t c_lo, c_hi;
};

__device__ __forceinline__ void blk_tables(const BlockArgs& a, BlkShared& sm) {
  const mcov_block_hdr& h = a.h;
  const uint32_t* jt = reinterpret_cast<const uint32_t*>(a.blk + h.off_jt);
  const uint32_t* dict_off = reinterpret_cast<const uint32_t*>(a.blk + h.off_dict_off);
  for (int k = threadIdx.x; k < 256; k += kBlkThreads) sm.jt[k] = k < h.n_jt ? jt[k] : ((0x4u << 8) | 128u);   // (255: an escape, patched below)
  for (int k = threadIdx.x; k < 128; k += kBlkThreads) sm.dn[k] = k < h.n_dict ? dict_off[k + 1] - dict_off[k] : 0u;
}

// the chunk's reads into registers: e[j] = flag << 8 | class, d[j] = position difference, escapes and exceptions applied
// (reads at and beyond n: class 128 = no ops, difference 0).  Ends with the CTA synchronised when it patched.
__device__ __forceinline__ void blk_load(const BlockArgs& a, BlkShared& sm, int64_t chunk, uint32_t (&e)[kBlkPer], int32_t (&d)[kBlkPer]) {
  const mcov_block_hdr& h = a.h;
  const int64_t n = h.n, c0 = chunk * kBlkChunk, i0 = c0 + (int64_t)threadIdx.x * kBlkPer;
  const uint8_t* fc = reinterpret_cast<const uint8_t*>(a.blk + h.off_fc);
  const uint8_t* dp = reinterpret_cast<const uint8_t*>(a.blk + h.off_dpos);
  if (h.nib) {
    // nibble form: one byte per read; a nibble of 15 sends to the next entry of a side list, whose place is the chunk's
    // offset (chunk table of the block) + the number of such nibbles in front of the read (one scan over the CTA)
    const uint8_t* nb = reinterpret_cast<const uint8_t*>(a.blk + h.off_nb);
    uint32_t lo[kBlkPer], hi[kBlkPer];
    if (i0 + kBlkPer <= n) {
      const uint2 q = *reinterpret_cast<const uint2*>(nb + i0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t b0 = (q.x >> (8 * j)) & 255u, b1 = (q.y >> (8 * j)) & 255u;
        lo[j] = b0 & 15u; hi[j] = b0 >> 4; lo[j + 4] = b1 & 15u; hi[j + 4] = b1 >> 4;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kBlkPer; ++j) {
        const uint32_t b = i0 + j < n ? (uint32_t)nb[i0 + j] : 0x1000u;       // (beyond n: no difference; index 256 = no ops)
        lo[j] = b & 15u; hi[j] = b >> 4;
      }
    }
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) cnt += (lo[j] == 15u ? 1u : 0u) + (hi[j] == 15u ? 0x10000u : 0u);
    uint32_t inc = cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) sm.w_n[warp] = inc;
    __syncthreads();
    uint32_t ex = inc - cnt;
    for (int w = 0; w < warp; ++w) ex += sm.w_n[w];
    uint32_t kd = 0, kf = 0;
    if (c0 < n) {
      const uint32_t* ct = reinterpret_cast<const uint32_t*>(a.blk + h.off_chunk);
      kd = ct[2 * chunk] + (ex & 0xFFFFu); kf = ct[2 * chunk + 1] + (ex >> 16);
    }
    const uint8_t* dq = reinterpret_cast<const uint8_t*>(a.blk + h.off_dq);
    const uint8_t* fq = reinterpret_cast<const uint8_t*>(a.blk + h.off_fq);
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) {
      uint32_t dv = lo[j], fi = hi[j];
      if (dv == 15u) { dv = (int64_t)kd < h.n_dq ? (uint32_t)dq[kd] : 0u; ++kd; }
      if (fi == 15u) { fi = (int64_t)kf < h.n_fq ? (uint32_t)fq[kf] : 255u; ++kf; }
      d[j] = (int32_t)dv;
      e[j] = fi < 256u ? sm.jt[fi] : 128u;
    }
  } else if (i0 + kBlkPer <= n) {
    const uint2 f = *reinterpret_cast<const uint2*>(fc + i0), q = *reinterpret_cast<const uint2*>(dp + i0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      e[j] = sm.jt[(f.x >> (8 * j)) & 255u]; e[j + 4] = sm.jt[(f.y >> (8 * j)) & 255u];
      d[j] = (int32_t)((q.x >> (8 * j)) & 255u); d[j + 4] = (int32_t)((q.y >> (8 * j)) & 255u);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) {
      const bool in = i0 + j < n;
      e[j] = in ? sm.jt[fc[i0 + j]] : 128u;
      d[j] = in ? (int32_t)dp[i0 + j] : 0;
    }
  }
  const uint32_t kf = a.esc_first[chunk], xf = a.exc_first[chunk];          // (uniform over the CTA)
  if (kf == kBlkNone && xf == kBlkNone) return;
  const int64_t c1 = c0 + kBlkChunk;
#pragma unroll
  for (int j = 0; j < kBlkPer; ++j) { sm.e[threadIdx.x * kBlkPer + j] = e[j]; sm.d[threadIdx.x * kBlkPer + j] = d[j]; }
  __syncthreads();
  if (kf != kBlkNone) {
    const uint32_t* qi = reinterpret_cast<const uint32_t*>(a.blk + h.off_esc_idx);
    const uint16_t* qf = reinterpret_cast<const uint16_t*>(a.blk + h.off_esc_flag);
    const uint8_t* qc = reinterpret_cast<const uint8_t*>(a.blk + h.off_esc_cls);
    for (int64_t k = (int64_t)kf + threadIdx.x; k < h.n_esc; k += kBlkThreads) {
      const int64_t i = qi[k];
      if (i >= c1 || i >= n) break;                                           // (ascending: the rest belongs to later chunks)
      if (i >= c0) sm.e[i - c0] = ((uint32_t)qf[k] << 8) | qc[k];
    }
  }
  if (xf != kBlkNone) {
    const uint32_t* ei = reinterpret_cast<const uint32_t*>(a.blk + h.off_exc_idx);
    const int32_t* ev = reinterpret_cast<const int32_t*>(a.blk + h.off_exc_val);
    for (int64_t k = (int64_t)xf + threadIdx.x; k < h.n_exc; k += kBlkThreads) {
      const int64_t i = ei[k];
      if (i >= c1 || i >= n) break;
      if (i >= c0) sm.d[i - c0] = ev[k];
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kBlkPer; ++j) { e[j] = sm.e[threadIdx.x * kBlkPer + j]; d[j] = sm.d[threadIdx.x * kBlkPer + j]; }
  __syncthreads();                                                            // (sm.e is reused for the contig marks)
}

// ops / explicit ops of a class
__device__ __forceinline__ void blk_counts(const BlkShared& sm, uint32_t cls, uint32_t& nc, uint32_t& nx) {
  if (cls < 128u) { nc = sm.dn[cls]; nx = 0; } else { nc = cls - 128u; nx = nc; }
}

__global__ void __launch_bounds__(kBlkThreads) k_block_reduce(const __grid_constant__ BlockArgs a) {
  __shared__ BlkShared sm;
  const mcov_block_hdr& h = a.h;
  const int64_t chunk = blockIdx.x, c0 = chunk * kBlkChunk, c1 = c0 + kBlkChunk;
  const int64_t* crs = reinterpret_cast<const int64_t*>(a.blk + h.off_crs);
  blk_tables(a, sm);
  if (threadIdx.x < 32) {                                 // the chunk's last contig start (if any)
    const int32_t c = blk_warp_search(crs, h.n_contigs, c1 - 1);
    if (threadIdx.x == 0) sm.c_hi = c;
  }
  __syncthreads();
  uint32_t e[kBlkPer];
  int32_t d[kBlkPer];
  blk_load(a, sm, chunk, e, d);
  const int64_t seg = crs[sm.c_hi];                       // <= c1 - 1
  const bool has_seg = seg >= c0;
  const int64_t i0 = c0 + (int64_t)threadIdx.x * kBlkPer;
  uint32_t s = 0, nc = 0, nx = 0;
#pragma unroll
  for (int j = 0; j < kBlkPer; ++j) {
    uint32_t c, x;
    blk_counts(sm, e[j] & 255u, c, x);
    nc += c; nx += x;
    if (i0 + j >= seg) s += (uint32_t)d[j];
  }
  unsigned long long cx = (unsigned long long)nc | ((unsigned long long)nx << 32);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); cx += __shfl_xor_sync(0xffffffffu, cx, o); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm.w_s[warp] = (int32_t)s; sm.w64[warp] = cx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t ts = 0;
    unsigned long long tc = 0;
    for (int w = 0; w < kBlkThreads / 32; ++w) { ts += (uint32_t)sm.w_s[w]; tc += sm.w64[w]; }
    a.agg[chunk] = make_uint4(ts, has_seg ? 1u : 0u, (uint32_t)tc, (uint32_t)(tc >> 32));
  }
}

// exclusive prefixes of the chunk totals: one CTA, every thread a contiguous run of chunks
constexpr int kBlkPrefixThreads = 1024;
__global__ void __launch_bounds__(kBlkPrefixThreads) k_block_prefix(const __grid_constant__ BlockArgs a) {
  __shared__ uint32_t s_s[32], s_f[32], s_c[32], s_x[32];
  const int64_t m = (a.n_chunks + kBlkPrefixThreads - 1) / kBlkPrefixThreads;
  const int64_t lo = (int64_t)threadIdx.x * m, hi = min(lo + m, a.n_chunks);
  uint32_t s = 0, f = 0, c = 0, x = 0;
  for (int64_t k = lo; k < hi; ++k) {
    const uint4 g = a.agg[k];
    s = g.y ? g.x : s + g.x; f |= g.y; c += g.z; x += g.w;
  }
  // inclusive scan over the threads: (s, f) under the segmented sum, c and x under +
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t is = s, jf = f, ic = c, ix = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t us = __shfl_up_sync(0xffffffffu, is, o), uf = __shfl_up_sync(0xffffffffu, jf, o);
    const uint32_t uc = __shfl_up_sync(0xffffffffu, ic, o), ux = __shfl_up_sync(0xffffffffu, ix, o);
    if (lane >= o) { if (!jf) is += us; jf |= uf; ic += uc; ix += ux; }
  }
  if (lane == 31) { s_s[warp] = is; s_f[warp] = jf; s_c[warp] = ic; s_x[warp] = ix; }
  __syncthreads();
  // this thread's exclusive prefix: the warps before it, then the lanes before it
  uint32_t ps = 0, pc = 0, px = 0;
  for (int w = 0; w < warp; ++w) { ps = s_f[w] ? s_s[w] : ps + s_s[w]; pc += s_c[w]; px += s_x[w]; }
  {
    const uint32_t es = __shfl_up_sync(0xffffffffu, is, 1), ef = __shfl_up_sync(0xffffffffu, jf, 1);
    const uint32_t ec = __shfl_up_sync(0xffffffffu, ic, 1), ex = __shfl_up_sync(0xffffffffu, ix, 1);
    if (lane > 0) { ps = ef ? es : ps + es; pc += ec; px += ex; }
  }
  for (int64_t k = lo; k < hi; ++k) {
    const uint4 g = a.agg[k];
    a.pre[k] = make_uint4(ps, 0u, pc, px);
    ps = g.y ? g.x : ps + g.x; pc += g.z; px += g.w;
  }
}

__global__ void __launch_bounds__(kBlkThreads) k_block_expand(const __grid_constant__ BlockArgs a) {
  __shared__ BlkShared sm;
  __shared__ uint32_t s_dict_off[129], s_dict_ops[512];
  const mcov_block_hdr& h = a.h;
  const int64_t n = h.n, chunk = blockIdx.x, c0 = chunk * kBlkChunk, c1 = c0 + kBlkChunk;
  const int64_t* crs = reinterpret_cast<const int64_t*>(a.blk + h.off_crs);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  blk_tables(a, sm);
  {
    const uint32_t* dict_off = reinterpret_cast<const uint32_t*>(a.blk + h.off_dict_off);
    const uint32_t* dict_ops = reinterpret_cast<const uint32_t*>(a.blk + h.off_dict_ops);
    for (int k = t; k < 129; k += kBlkThreads) s_dict_off[k] = k <= h.n_dict ? min(dict_off[k], 512u) : 0u;
    for (int k = t; k < 512; k += kBlkThreads) s_dict_ops[k] = k < h.n_dictops ? dict_ops[k] : 0u;
  }
  if (t < 32) {                                           // contig of the chunk's first read
    const int32_t c = blk_warp_search(crs, h.n_contigs, c0);
    if (t == 0) sm.c_lo = c;
  }
  __syncthreads();
  uint32_t e[kBlkPer];
  int32_t d[kBlkPer];
  blk_load(a, sm, chunk, e, d);
  // contig starts inside the chunk: mark = contig index + 1 at the start's place (empty contigs share a start: the last wins)
  const int32_t c_lo = sm.c_lo;
#pragma unroll
  for (int j = 0; j < kBlkPer; ++j) sm.e[t * kBlkPer + j] = 0u;
  __syncthreads();
  if (t == 0 && crs[c_lo] == c0) sm.e[0] = (uint32_t)c_lo + 1u;
  for (int64_t base = (int64_t)c_lo + 1;; base += kBlkThreads) {
    const int64_t c = base + t;
    const bool v = c <= h.n_contigs && crs[min(c, (int64_t)h.n_contigs)] < c1;
    if (v) atomicMax(&sm.e[crs[c] - c0], (uint32_t)c + 1u);
    if (!__syncthreads_and(v ? 1 : 0)) break;
  }
  // thread-local walks: contig (running maximum of the marks), segmented position sum, op counts
  uint32_t mk[kBlkPer], ps[kBlkPer];
  uint32_t m = 0, s = 0, f = 0, nc = 0, nx = 0;
  uint32_t cn[kBlkPer], xn[kBlkPer];
#pragma unroll
  for (int j = 0; j < kBlkPer; ++j) {
    const uint32_t q = sm.e[t * kBlkPer + j];
    if (q) { m = q; s = (uint32_t)d[j]; f |= 1u << j; } else s += (uint32_t)d[j];
    if (f) f |= 1u << j;
    mk[j] = m; ps[j] = s;
    blk_counts(sm, e[j] & 255u, cn[j], xn[j]);
    nc += cn[j]; nx += xn[j];
  }
  // scans over the threads of the CTA
  uint32_t im = m, is = s, jf = f ? 1u : 0u;
  unsigned long long icx = (unsigned long long)nc | ((unsigned long long)nx << 32);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t um = __shfl_up_sync(0xffffffffu, im, o), us = __shfl_up_sync(0xffffffffu, is, o), uf = __shfl_up_sync(0xffffffffu, jf, o);
    const unsigned long long uc = __shfl_up_sync(0xffffffffu, icx, o);
    if (lane >= o) { im = max(im, um); if (!jf) is += us; jf |= uf; icx += uc; }
  }
  if (lane == 31) { sm.w_m[warp] = (int32_t)im; sm.w_s[warp] = (int32_t)is; sm.w_f[warp] = (int32_t)jf; sm.w64[warp] = icx; }
  __syncthreads();
  const uint4 pre = a.pre[chunk];
  uint32_t pm = (uint32_t)c_lo + 1u, pp = pre.x;
  unsigned long long pcx = (unsigned long long)pre.z | ((unsigned long long)pre.w << 32);
  for (int w = 0; w < warp; ++w) {
    pm = max(pm, (uint32_t)sm.w_m[w]);
    pp = sm.w_f[w] ? (uint32_t)sm.w_s[w] : pp + (uint32_t)sm.w_s[w];
    pcx += sm.w64[w];
  }
  {
    const uint32_t em = __shfl_up_sync(0xffffffffu, im, 1), es = __shfl_up_sync(0xffffffffu, is, 1), ef = __shfl_up_sync(0xffffffffu, jf, 1);
    const unsigned long long ec = __shfl_up_sync(0xffffffffu, icx, 1);
    if (lane > 0) { pm = max(pm, em); pp = ef ? es : pp + es; pcx += ec; }
  }
  // outputs
  const int64_t i0 = c0 + (int64_t)t * kBlkPer;
  uint32_t o_c = (uint32_t)pcx, o_x = (uint32_t)(pcx >> 32);
  int32_t v_tid[kBlkPer], v_pos[kBlkPer];
  uint32_t v_off[kBlkPer];
#pragma unroll
  for (int j = 0; j < kBlkPer; ++j) {
    const uint32_t ci = max(pm, mk[j]) - 1u;
    v_tid[j] = ci < (uint32_t)h.n_contigs ? (int32_t)ci : -1;
    v_pos[j] = (int32_t)(((f >> j) & 1u) ? ps[j] : pp + ps[j]);
    v_off[j] = o_c;
    o_c += cn[j];
  }
  if (i0 + kBlkPer <= n) {
    int4* pt = reinterpret_cast<int4*>(a.tid + i0); int4* pq = reinterpret_cast<int4*>(a.pos + i0);
    pt[0] = make_int4(v_tid[0], v_tid[1], v_tid[2], v_tid[3]); pt[1] = make_int4(v_tid[4], v_tid[5], v_tid[6], v_tid[7]);
    pq[0] = make_int4(v_pos[0], v_pos[1], v_pos[2], v_pos[3]); pq[1] = make_int4(v_pos[4], v_pos[5], v_pos[6], v_pos[7]);
    *reinterpret_cast<uint4*>(a.flag + i0) = make_uint4((e[0] >> 8) | ((e[1] >> 8) << 16), (e[2] >> 8) | ((e[3] >> 8) << 16),
                                                        (e[4] >> 8) | ((e[5] >> 8) << 16), (e[6] >> 8) | ((e[7] >> 8) << 16));
    uint2 mq = make_uint2(0xffffffffu, 0xffffffffu);
    if (h.has_mapq) mq = *reinterpret_cast<const uint2*>(a.blk + h.off_mapq + i0);
    *reinterpret_cast<uint2*>(a.mapq + i0) = mq;
  } else {
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) {
      if (i0 + j < n) {
        a.tid[i0 + j] = v_tid[j]; a.pos[i0 + j] = v_pos[j]; a.flag[i0 + j] = (uint16_t)(e[j] >> 8);
        a.mapq[i0 + j] = h.has_mapq ? reinterpret_cast<const uint8_t*>(a.blk + h.off_mapq)[i0 + j] : (uint8_t)0xff;
      }
    }
  }
  if (i0 + kBlkPer <= a.off_len) {
    uint4* po = reinterpret_cast<uint4*>(a.cig_off + i0);
    po[0] = make_uint4(v_off[0], v_off[1], v_off[2], v_off[3]); po[1] = make_uint4(v_off[4], v_off[5], v_off[6], v_off[7]);
  } else {
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) if (i0 + j < a.off_len) a.cig_off[i0 + j] = v_off[j];
  }
  // ops: dictionary entries from shared memory, explicit ops from the side list (u16 or u32 per op)
  const uint32_t n_cig = (uint32_t)h.n_cigar, n_xops = (uint32_t)h.n_xops;
  const uint32_t* x32 = reinterpret_cast<const uint32_t*>(a.blk + h.off_xops);
  const uint16_t* x16 = reinterpret_cast<const uint16_t*>(a.blk + h.off_xops);
  const bool narrow = h.xop_bytes == 2;
#pragma unroll
  for (int j = 0; j < kBlkPer; ++j) {
    const uint32_t cls = e[j] & 255u, o0 = v_off[j];
    if (cls < 128u) {
      const uint32_t b = s_dict_off[cls];
      for (uint32_t k = 0; k < cn[j]; ++k) if (o0 + k < n_cig) a.cig[o0 + k] = s_dict_ops[min(b + k, 511u)];
    } else {
      for (uint32_t k = 0; k < cn[j]; ++k)
        if (o0 + k < n_cig && o_x + k < n_xops) a.cig[o0 + k] = narrow ? (uint32_t)x16[o_x + k] : x32[o_x + k];
      o_x += cn[j];
    }
  }
}

}  // namespace mcov
