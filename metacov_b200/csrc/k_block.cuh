// k_block.cuh -- the transport block (include/metacov_b200.h: mcov_block_hdr, version 4) widened into SoA columns on the
// device.
//
// One host-to-device copy brings the block; ONE kernel rebuilds the columns, every byte of the block read once and every
// column written once.  k_block_expand runs one CTA per CHUNK of 2 048 reads, eight consecutive reads per thread, and a
// chunk needs nothing from the others: the packer, which has the columns at hand, left a CHUNK TABLE in the block (op
// offset and position in front of the chunk; where the chunk's entries of the side lists, the explicit ops, the escapes
// and the exceptions begin).  Inside a chunk:
//   * the per-read bytes -- wide (dpos[], fc[]) or nibbles whose value 15 sends to a side list (its entry found by one
//     scan of the 15-counts over the CTA), as the packer chose -- with 64-bit loads; escapes and exceptions of the chunk
//     patched through shared memory;
//   * tid: the contig of the chunk's first read by a warp-wide 32-ary search in the read prefix crs[], the contig starts
//     inside the chunk marked in shared memory, a running maximum over the marks;
//   * pos: a SEGMENTED inclusive scan of the differences (one segment per contig), seeded with the chunk's carry;
//   * op offsets: a scan of the op counts (dictionary entry or explicit count);
//   * ops: dictionary entries and the chunk's explicit ops (loaded coalesced) are gathered into a shared-memory image
//     of the chunk's op range, which is then stored coalesced;
//   * tid, pos, flag, mapq, cig_off with 128-bit stores.
// Round 2's first version took eight launches (seed, two patches, counts, three look-back scans, finish) and 335 us for
// 10 M reads; a three-kernel version with device-side chunk sums 177 us.
#pragma once
#include <cstddef>
#include "common.cuh"

namespace mcov {

constexpr int kBlkThreads = 256;
constexpr int kBlkPer = 8;
constexpr int kBlkChunk = kBlkThreads * kBlkPer;          // 2 048 reads per CTA
static_assert(kBlkChunk == MCOV_BLOCK_CHUNK, "the block's chunk table is per CTA of the unpack kernel");
constexpr uint32_t kBlkNone = 0xFFFFFFFFu;
constexpr uint32_t kBlkEscape = 0xFFFFFFFFu;              // joint-table entry of an index beyond the table (no flag << 8 | class looks like it)
constexpr int kBlkOpCap = 2 * kBlkChunk;                  // ops of a chunk staged in shared memory (more: direct stores)
constexpr int kBlkXopCap = kBlkChunk;                     // explicit ops of a chunk staged in shared memory

struct BlockArgs {
  const char* blk;            // the block on the device
  mcov_block_hdr h;
  int64_t off_len;            // entries of cig_off[] that are written: n + 1 rounded up to a multiple of 4
  int32_t* tid; int32_t* pos; uint16_t* flag; uint8_t* mapq; uint32_t* cig_off; uint32_t* cig;
};

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

struct alignas(16) BlkShared {
  uint32_t jt[256];           // joint table: flag << 8 | class
  uint32_t dn[128];           // ops of a dictionary entry
  uint32_t dict_off[132];     // (129 used; padded so that what follows stays 16-byte aligned)
  uint32_t src[512 + kBlkXopCap];   // op sources of the chunk in one array: [0, 512) the dictionary's ops, behind them the chunk's explicit ops
  // per read: flag << 8 | class and position difference while the escapes / exceptions are patched; then e[] holds the
  // contig-start marks; then e[] and d[] together are the image of the chunk's op range
  alignas(16) uint32_t e[kBlkChunk];
  int32_t d[kBlkChunk];
  unsigned long long w64[kBlkThreads / 32];
  int32_t w_s[kBlkThreads / 32], w_f[kBlkThreads / 32], w_m[kBlkThreads / 32];
  uint32_t w_n[kBlkThreads / 32];
  int32_t c_lo, starts;
  uint32_t list_n;
  uint32_t list[kBlkChunk];   // reads of more than one op: {image offset, ops, source}
};
static_assert(offsetof(BlkShared, d) == offsetof(BlkShared, e) + sizeof(uint32_t) * kBlkChunk, "e[] and d[] form one array");

// the chunk's reads into registers: e[j] = flag << 8 | class, d[j] = position difference, escapes and exceptions applied
// (reads at and beyond n: class 128 = no ops, difference 0)
__device__ __forceinline__ void blk_load(const BlockArgs& a, BlkShared& sm, int64_t chunk, const mcov_block_chunk& ce,
                                         const bool marks_follow, const uint2 w0, const uint2 w1, uint32_t (&e)[kBlkPer],
                                         int32_t (&d)[kBlkPer]) {
  // w0 / w1: the thread's eight per-read bytes (nibble form: w0; wide form: w0 = fc, w1 = dpos), fetched by the caller before
  // its first barrier when all eight reads exist
  const mcov_block_hdr& h = a.h;
  const int64_t n = h.n, c0 = chunk * kBlkChunk, i0 = c0 + (int64_t)threadIdx.x * kBlkPer;
  if (h.nib) {
    // nibble form: one byte per read; a nibble of 15 sends to the next entry of a side list, whose place is the chunk's
    // offset (chunk table) + the number of such nibbles in front of the read (one scan over the CTA)
    const uint8_t* nb = reinterpret_cast<const uint8_t*>(a.blk + h.off_nb);
    const uint8_t* dq = reinterpret_cast<const uint8_t*>(a.blk + h.off_dq);
    const uint8_t* fq = reinterpret_cast<const uint8_t*>(a.blk + h.off_fq);
    const uint32_t n_dq = (uint32_t)h.n_dq, n_fq = (uint32_t)h.n_fq;            // (both at most n < 2^32)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool full = i0 + kBlkPer <= n;
    // the eight bytes as one 64-bit word: low / high nibbles spread to bytes, and one bit per nibble that is 15 (bit 0 of
    // a nibble of y: all four of its bits are set) -- the side-list entries are then patched in by a loop over those bits,
    // which a thread with none (most) skips, instead of eight predicated look-ups
    unsigned long long L = 0, H = 0, ml = 0, mh = 0;
    uint32_t lo[kBlkPer], hi[kBlkPer];                                         // (the ragged last chunk only)
    uint32_t cnt = 0;
    if (full) {
      const unsigned long long W = (unsigned long long)w0.x | ((unsigned long long)w0.y << 32);
      L = W & 0x0F0F0F0F0F0F0F0Full; H = (W >> 4) & 0x0F0F0F0F0F0F0F0Full;
      unsigned long long y = W & (W >> 1);
      y &= y >> 2;
      ml = y & 0x0101010101010101ull; mh = (y >> 4) & 0x0101010101010101ull;
      cnt = (uint32_t)__popcll(ml) | ((uint32_t)__popcll(mh) << 16);
    } else {
#pragma unroll
      for (int j = 0; j < kBlkPer; ++j) {
        const uint32_t b = i0 + j < n ? (uint32_t)nb[i0 + j] : 0x1000u;       // (beyond n: no difference; index 256 = no ops)
        lo[j] = b & 15u; hi[j] = b >> 4;
        cnt += (lo[j] == 15u ? 1u : 0u) + (hi[j] == 15u ? 0x10000u : 0u);
      }
    }
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) sm.w_n[warp] = inc;
    __syncthreads();
    uint32_t ex = inc - cnt;
    for (int w = 0; w < warp; ++w) ex += sm.w_n[w];
    uint32_t kd = ce.dq_off + (ex & 0xFFFFu), kf = ce.fq_off + (ex >> 16);
    if (full) {
      while (ml) {
        const int b = __ffsll((long long)ml) - 1;                             // bit 8j
        ml &= ml - 1;
        const unsigned long long v = kd < n_dq ? (unsigned long long)dq[kd] : 0ull;
        ++kd;
        L = (L & ~(0xFFull << b)) | (v << b);
      }
      while (mh) {
        const int b = __ffsll((long long)mh) - 1;
        mh &= mh - 1;
        const unsigned long long v = kf < n_fq ? (unsigned long long)fq[kf] : 255ull;
        ++kf;
        H = (H & ~(0xFFull << b)) | (v << b);
      }
      const uint32_t l0 = (uint32_t)L, l1 = (uint32_t)(L >> 32), h0 = (uint32_t)H, h1 = (uint32_t)(H >> 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        d[j] = (int32_t)((l0 >> (8 * j)) & 255u); d[j + 4] = (int32_t)((l1 >> (8 * j)) & 255u);
        e[j] = sm.jt[(h0 >> (8 * j)) & 255u]; e[j + 4] = sm.jt[(h1 >> (8 * j)) & 255u];
      }
    } else {
      if (cnt) {
#pragma unroll
        for (int j = 0; j < kBlkPer; ++j) {
          if (lo[j] == 15u) { lo[j] = kd < n_dq ? (uint32_t)dq[kd] : 0u; ++kd; }
          if (hi[j] == 15u) { hi[j] = kf < n_fq ? (uint32_t)fq[kf] : 255u; ++kf; }
        }
      }
#pragma unroll
      for (int j = 0; j < kBlkPer; ++j) { d[j] = (int32_t)lo[j]; e[j] = hi[j] < 256u ? sm.jt[hi[j]] : 128u; }
    }
  } else {
    const uint8_t* fc = reinterpret_cast<const uint8_t*>(a.blk + h.off_fc);
    const uint8_t* dp = reinterpret_cast<const uint8_t*>(a.blk + h.off_dpos);
    if (i0 + kBlkPer <= n) {
      const uint2 f = w0, q = w1;
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
  }
  const uint32_t kf = ce.esc_first, xf = ce.exc_first;                        // (uniform over the CTA)
  const int64_t c1 = c0 + kBlkChunk;
  if (xf == kBlkNone) {
    if (kf == kBlkNone) {
#pragma unroll
      for (int j = 0; j < kBlkPer; ++j) if (e[j] == kBlkEscape) e[j] = 128u;   // (an escape the list does not hold: a malformed block)
      return;
    }
    // no exception in the chunk (the common case): the few escapes are dropped at their reads' places in shared memory
    // and only the reads whose table index said "escape" look there
    {
      const uint32_t* qi = reinterpret_cast<const uint32_t*>(a.blk + h.off_esc_idx);
      const uint16_t* qf = reinterpret_cast<const uint16_t*>(a.blk + h.off_esc_flag);
      const uint8_t* qc = reinterpret_cast<const uint8_t*>(a.blk + h.off_esc_cls);
      for (int64_t k = (int64_t)kf + threadIdx.x; k < h.n_esc; k += kBlkThreads) {
        const int64_t i = qi[k];
        if (i >= c1 || i >= n) break;                                         // (ascending: the rest belongs to later chunks)
        if (i >= c0) { const uint32_t cls = qc[k]; sm.e[i - c0] = ((uint32_t)qf[k] << 8) | cls | ((cls < 128u ? sm.dn[cls] : cls - 128u) << 24); }
      }
    }
    __syncthreads();
    bool any = false;
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) any |= e[j] == kBlkEscape;
    if (any) {
#pragma unroll
      for (int j = 0; j < kBlkPer; ++j) if (e[j] == kBlkEscape) e[j] = sm.e[threadIdx.x * kBlkPer + j] & 0x7FFFFFFFu;
    }
    if (marks_follow) __syncthreads();                                        // (sm.e is rewritten at once by the contig marks; else
                                                                              //  only behind the next barriers, as the op image)
    return;
  }
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
      if (i >= c0) { const uint32_t cls = qc[k]; sm.e[i - c0] = ((uint32_t)qf[k] << 8) | cls | ((cls < 128u ? sm.dn[cls] : cls - 128u) << 24); }
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
  for (int j = 0; j < kBlkPer; ++j) {
    e[j] = sm.e[threadIdx.x * kBlkPer + j]; d[j] = sm.d[threadIdx.x * kBlkPer + j];
    if (e[j] == kBlkEscape) e[j] = 128u;                                      // (an escape the list does not hold: a malformed block)
  }
  __syncthreads();                                                            // (sm.e is reused for the contig marks)
}

__global__ void __launch_bounds__(kBlkThreads, 4) k_block_expand(const __grid_constant__ BlockArgs a) {
  __shared__ BlkShared sm;
  const mcov_block_hdr& h = a.h;
  const int64_t n = h.n, chunk = blockIdx.x, c0 = chunk * kBlkChunk, c1 = c0 + kBlkChunk;
  const int64_t* crs = reinterpret_cast<const int64_t*>(a.blk + h.off_crs);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // the chunk's row of the table (the chunk behind the last read only completes cig_off[])
  mcov_block_chunk ce;
  if (c0 < n) {
    const uint4* row = reinterpret_cast<const uint4*>(a.blk + h.off_chunk) + 2 * chunk;
    const uint4 r0 = row[0], r1 = row[1];
    ce.dq_off = r0.x; ce.fq_off = r0.y; ce.op_off = r0.z; ce.xop_off = r0.w;
    ce.pos_carry = (int32_t)r1.x; ce.esc_first = r1.y; ce.exc_first = r1.z; ce.reserved = 0;
  } else {
    ce.dq_off = 0; ce.fq_off = 0; ce.op_off = (uint32_t)h.n_cigar; ce.xop_off = (uint32_t)h.n_xops;
    ce.pos_carry = 0; ce.esc_first = kBlkNone; ce.exc_first = kBlkNone; ce.reserved = 0;
  }
  // the chunk's explicit ops into shared memory (coalesced) -- their number is known from the table (the next row's entry
  // point), so the loads are issued now and have landed long before the ops are gathered
  const uint32_t n_cig = (uint32_t)h.n_cigar, n_xops = (uint32_t)h.n_xops;
  const bool narrow = h.xop_bytes == 2;
  const uint32_t* x32 = reinterpret_cast<const uint32_t*>(a.blk + h.off_xops);
  const uint16_t* x16 = reinterpret_cast<const uint16_t*>(a.blk + h.off_xops);
  uint32_t n_xo_tab = 0;
  if (c0 < n) {
    const uint32_t x_end = c1 < n ? (reinterpret_cast<const uint4*>(a.blk + h.off_chunk) + 2 * (chunk + 1))->w : n_xops;
    n_xo_tab = x_end >= ce.xop_off ? x_end - ce.xop_off : 0u;
    if (n_xo_tab <= (uint32_t)kBlkXopCap) {
      for (uint32_t q = t; q < n_xo_tab; q += kBlkThreads) {
        const uint32_t g = ce.xop_off + q;
        sm.src[512 + q] = g < n_xops ? (narrow ? (uint32_t)x16[g] : x32[g]) : 0u;
      }
    }
  }
  if (t == 0) sm.list_n = 0;
  // the thread's own eight per-read bytes, in flight while the tables load
  uint2 w0 = make_uint2(0u, 0u), w1 = make_uint2(0u, 0u);
  {
    const int64_t i0 = c0 + (int64_t)t * kBlkPer;
    if (i0 + kBlkPer <= n) {
      if (h.nib) w0 = *reinterpret_cast<const uint2*>(a.blk + h.off_nb + i0);
      else { w0 = *reinterpret_cast<const uint2*>(a.blk + h.off_fc + i0); w1 = *reinterpret_cast<const uint2*>(a.blk + h.off_dpos + i0); }
    }
  }
  {
    const uint32_t* jt = reinterpret_cast<const uint32_t*>(a.blk + h.off_jt);
    const uint32_t* dict_off = reinterpret_cast<const uint32_t*>(a.blk + h.off_dict_off);
    const uint32_t* dict_ops = reinterpret_cast<const uint32_t*>(a.blk + h.off_dict_ops);
    // joint table entry: flag << 8 | class, and the read's op count in bits 24..30 (dictionary entry: its length; explicit:
    // class - 128), so that the count is one shift away
    for (int k = t; k < 256; k += kBlkThreads) {
      uint32_t en = kBlkEscape;                                               // (255: an escape, patched in blk_load)
      if (k < h.n_jt) {
        en = jt[k] & 0xFFFFFFu;
        const uint32_t cls = en & 255u;
        const uint32_t cnt = cls >= 128u ? cls - 128u : ((int)cls < h.n_dict ? min(dict_off[cls + 1] - dict_off[cls], 4u) : 0u);
        en |= cnt << 24;
      }
      sm.jt[k] = en;
    }
    for (int k = t; k < 128; k += kBlkThreads) sm.dn[k] = k < h.n_dict ? min(dict_off[k + 1] - dict_off[k], 4u) : 0u;
    for (int k = t; k < 129; k += kBlkThreads) sm.dict_off[k] = k <= h.n_dict ? min(dict_off[k], 508u) : 0u;
    for (int k = t; k < 512; k += kBlkThreads) sm.src[k] = k < h.n_dictops ? dict_ops[k] : 0u;
  }
  if (t < 32) {                                           // contig of the chunk's first read
    const int32_t c = blk_warp_search(crs, h.n_contigs, c0);
    if (t == 0) {
      sm.c_lo = c;
      sm.starts = (crs[c] == c0 || (c < h.n_contigs && crs[c + 1] < c1)) ? 1 : 0;    // does a contig start inside the chunk?
    }
  }
  __syncthreads();
  uint32_t e[kBlkPer];
  int32_t d[kBlkPer];
  blk_load(a, sm, chunk, ce, sm.starts != 0, w0, w1, e, d);
  // contig starts inside the chunk: mark = contig index + 1 at the start's place (empty contigs share a start: the last wins)
  const int32_t c_lo = sm.c_lo;
  const bool starts = sm.starts != 0;                     // (uniform: most chunks of deep data lie inside one contig)
  if (starts) {
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
  }
  // thread-local walks: contig (running maximum of the marks), segmented position sum, op counts
  uint32_t mk[kBlkPer], ps[kBlkPer], cn[kBlkPer];
  uint32_t m = 0, s = 0, f = 0, nc = 0, nx = 0;
#pragma unroll
  for (int j = 0; j < kBlkPer; ++j) {
    const uint32_t q = starts ? sm.e[t * kBlkPer + j] : 0u;
    if (q) { m = q; s = (uint32_t)d[j]; f |= 1u << j; } else s += (uint32_t)d[j];
    if (f) f |= 1u << j;
    mk[j] = m; ps[j] = s;
    const uint32_t cls = e[j] & 255u;
    cn[j] = e[j] >> 24;
    nc += cn[j]; nx += cls < 128u ? 0u : cn[j];
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
  __syncthreads();                                        // (also: every thread has read its marks -- e[] / d[] are free)
  uint32_t pm = (uint32_t)c_lo + 1u, pp = (uint32_t)ce.pos_carry;
  unsigned long long pcx = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kBlkThreads / 32; ++w) {
    const unsigned long long v = sm.w64[w];
    tot += v;
    if (w < warp) {
      pm = max(pm, (uint32_t)sm.w_m[w]);
      pp = sm.w_f[w] ? (uint32_t)sm.w_s[w] : pp + (uint32_t)sm.w_s[w];
      pcx += v;
    }
  }
  {
    const uint32_t em = __shfl_up_sync(0xffffffffu, im, 1), es = __shfl_up_sync(0xffffffffu, is, 1), ef = __shfl_up_sync(0xffffffffu, jf, 1);
    const unsigned long long ec = __shfl_up_sync(0xffffffffu, icx, 1);
    if (lane > 0) { pm = max(pm, em); pp = ef ? es : pp + es; pcx += ec; }
  }
  const uint32_t n_ops = (uint32_t)tot, n_xo = (uint32_t)(tot >> 32);       // of the chunk
  // (a table that disagrees with the classes -- a malformed block -- sends the chunk down the unstaged path)
  const bool staged = n_ops + 3u <= (uint32_t)kBlkOpCap && n_xo <= (uint32_t)kBlkXopCap && n_xo == n_xo_tab;
  const uint32_t pad = ce.op_off & 3u;                                        // the image starts `pad` words in: image index = global
                                                                              // op index mod 4, so both sides of the flush are 16-byte aligned
  // outputs
  const int64_t i0 = c0 + (int64_t)t * kBlkPer;
  uint32_t o_c = (uint32_t)pcx, o_x = (uint32_t)(pcx >> 32);                  // relative to the chunk
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
    // (bytes 1, 2 of an entry are the flag: one byte permute packs two flags)
    *reinterpret_cast<uint4*>(a.flag + i0) = make_uint4(__byte_perm(e[0], e[1], 0x6521), __byte_perm(e[2], e[3], 0x6521),
                                                        __byte_perm(e[4], e[5], 0x6521), __byte_perm(e[6], e[7], 0x6521));
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
    po[0] = make_uint4(ce.op_off + v_off[0], ce.op_off + v_off[1], ce.op_off + v_off[2], ce.op_off + v_off[3]);
    po[1] = make_uint4(ce.op_off + v_off[4], ce.op_off + v_off[5], ce.op_off + v_off[6], ce.op_off + v_off[7]);
  } else {
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) if (i0 + j < a.off_len) a.cig_off[i0 + j] = ce.op_off + v_off[j];
  }
  // ops
  if (staged) {
    // every read's ops -- dictionary entry or its slice of the chunk's explicit ops, both in shared memory -- gathered into the
    // image of the chunk's op range, then the image stored coalesced
    uint32_t* img = sm.e;                                                     // e[] and d[]: kBlkOpCap words (free since the barrier
                                                                              // behind the scans, which also covers xo[] and list_n)
    // up to three ops of every read here (predicated: a read of short-read data has one op, one in ten two or three);
    // the rare longer CIGARs go to a list that the CTA then works through together -- a loop `for k < c` per read makes
    // every warp wait for its longest CIGAR eight times
    // (o0 + c <= n_ops <= kBlkOpCap and o_x + c <= n_xo <= kBlkXopCap hold by construction: all four are sums of the same cn[])
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) {
      const uint32_t cls = e[j] & 255u, c = cn[j], o0 = v_off[j] + pad;
      const bool dict = cls < 128u;
      const uint32_t si = dict ? sm.dict_off[cls] : 512u + o_x;
      if (c > 0) img[o0] = sm.src[si];
      if (c > 1) img[o0 + 1] = sm.src[si + 1];
      if (c > 2) img[o0 + 2] = sm.src[si + 2];
      const uint32_t more = __ballot_sync(0xffffffffu, c > 3);
      if (more) {
        uint32_t base = 0;
        const int leader = __ffs(more) - 1;
        if (lane == leader) base = atomicAdd(&sm.list_n, (uint32_t)__popc(more));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (c > 3) sm.list[base + __popc(more & ((1u << lane) - 1u))] = o0 | (c << 12) | (si << 19);
      }
      if (!dict) o_x += c;
    }
    __syncthreads();
    for (uint32_t q = t; q < sm.list_n; q += kBlkThreads) {
      const uint32_t en = sm.list[q], o0 = en & 4095u, c = (en >> 12) & 127u, si = en >> 19;
      for (uint32_t k = 3; k < c; ++k) img[o0 + k] = sm.src[si + k];
    }
    __syncthreads();
    {
      // 128-bit loads from the image, 128-bit stores to cig[]; the first and the last vector may be partial
      uint32_t* gbase = a.cig + (ce.op_off - pad);                            // 16-byte aligned (cig[] is, and op_off - pad is a multiple of 4)
      const uint32_t gb = ce.op_off - pad;
      const uint32_t end = pad + n_ops, lim = n_cig > gb ? n_cig - gb : 0u;   // image indices [pad, end) hold ops; global room from gbase
      const uint32_t nvec = (end + 3u) >> 2;
      for (uint32_t v = t; v < nvec; v += kBlkThreads) {
        const uint32_t q = 4u * v;
        const uint4 x = *reinterpret_cast<const uint4*>(img + q);
        if (q >= pad && q + 4u <= end && q + 4u <= lim) *reinterpret_cast<uint4*>(gbase + q) = x;
        else {
          const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) if (q + u >= pad && q + u < end && q + u < lim) gbase[q + u] = xs[u];
        }
      }
    }
  } else {
    // a chunk of long CIGARs: every thread stores its reads' ops itself
#pragma unroll
    for (int j = 0; j < kBlkPer; ++j) {
      const uint32_t cls = e[j] & 255u, o0 = ce.op_off + v_off[j];
      if (cls < 128u) {
        const uint32_t b = sm.dict_off[cls];
        for (uint32_t k = 0; k < cn[j]; ++k) if (o0 + k < n_cig) a.cig[o0 + k] = sm.src[b + k];
      } else {
        const uint32_t g = ce.xop_off + o_x;
        for (uint32_t k = 0; k < cn[j]; ++k)
          if (o0 + k < n_cig && g + k < n_xops) a.cig[o0 + k] = narrow ? (uint32_t)x16[g + k] : x32[g + k];
        o_x += cn[j];
      }
    }
  }
}

}  // namespace mcov
