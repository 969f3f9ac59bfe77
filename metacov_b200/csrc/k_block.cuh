// k_block.cuh -- the transport block (include/metacov_b200.h: mcov_block_hdr) widened into SoA columns on the device.
//
// One host-to-device copy brings the block.  k_block_seed looks every read's (flag, CIGAR class) pair up in the
// joint table and seeds the position sum; k_block_patch applies the escapes (pairs outside the table) and
// k_delta_patch the position exceptions; k_block_counts turns the classes into op counts; three prefix sums
// (k_scan_inplace) give positions, op offsets and explicit-op offsets; k_block_finish writes pos[] and copies
// every read's ops from the dictionary or the explicit list.  The fused pass then runs on ordinary columns.
#pragma once
#include "common.cuh"

namespace mcov {

struct BlockArgs {
  const char* blk;            // the block on the device
  mcov_block_hdr h;
  int64_t off_len;            // length of the scanned arrays: n + 1 rounded up to a multiple of 4
  int32_t* S;                 // position prefix sums
  uint32_t* xoff;             // explicit-op offsets
  uint8_t* cls;               // CIGAR class of every read
  int32_t* tid; int32_t* pos; uint16_t* flag; uint8_t* mapq; uint32_t* cig_off; uint32_t* cig;
};

__global__ void k_block_seed(const __grid_constant__ BlockArgs a) {
  __shared__ uint32_t s_jt[256];
  const mcov_block_hdr& h = a.h;
  const uint32_t* jt = reinterpret_cast<const uint32_t*>(a.blk + h.off_jt);
  for (int k = threadIdx.x; k < 256; k += blockDim.x) s_jt[k] = k < h.n_jt ? jt[k] : ((0x4u << 8) | 128u);   // (escapes: patched next)
  __syncthreads();
  const int64_t n = h.n;
  const uint8_t* dpos = reinterpret_cast<const uint8_t*>(a.blk + h.off_dpos);
  const int64_t* crs = reinterpret_cast<const int64_t*>(a.blk + h.off_crs);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.off_len) a.S[i] = i < n ? (int32_t)dpos[i] : 0;
  if (i >= n) return;
  const uint32_t e = s_jt[reinterpret_cast<const uint8_t*>(a.blk + h.off_fc)[i]];
  a.flag[i] = (uint16_t)(e >> 8);
  a.cls[i] = (uint8_t)(e & 255u);
  a.mapq[i] = h.has_mapq ? reinterpret_cast<const uint8_t*>(a.blk + h.off_mapq)[i] : (uint8_t)0xff;
  // contig of read i: largest c with crs[c] <= i; reads past the last contig are unplaced
  if (i >= crs[h.n_contigs]) { a.tid[i] = -1; return; }
  int32_t lo = 0, hi = h.n_contigs;
  while (hi - lo > 1) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    if (crs[mid] <= i) lo = mid; else hi = mid;
  }
  a.tid[i] = lo;
}

__global__ void k_block_patch(const __grid_constant__ BlockArgs a) {
  const mcov_block_hdr& h = a.h;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= h.n_esc) return;
  const uint32_t i = reinterpret_cast<const uint32_t*>(a.blk + h.off_esc_idx)[k];
  if (i >= h.n) return;
  a.flag[i] = reinterpret_cast<const uint16_t*>(a.blk + h.off_esc_flag)[k];
  a.cls[i] = reinterpret_cast<const uint8_t*>(a.blk + h.off_esc_cls)[k];
}

__global__ void k_block_counts(const __grid_constant__ BlockArgs a) {
  __shared__ uint32_t s_dn[128];
  const mcov_block_hdr& h = a.h;
  const uint32_t* dict_off = reinterpret_cast<const uint32_t*>(a.blk + h.off_dict_off);
  for (int k = threadIdx.x; k < 128; k += blockDim.x) s_dn[k] = k < h.n_dict ? dict_off[k + 1] - dict_off[k] : 0u;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.off_len) return;
  uint32_t nc = 0, nx = 0;
  if (i >= 1 && i <= h.n) {
    const uint32_t c = a.cls[i - 1];
    if (c < 128u) nc = s_dn[c]; else { nc = c - 128u; nx = nc; }
  }
  a.cig_off[i] = nc;                                     // entry 0 and the padding are 0: an inclusive scan gives the offsets
  a.xoff[i] = nx;
}

__global__ void k_block_finish(const __grid_constant__ BlockArgs a) {
  const mcov_block_hdr& h = a.h;
  const int64_t n = h.n;
  const int64_t* crs = reinterpret_cast<const int64_t*>(a.blk + h.off_crs);
  const uint32_t* dict_off = reinterpret_cast<const uint32_t*>(a.blk + h.off_dict_off);
  const uint32_t* dict_ops = reinterpret_cast<const uint32_t*>(a.blk + h.off_dict_ops);
  const uint32_t* xops = reinterpret_cast<const uint32_t*>(a.blk + h.off_xops);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t t = a.tid[i];
    const int64_t first = crs[t >= 0 ? t : h.n_contigs];                  // unplaced reads: one more segment
    a.pos[i] = (int32_t)((uint32_t)a.S[i] - (first > 0 ? (uint32_t)a.S[first - 1] : 0u));
    const uint32_t c = a.cls[i];
    const uint32_t o0 = a.cig_off[i], cnt = a.cig_off[i + 1] - o0;
    const uint32_t* src = c < 128u ? dict_ops + dict_off[min(c, (uint32_t)max(h.n_dict - 1, 0))] : xops + a.xoff[i];
    for (uint32_t k = 0; k < cnt; ++k) a.cig[o0 + k] = src[k];
  }
}

}  // namespace mcov
