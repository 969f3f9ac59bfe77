// k_hist.cuh -- read-statistics scan: ByFlag grouping + insert-size histogram.
//
// Replaces `ByFlag.process_read` (reference metacov/scan.pyx:406-420: group
// index n = bits of the selected flags, first selected flag = most significant
// bit) and `IsizeHist.process_read` (scan.pyx:590-610: counts[abs(isize)] += 1
// for EVERY read) with the getter semantics of scan.pyx:267-271, 293-294
// (isize counts only when PROPER_PAIR is set, otherwise the read lands in bin
// 0).  The per-record iteration order of `scan_reads` (scan.pyx:653-667) does
// not matter for a histogram, so reads are processed in parallel.
//
// Histograms are privatised in shared memory when they fit (groups * bins <=
// kIsizeSmemBins) and flushed with one atomic per non-zero bin per CTA.
#pragma once
#include "common.cuh"

namespace mcov {

constexpr int kHistThreads = 256;
constexpr int kIsizeSmemBins = 8192;
constexpr int kMaxGroupFlags = 11;

struct IsizeArgs {
  int64_t n;
  const uint16_t* flag;
  const int32_t* isize;
  int32_t n_group_flags;
  uint16_t group_flags[kMaxGroupFlags];
  int32_t n_bins;
  uint32_t* hist;                 // [groups][n_bins]
  unsigned long long* group_cnt;  // [groups]
  int* max_isize;
};

__global__ void __launch_bounds__(kHistThreads)
k_isize_hist(IsizeArgs a, int use_smem) {
  __shared__ uint32_t s_hist[kIsizeSmemBins];
  const int groups = 1 << a.n_group_flags;
  const int total = groups * a.n_bins;
  if (use_smem) {
    for (int k = threadIdx.x; k < total; k += kHistThreads) s_hist[k] = 0;
    __syncthreads();
  }
  int mx = 0;
  const int64_t stride = (int64_t)gridDim.x * kHistThreads;
  for (int64_t i = (int64_t)blockIdx.x * kHistThreads + threadIdx.x; i < a.n; i += stride) {
    uint32_t f = a.flag[i];
    int g = 0;
    for (int k = 0; k < a.n_group_flags; ++k) g = (g << 1) | ((f & a.group_flags[k]) ? 1 : 0);
    int v = (f & 0x2u) ? a.isize[i] : 0;      // get_isize(): PROPER_PAIR only
    // abs() as the reference computes it (-INT_MIN stays negative there too)
    if (v < 0) v = -v;
    mx = max(mx, v);
    if (v >= 0 && v < a.n_bins) {
      if (use_smem) atomicAdd(&s_hist[g * a.n_bins + v], 1u);
      else atomicAdd(a.hist + (int64_t)g * a.n_bins + v, 1u);
    }
    // group sizes ride in an extra column handled by the caller (bin n_bins)
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(a.max_isize, mx);
  if (use_smem) {
    __syncthreads();
    for (int k = threadIdx.x; k < total; k += kHistThreads) {
      uint32_t c = s_hist[k];
      if (c) atomicAdd(a.hist + k, c);
    }
  }
}

// reads per ByFlag group (independent of the isize range)
__global__ void __launch_bounds__(kHistThreads)
k_group_count(IsizeArgs a) {
  __shared__ uint32_t s_cnt[1 << kMaxGroupFlags];
  const int groups = 1 << a.n_group_flags;
  for (int k = threadIdx.x; k < groups; k += kHistThreads) s_cnt[k] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * kHistThreads;
  for (int64_t i = (int64_t)blockIdx.x * kHistThreads + threadIdx.x; i < a.n; i += stride) {
    uint32_t f = a.flag[i];
    int g = 0;
    for (int k = 0; k < a.n_group_flags; ++k) g = (g << 1) | ((f & a.group_flags[k]) ? 1 : 0);
    // warp-aggregate: most reads of a warp share a group
    unsigned peers = __match_any_sync(__activemask(), g);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_cnt[g], (uint32_t)__popc(peers));
  }
  __syncthreads();
  for (int k = threadIdx.x; k < groups; k += kHistThreads) {
    uint32_t c = s_cnt[k];
    if (c) atomicAdd(a.group_cnt + k, (unsigned long long)c);
  }
}

}  // namespace mcov
