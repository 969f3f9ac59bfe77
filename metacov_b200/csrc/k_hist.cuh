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

// ---- k-mer histogram (KmerHist, reference metacov/scan.pyx:491-533) -------------------------------
// For each read with rlen >= OFFSET + STEP*NK (scan.pyx:510) and each i < NK: the k-mer starting at
// read position OFFSET + i*STEP, 2 bits per base with the FIRST base in the low bits (scan.pyx:520);
// a base that is not A/C/G/T sends the k-mer to bin 4^K (scan.pyx:517-519); counts[k][i] += 1.
// "Read position" is in the sequenced orientation: for reverse-strand records the stored SEQ is
// reverse-complemented back (scan.pyx:251-257).  The input is the per-read window of
// win_bases = OFFSET + (NK-1)*STEP + K bases produced by mcov_bam_seq_windows: window base j of a
// forward read is read base j; of a reverse read it is stored base (l_seq - win_bases + j), i.e. read
// base x = complement(window[win_bases-1-x]).
struct KmerArgs {
  int64_t n;
  const uint16_t* flag;
  const int32_t* l_seq;
  const uint8_t* win;       // [n][win_bytes]
  int32_t win_bytes, win_bases;
  int32_t K, NK, STEP, OFFSET;
  int32_t n_group_flags;
  uint16_t group_flags[kMaxGroupFlags];
  uint32_t* hist;           // [groups][4^K + 1][NK]
};

__device__ __forceinline__ int nt16_to_nt4(int c) {      // "=ACMGRSVTWYHKDBN" -> A0 C1 G2 T3, everything else 4
  return c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4;
}

__global__ void __launch_bounds__(kHistThreads)
k_kmer_hist(KmerArgs a) {
  const int64_t stride = (int64_t)gridDim.x * kHistThreads;
  const int64_t table = ((int64_t)1 << (2 * a.K)) + 1;
  for (int64_t i = (int64_t)blockIdx.x * kHistThreads + threadIdx.x; i < a.n; i += stride) {
    const int rlen = a.l_seq[i];
    if (rlen < a.OFFSET + a.STEP * a.NK) continue;
    const uint32_t f = a.flag[i];
    int g = 0;
    for (int k = 0; k < a.n_group_flags; ++k) g = (g << 1) | ((f & a.group_flags[k]) ? 1 : 0);
    const bool rev = (f & 0x10u) != 0;
    const uint8_t* w = a.win + i * a.win_bytes;
    uint32_t* h = a.hist + (int64_t)g * table * a.NK;
    for (int s = 0; s < a.NK; ++s) {
      int64_t kmer = 0;
      for (int j = 0; j < a.K; ++j) {
        int x = a.OFFSET + s * a.STEP + j;                 // read position (sequenced orientation)
        int c = 4;
        if (x < rlen && x < a.win_bases) {
          int wj = rev ? a.win_bases - 1 - x : x;
          int nib = (wj & 1) ? (w[wj >> 1] & 15) : (w[wj >> 1] >> 4);
          c = nt16_to_nt4(nib);
          if (rev && c < 4) c = 3 - c;                     // nt4_comp (scan.pyx:61-62); N stays N
        }
        if (c > 3) { kmer = table - 1; break; }
        kmer |= (int64_t)c << (2 * j);
      }
      atomicAdd(h + kmer * a.NK + s, 1u);
    }
  }
}

}  // namespace mcov
