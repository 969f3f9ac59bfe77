// k_stats.cuh -- K3/K4: per-region statistics of the depth array, one read of
// the depth per region.
//
// Replaces the seven reductions of `classic` (reference metacov/pileup.py:
// 18-26): min, max, sum, sum of squares (-> std), and the two order statistics
// the reference gets from `np.median` (pileup.py:21) and from
// `sorted(columns)[n//4 : n-n//4]` (pileup.py:24).  The order statistics come
// from an exact counting histogram of the depth values (one bin per value,
// kHistBins bins in shared memory): with htslib's default max_depth = 8000 a
// depth never reaches kHistBins = 8192, so a single pass is exact.  Regions
// whose depth does reach kHistBins are flagged (kStatOverflow) and finished by
// the radix path in mcov_api.cu.
//
// Work unit = one chunk (<= chunk_len slots) of one region; regions are
// arbitrary [start,end) slices and may overlap (reference cli.py:85-95 with
// tests/data/regions.blast7).  A region made of several chunks merges its
// partial histogram into a global per-region histogram; the last chunk to
// finish walks it.
//
// HBM bytes per launch (algorithmic): 4 * sum(region lengths) + 64 * G.
#pragma once
#include "common.cuh"

namespace mcov {

constexpr int kStatThreads = 512;
constexpr int kHistBins = 8192;
constexpr int kStatValid = 1;      // mcov_region_stats.flags bit0
constexpr int kStatOverflow = 2;   // bit1

struct StatTask {
  int64_t slot;     // first slot of the chunk inside the depth array
  int32_t n;        // slots in this chunk
  int32_t region;   // region index
};

struct StatArgs {
  const int32_t* depth;
  const StatTask* tasks;
  const int32_t* region_len;      // slots of the region inside its contig
  const int32_t* region_pad;      // positions of the region beyond the contig end: depth 0 there
  const int32_t* region_chunks;   // chunks per region
  const int32_t* region_hist;     // index into hist_pool (multi-chunk regions) or -1
  uint32_t* hist_pool;            // [n_multi][kHistBins], zeroed
  uint32_t* region_done;          // per-region arrival counter, zeroed
  mcov_region_stats* out;         // device; pre-initialised (min=INT_MAX, max=INT_MIN, rest 0)
  int32_t breadth_n;
};

struct Partial {
  long long sum;
  unsigned long long sumsq;
  long long ge1, geN;
  int mn, mx;
};

__device__ __forceinline__ void hist_add(uint32_t* hist, int v, uint32_t c) {
  atomicAdd(hist + min(v, kHistBins - 1), c);
}

// Walk a kHistBins histogram held in shared memory: order statistics of the
// multiset it describes.  n = number of values; all threads participate.
__device__ __forceinline__ void hist_walk(const uint32_t* hist, long long n, int lo_bin, int hi_bin,
                                          long long& iq_sum_out, int& med_lo_out, int& med_hi_out,
                                          unsigned long long* s_scratch /* >= kStatThreads/32 + 1 */,
                                          int* s_med /* 2 ints */) {
  // ranks (0-based, in sorted order)
  const long long k1 = n / 4, k2 = n - n / 4;                // interquartile [k1, k2)
  const long long m1 = (n - 1) / 2, m2 = n / 2;              // median pair
  constexpr int kPer = kHistBins / kStatThreads;             // 16 bins per thread
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b0 = t * kPer;
  unsigned long long local = 0;
  if (b0 <= hi_bin && b0 + kPer > lo_bin) {
#pragma unroll
    for (int k = 0; k < kPer; ++k) local += hist[b0 + k];
  }
  // block exclusive scan of `local`
  unsigned long long x = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_scratch[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned long long wv = (lane < kStatThreads / 32) ? s_scratch[lane] : 0ull, z = wv;
#pragma unroll
    for (int o = 1; o < kStatThreads / 32; o <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, z, o);
      if (lane >= o) z += y;
    }
    if (lane < kStatThreads / 32) s_scratch[lane] = z - wv;
  }
  __syncthreads();
  long long c = (long long)(s_scratch[warp] + x - local);     // values sorted before bin b0
  long long iq = 0;
  if (local) {
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      long long cnt = hist[b0 + k];
      if (cnt) {
        long long a = c > k1 ? c : k1, b = (c + cnt) < k2 ? (c + cnt) : k2;
        if (b > a) iq += (b - a) * (long long)(b0 + k);
        if (c <= m1 && m1 < c + cnt) s_med[0] = b0 + k;
        if (c <= m2 && m2 < c + cnt) s_med[1] = b0 + k;
        c += cnt;
      }
    }
  }
  __syncthreads();
  iq = warp_sum(iq);
  if (lane == 0) s_scratch[warp] = (unsigned long long)iq;
  __syncthreads();
  long long tot = 0;
  if (t == 0) {
    for (int k = 0; k < kStatThreads / 32; ++k) tot += (long long)s_scratch[k];
  }
  iq_sum_out = tot;               // valid on thread 0
  med_lo_out = s_med[0];
  med_hi_out = s_med[1];
  __syncthreads();
}

__global__ void __launch_bounds__(kStatThreads)
k_region_stats(StatArgs a) {
  __shared__ uint32_t s_hist[kHistBins];
  __shared__ unsigned long long s_red[4][kStatThreads / 32];
  __shared__ int s_mm[2][kStatThreads / 32];
  __shared__ unsigned long long s_scratch[kStatThreads / 32 + 1];
  __shared__ int s_med[2];
  __shared__ int s_last;
  __shared__ Partial s_tot;

  const StatTask task = a.tasks[blockIdx.x];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  {
    uint4* h4 = reinterpret_cast<uint4*>(s_hist);
    for (int k = t; k < kHistBins / 4; k += kStatThreads) h4[k] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  // ---- stream the chunk: aligned int4 window covering [slot, slot+n) ----
  const int64_t s0 = task.slot, s1 = task.slot + task.n;
  const int64_t a0 = s0 & ~(int64_t)3;
  const int64_t nvec = ((s1 - a0) + 3) >> 2;
  const int4* vp = reinterpret_cast<const int4*>(a.depth + a0);
  Partial p;
  p.sum = 0; p.sumsq = 0; p.ge1 = 0; p.geN = 0; p.mn = INT_MAX; p.mx = INT_MIN;
  const int bn = a.breadth_n;
  for (int64_t j = t; j < nvec; j += kStatThreads) {
    int4 q = ld_stream_int4(vp + j);
    int64_t e0 = a0 + (j << 2);
    int vals[4] = {q.x, q.y, q.z, q.w};
    int run_v = 0; uint32_t run_c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int64_t e = e0 + k;
      if (e >= s0 && e < s1) {
        int d = vals[k];
        p.sum += d;
        p.sumsq += (unsigned long long)((long long)d * (long long)d);
        p.ge1 += (d >= 1);
        p.geN += (d >= bn);
        p.mn = min(p.mn, d);
        p.mx = max(p.mx, d);
        if (run_c && d == run_v) ++run_c;
        else {
          if (run_c) hist_add(s_hist, run_v, run_c);
          run_v = d; run_c = 1;
        }
      }
    }
    if (run_c) hist_add(s_hist, run_v, run_c);
  }
  // ---- block reduce the partials ----
  p.sum = warp_sum(p.sum); p.sumsq = warp_sum(p.sumsq);
  p.ge1 = warp_sum(p.ge1); p.geN = warp_sum(p.geN);
  p.mn = warp_min(p.mn);   p.mx = warp_max(p.mx);
  if (lane == 0) {
    s_red[0][warp] = (unsigned long long)p.sum; s_red[1][warp] = p.sumsq;
    s_red[2][warp] = (unsigned long long)p.ge1; s_red[3][warp] = (unsigned long long)p.geN;
    s_mm[0][warp] = p.mn; s_mm[1][warp] = p.mx;
  }
  __syncthreads();
  if (t == 0) {
    Partial q; q.sum = 0; q.sumsq = 0; q.ge1 = 0; q.geN = 0; q.mn = INT_MAX; q.mx = INT_MIN;
    for (int k = 0; k < kStatThreads / 32; ++k) {
      q.sum += (long long)s_red[0][k]; q.sumsq += s_red[1][k];
      q.ge1 += (long long)s_red[2][k]; q.geN += (long long)s_red[3][k];
      q.mn = min(q.mn, s_mm[0][k]);    q.mx = max(q.mx, s_mm[1][k]);
    }
    s_tot = q;
  }
  __syncthreads();
  Partial tot = s_tot;
  const int g = task.region;
  const int n_chunks = a.region_chunks[g];
  const int pad = a.region_pad[g];
  const long long n_region = (long long)a.region_len[g] + pad;
  mcov_region_stats* out = a.out + g;

  if (n_chunks == 1) {
    if (pad > 0) {                       // zeros beyond the contig end join the multiset
      if (t == 0) s_hist[0] += (uint32_t)pad;
      tot.mn = min(tot.mn, 0); tot.mx = max(tot.mx, 0);
      tot.geN += (0 >= bn) ? pad : 0;
      __syncthreads();
    }
    long long iq; int ml, mh;
    int hi_bin = min(tot.mx, kHistBins - 1), lo_bin = max(min(tot.mn, kHistBins - 1), 0);
    hist_walk(s_hist, n_region, lo_bin, hi_bin, iq, ml, mh, s_scratch, s_med);
    if (t == 0) {
      mcov_region_stats r;
      r.sum = tot.sum; r.sumsq = tot.sumsq; r.iq_sum = iq; r.n_ge1 = tot.ge1; r.n_geN = tot.geN;
      r.min = tot.mn; r.max = tot.mx; r.med_lo = ml; r.med_hi = mh; r.reserved = 0;
      r.flags = (tot.mx >= kHistBins - 1 || tot.mn < 0) ? kStatOverflow : kStatValid;
      *out = r;
    }
    return;
  }

  // ---- multi-chunk region: merge into the global histogram ----
  uint32_t* gh = a.hist_pool + (int64_t)a.region_hist[g] * kHistBins;
  {
    int lo = max(min(tot.mn, kHistBins - 1), 0), hi = min(max(tot.mx, 0), kHistBins - 1);
    for (int b = lo + t; b <= hi; b += kStatThreads) {
      uint32_t c = s_hist[b];
      if (c) atomicAdd(gh + b, c);
    }
  }
  if (t == 0) {
    atomicAdd((unsigned long long*)&out->sum, (unsigned long long)tot.sum);
    atomicAdd((unsigned long long*)&out->sumsq, tot.sumsq);
    atomicAdd((unsigned long long*)&out->n_ge1, (unsigned long long)tot.ge1);
    atomicAdd((unsigned long long*)&out->n_geN, (unsigned long long)tot.geN);
    atomicMin(&out->min, tot.mn);
    atomicMax(&out->max, tot.mx);
  }
  __threadfence();
  __syncthreads();
  if (t == 0) {
    unsigned prev = atomicAdd(a.region_done + g, 1u);
    s_last = (prev == (unsigned)(n_chunks - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last chunk of the region: pull the merged histogram and finish
  int gmn = *((volatile int*)&out->min), gmx = *((volatile int*)&out->max);
  for (int b = t; b < kHistBins; b += kStatThreads) s_hist[b] = __ldcg(gh + b);
  __syncthreads();
  if (pad > 0) {
    if (t == 0) {
      s_hist[0] += (uint32_t)pad;
      out->min = min(gmn, 0); out->max = max(gmx, 0);
      if (0 >= bn) out->n_geN += pad;
    }
    gmn = min(gmn, 0); gmx = max(gmx, 0);
    __syncthreads();
  }
  long long iq; int ml, mh;
  hist_walk(s_hist, n_region, max(min(gmn, kHistBins - 1), 0), min(max(gmx, 0), kHistBins - 1), iq, ml, mh,
            s_scratch, s_med);
  if (t == 0) {
    out->iq_sum = iq; out->med_lo = ml; out->med_hi = mh; out->reserved = 0;
    out->flags = (gmx >= kHistBins - 1 || gmn < 0) ? kStatOverflow : kStatValid;
  }
}

__global__ void k_init_region_stats(mcov_region_stats* out, int64_t g) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g) {
    mcov_region_stats r;
    r.sum = 0; r.sumsq = 0; r.iq_sum = 0; r.n_ge1 = 0; r.n_geN = 0;
    r.min = INT_MAX; r.max = INT_MIN; r.med_lo = 0; r.med_hi = 0; r.reserved = 0; r.flags = 0;
    out[i] = r;
  }
}

// Fixed-window mean depth: one warp per window.
__global__ void k_window_sums(const int32_t* __restrict__ depth, const int64_t* __restrict__ win_slot,
                              const int32_t* __restrict__ win_n, int64_t n_win, long long* __restrict__ out) {
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= n_win) return;
  const int32_t* p = depth + win_slot[w];
  int n = win_n[w];
  long long s = 0;
  for (int k = lane; k < n; k += 32) s += p[k];
  s = warp_sum(s);
  if (lane == 0) out[w] = s;
}

}  // namespace mcov
