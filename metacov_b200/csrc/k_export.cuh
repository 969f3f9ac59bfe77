// k_export.cuh -- run-length export of the per-base depth (the rows of a bedGraph file).
//
// Additive feature (the reference has no depth export: `classic` keeps its
// `columns` vector private, reference metacov/pileup.py:10-26).  A run is a maximal stretch of equal
// depth inside one contig; run k is described by (tid, start, depth) and ends where the next run of
// the same contig starts (or at the contig end).
//
// Work unit = one chunk of one contig (kRunChunk slots, one CTA).  Two passes over the depth, no
// atomics, output in position order:
//   k_run_count   run starts per chunk
//   k_run_offsets exclusive scan of the chunk counts (one CTA, 64-bit: 5*10^9 slots can hold > 2^31 runs)
//   k_run_write   the same flags again, block-wide ballot scan, records to their final place
// HBM bytes (algorithmic): 8 per slot + 12 per run.
#pragma once
#include "common.cuh"

namespace mcov {

constexpr int kRunThreads = 256;
constexpr int kRunChunk = 8192;

struct RunTask {
  int64_t slot;     // first slot of the chunk
  int32_t n;        // positions in the chunk
  int32_t tid;
  int32_t pos0;     // contig position of the first slot
  int32_t reserved;
};

// run start at position p of the chunk: first position of the contig, or depth differs from the slot before
__device__ __forceinline__ bool run_starts(const int32_t* __restrict__ d, const RunTask& t, int p) {
  const int32_t v = __ldg(d + p);
  return (t.pos0 + p == 0) || (__ldg(d + p - 1) != v);
}

__global__ void __launch_bounds__(kRunThreads)
k_run_count(const int32_t* __restrict__ depth, const RunTask* __restrict__ tasks, long long* __restrict__ counts) {
  __shared__ int s_w[kRunThreads / 32];
  const RunTask t = tasks[blockIdx.x];
  const int32_t* d = depth + t.slot;
  int c = 0;
  for (int p = threadIdx.x; p < t.n; p += kRunThreads) c += run_starts(d, t, p) ? 1 : 0;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int k = 0; k < kRunThreads / 32; ++k) tot += s_w[k];
    counts[blockIdx.x] = tot;
  }
}

// counts[0..n) -> exclusive offsets in place, counts[n] = total.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024)
k_run_offsets(long long* __restrict__ counts, int64_t n) {
  __shared__ long long s_w[32];
  __shared__ long long s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const long long v = i < n ? counts[i] : 0;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    long long before = s_carry;
    for (int k = 0; k < warp; ++k) before += s_w[k];
    if (i < n) counts[i] = before + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[n] = s_carry;
}

__global__ void __launch_bounds__(kRunThreads)
k_run_write(const int32_t* __restrict__ depth, const RunTask* __restrict__ tasks, const long long* __restrict__ offsets,
            int32_t* __restrict__ out_tid, int32_t* __restrict__ out_start, int32_t* __restrict__ out_depth) {
  __shared__ int s_w[kRunThreads / 32];
  const RunTask t = tasks[blockIdx.x];
  const int32_t* d = depth + t.slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long base = offsets[blockIdx.x];
  for (int p0 = 0; p0 < t.n; p0 += kRunThreads) {
    const int p = p0 + threadIdx.x;
    const bool st = p < t.n && run_starts(d, t, p);
    const unsigned m = __ballot_sync(0xffffffffu, st);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kRunThreads / 32; ++k) { before += (k < warp) ? s_w[k] : 0; total += s_w[k]; }
    if (st) {
      const long long k = base + before + __popc(m & ((1u << lane) - 1u));
      out_tid[k] = t.tid; out_start[k] = t.pos0 + p; out_depth[k] = __ldg(d + p);
    }
    base += total;
    __syncthreads();
  }
}

// end of run k = start of run k+1 if it lies in the same contig, else the contig's length
__global__ void k_run_ends(const int32_t* __restrict__ tid, const int32_t* __restrict__ start, const int32_t* __restrict__ contig_len,
                           int64_t n, int32_t* __restrict__ end) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    end[k] = (k + 1 < n && tid[k + 1] == tid[k]) ? start[k + 1] : contig_len[tid[k]];
}

}  // namespace mcov
