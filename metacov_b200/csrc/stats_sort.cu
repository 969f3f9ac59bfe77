// stats_sort.cu -- statistics of a region whose depth reaches the counting
// histogram's range (k_stats.cuh, kHistBins): radix-sort a copy of the region
// on the GPU and read everything off the sorted vector.  Only reachable when
// max_depth has been raised above 8190 or disabled; kept on the GPU so that no
// result ever comes from a CPU path.
#include <cub/device/device_radix_sort.cuh>

#include "ctx.cuh"

namespace mcov {

// sorted[0..n) ascending, preceded (virtually) by `pad` zeros.  *out must be zeroed.
__global__ void k_sorted_stats(const int32_t* __restrict__ sorted, long long n, long long pad, int breadth_n,
                               mcov_region_stats* out) {
  const long long N = n + pad;
  const long long k1 = N / 4, k2 = N - N / 4, m1 = (N - 1) / 2, m2 = N / 2;
  long long iq = 0, sum = 0, ge1 = 0, geN = 0;
  unsigned long long sumsq = 0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    long long v = sorted[r];
    sum += v; sumsq += (unsigned long long)(v * v);
    ge1 += v >= 1; geN += v >= breadth_n;
    long long rank = r + pad;
    if (rank >= k1 && rank < k2) iq += v;
  }
  iq = warp_sum(iq); sum = warp_sum(sum); ge1 = warp_sum(ge1); geN = warp_sum(geN); sumsq = warp_sum(sumsq);
  if ((threadIdx.x & 31) == 0) {
    if (iq) atomicAdd((unsigned long long*)&out->iq_sum, (unsigned long long)iq);
    if (sum) atomicAdd((unsigned long long*)&out->sum, (unsigned long long)sum);
    if (sumsq) atomicAdd((unsigned long long*)&out->sumsq, sumsq);
    if (ge1) atomicAdd((unsigned long long*)&out->n_ge1, (unsigned long long)ge1);
    if (geN) atomicAdd((unsigned long long*)&out->n_geN, (unsigned long long)geN);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out->med_lo = m1 >= pad ? sorted[m1 - pad] : 0;
    out->med_hi = m2 >= pad ? sorted[m2 - pad] : 0;
    int mn = n > 0 ? sorted[0] : 0, mx = n > 0 ? sorted[n - 1] : 0;
    if (pad > 0) { mn = min(mn, 0); mx = max(mx, 0); if (0 >= breadth_n) atomicAdd((unsigned long long*)&out->n_geN, (unsigned long long)pad); }
    out->min = mn; out->max = mx;
    out->flags = kStatValidBit | kStatOverflowBit;
  }
}

}  // namespace mcov

// d_stat: device record, rewritten completely.
int mcov_order_stats_by_sort(mcov_ctx* ctx, const int32_t* d_region, int64_t n, int64_t pad, int breadth_n,
                             mcov_region_stats* d_stat, mcov::DevBuf& keys_out, mcov::DevBuf& temp) {
  using namespace mcov;
  if (n > INT32_MAX) return MCOV_ERR_RANGE;
  cudaStream_t s = ctx->stream;
  if (keys_out.ensure((size_t)std::max<int64_t>(n, 1) * 4) != cudaSuccess) return MCOV_ERR_NOMEM;
  size_t tb = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tb, d_region, keys_out.as<int32_t>(), (int)n, 0, 32, s);
  if (temp.ensure(tb + 16) != cudaSuccess) return MCOV_ERR_NOMEM;
  if (n > 0 &&
      cub::DeviceRadixSort::SortKeys(temp.p, tb, d_region, keys_out.as<int32_t>(), (int)n, 0, 32, s) != cudaSuccess)
    return MCOV_ERR_CUDA;
  if (cudaMemsetAsync(d_stat, 0, sizeof(mcov_region_stats), s) != cudaSuccess) return MCOV_ERR_CUDA;
  ctx->prof_begin(kKSortedStats);
  k_sorted_stats<<<ctx->n_sm, 256, 0, s>>>(keys_out.as<int32_t>(), n, pad, breadth_n, d_stat);
  ctx->prof_end();
  return cudaGetLastError() == cudaSuccess ? MCOV_OK : MCOV_ERR_CUDA;
}
