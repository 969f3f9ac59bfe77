// stats_sort.cu -- order statistics of a region whose depth reaches the
// counting histogram's range (k_stats.cuh, kHistBins): radix-sort a copy of
// the region on the GPU and read the ranks off the sorted vector.  Only
// reachable when max_depth has been raised above 8191 or disabled; kept on the
// GPU so that no result ever comes from a CPU path.
#include <cub/device/device_radix_sort.cuh>

#include "ctx.cuh"

namespace mcov {

// sorted[0..n) ascending, preceded (virtually) by `pad` zeros.
__global__ void k_sorted_order_stats(const int32_t* __restrict__ sorted, long long n, long long pad,
                                     mcov_region_stats* out) {
  const long long N = n + pad;
  const long long k1 = N / 4, k2 = N - N / 4, m1 = (N - 1) / 2, m2 = N / 2;
  long long acc = 0;
  for (long long r = k1 + (long long)blockIdx.x * blockDim.x + threadIdx.x; r < k2;
       r += (long long)gridDim.x * blockDim.x)
    if (r >= pad) acc += sorted[r - pad];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd((unsigned long long*)&out->iq_sum, (unsigned long long)acc);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out->med_lo = m1 >= pad ? sorted[m1 - pad] : 0;
    out->med_hi = m2 >= pad ? sorted[m2 - pad] : 0;
  }
}

}  // namespace mcov

// d_stat: device mcov_region_stats whose iq_sum/med fields are rewritten.
int mcov_order_stats_by_sort(mcov_ctx* ctx, const int32_t* d_region, int64_t n, int64_t pad,
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
  if (cudaMemsetAsync(&d_stat->iq_sum, 0, sizeof(int64_t), s) != cudaSuccess) return MCOV_ERR_CUDA;
  k_sorted_order_stats<<<kNumSMsB200, 256, 0, s>>>(keys_out.as<int32_t>(), n, pad, d_stat);
  return cudaGetLastError() == cudaSuccess ? MCOV_OK : MCOV_ERR_CUDA;
}
