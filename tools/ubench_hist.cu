// ubench_hist.cu -- microbenchmark: strategies for the per-chunk counting histogram of k_region_stats.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_hist tools/ubench_hist.cu && /tmp/ubench_hist
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
constexpr int T = 512, BINS = 8192, CHUNK = 16384;
__device__ __forceinline__ int hb(int v) { return (unsigned)v < (unsigned)(BINS - 1) ? v : BINS - 1; }
__device__ __forceinline__ void flush(uint32_t* s, unsigned long long* out) {
  unsigned long long acc = 0;
  for (int b = threadIdx.x; b < BINS; b += T) acc += (unsigned long long)s[b] * b;
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
// 0: loads only; 1: per-element atomics; 2: per-vector RLE (current); 3: warp register window; 4: thread window (8 values)
template <int MODE>
__global__ void __launch_bounds__(T, 3) k(const int* __restrict__ d, long long n, unsigned long long* out) {
  __shared__ __align__(16) uint32_t s[BINS];
  for (int k2 = threadIdx.x; k2 < BINS / 4; k2 += T) reinterpret_cast<uint4*>(s)[k2] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const int4* vp = reinterpret_cast<const int4*>(d + (long long)blockIdx.x * CHUNK);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int mn = 1 << 30;
  if (MODE == 3) {
    // each warp: 128 contiguous vectors (512 elements) per iteration, lane takes 4 vectors
    for (int base = warp * 128; base < CHUNK / 4; base += (T / 32) * 128) {
      int4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldcs(vp + base + lane + 32 * u);
      int b[16];
#pragma unroll
      for (int u = 0; u < 4; ++u) { b[4*u] = hb(q[u].x); b[4*u+1] = hb(q[u].y); b[4*u+2] = hb(q[u].z); b[4*u+3] = hb(q[u].w); }
      int lo = b[0], hi = b[0];
#pragma unroll
      for (int e = 1; e < 16; ++e) { lo = min(lo, b[e]); hi = max(hi, b[e]); }
      lo = __reduce_min_sync(~0u, lo); hi = __reduce_max_sync(~0u, hi);
      if (hi - lo < 32) {
        uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // 32 bins x 8 bit (<= 16 per lane)
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          int dd = b[e] - lo; uint32_t inc = 1u << ((dd & 3) << 3); int w = dd >> 2;
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) c[k2] += (w == k2) ? inc : 0u;
        }
        // warp reduce: widen 8-bit fields to 16-bit, butterfly add
        uint32_t tot = 0;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          uint32_t a0 = c[k2] & 0x00ff00ffu, a1 = (c[k2] >> 8) & 0x00ff00ffu;
#pragma unroll
          for (int o = 16; o; o >>= 1) { a0 += __shfl_xor_sync(~0u, a0, o); a1 += __shfl_xor_sync(~0u, a1, o); }
          // bins 4k2+0 (a0 lo), 4k2+1 (a1 lo), 4k2+2 (a0 hi), 4k2+3 (a1 hi)
          int mybin = lane - 4 * k2;
          if (mybin == 0) tot = a0 & 0xffffu; else if (mybin == 1) tot = a1 & 0xffffu;
          else if (mybin == 2) tot = a0 >> 16; else if (mybin == 3) tot = a1 >> 16;
        }
        if (tot) atomicAdd(&s[lo + lane], tot);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) atomicAdd(&s[b[e]], 1u);
      }
    }
  } else {
    for (int j = threadIdx.x; j < CHUNK / 4; j += 4 * T) {
      int4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldcs(vp + j + u * T);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int b0 = hb(q[u].x), b1 = hb(q[u].y), b2 = hb(q[u].z), b3 = hb(q[u].w);
        if (MODE == 0) { mn = min(mn, min(min(b0, b1), min(b2, b3))); }
        else if (MODE == 1) { atomicAdd(&s[b0], 1u); atomicAdd(&s[b1], 1u); atomicAdd(&s[b2], 1u); atomicAdd(&s[b3], 1u); }
        else if (MODE == 2) {
          if (b0 == b1 && b2 == b3) { if (b0 == b2) atomicAdd(&s[b0], 4u); else { atomicAdd(&s[b0], 2u); atomicAdd(&s[b2], 2u); } }
          else {
            uint32_t c0 = 1;
            if (b1 == b0) ++c0; else { atomicAdd(&s[b0], c0); b0 = b1; c0 = 1; }
            if (b2 == b0) ++c0; else { atomicAdd(&s[b0], c0); b0 = b2; c0 = 1; }
            if (b3 == b0) ++c0; else { atomicAdd(&s[b0], c0); b0 = b3; c0 = 1; }
            atomicAdd(&s[b0], c0);
          }
        } else if (MODE == 4) {
          // thread-local window of 8 values around b0, 8-bit fields in a 64-bit word
          int lo = b0 - 3; unsigned long long p = 0; int bb[4] = {b0, b1, b2, b3};
#pragma unroll
          for (int e = 0; e < 4; ++e) { unsigned dd = (unsigned)(bb[e] - lo); if (dd < 8u) p += 1ull << (dd * 8); else atomicAdd(&s[bb[e]], 1u); }
#pragma unroll
          for (int e = 0; e < 8; ++e) { uint32_t c = (uint32_t)(p >> (8 * e)) & 0xffu; if (c) atomicAdd(&s[lo + e], c); }
        }
      }
    }
  }
  if (MODE == 0 && mn == -12345) s[0] = 1;
  __syncthreads();
  flush(s, out);
}
template <int MODE> void run(const char* name, const int* d, long long n, unsigned long long* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int grid = (int)(n / CHUNK);
  float best = 1e9f; unsigned long long h = 0;
  for (int r = 0; r < 6; ++r) {
    cudaMemset(out, 0, 8);
    cudaEventRecord(a); k<MODE><<<grid, T>>>(d, n, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (r) best = ms < best ? ms : best;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  }
  printf("%-28s %8.3f ms  %7.1f GB/s  checksum %llu\n", name, best, 4.0 * n / best / 1e6, h);
}
int main(int argc, char** argv) {
  long long n = 64LL << 20;
  int depth = argc > 1 ? atoi(argv[1]) : 30;
  std::vector<int> h(n);
  unsigned long long st = 88172645463325252ull; int d = depth;
  double p = depth / 150.0;                       // start probability per position
  for (long long i = 0; i < n; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17; double u = (st >> 11) * (1.0 / 9007199254740992.0);
    st ^= st << 13; st ^= st >> 7; st ^= st << 17; double v = (st >> 11) * (1.0 / 9007199254740992.0);
    if (u < p) ++d; if (v < p * d / depth && d > 0) --d;
    h[i] = d;
  }
  int* dd; unsigned long long* out;
  cudaMalloc(&dd, n * 4); cudaMalloc(&out, 8);
  cudaMemcpy(dd, h.data(), n * 4, cudaMemcpyHostToDevice);
  printf("n=%lld depth~%d\n", n, depth);
  run<0>("loads only", dd, n, out);
  run<1>("atomic per element", dd, n, out);
  run<2>("per-vector RLE (current)", dd, n, out);
  run<4>("thread window 8", dd, n, out);
  run<3>("warp register window 32", dd, n, out);
  return 0;
}
