// synth.cu -- synthetic read generator entry points (host threads or device
// kernels over the same integer-only generator, include/mcov_synth.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "mcov_synth.h"
#include "metacov_b200.h"

namespace {

__global__ void k_synth_ncigar(mcov_synth_params P, int64_t i0, int64_t n, uint32_t* out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = mcov_synth_ncigar(&P, i0 + k);
}

template <typename OffT>
__global__ void k_synth_fill(mcov_synth_params P, int64_t i0, int64_t n, const int64_t* read_start,
                             const int32_t* contig_len, int32_t n_contigs, int32_t tid_base, const OffT* cig_off,
                             int32_t* tid, int32_t* pos, uint16_t* flag, uint8_t* mapq, int32_t* isize, uint32_t* cig,
                             int64_t* reflen_out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int32_t t, p, is;
  uint16_t f;
  uint8_t q;
  OffT o0 = cig_off[k], o1 = cig_off[k + 1];
  int64_t rl = mcov_synth_read(&P, i0 + k, read_start, contig_len, n_contigs, &t, &p, &f, &q, &is, cig + o0, (uint32_t)(o1 - o0));
  tid[k] = t - tid_base; pos[k] = p; flag[k] = f; mapq[k] = q; isize[k] = is;
  if (reflen_out) reflen_out[k] = rl;
}

template <typename F>
void parallel_for(int64_t n, F fn) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::max(1u, std::min(hw ? hw : 4u, 64u));
  if (n < 1 << 16) nt = 1;
  std::vector<std::thread> th;
  int64_t per = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    int64_t a = t * per, b = std::min(n, a + per);
    if (a >= b) break;
    th.emplace_back([=]() { for (int64_t k = a; k < b; ++k) fn(k); });
  }
  for (auto& t : th) t.join();
}

}  // namespace

template <typename OffT>
static int synth_gen_reads_impl(const mcov_synth_params* P, int64_t i0, int64_t n, const int64_t* read_start,
                    const int32_t* contig_len, int32_t n_contigs, int32_t tid_base, const OffT* cig_off, int32_t* tid,
                    int32_t* pos, uint16_t* flag, uint8_t* mapq, int32_t* isize, uint32_t* cig, int64_t* reflen_out,
                    int mem_kind, void* stream) {
  if (!P || n < 0 || n_contigs <= 0 || !read_start || !contig_len) return MCOV_ERR_ARG;
  if (n == 0) return MCOV_OK;
  if (!cig_off || !tid || !pos || !flag || !mapq || !isize || !cig) return MCOV_ERR_ARG;
  if (mem_kind == MCOV_MEM_DEVICE) {
    k_synth_fill<OffT><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        *P, i0, n, read_start, contig_len, n_contigs, tid_base, cig_off, tid, pos, flag, mapq, isize, cig, reflen_out);
    return cudaGetLastError() == cudaSuccess ? MCOV_OK : MCOV_ERR_CUDA;
  }
  mcov_synth_params Q = *P;
  parallel_for(n, [=](int64_t k) {
    int32_t t, p, is;
    uint16_t f;
    uint8_t q;
    OffT o0 = cig_off[k], o1 = cig_off[k + 1];
    int64_t rl = mcov_synth_read(&Q, i0 + k, read_start, contig_len, n_contigs, &t, &p, &f, &q, &is, cig + o0, (uint32_t)(o1 - o0));
    tid[k] = t - tid_base; pos[k] = p; flag[k] = f; mapq[k] = q; isize[k] = is;
    if (reflen_out) reflen_out[k] = rl;
  });
  return MCOV_OK;
}

extern "C" {

int mcov_synth_gen_ncigar(const mcov_synth_params* P, int64_t i0, int64_t n, uint32_t* out, int mem_kind, void* stream) {
  if (!P || n < 0 || (n > 0 && !out)) return MCOV_ERR_ARG;
  if (n == 0) return MCOV_OK;
  if (mem_kind == MCOV_MEM_DEVICE) {
    k_synth_ncigar<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*P, i0, n, out);
    return cudaGetLastError() == cudaSuccess ? MCOV_OK : MCOV_ERR_CUDA;
  }
  mcov_synth_params Q = *P;
  parallel_for(n, [=](int64_t k) { out[k] = mcov_synth_ncigar(&Q, i0 + k); });
  return MCOV_OK;
}

int mcov_synth_gen_reads(const mcov_synth_params* P, int64_t i0, int64_t n, const int64_t* read_start,
                    const int32_t* contig_len, int32_t n_contigs, int32_t tid_base, const uint32_t* cig_off, int32_t* tid,
                    int32_t* pos, uint16_t* flag, uint8_t* mapq, int32_t* isize, uint32_t* cig, int64_t* reflen_out,
                    int mem_kind, void* stream) {
  return synth_gen_reads_impl<uint32_t>(P, i0, n, read_start, contig_len, n_contigs, tid_base, cig_off, tid, pos, flag, mapq,
                                        isize, cig, reflen_out, mem_kind, stream);
}

int mcov_synth_gen_reads_wide(const mcov_synth_params* P, int64_t i0, int64_t n, const int64_t* read_start,
                    const int32_t* contig_len, int32_t n_contigs, int32_t tid_base, const uint64_t* cig_off, int32_t* tid,
                    int32_t* pos, uint16_t* flag, uint8_t* mapq, int32_t* isize, uint32_t* cig, int64_t* reflen_out,
                    int mem_kind, void* stream) {
  return synth_gen_reads_impl<uint64_t>(P, i0, n, read_start, contig_len, n_contigs, tid_base, cig_off, tid, pos, flag, mapq,
                                        isize, cig, reflen_out, mem_kind, stream);
}

}  // extern "C"
