// experimental.cu -- the second coverage definition of the reference,
// `metacov.pileup.experimental` (reference metacov/pileup.py:38-173), for a
// batch of regions on the GPU.
//
// What the reference does per region, restated as reductions (every read-level
// rule cites the line it follows):
//   reads   = bam.fetch(ref, start, end): every record overlapping the region,
//             unfiltered, file order (pileup.py:90)
//   secondary / improper: counted and skipped (pileup.py:92-99)
//   rcor    = k_cor[0 if is_read1 else 1][first k_len bases of the aligned
//             sequence]; missing key or 0 -> 1 (pileup.py:121-130)
//   interval: forward [rs, rs+reflen), REVERSE [rs-reflen, rs) (pileup.py:
//             132-137; the quirk is reproduced, not fixed), clamped to the region
//   cov     = mean over the region of the interval count      -> sum of clipped lengths
//   covc    = the same weighted by 1/rcor                      -> sum of clipped length / rcor
//   starts[rstart] = 1, cor[rstart] = 1/rcor (LAST read in file order wins),
//   nreads += 1 when 0 <= rstart < length (pileup.py:143-146) -> per-position "last writer"
//   mates   : reads are joined through a dict keyed by query_name (pileup.py:
//             101-118): the 1st/2nd, 3rd/4th ... read of a name form pairs; a pair
//             adds 1 to cov2[s-1:e+1] (Python slice semantics, negative indices
//             wrap) and 1/(k_cor[strand_a][kmer_a]*k_cor[strand_b][kmer_b]) to wnf
//             (missing key or zero product -> 1)
// Integer outputs are exact; the float sums are accumulated in a fixed order
// (deterministic), within 1e-12 relative of the reference's file-order sums.
//
// Mapping: k_exp_prep derives the per-read quantities once; k_exp_entries
// writes one (name hash, read) entry per counted read of every region; a CUB
// stable segmented sort brings equal names together (file order kept inside a
// name); k_exp_region (one CTA per region) does the reductions, resolves the
// last-writer table and walks the sorted entries pairing consecutive reads.
#include <cub/device/device_segmented_sort.cuh>

#include <algorithm>
#include <vector>

#include "ctx.cuh"

namespace mcov {

constexpr int kExpThreads = 256;
constexpr uint8_t kClsSecondary = 0, kClsImproper = 1, kClsCounted = 2;
constexpr uint8_t kStrandKeyOk = 0x80;      // k_cor[strand] holds the read's k-mer

struct ExpReads {
  int64_t n;
  const int32_t* pos; const uint16_t* flag;
  const uint32_t* cig_off; const uint32_t* cig;
  const int32_t* kmer;          // 2-bit code of the first k_len aligned bases, -1 = no such key
  const uint64_t* name_hash;
  const double* kcor;           // [2][n_kmers]
  const uint8_t* kcor_has;      // [2][n_kmers]
  int64_t n_kmers;
  // derived (k_exp_prep)
  int32_t* rs;                  // start of the interval the reference covers (may be negative)
  int32_t* rl;                  // pysam reference_length (>= 1), 0 = None
  int32_t* endpos;              // bam_endpos, for the fetch overlap test
  uint8_t* cls;                 // kCls* | kStrandKeyOk
  double* w;                    // 1 / rcor
  double* ks;                   // k_cor[1 if reverse else 0][kmer]
};

__global__ void k_exp_prep(ExpReads R) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R.n) return;
  const uint32_t f = R.flag[i];
  const uint32_t c0 = R.cig_off[i], c1 = R.cig_off[i + 1];
  long long reflen = 0;
  for (uint32_t k = c0; k < c1; ++k) reflen += cigar_ref_len(R.cig[k]);
  if (reflen > 0x7fffffffll) reflen = 0x7fffffffll;
  const int32_t p = R.pos[i];
  const bool unmapped = (f & 0x4u) != 0;
  // pysam: reference_length is None when unmapped or without CIGAR, else bam_endpos - pos (>= 1)
  const int32_t rl = (unmapped || c1 == c0) ? 0 : (reflen > 0 ? (int32_t)reflen : 1);
  const long long el = unmapped ? 0 : reflen;
  R.endpos[i] = (int32_t)min((long long)p + (el > 0 ? el : 1), 0x7fffffffll);
  R.rl[i] = rl;
  R.rs[i] = (f & 0x10u) ? p - rl : p;                      // pileup.py:132-137
  uint8_t cls = (f & 0x100u) ? kClsSecondary : ((f & 0x2u) ? kClsCounted : kClsImproper);   // pileup.py:92-99
  const int32_t km = R.kmer[i];
  double w = 1.0, ks = 0.0;
  if (km >= 0 && km < R.n_kmers) {
    const int readno = (f & 0x40u) ? 0 : 1;                  // pileup.py:121
    if (R.kcor_has[(int64_t)readno * R.n_kmers + km]) {
      double rcor = R.kcor[(int64_t)readno * R.n_kmers + km];
      if (rcor == 0.0) rcor = 1.0;                           // pileup.py:128-130
      w = 1.0 / rcor;
    }
    const int strand = (f & 0x10u) ? 1 : 0;                  // pileup.py:108-110
    if (R.kcor_has[(int64_t)strand * R.n_kmers + km]) { ks = R.kcor[(int64_t)strand * R.n_kmers + km]; cls |= kStrandKeyOk; }
  }
  R.cls[i] = cls;
  R.w[i] = w;
  R.ks[i] = ks;
}

struct ExpRegions {
  int32_t g;
  const int32_t* start; const int32_t* end;
  const int64_t* lb; const int64_t* ub;       // candidate reads [lb, ub) of each region (file order)
  const int64_t* seg;                         // [g+1] entry segment of each region (capacity ub-lb)
  const int64_t* last_off;                    // [g+1] offsets into the last-writer table
  int32_t* last;                              // -1 everywhere between runs
  uint64_t* key; int32_t* val;                // entries, then sorted entries
  mcov_exp_stats* out;
};

__global__ void k_exp_entries(ExpReads R, ExpRegions G) {
  const int g = blockIdx.x;
  const int64_t lb = G.lb[g], ub = G.ub[g], seg = G.seg[g];
  const int32_t start = G.start[g], end = G.end[g];
  for (int64_t i = lb + threadIdx.x; i < ub; i += blockDim.x) {
    const bool fetched = R.endpos[i] > start && R.pos[i] < end;
    const bool counted = fetched && (R.cls[i] & 3) == kClsCounted;
    G.key[seg + (i - lb)] = counted ? R.name_hash[i] : ~0ull;
    G.val[seg + (i - lb)] = counted ? (int32_t)i : -1;
  }
}

// Python / numpy slice a[i:j] of a length-L vector: number of elements
__device__ __forceinline__ long long py_slice_len(long long i, long long j, long long L) {
  if (i < 0) { i += L; if (i < 0) i = 0; } else if (i > L) i = L;
  if (j < 0) { j += L; if (j < 0) j = 0; } else if (j > L) j = L;
  return j > i ? j - i : 0;
}

template <typename T>
__device__ __forceinline__ T block_sum_fixed(T v, T* s_buf) {      // fixed-order tree: deterministic
  const int t = threadIdx.x;
  __syncthreads();
  s_buf[t] = v;
  __syncthreads();
  for (int o = kExpThreads / 2; o > 0; o >>= 1) {
    if (t < o) s_buf[t] += s_buf[t + o];
    __syncthreads();
  }
  return s_buf[0];
}

__global__ void __launch_bounds__(kExpThreads)
k_exp_region(ExpReads R, ExpRegions G) {
  __shared__ double s_d[kExpThreads];
  __shared__ long long s_l[kExpThreads];
  __shared__ int s_head[kExpThreads / 32];
  const int g = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t lb = G.lb[g], ub = G.ub[g];
  const int32_t start = G.start[g], end = G.end[g];
  const long long L = (long long)end - start;
  int32_t* last = G.last + G.last_off[g];

  // ---- pass over the fetched reads (pileup.py:90-146) ----
  long long cov = 0, nreads = 0, sec = 0, imp = 0, bad = 0;
  double covw = 0.0;
  for (int64_t i = lb + t; i < ub; i += kExpThreads) {
    if (!(R.endpos[i] > start && R.pos[i] < end)) continue;
    const uint8_t c = R.cls[i] & 3;
    if (c == kClsSecondary) { ++sec; continue; }
    if (c == kClsImproper) { ++imp; continue; }
    const int32_t rl = R.rl[i];
    if (rl == 0) { ++bad; continue; }                        // reference_length None: the reference raises TypeError
    const long long rstart = (long long)R.rs[i] - start, rend = rstart + rl;
    const long long a = rstart > 0 ? rstart : 0, b = rend < L ? rend : L;
    if (b > a) { cov += b - a; covw += (double)(b - a) * R.w[i]; }
    if (rstart >= 0 && rstart < L) { ++nreads; atomicMax(&last[rstart], (int32_t)i); }
  }
  cov = block_sum_fixed(cov, s_l); nreads = block_sum_fixed(nreads, s_l);
  sec = block_sum_fixed(sec, s_l); imp = block_sum_fixed(imp, s_l); bad = block_sum_fixed(bad, s_l);
  covw = block_sum_fixed(covw, s_d);          // (its barriers also order the atomicMax's before the sweep below)

  // ---- last-writer table -> starts / cor (pileup.py:143-145); leave it at -1 for the next run ----
  long long n_starts = 0;
  double cor = 0.0;
  for (long long p = t; p < L; p += kExpThreads) {
    const int32_t li = __ldcg(last + p);                      // written by atomics (L2): do not trust L1
    if (li >= 0) { ++n_starts; cor += R.w[li]; last[p] = -1; }
  }
  n_starts = block_sum_fixed(n_starts, s_l);
  cor = block_sum_fixed(cor, s_d);

  // ---- mates: consecutive reads of one name in the sorted entries (pileup.py:101-118) ----
  const int64_t s0 = G.seg[g], s1 = G.seg[g + 1];
  long long cov2 = 0, n_pairs = 0;
  double wnf = 0.0;
  long long carry_head = s0;                  // index of the head of the run that reaches into this chunk
  for (int64_t base = s0; base < s1; base += kExpThreads) {
    const int64_t p = base + t;
    const bool in = p < s1;
    const uint64_t k = in ? G.key[p] : ~0ull;
    const bool live = in && k != ~0ull;
    const bool head = live && (p == s0 || G.key[p - 1] != k);
    // chunk-relative index of the last head at or before this entry (max-scan), -1 = none in this chunk
    int hr = head ? t : -1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, hr, o); if (lane >= o) hr = max(hr, y); }
    __syncthreads();                          // s_head free (previous chunk's readers are done)
    if (lane == 31) s_head[warp] = hr;
    __syncthreads();
    for (int w2 = 0; w2 < warp; ++w2) hr = max(hr, s_head[w2]);
    const long long h = hr >= 0 ? base + hr : carry_head;    // the run started in an earlier chunk
    if (live) {
      const long long rank = p - h;
      if (rank & 1) {                         // second read of a pair: a = this (later in the file), b = the one before
        const int32_t ia = G.val[p], ib = G.val[p - 1];
        const long long pa = R.pos[ia], pb = R.pos[ib];
        const long long s = (pa < pb ? pa : pb) - start, e = (pa > pb ? pa : pb) - start;
        cov2 += py_slice_len(s - 1, e + 1, L);
        double term = 1.0;                    // KeyError -> 1 (pileup.py:112-113)
        if ((R.cls[ia] & kStrandKeyOk) && (R.cls[ib] & kStrandKeyOk)) {
          const double prod = R.ks[ia] * R.ks[ib];
          term = prod == 0.0 ? 1.0 : 1.0 / prod;             // ZeroDivisionError -> 1 (pileup.py:114-115)
        }
        wnf += term;
        ++n_pairs;
      }
    }
    // carry: the last head of this chunk, if it has one
    int last_head = -1;
    for (int w2 = 0; w2 < kExpThreads / 32; ++w2) last_head = max(last_head, s_head[w2]);
    if (last_head >= 0) carry_head = base + last_head;
  }
  cov2 = block_sum_fixed(cov2, s_l); n_pairs = block_sum_fixed(n_pairs, s_l);
  wnf = block_sum_fixed(wnf, s_d);
  if (t == 0) {
    mcov_exp_stats o;
    o.covw_sum = covw; o.cor_sum = cor; o.wnf_sum = wnf;
    o.cov_sum = cov; o.cov2_sum = cov2;
    o.n_starts = (int32_t)n_starts; o.nreads = (int32_t)nreads; o.secondary = (int32_t)sec; o.improper = (int32_t)imp;
    o.no_reflen = (int32_t)bad; o.n_pairs = (int32_t)n_pairs;
    G.out[g] = o;
  }
}

// cor_revsum[i] = sum_{j < min(L-i, n_w)} w[j] * cor_rev[i+j]   (pileup.py:78-83: np.inner per position)
__global__ void k_revsum(const double* __restrict__ cor_rev, const double* __restrict__ w, int64_t L, int32_t n_w,
                         double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const int64_t l = min(L - i, (int64_t)n_w);
  double acc = 0.0;
  for (int64_t j = 0; j < l; ++j) acc += w[j] * cor_rev[i + j];
  out[i] = acc;
}

}  // namespace mcov

using namespace mcov;

namespace {
struct Scratch {                      // freed on every exit path
  std::vector<void*> dev;
  ~Scratch() { for (void* p : dev) cudaFree(p); }
  template <typename T> cudaError_t get(T** p, size_t n) {
    *p = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) dev.push_back(*p);
    return e;
  }
};
int efail(mcov_ctx* c, int code, const char* what) { if (c) c->err = what; return code; }
}  // namespace

#define XCU(call) do { if ((call) != cudaSuccess) { cudaGetLastError(); return efail(ctx, MCOV_ERR_CUDA, #call); } } while (0)

extern "C" int mcov_experimental_run(mcov_ctx* ctx, int64_t n, const int32_t* pos, const uint16_t* flag,
                                     const uint32_t* cig_off, const uint32_t* cig, const uint64_t* name_hash,
                                     const int32_t* kmer_code, int32_t k_len, const double* kcor, const uint8_t* kcor_has,
                                     int32_t g, const int32_t* r_start, const int32_t* r_end, const int64_t* r_lb,
                                     const int64_t* r_ub, mcov_exp_stats* out) {
  if (!ctx) return MCOV_ERR_ARG;
  if (n < 0 || g < 0 || k_len < 0 || k_len > 12) return efail(ctx, MCOV_ERR_ARG, "mcov_experimental_run: bad sizes (k_len <= 12)");
  if (g == 0) return MCOV_OK;
  if (!r_start || !r_end || !r_lb || !r_ub || !out) return efail(ctx, MCOV_ERR_ARG, "mcov_experimental_run: null region array");
  if (n > 0 && (!pos || !flag || !cig_off || !name_hash || !kmer_code)) return efail(ctx, MCOV_ERR_ARG, "mcov_experimental_run: null read array");
  if (n > INT32_MAX) return efail(ctx, MCOV_ERR_RANGE, "mcov_experimental_run: more than 2^31-1 reads");
  XCU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int64_t n_kmers = (k_len > 0 && kcor && kcor_has) ? (1ll << (2 * k_len)) : 0;
  std::vector<int64_t> seg((size_t)g + 1), loff((size_t)g + 1);
  seg[0] = 0; loff[0] = 0;
  for (int32_t i = 0; i < g; ++i) {
    if (r_lb[i] < 0 || r_ub[i] < r_lb[i] || r_ub[i] > n || r_end[i] < r_start[i])
      return efail(ctx, MCOV_ERR_ARG, "mcov_experimental_run: bad region / read range");
    seg[i + 1] = seg[i] + (r_ub[i] - r_lb[i]);
    loff[i + 1] = loff[i] + ((int64_t)r_end[i] - r_start[i]);
  }
  const int64_t E = seg[g], LL = loff[g];
  if (E > INT32_MAX) return efail(ctx, MCOV_ERR_RANGE, "mcov_experimental_run: too many (region, read) entries in one call");
  const uint32_t n_cig = n > 0 ? cig_off[n] : 0;

  Scratch M;
  ExpReads R;
  R.n = n; R.n_kmers = n_kmers;
  int32_t *d_pos, *d_kmer, *d_rs, *d_rl, *d_end; uint16_t* d_flag; uint32_t *d_off, *d_cig; uint64_t* d_hash;
  double *d_kcor, *d_w, *d_ks; uint8_t *d_has, *d_cls;
  XCU(M.get(&d_pos, n)); XCU(M.get(&d_flag, n)); XCU(M.get(&d_off, n + 1)); XCU(M.get(&d_cig, n_cig));
  XCU(M.get(&d_hash, n)); XCU(M.get(&d_kmer, n)); XCU(M.get(&d_kcor, 2 * n_kmers)); XCU(M.get(&d_has, 2 * n_kmers));
  XCU(M.get(&d_rs, n)); XCU(M.get(&d_rl, n)); XCU(M.get(&d_end, n)); XCU(M.get(&d_cls, n)); XCU(M.get(&d_w, n)); XCU(M.get(&d_ks, n));
  if (n > 0) {
    XCU(cudaMemcpyAsync(d_pos, pos, n * 4, cudaMemcpyHostToDevice, s));
    XCU(cudaMemcpyAsync(d_flag, flag, n * 2, cudaMemcpyHostToDevice, s));
    XCU(cudaMemcpyAsync(d_off, cig_off, (n + 1) * 4, cudaMemcpyHostToDevice, s));
    if (n_cig) XCU(cudaMemcpyAsync(d_cig, cig, (size_t)n_cig * 4, cudaMemcpyHostToDevice, s));
    XCU(cudaMemcpyAsync(d_hash, name_hash, n * 8, cudaMemcpyHostToDevice, s));
    XCU(cudaMemcpyAsync(d_kmer, kmer_code, n * 4, cudaMemcpyHostToDevice, s));
  }
  if (n_kmers) {
    XCU(cudaMemcpyAsync(d_kcor, kcor, (size_t)2 * n_kmers * 8, cudaMemcpyHostToDevice, s));
    XCU(cudaMemcpyAsync(d_has, kcor_has, (size_t)2 * n_kmers, cudaMemcpyHostToDevice, s));
  }
  R.pos = d_pos; R.flag = d_flag; R.cig_off = d_off; R.cig = d_cig; R.kmer = d_kmer; R.name_hash = d_hash;
  R.kcor = d_kcor; R.kcor_has = d_has; R.rs = d_rs; R.rl = d_rl; R.endpos = d_end; R.cls = d_cls; R.w = d_w; R.ks = d_ks;

  ExpRegions G;
  G.g = g;
  int32_t *d_rstart, *d_rend, *d_last, *d_val, *d_val2; int64_t *d_lb, *d_ub, *d_seg, *d_loff; uint64_t *d_key, *d_key2;
  mcov_exp_stats* d_out;
  XCU(M.get(&d_rstart, g)); XCU(M.get(&d_rend, g)); XCU(M.get(&d_lb, g)); XCU(M.get(&d_ub, g));
  XCU(M.get(&d_seg, g + 1)); XCU(M.get(&d_loff, g + 1)); XCU(M.get(&d_last, LL));
  XCU(M.get(&d_key, E)); XCU(M.get(&d_val, E)); XCU(M.get(&d_key2, E)); XCU(M.get(&d_val2, E)); XCU(M.get(&d_out, g));
  XCU(cudaMemcpyAsync(d_rstart, r_start, (size_t)g * 4, cudaMemcpyHostToDevice, s));
  XCU(cudaMemcpyAsync(d_rend, r_end, (size_t)g * 4, cudaMemcpyHostToDevice, s));
  XCU(cudaMemcpyAsync(d_lb, r_lb, (size_t)g * 8, cudaMemcpyHostToDevice, s));
  XCU(cudaMemcpyAsync(d_ub, r_ub, (size_t)g * 8, cudaMemcpyHostToDevice, s));
  XCU(cudaMemcpyAsync(d_seg, seg.data(), ((size_t)g + 1) * 8, cudaMemcpyHostToDevice, s));
  XCU(cudaMemcpyAsync(d_loff, loff.data(), ((size_t)g + 1) * 8, cudaMemcpyHostToDevice, s));
  if (LL) XCU(cudaMemsetAsync(d_last, 0xff, (size_t)LL * 4, s));
  G.start = d_rstart; G.end = d_rend; G.lb = d_lb; G.ub = d_ub; G.seg = d_seg; G.last_off = d_loff; G.last = d_last;
  G.key = d_key; G.val = d_val; G.out = d_out;

  if (n > 0) {
    MCOV_LAUNCH(ctx, kKExpPrep, (k_exp_prep<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(R)));
    XCU(cudaGetLastError());
  }
  MCOV_LAUNCH(ctx, kKExpEntries, (k_exp_entries<<<(unsigned)g, 256, 0, s>>>(R, G)));
  XCU(cudaGetLastError());
  if (E > 0) {
    size_t tb = 0;
    XCU(cub::DeviceSegmentedSort::StableSortPairs(nullptr, tb, d_key, d_key2, d_val, d_val2, (int)E, (int)g, d_seg, d_seg + 1, s));
    void* d_tmp;
    XCU(M.get(reinterpret_cast<char**>(&d_tmp), tb + 16));
    XCU(cub::DeviceSegmentedSort::StableSortPairs(d_tmp, tb, d_key, d_key2, d_val, d_val2, (int)E, (int)g, d_seg, d_seg + 1, s));
    G.key = d_key2; G.val = d_val2;
  }
  MCOV_LAUNCH(ctx, kKExpRegion, (k_exp_region<<<(unsigned)g, kExpThreads, 0, s>>>(R, G)));
  XCU(cudaGetLastError());
  XCU(cudaMemcpyAsync(out, d_out, (size_t)g * sizeof(mcov_exp_stats), cudaMemcpyDeviceToHost, s));
  XCU(cudaStreamSynchronize(s));
  return MCOV_OK;
}

extern "C" int mcov_exp_revsum(mcov_ctx* ctx, int64_t L, const double* cor_rev, int32_t n_w, const double* w, double* out) {
  if (!ctx) return MCOV_ERR_ARG;
  if (L < 0 || n_w < 0 || (L > 0 && (!cor_rev || !out)) || (n_w > 0 && !w)) return efail(ctx, MCOV_ERR_ARG, "mcov_exp_revsum: bad arguments");
  if (L == 0) return MCOV_OK;
  XCU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  Scratch M;
  double *d_rev, *d_w, *d_out;
  XCU(M.get(&d_rev, L)); XCU(M.get(&d_w, n_w)); XCU(M.get(&d_out, L));
  XCU(cudaMemcpyAsync(d_rev, cor_rev, (size_t)L * 8, cudaMemcpyHostToDevice, s));
  if (n_w) XCU(cudaMemcpyAsync(d_w, w, (size_t)n_w * 8, cudaMemcpyHostToDevice, s));
  MCOV_LAUNCH(ctx, kKExpRevsum, (k_revsum<<<(unsigned)((L + 127) / 128), 128, 0, s>>>(d_rev, d_w, L, n_w, d_out)));
  XCU(cudaGetLastError());
  XCU(cudaMemcpyAsync(out, d_out, (size_t)L * 8, cudaMemcpyDeviceToHost, s));
  XCU(cudaStreamSynchronize(s));
  return MCOV_OK;
}
