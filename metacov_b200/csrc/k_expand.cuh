// k_expand.cuh -- K1: filter + CIGAR reduce + difference-array deltas.
//
// Replaces the per-record work of htslib's pileup engine that the reference
// reaches through `bam.pileup()` (reference metacov/pileup.py:13): which
// records count (pysam __advance_samtools) and the
// interval [pos, pos+reflen) each one covers (Appendix A-4).  A read is ONE
// interval however indel-heavy its CIGAR, so the "expansion" is a reduction
// over its ops followed by +1 @ pos and -1 @ pos+reflen.
//
// Mapping: one thread per read for short CIGARs (<= kThreadOps ops, the
// 150 bp case: adjacent threads read adjacent ops, so the loads coalesce);
// reads with longer CIGARs are handed to the whole warp, one read at a time,
// and reduced with 128-bit loads (long-read case, where the CIGAR stream is
// the dominant HBM traffic).  Deltas go to L2 with red.global.add; the
// coordinate-sorted order keeps a warp's updates within a few cache lines.
//
// HBM bytes per launch (algorithmic): 15*R + 4*sum(n_cigar of passing reads)
// + 8*R_pass (two 4-byte RMWs).
#pragma once
#include "common.cuh"

namespace mcov {

constexpr int kExpandThreads = 256;
constexpr uint32_t kThreadOps = 4;   // CIGARs up to this many ops are reduced by the owning thread

struct ExpandArgs {
  int64_t n;
  const int32_t* tid;
  const int32_t* pos;
  const uint16_t* flag;
  const uint8_t* mapq;
  const uint32_t* cig_off;
  const uint32_t* cig;
  const int64_t* contig_off;
  const int32_t* contig_len;
  int32_t n_contigs;
  mcov_filter filt;
  int32_t* delta;
  PassCounters* pc;
  int cig_aligned16;
};

// Per-thread part shared by the push path and the fused path: returns true if
// the read passes, with its clipped slot interval [s,e) and raw reflen.
__device__ __forceinline__ bool expand_one(const ExpandArgs& a, int64_t i, bool in_range, int lane,
                                           int64_t& s_slot, int64_t& e_slot, unsigned long long& reflen) {
  bool pass = false;
  int32_t t = -1, p = 0;
  uint32_t o0 = 0, o1 = 0;
  if (in_range) {
    uint32_t f = a.flag[i];
    uint32_t q = a.mapq[i];
    t = a.tid[i];
    pass = read_passes(f, q, a.filt) && t >= 0 && t < a.n_contigs;
    if (pass) {
      p = a.pos[i];
      o0 = a.cig_off[i];
      o1 = a.cig_off[i + 1];
    }
  }
  reflen = 0;
  uint32_t nc = o1 - o0;
  bool mine = pass && nc <= kThreadOps;
  if (mine) {
    for (uint32_t k = o0; k < o1; ++k) reflen += cigar_ref_len(__ldg(a.cig + k));
  }
  // warp-cooperative reduction of the long CIGARs in this warp
  unsigned todo = __ballot_sync(0xffffffffu, pass && !mine);
  while (todo) {
    int src = __ffs(todo) - 1;
    todo &= todo - 1;
    uint32_t b0 = __shfl_sync(0xffffffffu, o0, src);
    uint32_t b1 = __shfl_sync(0xffffffffu, o1, src);
    unsigned long long r = warp_cigar_reflen(a.cig, b0, b1, lane, a.cig_aligned16 != 0);
    if (lane == src) reflen = r;
  }
  if (!pass) return false;
  int64_t len = a.contig_len[t];
  int64_t s = p, e = (int64_t)p + (int64_t)reflen;
  s = s < 0 ? 0 : (s > len ? len : s);
  e = e < 0 ? 0 : (e > len ? len : e);
  if (e <= s) return false;          // reflen == 0 or entirely outside the contig
  int64_t base = a.contig_off[t];
  s_slot = base + s;
  e_slot = base + e;
  return true;
}

__global__ void __launch_bounds__(kExpandThreads)
k_expand(ExpandArgs a) {
  const int lane = threadIdx.x & 31;
  unsigned long long n_pass = 0, aligned = 0;
  int unsorted = 0;
  const int64_t stride = (int64_t)gridDim.x * kExpandThreads;
  // whole warps iterate together (the cooperative path needs all 32 lanes)
  const int64_t n_round = (a.n + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * kExpandThreads + threadIdx.x; i < n_round; i += stride) {
    bool in_range = i < a.n;
    int64_t s, e;
    unsigned long long reflen;
    bool ok = expand_one(a, i, in_range, lane, s, e, reflen);
    if (ok) {
      atomicAdd(a.delta + s, 1);
      atomicAdd(a.delta + e, -1);
      n_pass += 1;
      aligned += reflen;
    }
    if (in_range && i > 0) {
      int32_t t0 = a.tid[i - 1], t1 = a.tid[i];
      // unmapped-without-coordinates (tid -1) sort last in a BAM
      uint32_t u0 = (uint32_t)t0, u1 = (uint32_t)t1;
      if (u1 < u0 || (u1 == u0 && a.pos[i] < a.pos[i - 1])) unsorted = 1;
    }
  }
  n_pass = warp_sum(n_pass);
  aligned = warp_sum(aligned);
  unsorted = __any_sync(0xffffffffu, unsorted);
  __shared__ unsigned long long s_np[kExpandThreads / 32], s_al[kExpandThreads / 32];
  __shared__ int s_un[kExpandThreads / 32];
  int w = threadIdx.x >> 5;
  if (lane == 0) { s_np[w] = n_pass; s_al[w] = aligned; s_un[w] = unsorted; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long np = 0, al = 0; int un = 0;
    for (int k = 0; k < kExpandThreads / 32; ++k) { np += s_np[k]; al += s_al[k]; un |= s_un[k]; }
    if (np) atomicAdd(&a.pc->n_pass, np);
    if (al) atomicAdd(&a.pc->aligned_bases, al);
    if (un) atomicOr(&a.pc->unsorted, 1);
  }
}

}  // namespace mcov
