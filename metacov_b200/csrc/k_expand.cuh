// k_expand.cuh -- K1: filter + CIGAR reduce + difference-array deltas.
//
// Replaces the per-record work of htslib's pileup engine that the reference
// reaches through `bam.pileup()` (reference metacov/pileup.py:13): which
// records count (pysam __advance_samtools) and the
// interval [pos, pos+reflen) each one covers (Appendix A-4).  A read is ONE
// interval however indel-heavy its CIGAR, so the "expansion" is a reduction
// over its ops followed by +1 @ pos and -1 @ pos+reflen.
//
// Mapping: four consecutive reads per thread, every SoA column fetched with one
// 128-bit load and all of them issued before the first use (one thread per read
// with the loads chained flag -> pos/offsets -> ops -> contig table left the
// kernel latency-bound: ncu long_scoreboard 77 %, 221 us on C2); CIGARs of up
// to kThreadOps ops are reduced by the owning thread, longer ones by the whole
// warp, one read at a time, with 128-bit loads (long-read case, where the
// CIGAR stream is the dominant HBM traffic).  Deltas go to L2 with
// red.global.add -- measured at 17 % of the L2 red-sector peak on C2, so no
// shared-memory binning stage is needed in front of it; the coordinate-sorted
// order keeps a warp's updates within a few cache lines.
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
  const uint32_t* cig_off;       // 32-bit op offsets (n+1 entries), or nullptr when cig_off64 is set
  const uint64_t* cig_off64;     // 64-bit op offsets for batches of 2^32 or more ops (config C5 at full size)
  const uint32_t* cig;
  const int64_t* contig_off;
  const int32_t* contig_len;
  int32_t n_contigs;
  mcov_filter filt;
  int32_t* delta;
  PassCounters* pc;
  int cig_aligned16;
  int64_t n_cig;                 // host side only: CIGAR ops in the batch, -1 = not known (picks the prep kernel)
};

// (the kernel itself, k_expand, lives in k_fused.cuh: it shares the vectorised load + filter + CIGAR
// reduction of the fused path's first kernel)

}  // namespace mcov
