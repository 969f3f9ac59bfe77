/* mcov_synth.h -- deterministic synthetic read generator.
 *
 * Counter-based (Philox-4x32-10 keyed by (seed, stream), counter = read index)
 * and integer-only, so the host (gcc) and the device (nvcc) produce identical
 * reads from the same header and any read can be re-derived on its own.
 * Plain C99 / CUDA C++; every function is `static inline`.
 *
 * The reference has no generator for mapped reads (its simulator wraps the
 * external art_illumina, reference metacov/simulate.py:9-50, out of scope);
 * this replaces it for benchmarks and parity tests.
 */
#ifndef MCOV_SYNTH_H
#define MCOV_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MCOV_HD __host__ __device__ __forceinline__
#else
#define MCOV_HD static inline
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcov_synth_params {
  uint64_t seed;
  int32_t  mode;        /* 0 = short reads (fixed query length), 1 = long reads   */
  int32_t  read_len;    /* short: query length (150)                              */
  int32_t  span_min;    /* long: target reference span, uniform [span_min,span_max] */
  int32_t  span_max;
  int32_t  margin;      /* positions are drawn in [0, len - margin]; margin >= max reflen */
  int32_t  reserved;
} mcov_synth_params;

typedef struct mcov_philox4 { uint32_t v[4]; } mcov_philox4;

MCOV_HD mcov_philox4 mcov_philox4x32(uint64_t ctr_lo, uint32_t ctr_hi, uint32_t ctr_sub, uint64_t key)
{
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = ctr_hi, c3 = ctr_sub;
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  mcov_philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

/* contig of read i: largest c with read_start[c] <= i (read_start has n_contigs+1 entries) */
MCOV_HD int32_t mcov_synth_contig_of(int64_t i, const int64_t* read_start, int32_t n_contigs)
{
  int32_t lo = 0, hi = n_contigs;          /* invariant: read_start[lo] <= i < read_start[hi] */
  while (hi - lo > 1) {
    int32_t mid = lo + ((hi - lo) >> 1);
    if (read_start[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

/* BAM flag of read i from one random word (mix of SURVEY.md 8(d) C2). */
MCOV_HD uint16_t mcov_synth_flag(uint32_t w, int32_t mode)
{
  if (mode == 1) {                         /* long reads: unpaired */
    uint32_t r = w % 100u;
    uint16_t f = (uint16_t)(((w >> 16) & 1u) ? 16u : 0u);
    if (r == 0u) f |= 0x800u;              /* supplementary: kept by the default filter */
    else if (r == 1u) f |= 0x100u;         /* secondary: dropped                         */
    return f;
  }
  uint32_t r = w % 1000u, s = w >> 16;
  const uint16_t proper[4] = {99, 147, 83, 163};
  const uint16_t orphan[6] = {97, 145, 81, 161, 73, 137};
  const uint16_t unmapd[4] = {69, 133, 101, 165};
  if (r < 960u) return proper[s & 3u];
  if (r < 980u) return orphan[s % 6u];
  if (r < 988u) return unmapd[s & 3u];
  if (r < 993u) return (uint16_t)(proper[s & 3u] | 0x100u);
  if (r < 997u) return (uint16_t)(proper[s & 3u] | 0x400u);
  return (uint16_t)(proper[s & 3u] | 0x200u);
}

/* number of CIGAR ops of read i */
MCOV_HD uint32_t mcov_synth_ncigar(const mcov_synth_params* P, int64_t i)
{
  mcov_philox4 a = mcov_philox4x32((uint64_t)i, 0u, 0u, P->seed);
  uint16_t f = mcov_synth_flag(a.v[2], P->mode);
  if (f & 0x4u) return 0u;                 /* unmapped-with-coordinates: no CIGAR */
  if (P->mode == 0) {
    uint32_t r = a.v[1] % 1000u;
    if (r < 900u) return 1u;               /* 150M            */
    if (r < 950u) return 2u;               /* aS(150-a)M      */
    return 3u;                             /* aM bI cM / aM bD cM */
  }
  uint32_t span = (uint32_t)P->span_min + a.v[1] % (uint32_t)(P->span_max - P->span_min + 1);
  uint32_t m = span / 22u;                 /* (M,indel) pairs; mean 20+2 ref/query bases each */
  return 2u * m + 1u + (a.v[3] & 1u) + ((a.v[3] >> 1) & 1u);   /* optional leading / trailing S */
}

/* k-th CIGAR op of a long read (op index k in [0, n_cigar)) */
MCOV_HD uint32_t mcov_synth_long_op(const mcov_synth_params* P, int64_t i, uint32_t k, uint32_t n_cigar,
                                    uint32_t lead_s, uint32_t trail_s)
{
  if (lead_s && k == 0u) return ((17u + (uint32_t)(i & 63)) << 4) | 4u;
  if (trail_s && k == n_cigar - 1u) return ((9u + (uint32_t)(i & 31)) << 4) | 4u;
  uint32_t j = k - lead_s;                 /* index inside the M (I|D) M ... M body */
  mcov_philox4 b = mcov_philox4x32((uint64_t)i, 1u + (j >> 2), 1u, P->seed);
  uint32_t w = b.v[j & 3u];
  if ((j & 1u) == 0u) return ((1u + (w & 0xFFFFu) % 39u) << 4) | 0u;        /* M 1..39 */
  uint32_t len = 1u + ((w >> 8) % 3u);                                      /* 1..3    */
  return (len << 4) | ((w & 1u) ? 1u : 2u);                                 /* I or D  */
}

/* Fill the fixed-size fields and the CIGAR of read i.  cig points at this
 * read's first op (n_cigar ops).  Returns the reference length. */
MCOV_HD int64_t mcov_synth_read(const mcov_synth_params* P, int64_t i,
                                const int64_t* read_start, const int32_t* contig_len, int32_t n_contigs,
                                int32_t* tid, int32_t* pos, uint16_t* flag, uint8_t* mapq,
                                int32_t* isize, uint32_t* cig, uint32_t n_cigar)
{
  mcov_philox4 a = mcov_philox4x32((uint64_t)i, 0u, 0u, P->seed);
  int32_t c = mcov_synth_contig_of(i, read_start, n_contigs);
  uint64_t j = (uint64_t)(i - read_start[c]);
  uint64_t n_c = (uint64_t)(read_start[c + 1] - read_start[c]);
  int64_t span = (int64_t)contig_len[c] - P->margin + 1;
  if (span < 1) span = 1;
  /* pos = floor((j + u) * span / n_c), u = 12-bit fraction: monotone in j */
  uint64_t num = ((j << 12) + (a.v[0] >> 20)) * (uint64_t)span;
  int32_t p = (int32_t)(num / (n_c << 12));
  uint16_t f = mcov_synth_flag(a.v[2], P->mode);
  *tid = c; *pos = p; *flag = f; *mapq = (uint8_t)(a.v[3] >> 8) % 61u;
  int64_t reflen = 0;
  if (P->mode == 0) {
    uint32_t L = (uint32_t)P->read_len;
    uint32_t r = a.v[1] % 1000u, s = a.v[1] >> 16;
    if (n_cigar == 1u) { cig[0] = (L << 4) | 0u; reflen = L; }
    else if (n_cigar == 2u) {
      uint32_t x = 1u + s % 30u;
      cig[0] = (x << 4) | 4u; cig[1] = ((L - x) << 4) | 0u; reflen = L - x;
    } else if (n_cigar == 3u) {
      uint32_t b = 1u + (s & 0xFFu) % 5u;
      if (r < 975u) {                       /* aM bI cM : query L, ref L-b */
        uint32_t x = 10u + (s >> 8) % (L - 30u);
        cig[0] = (x << 4) | 0u; cig[1] = (b << 4) | 1u; cig[2] = ((L - b - x) << 4) | 0u;
        reflen = L - b;
      } else {                              /* aM bD cM : query L, ref L+b */
        uint32_t x = 10u + (s >> 8) % (L - 30u);
        cig[0] = (x << 4) | 0u; cig[1] = (b << 4) | 2u; cig[2] = ((L - x) << 4) | 0u;
        reflen = L + b;
      }
    }
    /* template length: proper pairs get a plausible insert, sign by strand */
    int32_t ins = 200 + (int32_t)((a.v[3] >> 16) % 400u);
    *isize = (f & 0x1u) ? ((f & 0x10u) ? -ins : ins) : 0;
  } else {
    uint32_t lead_s = a.v[3] & 1u, trail_s = (a.v[3] >> 1) & 1u;
    for (uint32_t k = 0; k < n_cigar; ++k) {
      uint32_t op = mcov_synth_long_op(P, i, k, n_cigar, lead_s, trail_s);
      cig[k] = op;
      if ((0x18Du >> (op & 15u)) & 1u) reflen += op >> 4;
    }
    *isize = 0;
  }
  return reflen;
}

#ifdef __cplusplus
}
#endif
#endif /* MCOV_SYNTH_H */
