/* coverage.c -- CPU restatement (plain C) of the coverage hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference arm; never by metacov_b200.
 * PARITY UNPINNED at the pysam/htslib boundary (see oracle/__init__.py).
 *
 * Follows
 *   - reference metacov/pileup.py:9-26 (`classic`: per-base depth of a region,
 *     then min/max/med/std/avg/q23/sum) and metacov/cli.py:85-95 (region loop);
 *   - pysam libcalignmentfile.pyx `__advance_samtools` (read filter) and htslib
 *     sam.c `bam_plp_push` / `bam_plp_next` (column engine incl. maxcnt cap),
 *     restated in SURVEY.md Appendix A (pysam/htslib are un-vendored, unpinned:
 *     reference requirements.txt:2).
 *
 * Two depth formulations, cross-checked in tests/:
 *   orc_depth_diff  difference array + running sum (cap assumed idle)
 *   orc_depth_plp   the sequential htslib machine, one column at a time
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct orc_filter {
  uint16_t flag_filter, flag_require;
  uint8_t min_mapq, ignore_orphans, count_del, reflen0_as_one;   /* the two switches of SURVEY.md Appendix A-4 / 8(b) */
  int32_t max_depth;
} orc_filter;

typedef struct orc_stats {   /* same layout as mcov_region_stats */
  int64_t sum; uint64_t sumsq; int64_t iq_sum, n_ge1, n_geN;
  int32_t min, max, med_lo, med_hi, reserved, flags;
} orc_stats;

/* pysam __advance_samtools + bam_plp_push's own UNMAP drop (Appendix A-2) */
static int orc_pass(uint32_t flag, uint32_t mapq, const orc_filter* f) {
  if (flag & f->flag_filter) return 0;
  if (f->flag_require && !(flag & f->flag_require)) return 0;
  if (f->min_mapq > 0 && mapq < f->min_mapq) return 0;
  if (f->ignore_orphans && (flag & 0x1) && !(flag & 0x2)) return 0;
  if (flag & 0x4) return 0;
  return 1;
}

/* htslib bam_cigar2rlen: M D N = X consume reference */
static int64_t orc_reflen(const uint32_t* cig, uint32_t n) {
  int64_t r = 0;
  for (uint32_t k = 0; k < n; ++k) {
    uint32_t op = cig[k] & 15u;
    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) r += cig[k] >> 4;
  }
  return r;
}

int64_t orc_reflen_all(int64_t n, const uint32_t* cig_off, const uint32_t* cig, int64_t* out) {
  int64_t tot = 0;
  for (int64_t i = 0; i < n; ++i) { out[i] = orc_reflen(cig + cig_off[i], cig_off[i + 1] - cig_off[i]); tot += out[i]; }
  return tot;
}

/* Layout of the depth output: contig c occupies depth[off[c] .. off[c]+len[c]] (len+1 slots). */

typedef struct {
  int64_t i0, i1;
  const int32_t *tid, *pos; const uint16_t* flag; const uint8_t* mapq;
  const uint32_t *cig_off, *cig; const orc_filter* f;
  int32_t n_contigs; const int32_t* len; const int64_t* off; int32_t* depth;
  int64_t n_pass, aligned;
} diff_job;

static void diff_add(diff_job* j) {
  for (int64_t i = j->i0; i < j->i1; ++i) {
    int32_t t = j->tid[i];
    if (t < 0 || t >= j->n_contigs || !orc_pass(j->flag[i], j->mapq[i], j->f)) continue;
    int64_t L = j->len[t];
    if (!j->f->count_del) {
      /* only M = X positions count: one interval per run of such ops; D / N advance the position uncounted */
      int64_t p = j->pos[i], counted = 0, hit = 0;
      for (uint32_t k = j->cig_off[i]; k < j->cig_off[i + 1]; ++k) {
        uint32_t op = j->cig[k] & 15u; int64_t ln = j->cig[k] >> 4;
        if (op == 0 || op == 7 || op == 8) {
          int64_t s = p, e = p + ln;
          if (s < 0) s = 0; if (s > L) s = L;
          if (e < 0) e = 0; if (e > L) e = L;
          if (e > s) { j->depth[j->off[t] + s] += 1; j->depth[j->off[t] + e] -= 1; hit = 1; }
          counted += ln; p += ln;
        } else if (op == 2 || op == 3) p += ln;
      }
      if (counted == 0 && j->f->reflen0_as_one && orc_reflen(j->cig + j->cig_off[i], j->cig_off[i + 1] - j->cig_off[i]) == 0) {
        int64_t s = j->pos[i];
        if (s >= 0 && s < L) { j->depth[j->off[t] + s] += 1; j->depth[j->off[t] + s + 1] -= 1; hit = 1; counted = 1; }
      }
      if (hit) { j->n_pass++; j->aligned += counted; }
      continue;
    }
    int64_t rl = orc_reflen(j->cig + j->cig_off[i], j->cig_off[i + 1] - j->cig_off[i]);
    if (rl == 0 && j->f->reflen0_as_one) rl = 1;
    int64_t s = j->pos[i], e = s + rl;
    if (s < 0) s = 0; if (s > L) s = L;
    if (e < 0) e = 0; if (e > L) e = L;
    if (e <= s) continue;
    j->depth[j->off[t] + s] += 1;
    j->depth[j->off[t] + e] -= 1;
    j->n_pass++; j->aligned += rl;
  }
}

/* Difference-array depth, single thread, any read order.  depth must hold off[n_contigs] ints. */
int64_t orc_depth_diff(int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq,
                       const uint32_t* cig_off, const uint32_t* cig, const orc_filter* f, int32_t n_contigs,
                       const int32_t* len, const int64_t* off, int32_t* depth, int64_t* aligned_out) {
  memset(depth, 0, (size_t)off[n_contigs] * sizeof(int32_t));
  diff_job j = {0, n, tid, pos, flag, mapq, cig_off, cig, f, n_contigs, len, off, depth, 0, 0};
  diff_add(&j);
  for (int32_t c = 0; c < n_contigs; ++c) {
    int32_t run = 0;
    int32_t* d = depth + off[c];
    for (int64_t k = 0; k <= len[c]; ++k) { run += d[k]; d[k] = run; }
  }
  if (aligned_out) *aligned_out = j.aligned;
  return j.n_pass;
}

/* ---- contig-parallel variant for sorted input (the strong CPU baseline) ---- */
typedef struct {
  diff_job j; int32_t c0, c1; const int64_t* read_lo;   /* read_lo[c] = first read of contig c */
} par_job;

static void* par_run(void* arg) {
  par_job* p = (par_job*)arg;
  diff_job* j = &p->j;
  for (int32_t c = p->c0; c < p->c1; ++c) {
    int32_t* d = j->depth + j->off[c];
    memset(d, 0, (size_t)(j->off[c + 1] - j->off[c]) * sizeof(int32_t));
    j->i0 = p->read_lo[c]; j->i1 = p->read_lo[c + 1];
    diff_add(j);
    int32_t run = 0;
    for (int64_t k = 0; k <= j->len[c]; ++k) { run += d[k]; d[k] = run; }
  }
  return NULL;
}

/* reads must be sorted by tid (unplaced last).  Returns n_pass or -1 if unsorted. */
int64_t orc_depth_diff_par(int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq,
                           const uint32_t* cig_off, const uint32_t* cig, const orc_filter* f, int32_t n_contigs,
                           const int32_t* len, const int64_t* off, int32_t* depth, int64_t* aligned_out, int n_threads) {
  int64_t* read_lo = (int64_t*)malloc(((size_t)n_contigs + 1) * sizeof(int64_t));
  int64_t i = 0;
  for (int32_t c = 0; c <= n_contigs; ++c) {
    read_lo[c] = i;
    if (c < n_contigs) while (i < n && tid[i] == c) ++i;
  }
  for (int64_t k = i; k < n; ++k) if (tid[k] >= 0 && tid[k] < n_contigs) { free(read_lo); return -1; }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_contigs) n_threads = n_contigs;
  par_job* jobs = (par_job*)calloc((size_t)n_threads, sizeof(par_job));
  pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
  /* balance on slots + reads */
  int64_t total = 0;
  for (int32_t c = 0; c < n_contigs; ++c) total += 3 * ((int64_t)len[c] + 1) + 8 * (read_lo[c + 1] - read_lo[c]);
  int32_t c = 0;
  for (int t = 0; t < n_threads; ++t) {
    diff_job j = {0, 0, tid, pos, flag, mapq, cig_off, cig, f, n_contigs, len, off, depth, 0, 0};
    jobs[t].j = j; jobs[t].read_lo = read_lo; jobs[t].c0 = c;
    int64_t want = total * (t + 1) / n_threads, acc = 0;
    for (int32_t k = 0; k < c; ++k) acc += 3 * ((int64_t)len[k] + 1) + 8 * (read_lo[k + 1] - read_lo[k]);
    while (c < n_contigs && (acc < want || t == n_threads - 1)) { acc += 3 * ((int64_t)len[c] + 1) + 8 * (read_lo[c + 1] - read_lo[c]); ++c; }
    jobs[t].c1 = c;
  }
  jobs[n_threads - 1].c1 = n_contigs;
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, par_run, &jobs[t]);
  int64_t np = 0, al = 0;
  for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); np += jobs[t].j.n_pass; al += jobs[t].j.aligned; }
  if (aligned_out) *aligned_out = al;
  free(jobs); free(th); free(read_lo);
  return np;
}

/* ---- the htslib column engine (sam.c bam_plp_push / bam_plp_next), whole-file iterator ----
 * Emits depth[off[tid]+pos] = n for every column; returns the number of reads the maxcnt cap
 * dropped, or -1 on unsorted input.  One iterator over the whole file, i.e. the pileup a
 * region-less `bam.pileup()` would run; per-region iterators differ only when the cap fires. */
typedef struct { int32_t tid; int64_t beg, end; } plp_node;

int64_t orc_depth_plp(int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq,
                      const uint32_t* cig_off, const uint32_t* cig, const orc_filter* f, int32_t n_contigs,
                      const int32_t* len, const int64_t* off, int32_t* depth) {
  memset(depth, 0, (size_t)off[n_contigs] * sizeof(int32_t));
  int64_t cap = 1024, nbuf = 0, dropped = 0;
  plp_node* buf = (plp_node*)malloc((size_t)cap * sizeof(plp_node));
  int32_t it_tid = 0, max_tid = -1;
  int64_t it_pos = 0, max_pos = -1;
  int is_eof = 0;
  int64_t i = 0;
  const int64_t maxcnt = f->max_depth > 0 ? f->max_depth : INT64_MAX / 4;
  for (;;) {
    /* bam_plp_next: assemble columns while the iterator is behind the newest read */
    while ((is_eof && nbuf > 0) || max_tid > it_tid || (max_tid == it_tid && max_pos > it_pos)) {
      int64_t n_plp = 0, w = 0;
      for (int64_t k = 0; k < nbuf; ++k) {
        plp_node* p = &buf[k];
        if (p->tid < it_tid || (p->tid == it_tid && p->end <= it_pos)) continue;      /* retire */
        if (p->tid == it_tid && p->beg <= it_pos) ++n_plp;
        buf[w++] = *p;
      }
      nbuf = w;
      if (n_plp && it_tid >= 0 && it_tid < n_contigs && it_pos >= 0 && it_pos < len[it_tid])
        depth[off[it_tid] + it_pos] = (int32_t)n_plp;
      if (nbuf > 0) {
        if (it_tid < buf[0].tid) { it_tid = buf[0].tid; it_pos = buf[0].beg; }
        else if (it_pos < buf[0].beg) it_pos = buf[0].beg;
        else ++it_pos;
      } else ++it_pos;
      if (is_eof && nbuf == 0) break;
    }
    if (is_eof) break;
    /* read the next passing record (pysam __advance_samtools) and bam_plp_push it */
    while (i < n && !(tid[i] >= 0 && orc_pass(flag[i], mapq[i], f))) ++i;
    if (i >= n) { is_eof = 1; continue; }
    int32_t t = tid[i];
    int64_t p = pos[i], e = p + orc_reflen(cig + cig_off[i], cig_off[i + 1] - cig_off[i]);
    ++i;
    if (it_tid == t && it_pos == p && (1 + nbuf) > maxcnt) { ++dropped; continue; }
    if (t < max_tid || (t == max_tid && p < max_pos)) { free(buf); return -1; }
    max_tid = t; max_pos = p;
    if (e > it_pos || t > it_tid) {
      if (nbuf == cap) { cap *= 2; buf = (plp_node*)realloc(buf, (size_t)cap * sizeof(plp_node)); }
      buf[nbuf].tid = t; buf[nbuf].beg = p; buf[nbuf].end = e; ++nbuf;
    }
  }
  free(buf);
  return dropped;
}

/* ---- region statistics: exact integers behind pileup.py:18-26 ---- */
static int cmp_i32(const void* a, const void* b) {
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

typedef struct {
  const int32_t* depth; const int64_t* off; const int32_t* len;
  int64_t g0, g1; const int32_t *tid, *start, *end; int32_t breadth_n; orc_stats* out;
} stat_job;

static void* stat_run(void* arg) {
  stat_job* j = (stat_job*)arg;
  int64_t capn = 0;
  int32_t* tmp = NULL;
  for (int64_t g = j->g0; g < j->g1; ++g) {
    int64_t n = (int64_t)j->end[g] - j->start[g];
    orc_stats s;
    memset(&s, 0, sizeof(s));
    if (n <= 0) { j->out[g] = s; continue; }
    if (n > capn) { capn = n; tmp = (int32_t*)realloc(tmp, (size_t)capn * sizeof(int32_t)); }
    int64_t L = j->len[j->tid[g]];
    const int32_t* d = j->depth + j->off[j->tid[g]];
    for (int64_t k = 0; k < n; ++k) {             /* positions past the contig stay 0 (pileup.py:11) */
      int64_t p = j->start[g] + k;
      tmp[k] = p < L ? d[p] : 0;
    }
    s.min = INT32_MAX; s.max = INT32_MIN;
    for (int64_t k = 0; k < n; ++k) {
      int32_t v = tmp[k];
      s.sum += v; s.sumsq += (uint64_t)((int64_t)v * v);
      s.n_ge1 += v >= 1; s.n_geN += v >= j->breadth_n;
      if (v < s.min) s.min = v;
      if (v > s.max) s.max = v;
    }
    qsort(tmp, (size_t)n, sizeof(int32_t), cmp_i32);       /* sorted(columns), pileup.py:24 */
    int64_t q = n / 4;
    for (int64_t k = q; k < n - q; ++k) s.iq_sum += tmp[k];
    s.med_lo = tmp[(n - 1) / 2]; s.med_hi = tmp[n / 2];
    s.flags = 1;
    j->out[g] = s;
  }
  free(tmp);
  return NULL;
}

void orc_region_stats(const int32_t* depth, const int64_t* off, const int32_t* len, int64_t g, const int32_t* tid,
                      const int32_t* start, const int32_t* end, int32_t breadth_n, orc_stats* out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > g) n_threads = (int)(g > 0 ? g : 1);
  stat_job* jobs = (stat_job*)calloc((size_t)n_threads, sizeof(stat_job));
  pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
  for (int t = 0; t < n_threads; ++t) {
    stat_job j = {depth, off, len, g * t / n_threads, g * (t + 1) / n_threads, tid, start, end, breadth_n, out};
    jobs[t] = j;
    pthread_create(&th[t], NULL, stat_run, &jobs[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(jobs); free(th);
}

/* ---- read-statistics scan: ByFlag + IsizeHist (scan.pyx:267-271, 406-420, 590-610) ---- */
int32_t orc_isize_hist(int64_t n, const uint16_t* flag, const int32_t* isize, int32_t n_group_flags,
                       const uint16_t* group_flags, int32_t n_bins, uint32_t* hist, uint64_t* group_cnt) {
  int32_t mx = 0;
  for (int64_t i = 0; i < n; ++i) {
    int g = 0;
    for (int k = 0; k < n_group_flags; ++k) { g <<= 1; if (flag[i] & group_flags[k]) g += 1; }
    int32_t v = (flag[i] & 0x2) ? isize[i] : 0;
    if (v < 0) v = -v;
    if (v > mx) mx = v;
    group_cnt[g] += 1;
    if (v >= 0 && v < n_bins) hist[(int64_t)g * n_bins + v] += 1;
  }
  return mx;
}
