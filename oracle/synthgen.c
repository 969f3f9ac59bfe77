/* oracle/synthgen.c -- the counter-based synthetic read generator (include/mcov_synth.h: plain C, shared by the
 * product's device generator) compiled for the host into liboracle.so, so that bench.py's reference arm and the
 * cpu_baseline leg can build BASELINE's workloads without loading the product library.  Test / bench
 * infrastructure only.  (The reference has no generator for mapped reads: metacov/simulate.py:9-50 wraps the
 * external art_illumina.) */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mcov_synth.h"

typedef struct {
  const mcov_synth_params* P;
  int64_t i0, a, b;
  const int64_t* read_start;
  const int32_t* contig_len;
  int32_t n_contigs, tid_base;
  const uint32_t* cig_off;
  int32_t *tid, *pos, *isize;
  uint16_t* flag;
  uint8_t* mapq;
  uint32_t* cig;
  uint32_t* ncig;
} gen_job;

static void* ncig_worker(void* p) {
  gen_job* j = (gen_job*)p;
  for (int64_t k = j->a; k < j->b; ++k) j->ncig[k] = mcov_synth_ncigar(j->P, j->i0 + k);
  return NULL;
}

static void* fill_worker(void* p) {
  gen_job* j = (gen_job*)p;
  for (int64_t k = j->a; k < j->b; ++k) {
    int32_t t, ps, is;
    uint16_t f;
    uint8_t q;
    const uint32_t o0 = j->cig_off[k], o1 = j->cig_off[k + 1];
    mcov_synth_read(j->P, j->i0 + k, j->read_start, j->contig_len, j->n_contigs, &t, &ps, &f, &q, &is, j->cig + o0, o1 - o0);
    j->tid[k] = t - j->tid_base; j->pos[k] = ps; j->flag[k] = f; j->mapq[k] = q; j->isize[k] = is;
  }
  return NULL;
}

static void run_jobs(gen_job* proto, int64_t n, int threads, void* (*fn)(void*)) {
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  if (n < (1 << 16)) threads = 1;
  pthread_t th[64];
  gen_job jobs[64];
  const int64_t per = (n + threads - 1) / threads;
  int started = 0;
  for (int t = 0; t < threads; ++t) {
    jobs[t] = *proto;
    jobs[t].a = t * per;
    jobs[t].b = jobs[t].a + per < n ? jobs[t].a + per : n;
    if (jobs[t].a >= jobs[t].b) break;
    pthread_create(&th[t], NULL, fn, &jobs[t]);
    ++started;
  }
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

void orc_synth_ncigar(const mcov_synth_params* P, int64_t i0, int64_t n, uint32_t* out, int threads) {
  gen_job j;
  memset(&j, 0, sizeof(j));
  j.P = P; j.i0 = i0; j.ncig = out;
  run_jobs(&j, n, threads, ncig_worker);
}

void orc_synth_reads(const mcov_synth_params* P, int64_t i0, int64_t n, const int64_t* read_start, const int32_t* contig_len,
                     int32_t n_contigs, int32_t tid_base, const uint32_t* cig_off, int32_t* tid, int32_t* pos, uint16_t* flag,
                     uint8_t* mapq, int32_t* isize, uint32_t* cig, int threads) {
  gen_job j;
  memset(&j, 0, sizeof(j));
  j.P = P; j.i0 = i0; j.read_start = read_start; j.contig_len = contig_len; j.n_contigs = n_contigs; j.tid_base = tid_base;
  j.cig_off = cig_off; j.tid = tid; j.pos = pos; j.flag = flag; j.mapq = mapq; j.isize = isize; j.cig = cig;
  run_jobs(&j, n, threads, fill_worker);
}
