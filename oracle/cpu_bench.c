/*
 * TEST / BENCH INFRASTRUCTURE ONLY -- multi-threaded driver that times a CPU LDPC decoder on the host cores the way
 * ldpctest.c:329-340 calls it (one blocking LDPCdecoder call per code block, decode_abort_t reset before each call,
 * re-entrant decoder, one pthread per core).  `fn` is the reference's LDPCdecoder (oracle/_ref/libref_ldpc_dec.so) or
 * NULL for the scalar oracle port.  Used only by bench.py's cpu_baseline / --impl reference leg.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "nrb200_oracle.h"
#include "../include/nrb200_ldpc.h"

typedef int32_t (*dec_fn_t)(nrb200_ldpc_dec_params_t *, uint8_t, uint8_t, uint8_t, int8_t *, int8_t *, void *, nrb200_decode_abort_t *);

typedef struct {
  dec_fn_t fn; const int8_t *llr; int n, stride, BG, Z, R, maxIter, tid, nthreads; double seconds; long count; long iters;
} job_t;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void *worker(void *arg)
{
  job_t *j = (job_t *)arg;
  nrb200_ldpc_dec_params_t p;
  memset(&p, 0, sizeof(p));
  p.BG = (uint8_t)j->BG; p.Z = (uint16_t)j->Z; p.R = (uint8_t)j->R; p.numMaxIter = (uint8_t)j->maxIter; p.outMode = NRB200_OUTMODE_BIT;
  p.E = (j->BG == 1 ? 22 : 10) * j->Z;
  nrb200_decode_abort_t ab;
  pthread_mutex_init(&ab.mutex_failure, NULL);
  int8_t *in = aligned_alloc(64, 27008), *out = aligned_alloc(64, 27008);
  const double t_end = now_s() + j->seconds;
  int i = j->tid;
  while (now_s() < t_end) {
    memcpy(in, j->llr + (size_t)(i % j->n) * j->stride, 27000);   /* callers hand the decoder their own aligned buffer */
    ab.failed = false;                                            /* set_abort(&dec_abort, false), ldpctest.c:330 */
    int it = j->fn ? j->fn(&p, 0, 0, 0, in, out, NULL, &ab)
                   : orc_ldpc_decode(j->BG, j->Z, j->R, j->maxIter, 0, in, out, 0, 0, 0, 0);
    j->iters += it;
    j->count++;
    i += j->nthreads;
  }
  free(in); free(out);
  return NULL;
}

/* returns total decodes; *elapsed = wall seconds; *mean_iters = mean returned iteration count */
long orc_bench_ldpc_decoder(void *fn, const int8_t *llr, int n, int stride, int BG, int Z, int R, int maxIter, int threads, double seconds,
                            double *elapsed, double *mean_iters)
{
  pthread_t *th = calloc((size_t)threads, sizeof(*th));
  job_t *jobs = calloc((size_t)threads, sizeof(*jobs));
  const double t0 = now_s();
  for (int t = 0; t < threads; t++) {
    jobs[t] = (job_t){(dec_fn_t)fn, llr, n, stride, BG, Z, R, maxIter, t, threads, seconds, 0, 0};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  long total = 0, its = 0;
  for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); total += jobs[t].count; its += jobs[t].iters; }
  *elapsed = now_s() - t0;
  *mean_iters = total ? (double)its / (double)total : 0.0;
  free(th); free(jobs);
  return total;
}

/* ---- same idea for the DFT library: `fn` is the reference's per-size entry point (e.g. idft4096 of oracle/_ref/libref_dfts.so) */
typedef void (*dft_fn_t)(int16_t *, int16_t *, unsigned char);
typedef struct { dft_fn_t fn; const int16_t *x; int N, n, tid, nthreads; double seconds; long count; } djob_t;
static void *dworker(void *arg)
{
  djob_t *j = (djob_t *)arg;
  int16_t *in = aligned_alloc(64, (size_t)j->N * 4 + 64), *out = aligned_alloc(64, (size_t)j->N * 4 + 64);
  const double t_end = now_s() + j->seconds;
  int i = j->tid;
  while (now_s() < t_end) {
    for (int r = 0; r < 16; r++) {
      memcpy(in, j->x + (size_t)(i % j->n) * j->N * 2, (size_t)j->N * 4);
      j->fn(in, out, 1);
      j->count++;
      i += j->nthreads;
    }
  }
  free(in); free(out);
  return NULL;
}
long orc_bench_dft(void *fn, const int16_t *x, int N, int n, int threads, double seconds, double *elapsed)
{
  pthread_t *th = calloc((size_t)threads, sizeof(*th));
  djob_t *jobs = calloc((size_t)threads, sizeof(*jobs));
  const double t0 = now_s();
  for (int t = 0; t < threads; t++) {
    jobs[t] = (djob_t){(dft_fn_t)fn, x, N, n, t, threads, seconds, 0};
    pthread_create(&th[t], NULL, dworker, &jobs[t]);
  }
  long total = 0;
  for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); total += jobs[t].count; }
  *elapsed = now_s() - t0;
  free(th); free(jobs);
  return total;
}

/* ---- decode EVERY block of a batch once on `threads` host threads (large-sample differential tests: tests/test_gpu_parity_scale.py).
 * `fn` = the reference's LDPCdecoder (or NULL for the scalar port), `crc_fn` = the reference's check_crc (crc_byte.c:314) or NULL for the
 * parity-check stop ldpctest uses.  Outputs land in out[n][out_stride] / iters[n]; returns n. */
typedef struct { dec_fn_t fn; void *crc_fn; const int8_t *llr; int n, stride, BG, Z, R, maxIter, crc_type, crc_len, tid, nthreads; uint8_t *out; int out_stride;
                 int32_t *iters; } ajob_t;
static void *aworker(void *arg)
{
  ajob_t *j = (ajob_t *)arg;
  nrb200_ldpc_dec_params_t p;
  memset(&p, 0, sizeof(p));
  p.BG = (uint8_t)j->BG; p.Z = (uint16_t)j->Z; p.R = (uint8_t)j->R; p.numMaxIter = (uint8_t)j->maxIter; p.outMode = NRB200_OUTMODE_BIT;
  p.E = j->crc_fn ? j->crc_len : (j->BG == 1 ? 22 : 10) * j->Z;
  p.crc_type = (uint8_t)j->crc_type;
  p.check_crc = (int (*)(uint8_t *, uint32_t, uint8_t))j->crc_fn;
  nrb200_decode_abort_t ab;
  pthread_mutex_init(&ab.mutex_failure, NULL);
  int8_t *in = aligned_alloc(64, 27008), *out = aligned_alloc(64, 27008);
  const int nb = j->out_stride < 27000 ? j->out_stride : 27000;
  for (int i = j->tid; i < j->n; i += j->nthreads) {
    memset(in, 0, 27008);
    memcpy(in, j->llr + (size_t)i * j->stride, (size_t)(j->stride < 27000 ? j->stride : 27000));
    memset(out, 0, 27008);
    ab.failed = false;
    j->iters[i] = j->fn ? j->fn(&p, 0, 0, 0, in, out, NULL, &ab)
                        : orc_ldpc_decode(j->BG, j->Z, j->R, j->maxIter, 0, in, out, j->crc_fn != NULL, (uint32_t)j->crc_len, j->crc_type, 0);
    memcpy(j->out + (size_t)i * j->out_stride, out, (size_t)nb);
  }
  free(in); free(out);
  return NULL;
}
long orc_decode_all(void *fn, void *crc_fn, const int8_t *llr, int n, int stride, int BG, int Z, int R, int maxIter, int crc_type, int crc_len,
                    uint8_t *out, int out_stride, int32_t *iters, int threads)
{
  pthread_t *th = calloc((size_t)threads, sizeof(*th));
  ajob_t *jobs = calloc((size_t)threads, sizeof(*jobs));
  for (int t = 0; t < threads; t++) {
    jobs[t] = (ajob_t){(dec_fn_t)fn, crc_fn, llr, n, stride, BG, Z, R, maxIter, crc_type, crc_len, t, threads, out, out_stride, iters};
    pthread_create(&th[t], NULL, aworker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return n;
}
