/*
 * Per-call benchmark of the OAI LDPC loader ABI, the way unmodified OAI host code drives it: the library is dlopen()ed with the loader's
 * flags and the four symbols are looked up by name (nrLDPC_load.c:46-71, load_module_shlib.c:160), then T host threads each make one
 * BLOCKING LDPCdecoder call per code block from their own aligned buffers, decode_abort_t reset before every call (ldpctest.c:329-340; the
 * tpool workers of nr_ulsch_decoding.c:435-468 do the same concurrently).  The same binary times libldpc_b200.so and the compiled reference
 * (oracle/_ref/libref_ldpc_dec.so) -- same harness, same inputs, same threads -- and checks every call's result against an expected file.
 *
 *   abi_bench <lib.so> <llr.bin> <n_blocks> <threads> <seconds> <BG> <Z> <R> <maxIter> [expected.bin]
 *     llr.bin       n_blocks rows of 27008 int8
 *     expected.bin  n_blocks rows of (int32 iterations, K/8 output bytes) -- K = 22Z (BG1) / 10Z (BG2); optional
 *   abi_bench enc <lib.so> <threads> <seconds> <BG> <Z> <n_segments>
 *     times LDPCencoder calls of n_segments segments each (ldpctest.c:269-284, nr_dlsch_coding.c:389-395) from `threads` host threads
 *   prints one JSON line.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/nrb200_ldpc.h"

typedef int32_t (*init_fn_t)(void);
typedef int32_t (*dec_fn_t)(nrb200_ldpc_dec_params_t *, uint8_t, uint8_t, uint8_t, int8_t *, int8_t *, nrb200_ldpc_time_stats_t *, nrb200_decode_abort_t *);

#define ROW 27008

typedef struct {
  dec_fn_t dec; const int8_t *llr; const uint8_t *expect; int n, kbytes, BG, Z, R, maxIter, tid, nthreads; double seconds;
  long count, iters, bad; float *lat; long lat_cap;
} job_t;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static pthread_barrier_t g_start;

static void *worker(void *arg)
{
  job_t *j = (job_t *)arg;
  nrb200_ldpc_dec_params_t p;
  memset(&p, 0, sizeof(p));
  p.BG = (uint8_t)j->BG; p.Z = (uint16_t)j->Z; p.R = (uint8_t)j->R; p.numMaxIter = (uint8_t)j->maxIter; p.outMode = NRB200_OUTMODE_BIT;
  p.E = 8 * j->kbytes;
  nrb200_decode_abort_t ab;
  pthread_mutex_init(&ab.mutex_failure, NULL);
  int8_t *in = aligned_alloc(64, ROW), *out = aligned_alloc(64, ROW);
  memset(out, 0, ROW);
  pthread_barrier_wait(&g_start);
  const double t_end = now_s() + j->seconds;
  int i = j->tid % j->n;
  double t = now_s();
  while (t < t_end) {
    memcpy(in, j->llr + (size_t)i * ROW, 27000);
    ab.failed = false;
    const int it = j->dec(&p, 0, 0, 0, in, out, NULL, &ab);
    const double t2 = now_s();
    if (j->count < j->lat_cap) j->lat[j->count] = (float)(1e6 * (t2 - t));
    t = t2;
    if (j->expect) {
      const uint8_t *e = j->expect + (size_t)i * (4 + j->kbytes);
      int32_t eit; memcpy(&eit, e, 4);
      if (it != eit || memcmp(out, e + 4, (size_t)j->kbytes) != 0) j->bad++;
    }
    j->iters += it;
    j->count++;
    i = (i + j->nthreads) % j->n;
  }
  free(in); free(out);
  return NULL;
}

typedef int32_t (*enc_fn_t)(uint8_t **, uint8_t **, nrb200_ldpc_enc_params_t *);
typedef struct { enc_fn_t enc; int BG, Z, nseg, tid; double seconds; long count; unsigned check; } ejob_t;
static void *eworker(void *arg)
{
  ejob_t *j = (ejob_t *)arg;
  const int K = (j->BG == 1 ? 22 : 10) * j->Z, nout = (j->BG == 1 ? 66 : 50) * j->Z;
  uint8_t *in[8], *out[8];
  for (int s = 0; s < j->nseg; s++) {
    in[s] = aligned_alloc(64, 1152); out[s] = aligned_alloc(64, 68 * 384);
    for (int i = 0; i < 1152; i++) in[s][i] = (uint8_t)(rand_r(&j->check) >> 3);
  }
  nrb200_ldpc_enc_params_t p; memset(&p, 0, sizeof(p));
  p.n_segments = (unsigned)j->nseg; p.macro_num = 0; p.Kb = j->BG == 1 ? 22 : 10; p.Zc = (unsigned)j->Z; p.BG = (uint8_t)j->BG; p.K = (unsigned)K; p.Kr = (unsigned)K;
  pthread_barrier_wait(&g_start);
  const double t_end = now_s() + j->seconds;
  unsigned acc = 0;
  while (now_s() < t_end) {
    in[0][0]++;
    j->enc(in, out, &p);
    acc += out[0][nout - 1] + out[j->nseg - 1][7];
    j->count++;
  }
  j->check = acc;
  return NULL;
}

static int enc_main(int argc, char **argv)
{
  if (argc < 8) return 2;
  const char *so = argv[2];
  const int threads = atoi(argv[3]);
  const double seconds = atof(argv[4]);
  const int BG = atoi(argv[5]), Z = atoi(argv[6]), nseg = atoi(argv[7]);
  void *h = dlopen(so, RTLD_LAZY | RTLD_NODELETE | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  init_fn_t init = (init_fn_t)dlsym(h, "LDPCinit");
  enc_fn_t enc = (enc_fn_t)dlsym(h, "LDPCencoder");
  if (!enc || nseg < 1 || nseg > 8) { fprintf(stderr, "missing loader symbols\n"); return 1; }
  dlclose(h);
  if (init && init() != 0) return 1;             /* oracle/_ref/libref_ldpc_enc.so carries the encoder only */
  pthread_t *th = calloc((size_t)threads, sizeof(*th));
  ejob_t *jobs = calloc((size_t)threads, sizeof(*jobs));
  { ejob_t w = {enc, BG, Z, nseg, 0, 0.05, 0, 1}; pthread_barrier_init(&g_start, NULL, 1); eworker(&w); }     /* warm-up */
  pthread_barrier_init(&g_start, NULL, (unsigned)threads + 1);
  for (int t = 0; t < threads; t++) { jobs[t] = (ejob_t){enc, BG, Z, nseg, t, seconds, 0, (unsigned)t + 7}; pthread_create(&th[t], NULL, eworker, &jobs[t]); }
  pthread_barrier_wait(&g_start);
  const double t0 = now_s();
  long total = 0;
  for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); total += jobs[t].count; }
  const double el = now_s() - t0;
  printf("{\"lib\": \"%s\", \"what\": \"LDPCencoder\", \"host_threads\": %d, \"segments_per_call\": %d, \"calls\": %ld, \"value\": %.1f, \"unit\": \"CB/s\", "
         "\"us_per_call_mean\": %.2f}\n", strrchr(so, '/') ? strrchr(so, '/') + 1 : so, threads, nseg, total, total * nseg / el, total ? 1e6 * el * threads / total : 0.0);
  return 0;
}

static int cmpf(const void *a, const void *b) { const float x = *(const float *)a, y = *(const float *)b; return x < y ? -1 : x > y; }

int main(int argc, char **argv)
{
  if (argc > 1 && strcmp(argv[1], "enc") == 0) return enc_main(argc, argv);
  if (argc < 10) { fprintf(stderr, "usage: %s lib.so llr.bin n threads seconds BG Z R maxIter [expected.bin]\n", argv[0]); return 2; }
  const char *so = argv[1];
  const int n = atoi(argv[3]), threads = atoi(argv[4]);
  const double seconds = atof(argv[5]);
  const int BG = atoi(argv[6]), Z = atoi(argv[7]), R = atoi(argv[8]), maxIter = atoi(argv[9]);
  const int kbytes = (BG == 1 ? 22 : 10) * Z / 8;
  void *h = dlopen(so, RTLD_LAZY | RTLD_NODELETE | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  init_fn_t init = (init_fn_t)dlsym(h, "LDPCinit"), shut = (init_fn_t)dlsym(h, "LDPCshutdown");
  dec_fn_t dec = (dec_fn_t)dlsym(h, "LDPCdecoder");
  if (!init || !shut || !dec) { fprintf(stderr, "missing loader symbols\n"); return 1; }
  void (*stats)(uint64_t *, uint64_t *) = (void (*)(uint64_t *, uint64_t *))dlsym(h, "nrb200_ll_stats");   /* this repo's library only */
  void (*timing)(uint64_t *) = (void (*)(uint64_t *))dlsym(h, "nrb200_ll_timing");
  dlclose(h);                                  /* the loader closes the handle after the lookup; RTLD_NODELETE keeps the library resident */
  if (init() != 0) { fprintf(stderr, "LDPCinit failed\n"); return 1; }
  int8_t *llr = aligned_alloc(64, (size_t)n * ROW);
  FILE *f = fopen(argv[2], "rb");
  if (!f || fread(llr, ROW, (size_t)n, f) != (size_t)n) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
  fclose(f);
  uint8_t *expect = NULL;
  if (argc > 10) {
    expect = malloc((size_t)n * (4 + kbytes));
    f = fopen(argv[10], "rb");
    if (!f || fread(expect, (size_t)(4 + kbytes), (size_t)n, f) != (size_t)n) { fprintf(stderr, "cannot read %s\n", argv[10]); return 1; }
    fclose(f);
  }
  pthread_t *th = calloc((size_t)threads, sizeof(*th));
  job_t *jobs = calloc((size_t)threads, sizeof(*jobs));
  pthread_barrier_init(&g_start, NULL, (unsigned)threads + 1);
  /* warm-up: graph tables, kernels, staging rows */
  {
    nrb200_ldpc_dec_params_t p; memset(&p, 0, sizeof(p));
    p.BG = (uint8_t)BG; p.Z = (uint16_t)Z; p.R = (uint8_t)R; p.numMaxIter = (uint8_t)maxIter; p.outMode = NRB200_OUTMODE_BIT; p.E = 8 * kbytes;
    nrb200_decode_abort_t ab; pthread_mutex_init(&ab.mutex_failure, NULL);
    int8_t *in = aligned_alloc(64, ROW), *out = aligned_alloc(64, ROW);
    for (int w = 0; w < 5; w++) { memcpy(in, llr, 27000); ab.failed = false; dec(&p, 0, 0, 0, in, out, NULL, &ab); }
    free(in); free(out);
  }
  const long lat_cap = 200000;
  for (int t = 0; t < threads; t++) {
    jobs[t] = (job_t){dec, llr, expect, n, kbytes, BG, Z, R, maxIter, t, threads, seconds, 0, 0, 0, malloc(sizeof(float) * (size_t)lat_cap), lat_cap};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  pthread_barrier_wait(&g_start);
  const double t0 = now_s();
  long total = 0, its = 0, bad = 0;
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  const double el = now_s() - t0;
  long nl = 0;
  for (int t = 0; t < threads; t++) { total += jobs[t].count; its += jobs[t].iters; bad += jobs[t].bad; nl += jobs[t].count < lat_cap ? jobs[t].count : lat_cap; }
  float *all = malloc(sizeof(float) * (size_t)(nl > 0 ? nl : 1));
  long k = 0;
  for (int t = 0; t < threads; t++) { const long c = jobs[t].count < lat_cap ? jobs[t].count : lat_cap; memcpy(all + k, jobs[t].lat, sizeof(float) * (size_t)c); k += c; }
  qsort(all, (size_t)nl, sizeof(float), cmpf);
  uint64_t launches = 0, blocks = 0;
  if (stats) stats(&launches, &blocks);
  uint64_t tm[5] = {0, 0, 0, 0, 0};
  if (timing) timing(tm);
  const double calls_all = blocks ? (double)blocks : 1.0;
  printf("{\"lib\": \"%s\", \"host_threads\": %d, \"calls\": %ld, \"seconds\": %.3f, \"value\": %.1f, \"unit\": \"CB/s\", \"us_per_call_mean\": %.2f, "
         "\"us_per_call_p50\": %.2f, \"us_per_call_p99\": %.2f, \"mean_returned_iterations\": %.3f, \"mismatches\": %ld, \"checked\": %s, "
         "\"blocks_per_launch\": %.2f, \"us_stage\": %.2f, \"us_queue_launch\": %.2f, \"us_wait_kernel\": %.2f, \"us_copy_out\": %.2f, \"us_device\": %.2f}\n",
         strrchr(so, '/') ? strrchr(so, '/') + 1 : so, threads, total, el, total / el, total ? 1e6 * el * threads / total : 0.0,
         nl ? all[nl / 2] : 0.0, nl ? all[(long)(nl * 0.99)] : 0.0, total ? (double)its / total : 0.0, bad, expect ? "true" : "false",
         launches ? (double)blocks / (double)launches : 0.0, 1e-3 * tm[0] / calls_all, 1e-3 * tm[1] / calls_all, 1e-3 * tm[2] / calls_all,
         1e-3 * tm[3] / calls_all, 1e-3 * tm[4] / calls_all);
  shut();
  return bad ? 3 : 0;
}
