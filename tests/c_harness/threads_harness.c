/* Concurrent callers of the drop-in boundary (include/laser_b200.h) from plain C.
 *
 * SURVEY.md 8(b) "Threading": the reference is called from one thread and forks its own OpenMP team
 * (gemm.nim:160); it has no global mutable state besides the cpuinfo initialisation (cpuinfo.nim:358-360), so
 * calls from different threads on distinct outputs are independent.  The replacement keeps that contract with a
 * per-device context (workspace, tensor-map cache, tile-scheduler slots) shared by every caller: this harness
 * runs T threads that each issue a mix of products -- exact kernel, tensor cores (default fp32 mode, both operand
 * major-nesses, split-K shapes), fp64, int64, the host-pointer entry and the device entry on the thread's own
 * stream (cudaStreamPerThread, (void*)0x2: no CUDA header needed) -- and requires every result to be BIT-IDENTICAL
 * to what the same call produced when the main thread ran it alone.
 *
 *   threads_harness [threads = 4] [rounds = 3] [scale = 1]
 * Exit code 0 = all checks passed; `--link-only` returns before touching the GPU. */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "laser_b200.h"

#define STREAM_PER_THREAD ((void *)0x2)

typedef struct {
  int kind; /* 0 f32 host entry, 1 f32 device entry, 2 f64 device entry, 3 i64 host entry */
  int64_t M, N, K;
  int a_transposed;
  float alpha, beta;
} job_t;

/* MNK > 128^3 takes the tensor cores (gemm.nim:140-141 is the reference's own switch) */
static const job_t kJobs[] = {
    {0, 96, 80, 64, 0, 1.0f, 0.0f},     /* exact kernel, host pointers */
    {1, 160, 192, 96, 0, 1.0f, 0.0f},   /* tensor cores, device pointers */
    {1, 264, 136, 200, 1, 0.5f, -1.25f},/* tensor cores, A given transposed, beta != 0 */
    {0, 300, 260, 72, 0, 1.0f, 0.0f},   /* tensor cores through the pipelined / staged host entry */
    {2, 70, 66, 130, 0, 1.0f, 0.0f},    /* fp64 */
    {3, 33, 17, 29, 0, 1.0f, 0.0f},     /* int64 (wrapping arithmetic) */
    {1, 128, 128, 1024, 0, 1.0f, 0.0f}, /* one tile, long K: the few-tiles split-K plan */
};
#define NJOBS ((int)(sizeof kJobs / sizeof kJobs[0]))

static uint64_t splitmix(uint64_t *s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static size_t elem_size(int kind) { return kind == 2 || kind == 3 ? 8 : 4; }

static void fill(void *p, size_t n, int kind, uint64_t seed) {
  for (size_t i = 0; i < n; ++i) {
    const double u = (double)(splitmix(&seed) >> 11) / 9007199254740992.0; /* [0, 1) */
    if (kind == 2) ((double *)p)[i] = u - 0.5;
    else if (kind == 3) ((int64_t *)p)[i] = (int64_t)(u * 2001) - 1000;
    else ((float *)p)[i] = (float)(u - 0.5);
  }
}

/* runs job j on the inputs derived from (j, salt); returns a malloc'd copy of C on the host, NULL on error */
static void *run_job(const job_t *jb, int j, uint64_t salt, int scale) {
  const int64_t M = jb->M * scale, N = jb->N * scale, K = jb->K * scale;
  const size_t es = elem_size(jb->kind), na = (size_t)(M * K), nb = (size_t)(K * N), nc = (size_t)(M * N);
  void *A = malloc(na * es), *B = malloc(nb * es), *C = malloc(nc * es);
  void *dA = NULL, *dB = NULL, *dC = NULL;
  int rc = 0;
  if (!A || !B || !C) { free(A); free(B); free(C); return NULL; }
  fill(A, na, jb->kind, 1000 * (uint64_t)j + salt);
  fill(B, nb, jb->kind, 1000 * (uint64_t)j + salt + 7);
  fill(C, nc, jb->kind, 1000 * (uint64_t)j + salt + 13);
  /* A is stored K x M when "transposed": rowStrideA = 1, colStrideA = M (BASELINE config 3's layout) */
  const int64_t rsA = jb->a_transposed ? 1 : K, csA = jb->a_transposed ? M : 1;
  switch (jb->kind) {
    case 0:
      rc = laser_b200_gemm_strided_f32(M, N, K, jb->alpha, A, rsA, csA, B, N, 1, jb->beta, C, N, 1);
      break;
    case 3:
      rc = laser_b200_gemm_strided_i64(M, N, K, 1, A, rsA, csA, B, N, 1, 0, C, N, 1);
      break;
    default:
      rc = laser_b200_malloc(&dA, na * es) || laser_b200_malloc(&dB, nb * es) || laser_b200_malloc(&dC, nc * es) ||
           laser_b200_memcpy_h2d(dA, A, na * es) || laser_b200_memcpy_h2d(dB, B, nb * es) || laser_b200_memcpy_h2d(dC, C, nc * es);
      if (!rc && jb->kind == 1)
        rc = laser_b200_gemm_strided_f32_dev(M, N, K, jb->alpha, dA, rsA, csA, dB, N, 1, jb->beta, dC, N, 1, LASER_B200_PATH_AUTO,
                                             STREAM_PER_THREAD);
      if (!rc && jb->kind == 2)
        rc = laser_b200_gemm_strided_f64_dev(M, N, K, 1.0, dA, rsA, csA, dB, N, 1, 0.0, dC, N, 1, STREAM_PER_THREAD);
      /* a blocking copy on the legacy stream waits for this thread's stream too (it is not a non-blocking stream) */
      if (!rc) rc = laser_b200_memcpy_d2h(C, dC, nc * es);
      laser_b200_free(dA); laser_b200_free(dB); laser_b200_free(dC);
  }
  free(A); free(B);
  if (rc) {
    fprintf(stderr, "job %d failed (%d): %s\n", j, rc, laser_b200_last_error());
    free(C);
    return NULL;
  }
  return C;
}

typedef struct {
  int tid, rounds, scale, failures;
  void **want; /* [NJOBS] results of the serial run */
} worker_t;

static void *worker(void *arg) {
  worker_t *w = arg;
  for (int r = 0; r < w->rounds; ++r)
    for (int q = 0; q < NJOBS; ++q) {
      const int j = (q + w->tid) % NJOBS; /* different threads are in different kernels at any moment */
      const job_t *jb = &kJobs[j];
      void *got = run_job(jb, j, 42, w->scale);
      const size_t bytes = (size_t)(jb->M * jb->N) * (size_t)w->scale * (size_t)w->scale * elem_size(jb->kind);
      if (!got || memcmp(got, w->want[j], bytes) != 0) {
        fprintf(stderr, "thread %d round %d job %d: %s\n", w->tid, r, j, got ? "result differs from the serial run" : "call failed");
        w->failures++;
      }
      free(got);
    }
  return NULL;
}

int main(int argc, char **argv) {
  if (argc > 1 && strcmp(argv[1], "--link-only") == 0) {
    printf("linked against liblaser_b200 version %d\n", laser_b200_version());
    return 0;
  }
  const int threads = argc > 1 ? atoi(argv[1]) : 4, rounds = argc > 2 ? atoi(argv[2]) : 3, scale = argc > 3 ? atoi(argv[3]) : 1;
  if (threads < 1 || threads > 64 || rounds < 1 || scale < 1 || scale > 16) { fprintf(stderr, "bad arguments\n"); return 64; }
  if (laser_b200_init() != LASER_B200_OK) { fprintf(stderr, "init: %s\n", laser_b200_last_error()); return 1; }
  void *want[NJOBS];
  for (int j = 0; j < NJOBS; ++j) {
    want[j] = run_job(&kJobs[j], j, 42, scale);
    if (!want[j]) return 1;
    /* the serial run itself must be repeatable, otherwise the comparison below proves nothing */
    void *again = run_job(&kJobs[j], j, 42, scale);
    const size_t bytes = (size_t)(kJobs[j].M * kJobs[j].N) * (size_t)scale * (size_t)scale * elem_size(kJobs[j].kind);
    if (!again || memcmp(again, want[j], bytes) != 0) { fprintf(stderr, "job %d is not repeatable on one thread\n", j); return 2; }
    free(again);
  }
  pthread_t th[64];
  worker_t w[64];
  for (int t = 0; t < threads; ++t) {
    w[t].tid = t; w[t].rounds = rounds; w[t].scale = scale; w[t].failures = 0; w[t].want = want;
    if (pthread_create(&th[t], NULL, worker, &w[t]) != 0) { fprintf(stderr, "pthread_create failed\n"); return 1; }
  }
  int failures = 0;
  for (int t = 0; t < threads; ++t) {
    pthread_join(th[t], NULL);
    failures += w[t].failures;
  }
  for (int j = 0; j < NJOBS; ++j) free(want[j]);
  printf("%d threads x %d rounds x %d jobs (scale %d): %d mismatches, %lld kernel launches\n", threads, rounds, NJOBS, scale, failures,
         (long long)laser_b200_launch_count());
  laser_b200_shutdown();
  return failures ? 3 : 0;
}
