/* Plain-C caller of the multi-GPU entry of the drop-in boundary (include/laser_b200.h):
 *   laser_b200_gemm_rowsharded_f32(ngpus, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC)
 * -- the reference signature (gemm.nim:184-193) plus a device count: what a Nim host that keeps its matrices in host memory
 * would call to spread the `ic` row blocks of gemm.nim:160-176 over the GPUs of the box.  Also exercises the explicit
 * communicator API the per-rank device entry uses.  usage: rowshard_harness <ngpus>;  exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "laser_b200.h"

static int check(int rc, const char *what) {
  if (rc != LASER_B200_OK) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, laser_b200_last_error());
    return 1;
  }
  return 0;
}

int main(int argc, char **argv) {
  const int ngpus = argc > 1 ? atoi(argv[1]) : 2;
  const int M = 3000, N = 520, K = 900;   /* 3000 rows: uneven panels (256-row granularity), the last rank shorter */
  float *A = malloc(sizeof(float) * M * K), *B = malloc(sizeof(float) * K * N), *C = malloc(sizeof(float) * M * N);
  float *C0 = malloc(sizeof(float) * M * N);
  srand(7);
  for (int i = 0; i < M * K; ++i) A[i] = (float)rand() / RAND_MAX - 0.5f;
  for (int i = 0; i < K * N; ++i) B[i] = (float)rand() / RAND_MAX - 0.5f;
  for (int i = 0; i < M * N; ++i) C0[i] = C[i] = (float)rand() / RAND_MAX;
  int64_t first, rows, covered = 0;
  for (int r = 0; r < ngpus; ++r) {
    laser_b200_rowshard_partition(M, ngpus, r, &first, &rows);
    if (first != covered) { fprintf(stderr, "partition not contiguous at rank %d\n", r); return 2; }
    covered += rows;
  }
  if (covered != M) { fprintf(stderr, "partition covers %lld of %d rows\n", (long long)covered, M); return 2; }
  if (check(laser_b200_gemm_rowsharded_f32(ngpus, M, N, K, 0.5f, A, K, 1, B, N, 1, -1.25f, C, N, 1), "gemm_rowsharded_f32")) return 1;
  double worst = 0, scale = 0;
  for (int i = 0; i < M; i += 41)
    for (int j = 0; j < N; j += 29) {
      double s = 0, sa = 0;
      for (int k = 0; k < K; ++k) { s += (double)A[i * K + k] * (double)B[k * N + j]; sa += fabs((double)A[i * K + k] * (double)B[k * N + j]); }
      const double want = 0.5 * s - 1.25 * C0[i * N + j];
      const double e = fabs(C[i * N + j] - want) / (0.5 * sa + 1.25 * fabs(C0[i * N + j]));
      if (e > worst) worst = e;
      if (sa > scale) scale = sa;
    }
  printf("%d GPUs, %d x %d x %d: max error relative to sum|a||b| on sampled entries %.3e\n", ngpus, M, N, K, worst);
  /* the explicit communicator API (what the per-rank device entry takes) */
  laser_b200_comm *comms[16] = {0};
  if (ngpus <= 16) {
    if (check(laser_b200_comm_init_all(comms, ngpus), "comm_init_all")) return 1;
    for (int r = 0; r < ngpus; ++r) {
      if (laser_b200_comm_rank(comms[r]) != r || laser_b200_comm_size(comms[r]) != ngpus) { fprintf(stderr, "rank/size mismatch\n"); return 2; }
      if (check(laser_b200_comm_destroy(comms[r]), "comm_destroy")) return 1;
    }
  }
  free(A); free(B); free(C); free(C0);
  laser_b200_shutdown();
  return worst < 1e-5 ? 0 : 3;
}
