/* Plain-C caller of the drop-in boundary (include/laser_b200.h), the stand-in for the Nim
 * call site that cannot be compiled in this image: same calls, same argument order as
 *   gemm_strided(M, N, K, 1, a, K, 1, b, N, 1, 0, c, N, 1)
 * in the reference's self-tests (gemm.nim:311-334) and bench (gemm_bench_float32.nim:184-189).
 * Exit code 0 = all checks passed.  `--link-only` returns before touching the GPU. */
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
  if (argc > 1 && strcmp(argv[1], "--link-only") == 0) {
    printf("linked against liblaser_b200 version %d\n", laser_b200_version());
    return 0;
  }
  /* known answer: gemm.nim:311-334 */
  const float a[6] = {1, 2, 3, 4, 5, 6}, b[6] = {7, 8, 9, 10, 11, 12};
  float c[4] = {-1, -1, -1, -1};
  if (check(laser_b200_gemm_strided_f32(2, 2, 3, 1.0f, a, 3, 1, b, 2, 1, 0.0f, c, 2, 1), "gemm_strided_f32")) return 1;
  if (c[0] != 58 || c[1] != 64 || c[2] != 139 || c[3] != 154) { fprintf(stderr, "known answer mismatch\n"); return 2; }
  /* int64 flavour of the same vector (the reference tests `int`) */
  const int64_t ai[6] = {-2, -3, -1, 3, 0, 4}, bi[12] = {1, 5, 2, -1, -3, 0, 3, 4, 6, -2, 7, -4};
  int64_t ci[8];
  const int64_t want[8] = {1, -8, -20, -6, 27, 7, 34, -19}; /* gemm.nim:336-360 */
  if (check(laser_b200_gemm_strided_i64(2, 4, 3, 1, ai, 3, 1, bi, 4, 1, 0, ci, 4, 1), "gemm_strided_i64")) return 1;
  for (int i = 0; i < 8; ++i) if (ci[i] != want[i]) { fprintf(stderr, "int64 mismatch at %d\n", i); return 2; }
  /* a larger strided product through the tensor-core path: A given transposed */
  const int M = 700, N = 520, K = 900;
  float *At = malloc(sizeof(float) * K * M), *B = malloc(sizeof(float) * K * N), *C = malloc(sizeof(float) * M * N);
  srand(42);
  for (int i = 0; i < K * M; ++i) At[i] = (float)rand() / RAND_MAX;
  for (int i = 0; i < K * N; ++i) B[i] = (float)rand() / RAND_MAX;
  for (int i = 0; i < M * N; ++i) C[i] = NAN; /* beta == 0 must not read it */
  if (check(laser_b200_gemm_strided_f32(M, N, K, 1.0f, At, 1, M, B, N, 1, 0.0f, C, N, 1), "strided gemm")) return 1;
  double worst = 0;
  for (int i = 0; i < M; i += 37)
    for (int j = 0; j < N; j += 29) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)At[k * M + i] * (double)B[k * N + j];
      const double e = fabs(C[i * N + j] - s) / s;
      if (e > worst) worst = e;
    }
  printf("path %d, max relative error on sampled entries %.3e, %lld kernel launches\n", laser_b200_last_path(), worst,
         (long long)laser_b200_launch_count());
  free(At); free(B); free(C);
  laser_b200_shutdown();
  return worst < 1e-4 ? 0 : 3;
}
