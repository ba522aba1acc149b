/*
 * laser_oracle.c -- TEST INFRASTRUCTURE ONLY (see laser_oracle.h).
 *
 * Numerics-faithful restatement of laser's gemm_strided, error metrics and the
 * shared counter-based input generator.  Compiled with -ffp-contract=off so
 * that every fused/unfused choice below is explicit.
 *
 * Order of operations being restated (all paths relative to /root/reference):
 *   - K is cut in blocks of kc = min(2048/sizeof(T), K)   gemm_tiling.nim:309-310
 *   - inside a block each C[i,j] gets AB = sum_k a*b accumulated from 0 by a
 *     k-sequential FMA chain                              gemm_ukernel_generator.nim:196-250
 *     (fused on AVX-512/FMA3 hosts, the dispatch order of gemm.nim:229-233;
 *     integer kernels use mullo+add, i.e. wrapping arithmetic)
 *   - beta is applied on the first block only, later blocks use beta' = 1
 *                                                          gemm.nim:158
 *   - epilogue: beta'==0 -> C is overwritten without being read; beta'!=1 ->
 *     C *= beta'; then C += AB (alpha==1) or C += alpha*AB
 *                                                          gemm_ukernel_generic.nim:53-76,97-126
 *   - K == 0: the pc loop never runs, C is left untouched  gemm.nim:150
 */
#include "laser_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define KC_BYTES 2048 /* gemm_tiling.nim:310 */

/* ------------------------------------------------------------------ */
/*                    floating point instantiations                    */
/* ------------------------------------------------------------------ */
#define DEFINE_ORACLE_FLOAT(NAME, T, FMA)                                                  \
  void NAME(int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t rsA,             \
            int64_t csA, const T *B, int64_t rsB, int64_t csB, T beta, T *C, int64_t rsC,  \
            int64_t csC) {                                                                 \
    if (M <= 0 || N <= 0 || K <= 0) return; /* gemm.nim:150: no pc iteration */            \
    const int64_t KC = (int64_t)(KC_BYTES / sizeof(T)) < K ? (int64_t)(KC_BYTES / sizeof(T)) : K; \
    _Pragma("omp parallel") {                                                              \
      T *acc = (T *)malloc(sizeof(T) * (size_t)N);                                         \
      _Pragma("omp for schedule(static)") for (int64_t i = 0; i < M; ++i) {                \
        for (int64_t pc = 0; pc < K; pc += KC) {                                           \
          const int64_t kc = (K - pc) < KC ? (K - pc) : KC;                                \
          for (int64_t j = 0; j < N; ++j) acc[j] = (T)0;                                   \
          for (int64_t k = 0; k < kc; ++k) {                                               \
            const T a = A[i * rsA + (pc + k) * csA];                                       \
            const T *brow = B + (pc + k) * rsB;                                            \
            if (csB == 1) {                                                                \
              for (int64_t j = 0; j < N; ++j) acc[j] = FMA(a, brow[j], acc[j]);            \
            } else {                                                                       \
              for (int64_t j = 0; j < N; ++j) acc[j] = FMA(a, brow[j * csB], acc[j]);      \
            }                                                                              \
          }                                                                                \
          const T b1 = (pc == 0) ? beta : (T)1;                                            \
          for (int64_t j = 0; j < N; ++j) {                                                \
            T *c = C + i * rsC + j * csC;                                                  \
            T v;                                                                           \
            if (b1 == (T)0) v = (T)0;                                                      \
            else if (b1 != (T)1) v = *c * b1;                                              \
            else v = *c;                                                                   \
            if (alpha == (T)1) v = v + acc[j];                                             \
            else v = v + alpha * acc[j];                                                   \
            *c = v;                                                                        \
          }                                                                                \
        }                                                                                  \
      }                                                                                    \
      free(acc);                                                                           \
    }                                                                                      \
  }

DEFINE_ORACLE_FLOAT(oracle_gemm_strided_f32, float, fmaf)
DEFINE_ORACLE_FLOAT(oracle_gemm_strided_f64, double, fma)

/* ------------------------------------------------------------------ */
/*        integer instantiations (wrapping, as _mm*_mullo + add)       */
/* ------------------------------------------------------------------ */
#define DEFINE_ORACLE_INT(NAME, T, UT)                                                     \
  void NAME(int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t rsA,             \
            int64_t csA, const T *B, int64_t rsB, int64_t csB, T beta, T *C, int64_t rsC,  \
            int64_t csC) {                                                                 \
    if (M <= 0 || N <= 0 || K <= 0) return;                                                \
    const int64_t KC = (int64_t)(KC_BYTES / sizeof(T)) < K ? (int64_t)(KC_BYTES / sizeof(T)) : K; \
    for (int64_t i = 0; i < M; ++i)                                                        \
      for (int64_t j = 0; j < N; ++j) {                                                    \
        T *c = C + i * rsC + j * csC;                                                      \
        for (int64_t pc = 0; pc < K; pc += KC) {                                           \
          const int64_t kc = (K - pc) < KC ? (K - pc) : KC;                                \
          UT acc = 0;                                                                      \
          for (int64_t k = 0; k < kc; ++k)                                                 \
            acc += (UT)A[i * rsA + (pc + k) * csA] * (UT)B[(pc + k) * rsB + j * csB];      \
          const T b1 = (pc == 0) ? beta : (T)1;                                            \
          UT v;                                                                            \
          if (b1 == 0) v = 0;                                                              \
          else if (b1 != 1) v = (UT)*c * (UT)b1;                                           \
          else v = (UT)*c;                                                                 \
          if (alpha == 1) v += acc;                                                        \
          else v += (UT)alpha * acc;                                                       \
          *c = (T)v;                                                                       \
        }                                                                                  \
      }                                                                                    \
  }

DEFINE_ORACLE_INT(oracle_gemm_strided_i32, int32_t, uint32_t)
DEFINE_ORACLE_INT(oracle_gemm_strided_i64, int64_t, uint64_t)

/* ------------------------------------------------------------------ */
/*                               bf16                                  */
/* ------------------------------------------------------------------ */
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* NaN */
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

void oracle_gemm_strided_bf16(int64_t M, int64_t N, int64_t K, float alpha, const uint16_t *A,
                              int64_t rsA, int64_t csA, const uint16_t *B, int64_t rsB,
                              int64_t csB, float beta, uint16_t *C, int64_t rsC, int64_t csC) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  float *a32 = (float *)malloc(sizeof(float) * (size_t)(M * K));
  float *b32 = (float *)malloc(sizeof(float) * (size_t)(K * N));
  float *c32 = (float *)malloc(sizeof(float) * (size_t)(M * N));
  for (int64_t i = 0; i < M; ++i)
    for (int64_t k = 0; k < K; ++k) a32[i * K + k] = bf16_to_f32(A[i * rsA + k * csA]);
  for (int64_t k = 0; k < K; ++k)
    for (int64_t j = 0; j < N; ++j) b32[k * N + j] = bf16_to_f32(B[k * rsB + j * csB]);
  for (int64_t i = 0; i < M; ++i)
    for (int64_t j = 0; j < N; ++j)
      c32[i * N + j] = (beta == 0.0f) ? 0.0f : bf16_to_f32(C[i * rsC + j * csC]);
  oracle_gemm_strided_f32(M, N, K, alpha, a32, K, 1, b32, N, 1, beta, c32, N, 1);
  for (int64_t i = 0; i < M; ++i)
    for (int64_t j = 0; j < N; ++j) C[i * rsC + j * csC] = f32_to_bf16_rne(c32[i * N + j]);
  free(a32);
  free(b32);
  free(c32);
}

/* ------------------------------------------------------------------ */
/*                 fp64-accumulated independent product                */
/* ------------------------------------------------------------------ */
void oracle_gemm_f32_in_f64(int64_t M, int64_t N, int64_t K, const float *A, int64_t rsA,
                            int64_t csA, const float *B, int64_t rsB, int64_t csB, double *C) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < M; ++i) {
    double *crow = C + i * N;
    for (int64_t j = 0; j < N; ++j) crow[j] = 0.0;
    for (int64_t k = 0; k < K; ++k) {
      const double a = (double)A[i * rsA + k * csA];
      for (int64_t j = 0; j < N; ++j) crow[j] += a * (double)B[k * rsB + j * csB];
    }
  }
}

/* ------------------------------------------------------------------ */
/*        error metrics (laser/private/error_functions.nim:6-34)       */
/* ------------------------------------------------------------------ */
double oracle_relative_error(double y, double y_true) {
  const double d = fmax(fabs(y_true), fabs(y));
  if (d == 0.0) return 0.0;
  return fabs(y_true - y) / d;
}
double oracle_mean_relative_error_f32(const float *y, const float *y_true, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += oracle_relative_error((double)y[i], (double)y_true[i]);
  return n > 0 ? s / (double)n : 0.0;
}
/* max_i |y - y_true| / |y_true| -- BASELINE.json's gate "max |ours-ref|/|ref|". */
double oracle_max_relative_error_f32(const float *y, const float *y_true, int64_t n) {
  double m = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double t = fabs((double)y_true[i]);
    const double e = fabs((double)y[i] - (double)y_true[i]);
    const double r = (t == 0.0) ? (e == 0.0 ? 0.0 : INFINITY) : e / t;
    if (r > m) m = r;
  }
  return m;
}
double oracle_normwise_relative_error_f32(const float *y, const float *y_true, int64_t n) {
  double num = 0.0, den = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double e = (double)y[i] - (double)y_true[i];
    num += e * e;
    den += (double)y_true[i] * (double)y_true[i];
  }
  return den == 0.0 ? (num == 0.0 ? 0.0 : INFINITY) : sqrt(num / den);
}

/* ------------------------------------------------------------------ */
/*     counter-based inputs, identical on host and device  */
/* ------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
void oracle_fill_uniform_f32(float *dst, int64_t n, uint64_t seed, float lo, float hi) {
  const float span = hi - lo;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i);
    const float u = (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f); /* 24 bits, exact */
    dst[i] = fmaf(u, span, lo);
  }
}
