/*
 * TEST INFRASTRUCTURE ONLY (see laser_oracle.h).
 * AVX+FMA3 fp32 micro-kernel, 6 x 16 (MR=6, NbVecs=2 x 8 lanes), following the
 * generated code shape of gemm_ukernel_generator.nim:140-250 instantiated by
 * gemm_ukernel_avx_fma.nim:10-23 (_mm256_fmadd_ps): per k, load NbVecs B
 * vectors, broadcast each of the MR A scalars, one FMA per (i, vec).
 * Compiled with -mavx -mfma only on this TU (nim.cfg:24-30).
 */
#include <immintrin.h>
#include <stdint.h>

void laser_ukernel_f32_avx_fma(int64_t kc, const float *pa, const float *pb, float *AB) {
  enum { MR = 6, NV = 2, NR = 16 };
  __m256 ab[MR][NV];
  for (int i = 0; i < MR; ++i)
    for (int v = 0; v < NV; ++v) ab[i][v] = _mm256_setzero_ps();
  for (int64_t k = 0; k < kc; ++k) {
    const __m256 b0 = _mm256_load_ps(pb + k * NR);
    const __m256 b1 = _mm256_load_ps(pb + k * NR + 8);
#pragma GCC unroll 6
    for (int i = 0; i < MR; ++i) {
      const __m256 a = _mm256_broadcast_ss(pa + k * MR + i);
      ab[i][0] = _mm256_fmadd_ps(a, b0, ab[i][0]);
      ab[i][1] = _mm256_fmadd_ps(a, b1, ab[i][1]);
    }
  }
  for (int i = 0; i < MR; ++i)
    for (int v = 0; v < NV; ++v) _mm256_store_ps(AB + i * NR + v * 8, ab[i][v]);
}
