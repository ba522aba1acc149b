/*
 * TEST INFRASTRUCTURE ONLY (see laser_oracle.h).
 * AVX-512 fp32 micro-kernel, 14 x 32 (MR=14, NbVecs=2 x 16 lanes: 28 zmm
 * accumulators), following gemm_ukernel_generator.nim:140-250 instantiated by
 * gemm_ukernel_avx512.nim:10-23 (_mm512_fmadd_ps).  Compiled with -mavx512f
 * -mavx512dq only on this TU (nim.cfg:24-30).
 */
#include <immintrin.h>
#include <stdint.h>

void laser_ukernel_f32_avx512(int64_t kc, const float *pa, const float *pb, float *AB) {
  enum { MR = 14, NV = 2, NR = 32 };
  __m512 ab[MR][NV];
  for (int i = 0; i < MR; ++i)
    for (int v = 0; v < NV; ++v) ab[i][v] = _mm512_setzero_ps();
  for (int64_t k = 0; k < kc; ++k) {
    const __m512 b0 = _mm512_load_ps(pb + k * NR);
    const __m512 b1 = _mm512_load_ps(pb + k * NR + 16);
    _mm_prefetch((const char *)(pb + (k + 1) * NR), _MM_HINT_T0);
#pragma GCC unroll 14
    for (int i = 0; i < MR; ++i) {
      const __m512 a = _mm512_set1_ps(pa[k * MR + i]);
      ab[i][0] = _mm512_fmadd_ps(a, b0, ab[i][0]);
      ab[i][1] = _mm512_fmadd_ps(a, b1, ab[i][1]);
    }
  }
  for (int i = 0; i < MR; ++i)
    for (int v = 0; v < NV; ++v) _mm512_store_ps(AB + i * NR + v * 16, ab[i][v]);
}
