/*
 * TEST INFRASTRUCTURE ONLY (see laser_oracle.h).
 * Scalar micro-kernel, x86_Generic configuration MR=2, NR=1
 * (gemm_tiling.nim:147-219: regs=2, NbVecs=1, nb_scalars=1) following
 * ukernel_generic_impl, gemm_ukernel_generic.nim:21-35 (AB += a*b, unfused:
 * this TU is compiled with -ffp-contract=off and no ISA flags).
 */
#include <stdint.h>

void laser_ukernel_f32_generic(int64_t kc, const float *pa, const float *pb, float *AB) {
  enum { MR = 2, NR = 1 };
  float ab[MR][NR] = {{0.0f}, {0.0f}};
  for (int64_t k = 0; k < kc; ++k)
    for (int i = 0; i < MR; ++i)
      for (int j = 0; j < NR; ++j) ab[i][j] += pa[k * MR + i] * pb[k * NR + j];
  for (int i = 0; i < MR; ++i)
    for (int j = 0; j < NR; ++j) AB[i * NR + j] = ab[i][j];
}
