/*
 * laser_cpu_gemm.c -- TEST INFRASTRUCTURE ONLY (see laser_oracle.h).
 *
 * Structure-faithful C restatement of laser's CPU gemm_strided for fp32: the
 * BLIS/Goto five-loop algorithm with packing, a register micro-kernel and
 * OpenMP, written from the behaviour of (paths relative to /root/reference):
 *   laser/primitives/matrix_multiplication/gemm.nim:48-101   gebp_mkernel  (loops jr, ir)
 *   .../gemm.nim:109-176                                      gemm_impl     (loops pc, ic)
 *   .../gemm.nim:184-247                                      gemm_strided  (ISA dispatch)
 *   .../gemm_tiling.nim:276-341                               partitionMNK / newTiles
 *   .../gemm_packing.nim:24-94                                pack_A_mc_kc / pack_B_kc_nc
 *   .../gemm_ukernel_generic.nim:53-126                       epilogues
 * The micro-kernels live in their own translation units so that ISA flags are
 * applied per file, as the reference's nim.cfg:24-30 does.
 *
 * This is the implementation TIMED as the CPU baseline ("port"): the reference
 * itself is Nim and cannot be compiled in this image.
 */
#include "laser_oracle.h"

#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* micro-kernels: AB[MR][NR] (row-major, 64-B aligned) = sum_k packA[k*MR+i] * packB[k*NR+j] */
void laser_ukernel_f32_generic(int64_t kc, const float *pa, const float *pb, float *AB);
void laser_ukernel_f32_avx_fma(int64_t kc, const float *pa, const float *pb, float *AB);
void laser_ukernel_f32_avx512(int64_t kc, const float *pa, const float *pb, float *AB);

typedef void (*ukernel_fn)(int64_t, const float *, const float *, float *);

typedef struct {
  int mr, nr; /* gemm_tiling.nim:147-219 */
  ukernel_fn fn;
} microkernel_t;

enum { ISA_AUTO = 0, ISA_GENERIC = 1, ISA_AVX_FMA = 2, ISA_AVX512 = 3 };

int laser_cpu_detect_isa(void) {
  /* dispatch order of gemm.nim:229-233 (fp32); SSE / AVX-without-FMA hosts are
   * not restated: they map to the generic kernel here. */
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return ISA_AVX512;
  if (__builtin_cpu_supports("fma") && __builtin_cpu_supports("avx")) return ISA_AVX_FMA;
  return ISA_GENERIC;
}

int laser_cpu_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void laser_cpu_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static microkernel_t select_ukernel(int isa) {
  microkernel_t u;
  switch (isa) {
    case ISA_AVX512: u.mr = 14; u.nr = 32; u.fn = laser_ukernel_f32_avx512; break;
    case ISA_AVX_FMA: u.mr = 6; u.nr = 16; u.fn = laser_ukernel_f32_avx_fma; break;
    default: u.mr = 2; u.nr = 1; u.fn = laser_ukernel_f32_generic; break;
  }
  return u;
}

static inline int64_t round_up(int64_t x, int64_t step) { return (x + step - 1) / step * step; }
static inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

static void *alloc64(size_t bytes) {
  void *p = NULL;
  if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) return NULL;
  return p;
}

/* gemm_packing.nim:24-55: [mc/MR] micro-panels, each [kc][MR], zero-padded tail */
static void pack_A_mc_kc(float *restrict buf, int64_t mc, int64_t kc, const float *A, int64_t rs,
                         int64_t cs, int MR) {
  const int64_t stop = mc - mc % MR;
  for (int64_t i = 0; i < stop; i += MR)
    for (int64_t k = 0; k < kc; ++k)
      for (int ii = 0; ii < MR; ++ii) buf[i * kc + k * MR + ii] = A[(i + ii) * rs + k * cs];
  const int64_t rem = mc - stop;
  if (rem > 0) {
    float *off = buf + kc * stop;
    for (int64_t k = 0; k < kc; ++k) {
      for (int64_t i = 0; i < rem; ++i) off[k * MR + i] = A[(stop + i) * rs + k * cs];
      for (int64_t i = rem; i < MR; ++i) off[k * MR + i] = 0.0f;
    }
  }
}

/* gemm_packing.nim:63-94: [nc/NR] micro-panels, each [kc][NR]; own parallel-for */
static void pack_B_kc_nc(float *restrict buf, int64_t kc, int64_t nc, const float *B, int64_t rs,
                         int64_t cs, int NR) {
  const int64_t stop = nc - nc % NR;
#pragma omp parallel for
  for (int64_t j = 0; j < stop; j += NR)
    for (int64_t k = 0; k < kc; ++k)
      for (int jj = 0; jj < NR; ++jj) buf[j * kc + k * NR + jj] = B[k * rs + (j + jj) * cs];
  const int64_t rem = nc - stop;
  if (rem > 0) {
    float *off = buf + kc * stop;
    for (int64_t k = 0; k < kc; ++k) {
      for (int64_t j = 0; j < rem; ++j) off[k * NR + j] = B[k * rs + (stop + j) * cs];
      for (int64_t j = rem; j < NR; ++j) off[k * NR + j] = 0.0f;
    }
  }
}

/* gemm_ukernel_generic.nim:53-76 (full tile) and :97-126 (edge) */
static inline void epilogue(int mr, int nr, int NR, float alpha, const float *AB, float beta,
                            float *C, int64_t rs, int64_t cs, int full) {
  if (full) {
    if (beta == 0.0f) {
      for (int i = 0; i < mr; ++i)
        for (int j = 0; j < nr; ++j) C[i * rs + j * cs] = 0.0f;
    } else if (beta != 1.0f) {
      for (int i = 0; i < mr; ++i)
        for (int j = 0; j < nr; ++j) C[i * rs + j * cs] *= beta;
    }
    if (alpha == 1.0f) {
      for (int i = 0; i < mr; ++i)
        for (int j = 0; j < nr; ++j) C[i * rs + j * cs] += AB[i * NR + j];
    } else {
      for (int i = 0; i < mr; ++i)
        for (int j = 0; j < nr; ++j) C[i * rs + j * cs] += alpha * AB[i * NR + j];
    }
  } else if (beta == 0.0f) {
    if (alpha == 1.0f) {
      for (int i = 0; i < mr; ++i)
        for (int j = 0; j < nr; ++j) C[i * rs + j * cs] = AB[i * NR + j];
    } else {
      for (int i = 0; i < mr; ++i)
        for (int j = 0; j < nr; ++j) C[i * rs + j * cs] = alpha * AB[i * NR + j];
    }
  } else {
    /* note: the edge epilogue multiplies by beta even when beta == 1 */
    for (int i = 0; i < mr; ++i)
      for (int j = 0; j < nr; ++j) C[i * rs + j * cs] *= beta;
    if (alpha == 1.0f) {
      for (int i = 0; i < mr; ++i)
        for (int j = 0; j < nr; ++j) C[i * rs + j * cs] += AB[i * NR + j];
    } else {
      for (int i = 0; i < mr; ++i)
        for (int j = 0; j < nr; ++j) C[i * rs + j * cs] += alpha * AB[i * NR + j];
    }
  }
}

/* gemm.nim:48-101 */
static void gebp_mkernel(microkernel_t u, int64_t mc, int64_t nc, int64_t kc, float alpha,
                         const float *packA, const float *packB, float beta, float *C, int64_t rs,
                         int64_t cs) {
  const int MR = u.mr, NR = u.nr;
#pragma omp taskloop firstprivate(u, mc, nc, kc, alpha, packA, packB, beta, C, rs, cs)
  for (int64_t jr = 0; jr < nc; jr += NR) {
    const int nr = (int)min64(nc - jr, NR);
    float AB[14 * 32] __attribute__((aligned(64)));
    for (int64_t ir = 0; ir < mc; ir += MR) {
      const int mr = (int)min64(mc - ir, MR);
      u.fn(kc, packA + ir * kc, packB + jr * kc, AB);
      epilogue(mr, nr, NR, alpha, AB, beta, C + ir * rs + jr * cs, rs, cs, (mr == MR && nr == NR));
    }
  }
}

int laser_cpu_gemm_strided_f32(int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                               int64_t rsA, int64_t csA, const float *B, int64_t rsB, int64_t csB,
                               float beta, float *C, int64_t rsC, int64_t csC, int isa) {
  if (isa == ISA_AUTO) isa = laser_cpu_detect_isa();
  if (M <= 0 || N <= 0 || K <= 0) return isa;
  const microkernel_t u = select_ukernel(isa);

  /* gemm_tiling.nim:309-341 */
  const int64_t mc = min64(768 / (int64_t)sizeof(float), M);
  const int64_t kc_t = min64(2048 / (int64_t)sizeof(float), K);
  const int64_t nc = N;
  const int64_t ic_num_tasks = (M + mc - 1) / mc;
  const int64_t upanelA_size = kc_t * round_up(mc, u.mr);
  float *tiles_a = (float *)alloc64(sizeof(float) * (size_t)(upanelA_size * ic_num_tasks));
  float *tiles_b = (float *)alloc64(sizeof(float) * (size_t)(kc_t * round_up(nc, u.nr)));
  if (!tiles_a || !tiles_b) { free(tiles_a); free(tiles_b); return -1; }

  const int parallelize = (M * N * K > (int64_t)128 * 128 * 128); /* gemm.nim:140-141 */

  for (int64_t pc = 0; pc < K; pc += kc_t) { /* gemm.nim:150 */
    const int64_t kc = min64(K - pc, kc_t);
    pack_B_kc_nc(tiles_b, kc, nc, B + pc * rsB, rsB, csB, u.nr);
    const float beta1 = (pc == 0) ? beta : 1.0f; /* gemm.nim:158 */
#pragma omp parallel if (parallelize)
    {
#pragma omp for nowait
      for (int64_t icb = 0; icb < ic_num_tasks; ++icb) { /* gemm.nim:163 */
        float *packA = tiles_a + icb * upanelA_size;
        const int64_t ic = icb * mc;
        const int64_t mcur = min64(M - ic, mc);
        pack_A_mc_kc(packA, mcur, kc, A + ic * rsA + pc * csA, rsA, csA, u.mr);
        gebp_mkernel(u, mcur, nc, kc, alpha, packA, tiles_b, beta1, C + ic * rsC, rsC, csC);
      }
    }
  }
  free(tiles_a);
  free(tiles_b);
  return isa;
}
