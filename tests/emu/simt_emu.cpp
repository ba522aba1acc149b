// simt_emu.cpp -- TEST INFRASTRUCTURE: the exact CUDA-core GEMM kernel of
// laser_b200/csrc/gemm_simt.cuh (gemm_simt_kernel, the path that must be bit-identical to the CPU
// reference) compiled for the host (cuda_emu.h) with the library's own launch planning
// (simt_plan) and the library's tile configurations (8x8x16 for 4-byte, 4x4x16 for 8-byte types).
#define LB200_HOST_EMULATION 1
#include "cuda_emu.h"

#include "ptx_emu.h"
#include "../../laser_b200/csrc/gemm_simt.cuh"
#include "../../laser_b200/csrc/gemm_dmma.cuh"

using namespace lb200;

template <typename T, int TM, int TN, int BK>
static int run(int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t rsA, int64_t csA, const T *B, int64_t rsB,
               int64_t csB, T beta, T *C, int64_t rsC, int64_t csC, int grid, const float *bias, int bias_per_row,
               int act, int64_t batch = 1, int64_t bsA = 0, int64_t bsB = 0, int64_t bsC = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;   // gemm.nim:150: nothing to do, C untouched
  SimtParams<T> p;
  const int64_t tiles = simt_plan<T, TM, TN>(p, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
  p.bias = bias; p.bias_per_row = bias_per_row; p.act = act;
  p.batch = batch; p.bsA = bsA; p.bsB = bsB; p.bsC = bsC;
  if (batch <= 0) return 0;
  if (grid <= 0 || grid > tiles * batch) grid = static_cast<int>(tiles * batch);
  if (batch > 1) emu::launch(static_cast<unsigned>(grid), 256, [=]() { gemm_simt_batched_kernel<T, TM, TN, BK>(p); });
  else emu::launch(static_cast<unsigned>(grid), 256, [=]() { gemm_simt_kernel<T, TM, TN, BK>(p); });
  return static_cast<int>(tiles);
}

extern "C" {
int emu_gemm_simt_f32(int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t rsA, int64_t csA,
                      const float *B, int64_t rsB, int64_t csB, float beta, float *C, int64_t rsC, int64_t csC,
                      int grid, const float *bias, int bias_per_row, int act) {
  return run<float, 8, 8, 16>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, grid, bias, bias_per_row, act);
}
int emu_gemm_simt_batched_f32(int64_t batch, int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t rsA,
                              int64_t csA, int64_t bsA, const float *B, int64_t rsB, int64_t csB, int64_t bsB, float beta,
                              float *C, int64_t rsC, int64_t csC, int64_t bsC, int grid) {
  return run<float, 8, 8, 16>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, grid, nullptr, 0, 0, batch, bsA,
                              bsB, bsC);
}
int emu_gemm_simt_f64(int64_t M, int64_t N, int64_t K, double alpha, const double *A, int64_t rsA, int64_t csA,
                      const double *B, int64_t rsB, int64_t csB, double beta, double *C, int64_t rsC, int64_t csC,
                      int grid) {
  return run<double, 4, 4, 16>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, grid, nullptr, 0, 0);
}
// the fp64 tensor-core kernel (gemm_dmma.cuh), planned and launched as capi.cu: gemm_simt<double> does
int emu_gemm_dmma_f64(int64_t batch, int64_t M, int64_t N, int64_t K, double alpha, const double *A, int64_t rsA, int64_t csA,
                      int64_t bsA, const double *B, int64_t rsB, int64_t csB, int64_t bsB, double beta, double *C, int64_t rsC,
                      int64_t csC, int64_t bsC, int grid) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return 0;
  SimtParams<double> p;
  const int64_t tiles = dmma_plan(p, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
  p.batch = batch; p.bsA = bsA; p.bsB = bsB; p.bsC = bsC;
  if (grid <= 0 || grid > tiles * batch) grid = static_cast<int>(tiles * batch);
  if (DMMA_SMEM_BYTES > emu::kDynSmemBytes) return -1;
  emu::launch(static_cast<unsigned>(grid), 256, [=]() { gemm_dmma_kernel(p); });
  return static_cast<int>(tiles);
}
// few rows, wide N (gemm_simt.cuh: gemm_skinny_m_kernel), launched as capi.cu: gemm_simt<float> does
int emu_gemm_skinny_m_f32(int64_t batch, int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t rsA, int64_t csA,
                          int64_t bsA, const float *B, int64_t rsB, int64_t csB, int64_t bsB, float beta, float *C, int64_t rsC,
                          int64_t csC, int64_t bsC, int grid, const float *bias, int bias_per_row, int act) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0 || M > 32) return 0;
  SimtParams<float> p;
  simt_plan<float, 8, 8>(p, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
  p.bias = bias; p.bias_per_row = bias_per_row; p.act = act;
  p.batch = batch; p.bsA = bsA; p.bsB = bsB; p.bsC = bsC;
  if (grid <= 0) grid = 2;
  // capi.cu: gemm_simt<float> -- B with unit column stride and 16-byte aligned rows takes the cp.async-staged kernel
  const bool async_ok = csB == 1 && rsB % 4 == 0 && (batch == 1 || bsB % 4 == 0) && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
  if (async_ok && ska_smem_bytes<32>() <= emu::kDynSmemBytes) {
    if (M <= 8) emu::launch(grid, 256, [=]() { gemm_skinny_m_async_kernel<8>(p); });
    else if (M <= 16) emu::launch(grid, 256, [=]() { gemm_skinny_m_async_kernel<16>(p); });
    else if (M <= 24) emu::launch(grid, 256, [=]() { gemm_skinny_m_async_kernel<24>(p); });
    else emu::launch(grid, 256, [=]() { gemm_skinny_m_async_kernel<32>(p); });
    return 2;
  }
  if (M <= 8) emu::launch(grid, 256, [=]() { gemm_skinny_m_kernel<8, 4>(p); });
  else if (M <= 16) emu::launch(grid, 256, [=]() { gemm_skinny_m_kernel<16, 4>(p); });
  else if (M <= 24) emu::launch(grid, 256, [=]() { gemm_skinny_m_kernel<24, 2>(p); });
  else emu::launch(grid, 256, [=]() { gemm_skinny_m_kernel<32, 2>(p); });
  return 1;
}
int emu_gemm_simt_i32(int64_t M, int64_t N, int64_t K, int32_t alpha, const int32_t *A, int64_t rsA, int64_t csA,
                      const int32_t *B, int64_t rsB, int64_t csB, int32_t beta, int32_t *C, int64_t rsC, int64_t csC,
                      int grid) {
  return run<int32_t, 8, 8, 16>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, grid, nullptr, 0, 0);
}
int emu_gemm_simt_i64(int64_t M, int64_t N, int64_t K, int64_t alpha, const int64_t *A, int64_t rsA, int64_t csA,
                      const int64_t *B, int64_t rsB, int64_t csB, int64_t beta, int64_t *C, int64_t rsC, int64_t csC,
                      int grid) {
  return run<int64_t, 4, 4, 16>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, grid, nullptr, 0, 0);
}
int emu_gemm_simt_batched_f64(int64_t batch, int64_t M, int64_t N, int64_t K, double alpha, const double *A, int64_t rsA,
                              int64_t csA, int64_t bsA, const double *B, int64_t rsB, int64_t csB, int64_t bsB,
                              double beta, double *C, int64_t rsC, int64_t csC, int64_t bsC, int grid) {
  return run<double, 4, 4, 16>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, grid, nullptr, 0, 0, batch, bsA,
                               bsB, bsC);
}
int emu_gemm_simt_batched_i64(int64_t batch, int64_t M, int64_t N, int64_t K, int64_t alpha, const int64_t *A, int64_t rsA,
                              int64_t csA, int64_t bsA, const int64_t *B, int64_t rsB, int64_t csB, int64_t bsB,
                              int64_t beta, int64_t *C, int64_t rsC, int64_t csC, int64_t bsC, int grid) {
  return run<int64_t, 4, 4, 16>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, grid, nullptr, 0, 0, batch,
                                bsA, bsB, bsC);
}
// skinny GEMM (N <= 4): variant 0 = scalar loads, 1 = float4 loads of A, 2 = B staged in shared memory
int emu_gemv(int variant, int NV, int64_t M, int64_t K, float alpha, const float *A, int64_t rsA, int64_t csA,
             const float *B, int64_t rsB, int64_t csB, float beta, float *C, int64_t rsC, int64_t csC, int grid) {
#define GEMV(NVC)                                                                                                  \
  do {                                                                                                             \
    if (variant == 2) emu::launch(grid, 256, [=]() { gemv_warp_smem_kernel<NVC>(M, K, alpha, A, rsA, B, rsB, csB, beta, C, rsC, csC); }); \
    else if (variant == 1) emu::launch(grid, 256, [=]() { gemv_warp_kernel<NVC, true>(M, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC); }); \
    else emu::launch(grid, 256, [=]() { gemv_warp_kernel<NVC, false>(M, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC); }); \
  } while (0)
  if (variant == 2 && static_cast<size_t>(NV) * K * 4 > 96 * 1024) return -1;
  if (NV == 1) GEMV(1); else if (NV == 2) GEMV(2); else if (NV == 3) GEMV(3); else if (NV == 4) GEMV(4); else return -1;
#undef GEMV
  return 0;
}
}  // extern "C"
