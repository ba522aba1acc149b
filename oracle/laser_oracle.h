/*
 * laser_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of mratsim/laser's strided GEMM hot path
 * (laser/primitives/matrix_multiplication).  It exists to CHECK the CUDA
 * product path in laser_b200/; nothing under laser_b200/ may call, link or
 * import it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * The reference is Nim; no Nim toolchain exists in this image, so the
 * reference itself cannot be compiled here (oracle/_ref is therefore absent,
 * cpu_baseline.kind == "port").  Parity of this restatement is PINNED by the
 * reference's own known-answer vectors (gemm.nim:255-507,
 * gemm_prepacked.nim:354-367), committed as tests/golden/known_answer.json and
 * checked by tests/test_oracle.py for every dtype and for both flavours below.
 * fp32 at scale / non-unit strides / alpha,beta != (1,0) have no golden vector
 * in the reference ("parity unpinned by the reference's tests" for those
 * cases); they are pinned by cross-checking the two independent flavours
 * against each other bit-for-bit and against an fp64 product.
 *
 * Two flavours:
 *   oracle_gemm_strided_*   numerics-faithful: per output element, an FMA
 *                           chain over k inside consecutive kc blocks
 *                           (kc = min(2048/sizeof T, K), gemm_tiling.nim:309-310),
 *                           blocks added in order, reference epilogue rules.
 *   laser_cpu_gemm_strided_f32
 *                           structure-faithful: 5-loop BLIS/Goto with packing,
 *                           MR x NR register micro-kernel and OpenMP, following
 *                           gemm.nim:48-176, gemm_packing.nim:24-94,
 *                           gemm_ukernel_generator.nim:140-250.  This is the one
 *                           that is TIMED as the CPU baseline.
 */
#ifndef LASER_ORACLE_H
#define LASER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- numerics-faithful oracle (gemm.nim:109-176 order of operations) ---- */
void oracle_gemm_strided_f32(int64_t M, int64_t N, int64_t K, float alpha,
                             const float *A, int64_t rsA, int64_t csA,
                             const float *B, int64_t rsB, int64_t csB,
                             float beta, float *C, int64_t rsC, int64_t csC);
void oracle_gemm_strided_f64(int64_t M, int64_t N, int64_t K, double alpha,
                             const double *A, int64_t rsA, int64_t csA,
                             const double *B, int64_t rsB, int64_t csB,
                             double beta, double *C, int64_t rsC, int64_t csC);
void oracle_gemm_strided_i32(int64_t M, int64_t N, int64_t K, int32_t alpha,
                             const int32_t *A, int64_t rsA, int64_t csA,
                             const int32_t *B, int64_t rsB, int64_t csB,
                             int32_t beta, int32_t *C, int64_t rsC, int64_t csC);
void oracle_gemm_strided_i64(int64_t M, int64_t N, int64_t K, int64_t alpha,
                             const int64_t *A, int64_t rsA, int64_t csA,
                             const int64_t *B, int64_t rsB, int64_t csB,
                             int64_t beta, int64_t *C, int64_t rsC, int64_t csC);

/* bf16 restatement used for the bf16 GEMM config (type does not exist in the
 * reference): inputs are bf16 bit patterns, promoted to fp32, run through the
 * fp32 reference order, result rounded RNE to bf16. */
void oracle_gemm_strided_bf16(int64_t M, int64_t N, int64_t K, float alpha,
                              const uint16_t *A, int64_t rsA, int64_t csA,
                              const uint16_t *B, int64_t rsB, int64_t csB,
                              float beta, uint16_t *C, int64_t rsC, int64_t csC);

/* fp64-accumulated product of fp32 inputs (independent cross-check only). */
void oracle_gemm_f32_in_f64(int64_t M, int64_t N, int64_t K,
                            const float *A, int64_t rsA, int64_t csA,
                            const float *B, int64_t rsB, int64_t csB,
                            double *C /* M x N row-major */);

/* ---- structure-faithful restatement (the timed CPU baseline) ---- */
/* isa: 0 = runtime detect (gemm.nim:229-233), 1 = generic scalar (2x1... see .c),
 *      2 = AVX+FMA 6x16, 3 = AVX-512 14x32.  Returns the isa actually used. */
int laser_cpu_gemm_strided_f32(int64_t M, int64_t N, int64_t K, float alpha,
                               const float *A, int64_t rsA, int64_t csA,
                               const float *B, int64_t rsB, int64_t csB,
                               float beta, float *C, int64_t rsC, int64_t csC,
                               int isa);
int laser_cpu_detect_isa(void);
int laser_cpu_num_threads(void);
void laser_cpu_set_num_threads(int n);

/* ---- error metrics (laser/private/error_functions.nim:6-34) ---- */
double oracle_relative_error(double y, double y_true);
double oracle_mean_relative_error_f32(const float *y, const float *y_true, int64_t n);
double oracle_max_relative_error_f32(const float *y, const float *y_true, int64_t n);
double oracle_normwise_relative_error_f32(const float *y, const float *y_true, int64_t n);

/* counter-based generator shared bit-for-bit with the device side
 * (laser_b200_fill_uniform_f32): value(seed, idx) in [lo, hi). */
void oracle_fill_uniform_f32(float *dst, int64_t n, uint64_t seed, float lo, float hi);

/* ---- the steps either side of the GEMM (laser_layers.c) ---- */
void oracle_transpose2d_copy(void *dst, const void *src, int64_t NR, int64_t NC, int elem_size);
void oracle_transpose2d_batched(void *dst, const void *src, int64_t N, int64_t NR, int64_t NC, int elem_size);
void oracle_nchw2nhwc(void *dst, const void *src, int64_t N, int64_t C, int64_t H, int64_t W, int elem_size);
void oracle_nhwc2nchw(void *dst, const void *src, int64_t N, int64_t C, int64_t H, int64_t W, int elem_size);
int oracle_conv2d_out_shape(const int64_t ishape[4], const int64_t kshape[4], const int64_t padding[2],
                            const int64_t strides[2], int64_t out[4]);
int64_t oracle_im2col_workspace_size(const int64_t ishape[4], const int64_t kshape[4], const int64_t padding[2],
                                     const int64_t strides[2]);
void oracle_im2col_f32(float *workspace, int64_t outH, int64_t outW, const float *input, int64_t C, int64_t H,
                       int64_t W, int64_t kH, int64_t kW, int64_t pH, int64_t pW, int64_t sH, int64_t sW);
int oracle_conv2d_im2col_f32(float *output, const float *input, const int64_t ishape[4], const float *kernel,
                             const int64_t kshape[4], const int64_t padding[2], const int64_t strides[2],
                             float *workspace);
int oracle_conv2d_direct_f32(float *output, const float *input, const int64_t ishape[4], const float *kernel,
                             const int64_t kshape[4], const int64_t padding[2], const int64_t strides[2]);
void oracle_gemm_strided_batched_f32(int64_t batch, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                                     int64_t rsA, int64_t csA, int64_t bsA, const float *B, int64_t rsB,
                                     int64_t csB, int64_t bsB, float beta, float *C, int64_t rsC, int64_t csC,
                                     int64_t bsC);

#ifdef __cplusplus
}
#endif
#endif
