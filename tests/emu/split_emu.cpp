// split_emu.cpp -- TEST INFRASTRUCTURE: the operand-preparation kernels of
// laser_b200/csrc/split.cuh (split_rows_tf32 / split_rows_mixed / pack_general / splitk_reduce /
// fill_uniform) compiled for the host (cuda_emu.h; cvt.rna.tf32 replaced by its software
// definition) behind a C interface for ctypes.
#define LB200_HOST_EMULATION 1
#include "cuda_emu.h"

#include "../../laser_b200/csrc/split.cuh"

using namespace lb200;

extern "C" {

void emu_split_rows_tf32(const float *src, int64_t R, int64_t Cc, int64_t src_ld, float *hi, float *lo, int64_t dst_ld,
                         int grid) {
  emu::launch(grid, 256, [=]() { split_rows_tf32_kernel(src, R, Cc, src_ld, hi, lo, dst_ld); });
}
void emu_split_rows_mixed(const float *src, int64_t R, int64_t Cc, int64_t src_ld, float *hi, int64_t dst_ld,
                          uint16_t *xb, uint16_t *lb, int64_t ld_b, int grid) {
  emu::launch(grid, 256, [=]() { split_rows_mixed_kernel(src, R, Cc, src_ld, hi, dst_ld, xb, lb, ld_b); });
}
void emu_split_rows_bf16x2(const float *src, int64_t R, int64_t Cc, int64_t src_ld, uint16_t *hb, uint16_t *lb, int64_t ld_b,
                           int grid) {
  emu::launch(grid, 256, [=]() { split_rows_bf16x2_kernel(src, R, Cc, src_ld, hb, lb, ld_b); });
}
// per_col 0: one abs-max word per row (K-major operand), 1: one per column (MN-major operand)
void emu_absmax_mn(int per_col, const float *src, int64_t R, int64_t Cc, int64_t src_ld, uint32_t *out, int grid) {
  if (per_col) emu::launch(grid, 256, [=]() { absmax_mn_kernel<true>(src, R, Cc, src_ld, out); });
  else emu::launch(grid, 256, [=]() { absmax_mn_kernel<false>(src, R, Cc, src_ld, out); });
}
void emu_split_rows_f16x2(int per_col, const float *src, int64_t R, int64_t Cc, int64_t src_ld, uint16_t *hb, uint16_t *lb,
                          int64_t ld_b, const uint32_t *absmax, int grid) {
  if (per_col) emu::launch(grid, 256, [=]() { split_rows_f16x2_kernel<true>(src, R, Cc, src_ld, hb, lb, ld_b, absmax); });
  else emu::launch(grid, 256, [=]() { split_rows_f16x2_kernel<false>(src, R, Cc, src_ld, hb, lb, ld_b, absmax); });
}
// mode 0: copy, 1: tf32 hi/lo, 2: mixed (hi fp32 + xb/lb bf16), 3: two bf16 pieces (xb, lb)
void emu_pack_general_f32(int mode, const float *src, int64_t R, int64_t Cc, int64_t sr, int64_t sc, float *dst,
                          float *dst_lo, int64_t ld, int read_along_r, uint16_t *xb, uint16_t *lb, int64_t ld_b, int grid) {
  if (mode == 0) emu::launch(grid, 256, [=]() { pack_general_kernel<float, 0>(src, R, Cc, sr, sc, dst, dst_lo, ld, read_along_r, xb, lb, ld_b); });
  else if (mode == 1) emu::launch(grid, 256, [=]() { pack_general_kernel<float, 1>(src, R, Cc, sr, sc, dst, dst_lo, ld, read_along_r, xb, lb, ld_b); });
  else if (mode == 3) emu::launch(grid, 256, [=]() { pack_general_kernel<float, 3>(src, R, Cc, sr, sc, dst, dst_lo, ld, read_along_r, xb, lb, ld_b); });
  else emu::launch(grid, 256, [=]() { pack_general_kernel<float, 2>(src, R, Cc, sr, sc, dst, dst_lo, ld, read_along_r, xb, lb, ld_b); });
}
void emu_pack_general_u16(const uint16_t *src, int64_t R, int64_t Cc, int64_t sr, int64_t sc, uint16_t *dst, int64_t ld,
                          int read_along_r, int grid) {
  emu::launch(grid, 256, [=]() { pack_general_kernel<uint16_t, 0>(src, R, Cc, sr, sc, dst, nullptr, ld, read_along_r, nullptr, nullptr, 0); });
}
void emu_splitk_reduce(const float *ws, int S, int64_t M, int64_t N, int64_t ld, int64_t plane, float alpha, float beta,
                       float *C, int64_t rsC, int64_t csC, const float *bias, int bias_per_row, int act, int grid) {
  emu::launch(grid, 256, [=]() { splitk_reduce_kernel(ws, S, M, N, ld, plane, alpha, beta, C, rsC, csC, bias, bias_per_row, act); });
}
void emu_fill_uniform_f32(float *dst, int64_t n, uint64_t seed, float lo, float hi, int grid) {
  emu::launch(grid, 256, [=]() { fill_uniform_f32_kernel(dst, n, seed, lo, hi); });
}

}  // extern "C"
