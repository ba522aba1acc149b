/*
 * laser_layers.c -- TEST INFRASTRUCTURE ONLY (see laser_oracle.h).
 *
 * CPU restatement of the steps either side of the GEMM in the reference's intended use
 * (SURVEY.md section 8f, rank 4):
 *   - physical transposition of contiguous matrices and NCHW <-> NHWC conversion
 *     (laser/primitives/swapaxes.nim:16-112),
 *   - im2col + GEMM convolution (benchmarks/convolution/conv2d_im2col.nim:8-166,
 *     shapes from benchmarks/convolution/conv2d_common.nim:15-45),
 *   - a direct convolution used only as an independent cross-check
 *     (benchmarks/convolution/conv2d_direct_convolution.nim:8-76),
 *   - a batch of strided GEMMs (the reference has no batched entry: README.md:253-263 lists it
 *     as roadmap; restated as a loop over gemm_strided).
 * Parity is pinned by the reference's own convolution known-answer vectors
 * (conv2d_common.nim:128-283), committed under tests/golden/conv2d_known_answer.json.
 */
#include <stdint.h>
#include <string.h>

#include "laser_oracle.h"

/* ---- swapaxes.nim:16-54: dst[j][i] = src[i][j]; 74-81: the same per matrix of a batch ---- */
void oracle_transpose2d_batched(void *dst, const void *src, int64_t N, int64_t NR, int64_t NC,
                                int elem_size) {
  const char *ps = (const char *)src;
  char *pd = (char *)dst;
  const int64_t blck = 32; /* swapaxes.nim:41 */
  for (int64_t n = 0; n < N; ++n)
    for (int64_t j = 0; j < NC; j += blck)
      for (int64_t i = 0; i < NR; i += blck)
        for (int64_t jj = j; jj < (j + blck < NC ? j + blck : NC); ++jj)
          for (int64_t ii = i; ii < (i + blck < NR ? i + blck : NR); ++ii)
            memcpy(pd + ((n * NC + jj) * NR + ii) * elem_size, ps + ((n * NR + ii) * NC + jj) * elem_size,
                   (size_t)elem_size);
}
void oracle_transpose2d_copy(void *dst, const void *src, int64_t NR, int64_t NC, int elem_size) {
  oracle_transpose2d_batched(dst, src, 1, NR, NC, elem_size);
}
/* swapaxes.nim:83-97 / 99-112 */
void oracle_nchw2nhwc(void *dst, const void *src, int64_t N, int64_t C, int64_t H, int64_t W, int elem_size) {
  oracle_transpose2d_batched(dst, src, N, C, H * W, elem_size);
}
void oracle_nhwc2nchw(void *dst, const void *src, int64_t N, int64_t C, int64_t H, int64_t W, int elem_size) {
  oracle_transpose2d_batched(dst, src, N, H * W, C, elem_size);
}

/* ---- conv2d_common.nim:15-45 (no dilation); shapes are (n, c, h, w) / (c_out, c_in, kH, kW) ---- */
int oracle_conv2d_out_shape(const int64_t ishape[4], const int64_t kshape[4], const int64_t padding[2],
                            const int64_t strides[2], int64_t out[4]) {
  const int64_t iH = ishape[2], iW = ishape[3], kH = kshape[2], kW = kshape[3];
  if (!(0 < strides[0] && strides[0] < iH) || !(0 < strides[1] && strides[1] < iW)) return -1; /* :35-36 */
  out[0] = ishape[0];
  out[1] = kshape[0];
  out[2] = 1 + (iH + 2 * padding[0] - kH) / strides[0];
  out[3] = 1 + (iW + 2 * padding[1] - kW) / strides[1];
  return 0;
}
/* conv2d_im2col.nim:8-18 (elements, one image) */
int64_t oracle_im2col_workspace_size(const int64_t ishape[4], const int64_t kshape[4], const int64_t padding[2],
                                     const int64_t strides[2]) {
  int64_t o[4];
  if (oracle_conv2d_out_shape(ishape, kshape, padding, strides, o)) return -1;
  return ishape[1] * kshape[2] * kshape[3] * o[2] * o[3];
}

/* conv2d_im2col.nim:44-93: one image [C][H][W] -> [C*kH*kW][outH*outW], zero outside the image */
void oracle_im2col_f32(float *workspace, int64_t outH, int64_t outW, const float *input, int64_t C, int64_t H,
                       int64_t W, int64_t kH, int64_t kW, int64_t pH, int64_t pW, int64_t sH, int64_t sW) {
  float *w = workspace;
  for (int64_t c = 0; c < C; ++c) {
    const float *in = input + c * H * W;
    for (int64_t krow = 0; krow < kH; ++krow)
      for (int64_t kcol = 0; kcol < kW; ++kcol) {
        int64_t row = -pH + krow;
        for (int64_t oh = 0; oh < outH; ++oh, row += sH) {
          if (row < 0 || row >= H) {
            for (int64_t ow = 0; ow < outW; ++ow) *w++ = 0.0f;
          } else {
            int64_t col = -pW + kcol;
            for (int64_t ow = 0; ow < outW; ++ow, col += sW) *w++ = (col >= 0 && col < W) ? in[row * W + col] : 0.0f;
          }
        }
      }
  }
}

/* conv2d_im2col.nim:95-166: per image, O[C_out x outH*outW] = F[C_out x K] * W[K x outH*outW]
 * (alpha 1, beta 0), 1x1 kernels skip im2col and read the image in place (:121,145-149).  The
 * reference calls a BLAS sgemm here; the restatement calls the Laser GEMM oracle. */
int oracle_conv2d_im2col_f32(float *output, const float *input, const int64_t ishape[4], const float *kernel,
                             const int64_t kshape[4], const int64_t padding[2], const int64_t strides[2],
                             float *workspace) {
  int64_t o[4];
  if (oracle_conv2d_out_shape(ishape, kshape, padding, strides, o)) return -1;
  if (ishape[1] != kshape[1]) return -2;
  const int64_t B = ishape[0], C = ishape[1], H = ishape[2], W = ishape[3], kH = kshape[2], kW = kshape[3];
  const int64_t M = kshape[0], K = C * kH * kW, N = o[2] * o[3];
  const int is1x1 = kH * kW == 1;
  for (int64_t n = 0; n < B; ++n) {
    const float *in = input + n * C * H * W;
    const float *rhs = in;
    if (!is1x1) {
      oracle_im2col_f32(workspace, o[2], o[3], in, C, H, W, kH, kW, padding[0], padding[1], strides[0], strides[1]);
      rhs = workspace;
    }
    oracle_gemm_strided_f32(M, N, K, 1.0f, kernel, K, 1, rhs, N, 1, 0.0f, output + n * M * N, N, 1);
  }
  return 0;
}

/* conv2d_direct_convolution.nim:8-76 -- cross-check only (output must be zero-initialised by the
 * caller as in the reference; unfused multiply-add in ci, krow, kcol order).  The reference
 * multiplies the output column by strides.h (:59); square strides only are cross-checked. */
int oracle_conv2d_direct_f32(float *output, const float *input, const int64_t ishape[4], const float *kernel,
                             const int64_t kshape[4], const int64_t padding[2], const int64_t strides[2]) {
  int64_t o[4];
  if (oracle_conv2d_out_shape(ishape, kshape, padding, strides, o)) return -1;
  const int64_t Nb = ishape[0], Cin = kshape[1], H = ishape[2], W = ishape[3], Cout = kshape[0];
  const int64_t kH = kshape[2], kW = kshape[3], outH = o[2], outW = o[3];
  for (int64_t n = 0; n < Nb; ++n)
    for (int64_t co = 0; co < Cout; ++co)
      for (int64_t ci = 0; ci < Cin; ++ci)
        for (int64_t oh = 0; oh < outH; ++oh)
          for (int64_t ow = 0; ow < outW; ++ow) {
            const int64_t ih = strides[0] * oh, iw = strides[1] * ow;
            float *dst = output + ow + outW * (oh + outH * (co + Cout * n));
            for (int64_t krow = 0; krow < kH; ++krow) {
              const int64_t row = ih + krow - padding[0];
              if (row < 0 || row >= H) continue;
              for (int64_t kcol = 0; kcol < kW; ++kcol) {
                const int64_t col = iw + kcol - padding[1];
                if (col < 0 || col >= W) continue;
                const float prod = input[col + W * (row + H * (ci + Cin * n))] *
                                   kernel[kcol + kW * (krow + kH * (ci + Cin * co))];
                *dst = *dst + prod;
              }
            }
          }
  return 0;
}

/* batch of strided GEMMs: problem b uses A + b*batchStrideA etc.; each one is the reference's
 * gemm_strided (gemm.nim:184-193) */
void oracle_gemm_strided_batched_f32(int64_t batch, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                                     int64_t rsA, int64_t csA, int64_t bsA, const float *B, int64_t rsB,
                                     int64_t csB, int64_t bsB, float beta, float *C, int64_t rsC, int64_t csC,
                                     int64_t bsC) {
  for (int64_t b = 0; b < batch; ++b)
    oracle_gemm_strided_f32(M, N, K, alpha, A + b * bsA, rsA, csA, B + b * bsB, rsB, csB, beta, C + b * bsC, rsC, csC);
}
