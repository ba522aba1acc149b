// f16_scale.cuh -- range handling of the opt-in fp32 mode LASER_B200_PATH_F16X3 (shared by split.cuh, which
// scales and splits the operands, and by gemm_tc_f16_kernel, whose epilogue undoes the scales).
//
// fp16 has 11 significant bits (bf16: 8) but only 5 exponent bits, so every row of A and every column of B (an
// "mn index": the vectors that meet in one output element) is first multiplied by its own power of two 2^s, which
// puts its largest finite |x| into [2^14, 2^15), the top binade below fp16's maximum 65504.  s is derived ON THE
// DEVICE from the abs-max word an earlier kernel of the same stream produced for that mn index (no host
// synchronisation), by the split kernel and again by the GEMM epilogue, which multiplies output (i, j) by
// 2^-sA[i] * 2^-sB[j].  Elements down to 2^-17 of their row's / column's maximum keep all 22 bits of the two pieces
// (the low piece, <= 2^-11 of the element, is then still rounded at or above fp16's subnormal spacing 2^-24);
// smaller ones keep an ABSOLUTE precision of 2^-39 of that maximum -- the usual row/column-norm error model of a
// blocked GEMM.  See DESIGN.md.
#pragma once

#include <stdint.h>
#ifdef LB200_HOST_EMULATION
#include <cmath>
#include <limits>
#endif

namespace lb200 {

// unbiased exponent s of the scale for a row / column whose largest finite |x| has fp32 bits `absmax_bits`
__device__ __forceinline__ int f16x2_scale_exp(uint32_t absmax_bits) {
  const int e = static_cast<int>(absmax_bits >> 23);  // biased exponent (the sign bit is clear)
  if (e == 0) return 0;                               // all zero / subnormal: leave as is
  int s = 14 - (e - 127);
  if (s > 126) s = 126;                               // both 2^s and 2^-s must be normal fp32 numbers
  if (s < -126) s = -126;
  return s;
}
__device__ __forceinline__ float f16x2_pow2(int s) { return __uint_as_float(static_cast<uint32_t>(127 + s) << 23); }
__device__ __forceinline__ float f16x2_scale(uint32_t absmax_bits) { return f16x2_pow2(f16x2_scale_exp(absmax_bits)); }
__device__ __forceinline__ float f16x2_unscale(uint32_t absmax_bits) { return f16x2_pow2(-f16x2_scale_exp(absmax_bits)); }

#ifndef LB200_HOST_EMULATION
__device__ __forceinline__ uint16_t f16_rn_bits(float x) {
  uint16_t h;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
__device__ __forceinline__ float f16_bits_to_f32(uint16_t h) {
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
  return f;
}
#else   // tests/emu: IEEE binary16, round to nearest even, subnormals kept, overflow to infinity
inline uint16_t f16_rn_bits(float x) {
  const uint32_t u = __float_as_uint(x);
  const uint16_t sign = static_cast<uint16_t>((u >> 16) & 0x8000u);
  const uint32_t a = u & 0x7fffffffu;
  if (a > 0x7f800000u) return static_cast<uint16_t>(sign | 0x7e00u);
  if (a == 0x7f800000u) return static_cast<uint16_t>(sign | 0x7c00u);
  const double ax = static_cast<double>(__uint_as_float(a));
  if (ax == 0.0) return sign;
  int e;
  std::frexp(ax, &e);                       // ax = m * 2^e, m in [0.5, 1)
  int E = e - 1;                            // floor(log2 ax)
  if (E < -14) E = -14;                     // subnormal range: fixed quantum 2^-24
  const double q = std::ldexp(1.0, E - 10);
  const double r = std::nearbyint(ax / q) * q;   // default rounding mode: to nearest even
  if (r >= 65520.0) return static_cast<uint16_t>(sign | 0x7c00u);
  if (r < std::ldexp(1.0, -14)) return static_cast<uint16_t>(sign | static_cast<uint16_t>(r / std::ldexp(1.0, -24)));
  int e2;
  const double m = std::frexp(r, &e2);      // r = m * 2^e2, m in [0.5, 1)
  const uint32_t mant = static_cast<uint32_t>(m * 2048.0) - 1024u;   // 10 stored bits
  return static_cast<uint16_t>(sign | (static_cast<uint32_t>(e2 - 1 + 15) << 10) | mant);
}
inline float f16_bits_to_f32(uint16_t h) {
  const int e = (h >> 10) & 31;
  const int m = h & 1023;
  double v;
  if (e == 0) v = std::ldexp(static_cast<double>(m), -24);
  else if (e == 31) v = m ? std::numeric_limits<double>::quiet_NaN() : std::numeric_limits<double>::infinity();
  else v = std::ldexp(static_cast<double>(1024 + m), e - 25);
  return static_cast<float>((h & 0x8000) ? -v : v);
}
#endif

// The two fp16 pieces of TWO already scaled values, packed (element 0 in the low half): h = fp16(x), l = fp16(x - h) -- the
// difference is exact in fp32 -- and l = 0 when h is not finite (x = +-inf / NaN: x - h would be NaN).  The conversion pipe
// (XU, a quarter of the FP32 rate) is what bounds the preparation kernels, so: ONE packed cvt.rn.f16x2.f32 for the two high
// pieces, their fp32 values rebuilt with integer ops (no cvt.f32.f16), one packed cvt for the two low pieces -- one XU
// instruction per element instead of three.
__device__ __forceinline__ float f16_bits_to_f32_alu(uint32_t v) {   // v: 16 bits, finite
  const uint32_t u = v & 0x7fffu;
  const float normal = __uint_as_float((u << 13) + 0x38000000u);                       // (e + 112) << 23 | m << 13
  const float subnormal = __uint_as_float(0x38800000u + (u << 13)) - __uint_as_float(0x38800000u);   // m * 2^-24, exact
  const float mag = (u < 0x0400u) ? subnormal : normal;
  return (v & 0x8000u) ? -mag : mag;
}
__device__ __forceinline__ void f16x2_pieces2(float x0, float x1, uint32_t &h2, uint32_t &l2) {
#ifndef LB200_HOST_EMULATION
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(x1), "f"(x0));   // d = {upper: first source, lower: second source}
  const uint32_t v0 = h2 & 0xffffu, v1 = h2 >> 16;
  float d0 = __fsub_rn(x0, f16_bits_to_f32_alu(v0)), d1 = __fsub_rn(x1, f16_bits_to_f32_alu(v1));
  if ((v0 & 0x7c00u) == 0x7c00u) d0 = 0.0f;
  if ((v1 & 0x7c00u) == 0x7c00u) d1 = 0.0f;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(d1), "f"(d0));
#else
  const uint16_t a0 = f16_rn_bits(x0), a1 = f16_rn_bits(x1);
  const uint16_t b0 = ((a0 & 0x7c00u) == 0x7c00u) ? static_cast<uint16_t>(0) : f16_rn_bits(__fsub_rn(x0, f16_bits_to_f32_alu(a0)));
  const uint16_t b1 = ((a1 & 0x7c00u) == 0x7c00u) ? static_cast<uint16_t>(0) : f16_rn_bits(__fsub_rn(x1, f16_bits_to_f32_alu(a1)));
  h2 = a0 | (static_cast<uint32_t>(a1) << 16);
  l2 = b0 | (static_cast<uint32_t>(b1) << 16);
#endif
}

}  // namespace lb200
