// gemm_tc.cuh -- the tcgen05 / TMEM / TMA strided GEMM for sm_100a.
//
// What it replaces in the reference (mratsim/laser, paths relative to
// laser/primitives/matrix_multiplication/):
//   pack_A_mc_kc / pack_B_kc_nc (gemm_packing.nim:24-94)  -> TMA tensor maps: the
//       copy engine resolves the operand's strides and lands 128-byte-swizzled
//       tiles in shared memory; no packing buffers exist.
//   gebb_ukernel register micro-kernel (gemm_ukernel_generator.nim:140-250)
//       -> tcgen05.mma (kind::tf32 / kind::f16) issued by ONE thread per CTA,
//       accumulators in TMEM (128 lanes x 256 columns fp32 per tile).
//   gemm_impl loop pc (gemm.nim:150-158: K is cut in kc blocks, every block's partial
//       product is ADDED to C in fp32) -> K is cut in accumulation blocks of `kb_per_block`
//       k-tiles: the tensor core accumulates one block in TMEM, the epilogue warps drain
//       it and add it (IEEE round-to-nearest FADD) to running sums held in registers
//       while the tensor core is already working on the next block in the other TMEM
//       stage.  This matters numerically: the tensor core's own accumulator truncates
//       (measured on B200: ~0.3 ulp of bias per MMA instruction, i.e. 5.5e-5 relative at
//       K = 8192 for positive inputs), so long chains must not live in TMEM.
//   gebp_mkernel loops jr/ir + loop ic (gemm.nim:48-176)
//       -> persistent CTAs (or CTA pairs, cta_group::2) walking 128 x 256 (256 x 256) output
//       tiles; the k loop is a 4- (6-) deep mbarrier ring between the TMA producer thread
//       and the MMA thread; few-tile / long-K problems split K across the idle SMs.
//   epilogues (gemm_ukernel_generic.nim:53-126)
//       -> alpha/beta in fp32 from the running sums; beta == 0 never reads C; optional fused
//       bias + activation (the reference's TODO at gemm.nim:196).
//
// Operand "major-ness" (which of the two strides is 1) is a template parameter:
// UMMA reads K-major and MN-major tiles natively, so A^T*B, A*B^T ... need no
// data movement.  The fp32-faithful modes run, per k-tile, the small cross terms first and
// the hi*hi product last over arrays produced by split.cuh: either three tf32 passes
// (hi*lo, lo*hi, hi*hi) or -- the default -- two bf16 passes for the cross terms at twice
// the rate plus one tf32 pass (npass = 2).
// The bf16 instantiation with fp32 output also serves the opt-in BF16X3 mode (fp32 operands as two bf16 arrays each,
// npass = 3), and gemm_tc_f16_kernel -- the fourth flavour of gemm_tc_kernel.inc -- the opt-in F16X3 mode (two fp16 arrays
// of the power-of-two-scaled operand; the epilogue undoes the scales, f16_scale.cuh).
#pragma once

#include <type_traits>

#include "f16_scale.cuh"
#include "ptx.cuh"

namespace lb200 {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_N = 256;
constexpr int TC_ROW_BYTES = 128;  // one swizzle row; BLOCK_K = 128 / sizeof(element)
constexpr int TC_A_STAGE_BYTES = TC_BLOCK_M * TC_ROW_BYTES;  // 16 KB
#ifndef TC_PAIR_STAGES
#define TC_PAIR_STAGES 6
#endif
// Single-CTA kernel: the CTA stages all 256 B columns (32 KB) -> 48 KB stages, 4 of them.
// CTA-pair kernel (cta_group::2, 256 x 256 tile per pair): each CTA stages its own 128 rows
// of A and HALF of the B columns (16 KB) -> 32 KB stages, 6 of them; the tensor cores of both
// SMs read both halves.
template <bool PAIR> struct TcCfg {
  static constexpr int B_COLS = PAIR ? TC_BLOCK_N / 2 : TC_BLOCK_N;
  static constexpr int B_STAGE_BYTES = B_COLS * TC_ROW_BYTES;
  static constexpr int STAGE_BYTES = TC_A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = PAIR ? TC_PAIR_STAGES : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};
constexpr int TC_ACC_STAGES = 2;
constexpr int TC_TMEM_COLS = TC_ACC_STAGES * TC_BLOCK_N;  // 512: all of TMEM
// warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle | warps 4-11: epilogue (2 warpgroups)
constexpr int TC_THREADS = 384;
constexpr int TC_EPI_THREADS = 256;
constexpr int TC_EPI_COLS = TC_BLOCK_N / 2;  // columns owned by one epilogue thread
constexpr int TC_REGS_CTRL = 56;   // setmaxnreg for the producer/MMA warpgroup
constexpr int TC_REGS_EPI = 216;   // ... and for the epilogue warpgroups (running sums)

// fused epilogue: v -> act(v + bias)   (gemm.nim:196 "elementwise epilogue fusion")
struct Epilogue {
  const float *bias = nullptr;
  int bias_per_row = 0;
  int act = 0;  // 0 none, 1 relu, 2 tanh, 3 sigmoid
};
__device__ __forceinline__ float epi_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return tanhf(v);
  if (act == 3) return 1.0f / (1.0f + expf(-v));
  return v;
}

struct TcParams {
  int64_t M, N, K;
  float alpha, beta;
  void *C;
  int64_t rsC, csC;
  int npass;          // 1 single pass | 3 tf32 hi/lo split | 2 mixed: tf32 hi*hi + bf16 cross terms
  int kb_per_block;   // k-tiles per TMEM accumulation block (>= 1)
  uint32_t zero;      // always 0; opaque to the compiler (see the epilogue)
  int raster_g;       // m-blocks per raster group (see tile_coords)
  Epilogue epi;
  // split-K (few output tiles, long K): unit u = (tile, split) covers the K range of one split
  // and writes its raw partial sums to plane `split` of a workspace (C points at it, alpha = 1,
  // beta = 0); splitk_reduce_kernel then adds the planes in order and applies alpha/beta/epilogue
  int k_splits;          // >= 1
  int kb_per_split;      // scheduling units (k-tiles / 64-groups) per split
  int64_t split_plane;   // elements between consecutive planes
  int num_m_blocks, num_n_blocks;  // output tiles: 128 x 256, or 256 x 256 per CTA pair
};
// gemm_tc_batched_kernel: `batch` problems of one shape per launch; the tensor maps are 3-d, problem b
// reads matrix b of an operand (matrix 0 if that operand is shared) and writes C + b * bsC.  (A separate
// struct: the parameter block of the measured single-problem kernel must not change size.)
// gemm_tc_hint_kernel: the single-problem kernel with L2 eviction-priority hints (ptx::kEvict*) on the
// A / B tile loads
struct TcHintParams : TcParams {
  uint64_t hint_a = 0, hint_b = 0;
};
// gemm_tc_f16_kernel: fp16 operands (two scaled pieces per fp32 operand); amax_a[i] / amax_b[j] = fp32 bits of the largest
// finite |a| of row i of A / |b| of column j of B, from which the epilogue derives the unscale factors (f16_scale.cuh)
struct TcF16Params : TcParams {
  const uint32_t *amax_a = nullptr, *amax_b = nullptr;
};
struct F16Scales {     // host side: where the two abs-max vectors of the current call live (device memory)
  const uint32_t *a = nullptr, *b = nullptr;
};
struct TcBatchedParams : TcParams {
  int batch = 1;
  int a_shared = 0, b_shared = 0;
  int64_t bsC = 0;
};

// host side: the part of TcParams that depends only on the problem (p.M, p.N, p.K set by the
// caller) and on the configuration
struct TcPlanCfg {
  int kc_faithful;      // K extent per TMEM accumulation block in the fp32-faithful modes
  int raster_g;         // 0 = default
  bool splitk_enabled;
  int sm_count;
};
template <int ESZ, bool OUT_F32>
inline void tc_plan(TcParams &p, int npass, bool pair, const TcPlanCfg &cfg) {
  {
    // K extent accumulated inside the tensor core before the epilogue warps add the block
    // to their fp32 running sums (the analogue of the reference's kc, gemm_tiling.nim:310).
    // Only the fp32-faithful modes need short chains.
    const int block_k = (npass == 2) ? 64 : TC_ROW_BYTES / ESZ;  // scheduling unit along K
    const int num_kb = static_cast<int>((p.K + block_k - 1) / block_k);
    int kc = (npass == 3 || npass == 2) ? cfg.kc_faithful : 0;
    p.kb_per_block = (kc > 0) ? (kc + block_k - 1) / block_k : num_kb;
    if (p.kb_per_block < 1) p.kb_per_block = 1;
    if (p.kb_per_block > num_kb) p.kb_per_block = num_kb;
  }
  p.raster_g = cfg.raster_g > 0 ? cfg.raster_g : (pair ? 8 : 16);
  const int tile_m = pair ? 2 * TC_BLOCK_M : TC_BLOCK_M;
  p.num_m_blocks = static_cast<int>((p.M + tile_m - 1) / tile_m);
  p.num_n_blocks = static_cast<int>((p.N + TC_BLOCK_N - 1) / TC_BLOCK_N);
  // ---- split-K: too few output tiles to fill the machine and a long K (fp32 output only) ----
  p.k_splits = 1;
  p.split_plane = 0;
  {
    const int block_k = (npass == 2) ? 64 : TC_ROW_BYTES / ESZ;
    const int num_kb = static_cast<int>((p.K + block_k - 1) / block_k);
    p.kb_per_split = num_kb;
    const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;
    const int units = pair ? cfg.sm_count / 2 : cfg.sm_count;
    if constexpr (OUT_F32) {
      const int blocks = (num_kb + p.kb_per_block - 1) / p.kb_per_block;   // accumulation blocks along K
      int S = static_cast<int>(units / (tiles > 0 ? tiles : 1));
      if (S > blocks / 4) S = blocks / 4;     // every split keeps >= 4 accumulation blocks (>= 512 K-elements)
      if (S > 16) S = 16;
      if (cfg.splitk_enabled && S >= 2) {
        const int blocks_per_split = (blocks + S - 1) / S;
        p.k_splits = (blocks + blocks_per_split - 1) / blocks_per_split;
        p.kb_per_split = blocks_per_split * p.kb_per_block;
      }
    }
  }
}

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int G, int &mb, int &nb) {
  // groups of G m-blocks sweep n together: the concurrently resident tiles (148 of 128 x 256, or
  // 74 pairs of 256 x 256) then cover a near-square patch, which minimises the A + B panels one
  // wave pulls through L2 (G = 16 single-CTA tiles / 8 pair tiles = 2048 rows)
  const int per_group = G * num_n;
  const int g = t / per_group;
  const int first_m = g * G;
  const int gsz = min(G, num_m - first_m);
  const int r = t - g * per_group;
  mb = first_m + r % gsz;
  nb = r / gsz;
}

__device__ __forceinline__ float bf16_bits_to_f32(uint16_t h) {
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

#define LB200_TC_KERNEL_NAME gemm_tc_kernel
#define LB200_TC_BATCHED 0
#define LB200_TC_HINT 0
#define LB200_TC_F16 0
#define LB200_TC_FMT16 ptx::kFmtBF16
#include "gemm_tc_kernel.inc"
#undef LB200_TC_KERNEL_NAME
#undef LB200_TC_HINT
#define LB200_TC_KERNEL_NAME gemm_tc_hint_kernel
#define LB200_TC_HINT 1
#include "gemm_tc_kernel.inc"
#undef LB200_TC_KERNEL_NAME
#undef LB200_TC_HINT
#undef LB200_TC_BATCHED
#define LB200_TC_KERNEL_NAME gemm_tc_batched_kernel
#define LB200_TC_BATCHED 1
#define LB200_TC_HINT 0
#include "gemm_tc_kernel.inc"
#undef LB200_TC_KERNEL_NAME
#undef LB200_TC_BATCHED
#undef LB200_TC_F16
#undef LB200_TC_FMT16
#define LB200_TC_KERNEL_NAME gemm_tc_f16_kernel
#define LB200_TC_BATCHED 0
#define LB200_TC_F16 1
#define LB200_TC_FMT16 ptx::kFmtF16
#include "gemm_tc_kernel.inc"
#undef LB200_TC_FMT16
#undef LB200_TC_F16
#undef LB200_TC_HINT
#undef LB200_TC_KERNEL_NAME
#undef LB200_TC_BATCHED

}  // namespace lb200
