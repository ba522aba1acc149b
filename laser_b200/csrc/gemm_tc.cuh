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
#pragma once

#include <type_traits>

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

// ESZ: element size of A/B in bytes (4 = tf32 containers, 2 = bf16).
// OutT: float or uint16_t (bf16 bits).
// PAIR: launched as clusters of 2 CTAs; CTA rank r owns rows [128r, 128r+128) of the 256-row tile.
template <int ESZ, bool A_MN, bool B_MN, typename OutT, bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
               const __grid_constant__ CUtensorMap mapB2, const __grid_constant__ CUtensorMap mapB3,
               const TcParams p) {
  // K extent of one scheduling unit: a 32-element k-tile (tf32 input), a 64-element k-tile
  // (bf16 input) or, in the mixed mode, a 64-element group = 2 bf16 correction tiles + 2 tf32 tiles
  const bool mixed = (ESZ == 4) && (p.npass == 2);
  const int unit_k = mixed ? 64 : TC_ROW_BYTES / ESZ;
  using Cfg = TcCfg<PAIR>;
  constexpr int TC_STAGES = Cfg::STAGES;
  constexpr int TC_B_STAGE_BYTES = Cfg::B_STAGE_BYTES;
  constexpr int TC_STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int TILE_M = PAIR ? 2 * TC_BLOCK_M : TC_BLOCK_M;  // rows of one scheduled tile
  const uint32_t cta_rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const int sched_id = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int sched_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  LB200_DYN_SMEM(uint8_t, smem_raw);
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + TC_STAGES * TC_A_STAGE_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t *full_bar = bars;                        // [TC_STAGES]
  uint64_t *empty_bar = bars + TC_STAGES;           // [TC_STAGES]
  uint64_t *tmem_full = bars + 2 * TC_STAGES;       // [TC_ACC_STAGES]
  uint64_t *tmem_empty = tmem_full + TC_ACC_STAGES; // [TC_ACC_STAGES]
  uint32_t *tmem_base_smem = reinterpret_cast<uint32_t *>(tmem_empty + TC_ACC_STAGES);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  const int num_kb = static_cast<int>((p.K + unit_k - 1) / unit_k);       // scheduling units along K
  const int num_units = num_tiles * p.k_splits;  // work units of the persistent scheduler
  // K range [kb_lo, kb_hi) of split sp, in scheduling units
  auto split_range = [&](int sp, int &kb_lo, int &kb_hi) {
    kb_lo = min(num_kb, sp * p.kb_per_split);
    kb_hi = min(num_kb, kb_lo + p.kb_per_split);
  };

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&mapA0);
    ptx::prefetch_tensormap(&mapB0);
    if (p.npass == 3) {
      ptx::prefetch_tensormap(&mapA1);
      ptx::prefetch_tensormap(&mapB1);
    }
    if (mixed) {
      ptx::prefetch_tensormap(&mapA2);
      ptx::prefetch_tensormap(&mapA3);
      ptx::prefetch_tensormap(&mapB2);
      ptx::prefetch_tensormap(&mapB3);
    }
  }
  if (threadIdx.x == 32) {
    for (int i = 0; i < TC_STAGES; ++i) {
      // pair: the leader's full barrier takes its own arrive.expect_tx plus the peer's arrive
      ptx::mbar_init(&full_bar[i], PAIR ? 2 : 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < TC_ACC_STAGES; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      // pair: the epilogue threads of BOTH CTAs release the leader's accumulator stage
      ptx::mbar_init(&tmem_empty[i], PAIR ? 2 * TC_EPI_THREADS : TC_EPI_THREADS);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 2) {
    if constexpr (PAIR) ptx::tmem_alloc_pair<TC_TMEM_COLS>(tmem_base_smem);
    else ptx::tmem_alloc<TC_TMEM_COLS>(tmem_base_smem);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();                          // CTA-level: barrier inits + TMEM base visible to all warps
  if constexpr (PAIR) ptx::cluster_sync();  // peer barriers must exist before any remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp_idx < 4) {
    ptx::setmaxnreg_dec<TC_REGS_CTRL>();  // hand registers to the epilogue warpgroups
    if (warp_idx == 0 && lane == 0) {
      // ===================== TMA producer (one thread) =====================
      int stage = 0;
      uint32_t phase = 0;
      auto tma = [&](void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
        if constexpr (PAIR) ptx::tma_load_2d_pair(dst, m, bar, c0, c1);  // bytes -> leader's barrier
        else ptx::tma_load_2d(dst, m, bar, c0, c1);
      };
      // E = element size of the tiles of THIS stage (4: fp32/tf32, 2: bf16)
      auto load_stage = [&](auto esz_tag, const CUtensorMap *ma, const CUtensorMap *mbp, int m0, int n0, int k0) {
        constexpr int E = decltype(esz_tag)::value;
        constexpr int BLOCK_K = TC_ROW_BYTES / E;             // 32 or 64 k-elements per tile
        [[maybe_unused]] constexpr int MN_ATOM = TC_ROW_BYTES / E;             // elements per 128-byte MN chunk
        [[maybe_unused]] constexpr int MN_BOX_BYTES = BLOCK_K * TC_ROW_BYTES;  // one [BLOCK_K][128 B] TMA box
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if constexpr (PAIR) {
          if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * TC_STAGE_BYTES);
          else ptx::mbar_arrive_leader(&full_bar[stage]);
        } else {
          ptx::mbar_arrive_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
        }
        uint8_t *sa = smem_a + stage * TC_A_STAGE_BYTES;
        uint8_t *sb = smem_b + stage * TC_B_STAGE_BYTES;
        if constexpr (!A_MN) {
          tma(sa, ma, &full_bar[stage], k0, m0);  // box {BLOCK_K, 128}
        } else {
#pragma unroll
          for (int c = 0; c < TC_BLOCK_M / MN_ATOM; ++c)  // boxes {MN_ATOM, BLOCK_K}
            tma(sa + c * MN_BOX_BYTES, ma, &full_bar[stage], m0 + c * MN_ATOM, k0);
        }
        if constexpr (!B_MN) {
          tma(sb, mbp, &full_bar[stage], k0, n0);  // box {BLOCK_K, B_COLS}
        } else {
#pragma unroll
          for (int c = 0; c < Cfg::B_COLS / MN_ATOM; ++c)
            tma(sb + c * MN_BOX_BYTES, mbp, &full_bar[stage], n0 + c * MN_ATOM, k0);
        }
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      };
      using E_in = std::integral_constant<int, ESZ>;
      using E_bf = std::integral_constant<int, 2>;
      for (int u = sched_id; u < num_units; u += sched_stride) {
        const int t = u / p.k_splits;
        int mb, nb, kb_lo, kb_hi;
        tile_coords(t, p.num_m_blocks, p.num_n_blocks, p.raster_g, mb, nb);
        split_range(u - t * p.k_splits, kb_lo, kb_hi);
        // pair: this CTA's 128 rows of A and its half of the B columns
        const int m0 = mb * TILE_M + static_cast<int>(cta_rank) * TC_BLOCK_M;
        const int n0 = nb * TC_BLOCK_N + static_cast<int>(cta_rank) * (TC_BLOCK_N - Cfg::B_COLS);
        for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += p.kb_per_block) {
          const int kb1 = min(kb_hi, kb0 + p.kb_per_block);
          // the small cross terms first (the accumulator is still small, so its truncation
          // does not touch them), then the hi*hi chain
          if (p.npass == 3) {
            for (int kb = kb0; kb < kb1; ++kb) {
              load_stage(E_in{}, &mapA0, &mapB1, m0, n0, kb * unit_k);  // A_hi * B_lo
              load_stage(E_in{}, &mapA1, &mapB0, m0, n0, kb * unit_k);  // A_lo * B_hi
            }
          }
          if constexpr (ESZ == 4) {
            if (mixed) {
              for (int kb = kb0; kb < kb1; ++kb) {
                load_stage(E_bf{}, &mapA2, &mapB3, m0, n0, kb * 64);  // bf16(A) * bf16(B_lo)
                load_stage(E_bf{}, &mapA3, &mapB2, m0, n0, kb * 64);  // bf16(A_lo) * bf16(B)
              }
              for (int kb = kb0; kb < kb1; ++kb) {
                load_stage(E_in{}, &mapA0, &mapB0, m0, n0, kb * 64);
                if (kb * 64 + 32 < p.K) load_stage(E_in{}, &mapA0, &mapB0, m0, n0, kb * 64 + 32);
              }
              continue;
            }
          }
          for (int kb = kb0; kb < kb1; ++kb) load_stage(E_in{}, &mapA0, &mapB0, m0, n0, kb * unit_k);
        }
      }
    } else if (warp_idx == 1 && lane == 0 && cta_rank == 0) {
      // ===================== MMA issuer (one thread; pair: the leader CTA only) =============
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      uint32_t d_tmem = 0;
      bool fresh = true;  // next MMA overwrites the accumulator (start of an accumulation block)
      auto mma_stage = [&](auto esz_tag) {
        constexpr int E = decltype(esz_tag)::value;
        constexpr int BLOCK_K = TC_ROW_BYTES / E;
        constexpr int UMMA_K = 32 / E;                        // 8 or 16 elements = 32 bytes
        constexpr int K_STEPS = BLOCK_K / UMMA_K;             // 4
        constexpr int MN_BOX_BYTES = BLOCK_K * TC_ROW_BYTES;
        // MN-major 32-bit operands must use the 128B-swizzle-with-32B-atoms layout (4 k-rows per atom)
        constexpr uint32_t MN_LAYOUT = E == 4 ? ptx::kLayoutSw128Base32 : ptx::kLayoutSw128;
        constexpr uint32_t MN_SBO = E == 4 ? 512 : 1024;
        constexpr uint32_t IDESC = ptx::make_idesc(E == 4 ? ptx::kFmtTF32 : ptx::kFmtBF16, A_MN ? 1 : 0,
                                                   B_MN ? 1 : 0, TILE_M, TC_BLOCK_N);
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(smem_a + stage * TC_A_STAGE_BYTES);
        const uint32_t b_addr = ptx::smem_u32(smem_b + stage * TC_B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < K_STEPS; ++k) {
          // K-major: step 32 bytes inside the 128-byte swizzle row.
          // MN-major: step UMMA_K k-rows of 128 bytes.
          const uint64_t ad =
              A_MN ? ptx::make_smem_desc(a_addr + k * UMMA_K * TC_ROW_BYTES, MN_BOX_BYTES, MN_SBO, MN_LAYOUT)
                   : ptx::make_smem_desc(a_addr + k * 32, 0, 1024, ptx::kLayoutSw128);
          const uint64_t bd =
              B_MN ? ptx::make_smem_desc(b_addr + k * UMMA_K * TC_ROW_BYTES, MN_BOX_BYTES, MN_SBO, MN_LAYOUT)
                   : ptx::make_smem_desc(b_addr + k * 32, 0, 1024, ptx::kLayoutSw128);
          const uint32_t accum = (fresh && k == 0) ? 0u : 1u;
          if constexpr (PAIR) {
            if constexpr (E == 4) ptx::mma_tf32_ss_pair(d_tmem, ad, bd, IDESC, accum);
            else ptx::mma_f16_ss_pair(d_tmem, ad, bd, IDESC, accum);
          } else {
            if constexpr (E == 4) ptx::mma_tf32_ss(d_tmem, ad, bd, IDESC, accum);
            else ptx::mma_f16_ss(d_tmem, ad, bd, IDESC, accum);
          }
        }
        fresh = false;
        // frees the smem slot (in both CTAs of a pair) when these MMAs retire
        if constexpr (PAIR) ptx::mma_commit_pair(&empty_bar[stage]);
        else ptx::mma_commit(&empty_bar[stage]);
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      };
      using E_in = std::integral_constant<int, ESZ>;
      using E_bf = std::integral_constant<int, 2>;
      for (int u = sched_id; u < num_units; u += sched_stride) {
        int kb_lo, kb_hi;
        split_range(u % p.k_splits, kb_lo, kb_hi);
        for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += p.kb_per_block) {
          const int kb1 = min(kb_hi, kb0 + p.kb_per_block);
          ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          ptx::tc_fence_after_sync();
          d_tmem = tmem_base + acc * TC_BLOCK_N;
          fresh = true;
          bool done = false;
          if (p.npass == 3)
            for (int kb = kb0; kb < kb1; ++kb) { mma_stage(E_in{}); mma_stage(E_in{}); }
          if constexpr (ESZ == 4) {
            if (mixed) {
              for (int kb = kb0; kb < kb1; ++kb) { mma_stage(E_bf{}); mma_stage(E_bf{}); }
              for (int kb = kb0; kb < kb1; ++kb) {
                mma_stage(E_in{});
                if (kb * 64 + 32 < p.K) mma_stage(E_in{});
              }
              done = true;
            }
          }
          if (!done)
            for (int kb = kb0; kb < kb1; ++kb) mma_stage(E_in{});
          // block complete -> the epilogue warps (of both CTAs) drain it
          if constexpr (PAIR) ptx::mma_commit_pair(&tmem_full[acc]);
          else ptx::mma_commit(&tmem_full[acc]);
          if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ============ epilogue: 8 warps; warp w owns TMEM lanes 32*(w%4).. and 128 columns ============
    ptx::setmaxnreg_inc<TC_REGS_EPI>();
    const int q = warp_idx & 3;          // TMEM lane quarter this warp may read
    const int h = (warp_idx - 4) >> 2;   // column half
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (p.csC == 1) && ((p.rsC * sizeof(OutT)) % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        ((p.split_plane * sizeof(OutT)) % 16 == 0);
    for (int u = sched_id; u < num_units; u += sched_stride) {
      const int t = u / p.k_splits;
      int mb, nb, kb_lo, kb_hi;
      tile_coords(t, p.num_m_blocks, p.num_n_blocks, p.raster_g, mb, nb);
      split_range(u - t * p.k_splits, kb_lo, kb_hi);
      const int num_blocks = (kb_hi - kb_lo + p.kb_per_block - 1) / p.kb_per_block;  // accumulation blocks
      OutT *__restrict__ C = reinterpret_cast<OutT *>(p.C) + (u - t * p.k_splits) * p.split_plane;
      const int64_t row = static_cast<int64_t>(mb) * TILE_M + cta_rank * TC_BLOCK_M + q * 32 + lane;
      const int64_t col0 = static_cast<int64_t>(nb) * TC_BLOCK_N + h * TC_EPI_COLS;
      if (p.beta != 0.0f && row < p.M && col0 < p.N && p.csC == 1) {
        // beta != 0: pull this thread's 512 bytes of old C into L2 now; they are needed only
        // after the whole K loop of the tile, so the latency is free
        const OutT *cp = C + row * p.rsC + col0;
#pragma unroll
        for (int l = 0; l < TC_EPI_COLS * static_cast<int>(sizeof(OutT)) / 128; ++l) {
          if (col0 + l * (128 / static_cast<int>(sizeof(OutT))) < p.N)
            ptx::prefetch_l2(cp + l * (128 / sizeof(OutT)));
        }
      }
      float run[TC_EPI_COLS];  // running sums of this thread's row segment (registers)
#pragma unroll
      for (int j = 0; j < TC_EPI_COLS; ++j) run[j] = 0.0f;
      for (int blk = 0; blk < num_blocks; ++blk) {
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        ptx::tc_fence_after_sync();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * TC_BLOCK_N + h * TC_EPI_COLS;
        // Eight 16-column chunks.  `dep` (always 0 at run time: p.zero is 0, but the compiler
        // cannot know) makes the address of chunk c+1 depend on an addition of chunk c, so the
        // scheduler cannot issue all eight loads first and keep 128 extra registers in flight.
        uint32_t dep = 0;
#pragma unroll
        for (int c = 0; c < TC_EPI_COLS / 16; ++c) {
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(t_addr + c * 16 + dep, r);
          ptx::tmem_ld_wait(r);
#pragma unroll
          for (int j = 0; j < 16; ++j) run[c * 16 + j] = __fadd_rn(run[c * 16 + j], __uint_as_float(r[j]));
          dep = (__float_as_uint(run[c * 16]) | __float_as_uint(run[c * 16 + 15])) & p.zero;
        }
        // this thread's TMEM reads of the block are done: hand the stage back to the MMA thread
        ptx::tc_fence_before_sync();
        if constexpr (PAIR) ptx::mbar_arrive_leader(&tmem_empty[acc]);
        else ptx::mbar_arrive(&tmem_empty[acc]);
        if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
      // ---- C <- alpha * sum + beta * C  (gemm_ukernel_generic.nim:53-76 semantics) ----
      if (row < p.M && col0 < p.N) {
        OutT *crow = C + row * p.rsC;
        const bool has_epi = (p.epi.bias != nullptr) || (p.epi.act != 0);
        const float row_bias = (p.epi.bias && p.epi.bias_per_row) ? p.epi.bias[row] : 0.0f;
        if (vec_ok && col0 + TC_EPI_COLS <= p.N) {
          if constexpr (sizeof(OutT) == 4) {
            float4 *dst = reinterpret_cast<float4 *>(crow + col0);
            // batches of 4 x 16 B: with beta != 0 the four loads of a batch are in flight
            // together (the old C lines were prefetched into L2 when the tile started)
#pragma unroll
            for (int b8 = 0; b8 < TC_EPI_COLS / 16; ++b8) {
              float4 o[4];
              if (p.beta != 0.0f) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = dst[b8 * 4 + e];
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int v4 = b8 * 4 + e;
                float4 v;
                v.x = p.alpha * run[4 * v4 + 0];
                v.y = p.alpha * run[4 * v4 + 1];
                v.z = p.alpha * run[4 * v4 + 2];
                v.w = p.alpha * run[4 * v4 + 3];
                if (p.beta != 0.0f) {
                  v.x = fmaf(p.beta, o[e].x, v.x);
                  v.y = fmaf(p.beta, o[e].y, v.y);
                  v.z = fmaf(p.beta, o[e].z, v.z);
                  v.w = fmaf(p.beta, o[e].w, v.w);
                }
                if (has_epi) {
                  float4 bv = make_float4(row_bias, row_bias, row_bias, row_bias);
                  if (p.epi.bias && !p.epi.bias_per_row) {
                    const float *bp = p.epi.bias + col0 + 4 * v4;
                    if ((reinterpret_cast<uintptr_t>(bp) & 15) == 0) bv = *reinterpret_cast<const float4 *>(bp);
                    else bv = make_float4(bp[0], bp[1], bp[2], bp[3]);
                  }
                  v.x = epi_act(v.x + bv.x, p.epi.act);
                  v.y = epi_act(v.y + bv.y, p.epi.act);
                  v.z = epi_act(v.z + bv.z, p.epi.act);
                  v.w = epi_act(v.w + bv.w, p.epi.act);
                }
                dst[v4] = v;
              }
            }
          } else {
            uint4 *dst = reinterpret_cast<uint4 *>(crow + col0);
#pragma unroll
            for (int v8 = 0; v8 < TC_EPI_COLS / 8; ++v8) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = p.alpha * run[8 * v8 + e];
              if (p.beta != 0.0f) {
                const uint4 o = dst[v8];
                const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  f[2 * e] = fmaf(p.beta, bf16_bits_to_f32(static_cast<uint16_t>(ow[e] & 0xffff)), f[2 * e]);
                  f[2 * e + 1] = fmaf(p.beta, bf16_bits_to_f32(static_cast<uint16_t>(ow[e] >> 16)), f[2 * e + 1]);
                }
              }
              if (has_epi) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float bv = (p.epi.bias && !p.epi.bias_per_row) ? p.epi.bias[col0 + 8 * v8 + e] : row_bias;
                  f[e] = epi_act(f[e] + bv, p.epi.act);
                }
              }
              uint4 w;
              w.x = f32_to_bf16_bits(f[0]) | (static_cast<uint32_t>(f32_to_bf16_bits(f[1])) << 16);
              w.y = f32_to_bf16_bits(f[2]) | (static_cast<uint32_t>(f32_to_bf16_bits(f[3])) << 16);
              w.z = f32_to_bf16_bits(f[4]) | (static_cast<uint32_t>(f32_to_bf16_bits(f[5])) << 16);
              w.w = f32_to_bf16_bits(f[6]) | (static_cast<uint32_t>(f32_to_bf16_bits(f[7])) << 16);
              dst[v8] = w;
            }
          }
        } else {
          // any C strides / ragged right edge: scalar, predicated; one running pointer so
          // that the unrolled loop does not keep 128 addresses live
          OutT *dst = crow + col0 * p.csC;
          const int64_t ncols = p.N - col0;
#pragma unroll
          for (int j = 0; j < TC_EPI_COLS; ++j) {
            if (j < ncols) {
              float v = p.alpha * run[j];
              if (p.beta != 0.0f) {
                if constexpr (sizeof(OutT) == 4) v = fmaf(p.beta, *dst, v);
                else v = fmaf(p.beta, bf16_bits_to_f32(*dst), v);
              }
              if (has_epi) {
                const float bv = (p.epi.bias && !p.epi.bias_per_row) ? p.epi.bias[col0 + j] : row_bias;
                v = epi_act(v + bv, p.epi.act);
              }
              if constexpr (sizeof(OutT) == 4) *dst = v;
              else *dst = f32_to_bf16_bits(v);
            }
            dst += p.csC;
          }
        }
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before_sync();
  if constexpr (PAIR) ptx::cluster_sync();  // neither CTA may leave while its peer still uses its smem/TMEM
  else __syncthreads();
  ptx::tc_fence_after_sync();
  if (warp_idx == 2) {
    if constexpr (PAIR) ptx::tmem_dealloc_pair<TC_TMEM_COLS>(tmem_base);
    else ptx::tmem_dealloc<TC_TMEM_COLS>(tmem_base);
  }
}

}  // namespace lb200
