// tc_params.h -- host-visible part of the tensor-core GEMM: tile constants, the kernel's parameter block and the launch
// planning (no device code: capi.cu includes this without instantiating the kernel, see tc_launch.h)
#pragma once

#include <stdint.h>

#if defined(__CUDACC__) || defined(LB200_HOST_EMULATION)
#define LB200_HD __host__ __device__
#else
#define LB200_HD
#endif

namespace lb200 {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_N = 256;
constexpr int TC_ROW_BYTES = 128;  // one swizzle row; BLOCK_K = 128 / sizeof(element)
constexpr int TC_A_TILE_BYTES = TC_BLOCK_M * TC_ROW_BYTES;  // 16 KB: 128 rows x 128 B
template <int NPASS, bool PAIR> struct TcCfg {
  static constexpr int PIECES = (NPASS == 3) ? 2 : 1;
  static constexpr int B_COLS = PAIR ? TC_BLOCK_N / 2 : TC_BLOCK_N;
  static constexpr int B_TILE_BYTES = B_COLS * TC_ROW_BYTES;
  static constexpr int A_STAGE_BYTES = PIECES * TC_A_TILE_BYTES;
  static constexpr int B_STAGE_BYTES = PIECES * B_TILE_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // 32 / 48 KB (1 pass), 64 / 96 KB (3 passes)
#ifdef TC_STAGES_OVERRIDE
  static constexpr int STAGES = TC_STAGES_OVERRIDE;
#else
  static constexpr int STAGES = (NPASS == 3) ? (PAIR ? 3 : 2) : (PAIR ? 6 : 4);   // 192 KB of tiles in every variant
#endif
  // after the operand stages: one 4 KB staging buffer per epilogue warp (32 rows x 128 B, 128B-swizzled) from which the
  // tile leaves through cp.async.bulk.tensor stores, then the barriers
  static constexpr int STORE_STAGING_BYTES = 8 * 32 * 128;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_STAGING_BYTES + 1024 /*align*/ + 512 /*barriers, scheduler slot*/;
};
constexpr int TC_ACC_STAGES = 2;
constexpr int TC_TMEM_COLS = TC_ACC_STAGES * TC_BLOCK_N;  // 512: all of TMEM
// warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 tile scheduler | warps 4-11: epilogue (2 warpgroups)
constexpr int TC_THREADS = 384;
constexpr int TC_EPI_THREADS = 256;
constexpr int TC_EPI_WARPS = TC_EPI_THREADS / 32;
constexpr int TC_EPI_COLS = TC_BLOCK_N / 2;  // columns owned by one epilogue thread
constexpr int TC_REGS_CTRL = 56;   // setmaxnreg for the producer/MMA warpgroup
constexpr int TC_REGS_EPI = 224;   // ... and for the epilogue warpgroups (running sums): 4 x 32 x 56 + 8 x 32 x 224 = the 64512 of the launch

// fused epilogue: v -> act(v + bias)   (gemm.nim:196 "elementwise epilogue fusion")
struct Epilogue {
  const float *bias = nullptr;
  int bias_per_row = 0;
  int act = 0;  // 0 none, 1 relu, 2 tanh, 3 sigmoid
};
struct TcParams {
  int64_t M, N, K;
  float alpha, beta;
  void *C;
  int64_t rsC, csC;
  int kb_per_block;   // k-tiles per TMEM accumulation block (>= 1)
  uint32_t zero;      // always 0; opaque to the compiler (see the epilogue)
  int raster_g;       // m-blocks per raster group (see tile_coords)
  Epilogue epi;
  // split-K.  Tiles [0, n_direct) (raster order, tile_coords) are computed over all of K and stored through the epilogue
  // into C.  Each of the remaining `num tiles - n_direct` tiles is cut in k_splits K-ranges of kb_per_split k-tiles; unit
  // n_direct + i * k_splits + s computes range s of tile n_direct + i and writes its raw partial sums (scales undone, no
  // alpha / beta / bias / activation) to the tile-local plane [s][i] of split_ws; splitk_tail_reduce_kernel then adds the
  // planes in order and applies alpha / beta / epilogue.  Two uses: few output tiles and a long K (n_direct = 0: every
  // tile is split), and the partial last wave of a persistent launch (n_direct = the tiles of the full waves: the
  // remainder of 4096^3's 256 pair-tiles on 74 SM pairs runs as 68 half-K units instead of 34 full ones: 3.5 waves, not 4)
  int n_direct;          // == num_m_blocks * num_n_blocks when nothing is split
  int k_splits;          // >= 1
  int kb_per_split;      // k-tiles per split
  float *split_ws = nullptr;   // [k_splits][split tiles][tile rows][TC_BLOCK_N] fp32
  int num_m_blocks, num_n_blocks;  // output tiles: 128 x 256, or 256 x 256 per CTA pair
  // fp32 C with unit column stride and 16-byte aligned rows leaves through TMA (mapC of the launch: box 32 columns x 32 rows,
  // 128B swizzle; rows / columns past M / N are clipped by the copy engine); otherwise, and for split units, plain stores
  int c_tma = 0;
  // SCALED: fp32 bits of the largest finite |a| of row i of A / |b| of column j of B (f16_scale.cuh)
  const uint32_t *amax_a = nullptr, *amax_b = nullptr;
  // tile scheduler: word 0 = next unit (atomicAdd), word 1 = pairs that have drawn their last unit (the last one zeroes
  // both words for the next launch that uses this slot).  nullptr: static round-robin (unit = pair index + i * pairs)
  unsigned int *sched = nullptr;
};

// raster order of the output tiles: groups of G m-blocks sweep n together, so that the concurrently resident tiles (148 of
// 128 x 256, or 74 pairs of 256 x 256) cover a near-square patch, which minimises the A + B panels one wave pulls through
// L2 (G = 16 single-CTA tiles / 8 pair tiles = 2048 rows)
LB200_HD inline void tile_coords(int t, int num_m, int num_n, int G, int &mb, int &nb) {
  const int per_group = G * num_n;
  const int g = t / per_group;
  const int first_m = g * G;
  const int gsz = (G < num_m - first_m) ? G : (num_m - first_m);
  const int r = t - g * per_group;
  mb = first_m + r % gsz;
  nb = r / gsz;
}

// host side: the part of TcParams that depends only on the problem (p.M, p.N, p.K set by the
// caller) and on the configuration
struct TcPlanCfg {
  int kc_faithful;      // K extent per TMEM accumulation block in the fp32-faithful modes
  int raster_g;         // 0 = default
  bool splitk_enabled;
  int sm_count;
  int tail_min_k = 2048;   // shortest K for which the remainder behind full waves is split
};
template <int ESZ, bool OUT_F32>
inline void tc_plan(TcParams &p, int npass, bool pair, const TcPlanCfg &cfg) {
  const int block_k = TC_ROW_BYTES / ESZ;  // k-tile
  const int num_kb = static_cast<int>((p.K + block_k - 1) / block_k);
  {
    // K extent accumulated inside the tensor core before the epilogue warps add the block
    // to their fp32 running sums (the analogue of the reference's kc, gemm_tiling.nim:310).
    // Only the fp32-faithful modes need short chains.
    const int kc = (npass == 3) ? cfg.kc_faithful : 0;
    p.kb_per_block = (kc > 0) ? (kc + block_k - 1) / block_k : num_kb;
    if (p.kb_per_block < 1) p.kb_per_block = 1;
    if (p.kb_per_block > num_kb) p.kb_per_block = num_kb;
  }
  p.raster_g = cfg.raster_g > 0 ? cfg.raster_g : (pair ? 8 : 16);
  const int tile_m = pair ? 2 * TC_BLOCK_M : TC_BLOCK_M;
  p.num_m_blocks = static_cast<int>((p.M + tile_m - 1) / tile_m);
  p.num_n_blocks = static_cast<int>((p.N + TC_BLOCK_N - 1) / TC_BLOCK_N);
  // ---- split-K (fp32 output only): too few output tiles to fill the machine and a long K, or a thin last wave ----
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;
  p.k_splits = 1;
  p.kb_per_split = num_kb;
  p.n_direct = static_cast<int>(tiles);
  p.split_ws = nullptr;
  const int units = pair ? cfg.sm_count / 2 : cfg.sm_count;
  if constexpr (OUT_F32) {
    const int blocks = (num_kb + p.kb_per_block - 1) / p.kb_per_block;   // accumulation blocks along K
    // every split keeps >= 512 K-elements
    const int min_tiles = 512 / block_k;
    const int min_blocks = (min_tiles + p.kb_per_block - 1) / p.kb_per_block;
    const int max_s = blocks / min_blocks < 16 ? blocks / min_blocks : 16;
    // tiles that do not fill a wave of their own: all of them when there are fewer than `units`, else the remainder
    const int64_t rem = tiles < units ? tiles : tiles % units;
    int S = rem > 0 ? static_cast<int>(units / rem) : 1;
    if (S > max_s) S = max_s;
    // a remainder behind full waves (S >= 2: at most half a wave of tiles) is split when a tile is long enough for half
    // of it to outweigh the reduce kernel
    const bool worth = S >= 2 && (tiles < units || p.K >= cfg.tail_min_k);
    if (cfg.splitk_enabled && worth && units > 0) {
      const int blocks_per_split = (blocks + S - 1) / S;
      p.k_splits = (blocks + blocks_per_split - 1) / blocks_per_split;
      p.kb_per_split = blocks_per_split * p.kb_per_block;
      if (p.k_splits >= 2) p.n_direct = static_cast<int>(tiles - rem);
      else { p.k_splits = 1; p.kb_per_split = num_kb; }
    }
  }
}
// work units of a planned launch, and the fp32 words of its split-K workspace (0: nothing is split)
inline int64_t tc_units(const TcParams &p) {
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;
  return p.n_direct + (tiles - p.n_direct) * p.k_splits;
}
inline int64_t tc_split_ws_floats(const TcParams &p, bool pair) {
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;
  const int64_t tile_m = pair ? 2 * TC_BLOCK_M : TC_BLOCK_M;
  return (tiles - p.n_direct) * p.k_splits * tile_m * TC_BLOCK_N;
}

}  // namespace lb200
