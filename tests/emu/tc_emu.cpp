// tc_emu.cpp -- TEST INFRASTRUCTURE: the tcgen05 GEMM kernel (laser_b200/csrc/gemm_tc.cuh) compiled
// by g++ on top of ptx_emu.h (a functional model of mbarrier / TMA / tcgen05 / TMEM / CTA pairs) and
// launched on host threads with the library's own planning (tc_plan) and the tensor-map parameters
// capi.cu passes to cuTensorMapEncodeTiled (operand_map).  See ptx_emu.h for what this can and
// cannot prove.
#define LB200_HOST_EMULATION 1
#include "cuda_emu.h"

#include "../../laser_b200/csrc/gemm_tc.cuh"

using namespace lb200;

static CUtensorMap make_map(int esz, const void *base, int64_t inner, int64_t outer, int64_t outer_stride_elems,
                            int box_inner, int box_outer) {
  CUtensorMap out;
  std::memset(&out, 0, sizeof out);
  if (!base) return out;
  emu::TensorMap2D m;
  m.magic = emu::kMapMagic;
  m.base = static_cast<const unsigned char *>(base);
  m.dim0 = inner; m.dim1 = outer;
  m.stride1_bytes = outer_stride_elems * esz;
  m.esz = esz; m.box0 = box_inner; m.box1 = box_outer;
  std::memcpy(&out, &m, sizeof m);
  return out;
}
// capi.cu: operand_map -- an operand seen as [mn][k]; K-major: array [mn][ld], MN-major: array [k][ld]
static CUtensorMap operand_map(int esz, const void *base, bool mn_major, int64_t mn, int64_t k, int64_t ld, int block_mn) {
  const int block_k = TC_ROW_BYTES / esz, mn_atom = TC_ROW_BYTES / esz;
  if (!mn_major) return make_map(esz, base, k, mn, ld, block_k, block_mn);
  return make_map(esz, base, mn, k, ld, mn_atom, block_k);
}

struct Args {
  int64_t M, N, K;
  float alpha, beta;
  const void *A[4], *B[4];   // hi, lo (fp32 containers) and xb, lb (bf16), see capi.cu: OperandMaps
  int64_t ldA, ldA_b, ldB, ldB_b;
  void *C;
  int64_t rsC, csC;
  int npass, kc_faithful, raster_g, splitk_enabled, sm_count;
  const float *bias;
  int bias_per_row, act;
  float *splitk_ws;          // k_splits planes of M x round_up(N, 4) when split-K triggers
  int *out_k_splits, *out_grid;
  int f16 = 0;               // 16-bit operands with fp32 output: 0 = bf16 pieces (gemm_tc_kernel), 1 = fp16 pieces (gemm_tc_f16_kernel)
  const uint32_t *amax_a = nullptr, *amax_b = nullptr;   // f16: abs-max words per row of A / column of B (f16_scale.cuh)
};

template <int ESZ, bool A_MN, bool B_MN, typename OutT, bool PAIR>
static int run(const Args &a) {
  TcParams p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.alpha = a.alpha; p.beta = a.beta;
  p.C = a.C; p.rsC = a.rsC; p.csC = a.csC; p.npass = a.npass; p.zero = 0;
  p.epi.bias = a.bias; p.epi.bias_per_row = a.bias_per_row; p.epi.act = a.act;
  tc_plan<ESZ, std::is_same<OutT, float>::value>(p, a.npass, PAIR, TcPlanCfg{a.kc_faithful, a.raster_g, a.splitk_enabled != 0, a.sm_count});
  if (a.out_k_splits) *a.out_k_splits = p.k_splits;
  if (p.k_splits > 1) {   // capi.cu: tc_run -- raw partial sums into the planes, reduced by splitk_reduce_kernel
    if (!a.splitk_ws) return -2;
    const int64_t ld = (a.N + 3) / 4 * 4;
    p.C = a.splitk_ws; p.rsC = ld; p.csC = 1; p.alpha = 1.0f; p.beta = 0.0f; p.epi = Epilogue();
    p.split_plane = a.M * ld;
  }
  const int b_block = PAIR ? TC_BLOCK_N / 2 : TC_BLOCK_N;
  const CUtensorMap mA0 = operand_map(ESZ, a.A[0], A_MN, a.M, a.K, a.ldA, TC_BLOCK_M), mA1 = operand_map(ESZ, a.A[1], A_MN, a.M, a.K, a.ldA, TC_BLOCK_M);
  const CUtensorMap mA2 = operand_map(2, a.A[2], A_MN, a.M, a.K, a.ldA_b, TC_BLOCK_M), mA3 = operand_map(2, a.A[3], A_MN, a.M, a.K, a.ldA_b, TC_BLOCK_M);
  const CUtensorMap mB0 = operand_map(ESZ, a.B[0], B_MN, a.N, a.K, a.ldB, b_block), mB1 = operand_map(ESZ, a.B[1], B_MN, a.N, a.K, a.ldB, b_block);
  const CUtensorMap mB2 = operand_map(2, a.B[2], B_MN, a.N, a.K, a.ldB_b, b_block), mB3 = operand_map(2, a.B[3], B_MN, a.N, a.K, a.ldB_b, b_block);
  // capi.cu: launch_tc -- persistent: one CTA (pair) per SM (pair of SMs), never more than work units
  const int64_t units_total = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks * p.k_splits;
  const int units = PAIR ? a.sm_count / 2 : a.sm_count;
  const int sched = static_cast<int>(units_total < units ? units_total : units);
  const unsigned grid = PAIR ? 2 * sched : sched;
  if (a.out_grid) *a.out_grid = static_cast<int>(grid);
  if (TcCfg<PAIR>::SMEM_BYTES > static_cast<int>(emu::kDynSmemBytes)) return -3;
  emu::reset_state();
  if constexpr (ESZ == 2 && std::is_same<OutT, float>::value) {
    if (a.f16) {   // capi.cu: launch_tc_f16
      TcF16Params pf;
      static_cast<TcParams &>(pf) = p;
      pf.amax_a = a.amax_a; pf.amax_b = a.amax_b;
      emu::launch(grid, TC_THREADS,
                  [=]() { gemm_tc_f16_kernel<2, A_MN, B_MN, float, PAIR>(mA0, mA1, mB0, mB1, mA2, mA3, mB2, mB3, pf); },
                  PAIR ? 2 : 1);
      return 0;
    }
  }
  emu::launch(grid, TC_THREADS,
              [=]() { gemm_tc_kernel<ESZ, A_MN, B_MN, OutT, PAIR>(mA0, mA1, mB0, mB1, mA2, mA3, mB2, mB3, p); },
              PAIR ? 2 : 1);
  return 0;
}

template <int ESZ, typename OutT>
static int dispatch(int a_mn, int b_mn, int pair, const Args &a) {
#define GO(AMN, BMN) (pair ? run<ESZ, AMN, BMN, OutT, true>(a) : run<ESZ, AMN, BMN, OutT, false>(a))
  if (!a_mn && !b_mn) return GO(false, false);
  if (!a_mn && b_mn) return GO(false, true);
  if (a_mn && !b_mn) return GO(true, false);
  return GO(true, true);
#undef GO
}

extern "C" int emu_gemm_tc(int esz, int a_mn, int b_mn, int pair, int64_t M, int64_t N, int64_t K, float alpha, float beta,
                           const void *A0, const void *A1, const void *A2, const void *A3, int64_t ldA, int64_t ldA_b,
                           const void *B0, const void *B1, const void *B2, const void *B3, int64_t ldB, int64_t ldB_b,
                           void *C, int64_t rsC, int64_t csC, int npass, int kc_faithful, int raster_g, int splitk_enabled,
                           int sm_count, const float *bias, int bias_per_row, int act, float *splitk_ws, int *out_k_splits,
                           int *out_grid) {
  Args a{M, N, K, alpha, beta, {A0, A1, A2, A3}, {B0, B1, B2, B3}, ldA, ldA_b, ldB, ldB_b, C, rsC, csC, npass, kc_faithful,
         raster_g, splitk_enabled, sm_count, bias, bias_per_row, act, splitk_ws, out_k_splits, out_grid};
  if (esz == 4) return dispatch<4, float>(a_mn, b_mn, pair, a);
  if (esz == 2) return dispatch<2, uint16_t>(a_mn, b_mn, pair, a);
  return -1;
}

// the two-piece fp32 modes (capi.cu: PATH_BF16X3 / PATH_F16X3): 16-bit hi / lo arrays per operand, three passes, fp32 output
extern "C" int emu_gemm_tc16x3(int f16, int a_mn, int b_mn, int pair, int64_t M, int64_t N, int64_t K, float alpha, float beta,
                               const void *A0, const void *A1, int64_t ldA, const void *B0, const void *B1, int64_t ldB,
                               void *C, int64_t rsC, int64_t csC, int kc_faithful, int raster_g, int splitk_enabled, int sm_count,
                               const float *bias, int bias_per_row, int act, float *splitk_ws, int *out_k_splits, int *out_grid,
                               const uint32_t *amax_a, const uint32_t *amax_b) {
  Args a{M, N, K, alpha, beta, {A0, A1, nullptr, nullptr}, {B0, B1, nullptr, nullptr}, ldA, ldA, ldB, ldB, C, rsC, csC, 3, kc_faithful,
         raster_g, splitk_enabled, sm_count, bias, bias_per_row, act, splitk_ws, out_k_splits, out_grid};
  a.f16 = f16; a.amax_a = amax_a; a.amax_b = amax_b;
  return dispatch<2, float>(a_mn, b_mn, pair, a);
}
