// capi.cu -- host side of liblaser_b200.so: the C ABI declared in include/laser_b200.h.
//
// Mirrors the host half of the reference's gemm_strided (gemm.nim:184-247): build the
// three matrix views, pick a kernel family (the reference picks an ISA micro-kernel at
// run time, gemm.nim:228-247; here: exact SIMT vs tcgen05), prepare the operands
// (the reference allocates packing Tiles per call, gemm_tiling.nim:312-341; here: TMA
// tensor maps, plus the hi/lo split workspace for the fp32-faithful mode) and launch.
// There is no CPU fallback anywhere in this file.
#include "../../include/laser_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "layers.cuh"
#include "split.cuh"

namespace {

using namespace lb200;

thread_local std::string g_last_error;
thread_local int g_last_path = 0;
thread_local Epilogue g_epi;   // fused epilogue of the call in flight on this thread (default: none)
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_f32_mode{-1};

int set_error(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      cudaGetLastError();                                                                   \
      return set_error(_e == cudaErrorMemoryAllocation ? LASER_B200_ENOMEM : LASER_B200_ECUDA, \
                       "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                       __LINE__);                                                           \
    }                                                                                       \
  } while (0)

struct Buffer {
  void *ptr = nullptr;
  size_t bytes = 0;
};

struct EventPair {
  cudaEvent_t a, b;
  int kind;  // 0 = tensor-core GEMM kernel, 1 = operand preparation kernels
  int launches;
};

// cache of encoded tensor maps (cuTensorMapEncodeTiled costs microseconds; the reference's
// analogue is re-using its Tiles object): keyed by everything the encoding depends on
struct MapKey {
  const void *base;
  int64_t inner, outer, stride;
  int esz, box_inner, box_outer, swz;
  bool operator==(const MapKey &o) const {
    return base == o.base && inner == o.inner && outer == o.outer && stride == o.stride && esz == o.esz &&
           box_inner == o.box_inner && box_outer == o.box_outer && swz == o.swz;
  }
};
struct MapCacheEntry {
  MapKey key;
  CUtensorMap map;
  bool valid = false;
};
constexpr int kMapCacheSize = 64;

struct Ctx {
  MapCacheEntry map_cache[kMapCacheSize];
  int map_cache_next = 0;
  int raster_g = 0;       // env LASER_B200_RASTER (0 = default)
  bool splitk_enabled = true;  // env LASER_B200_SPLITK=0 disables split-K
  int64_t panel_rows = 1024;  // env LASER_B200_PANEL_ROWS: row-panel height of the pipelined host-pointer entry
  int l2_hint = 0;            // env LASER_B200_L2HINT: 0 = the measured kernel (default); 1 A evict_last / B evict_first, 2 the reverse (gemm_tc_hint_kernel, unmeasured)
  bool tc_batched = false;    // env LASER_B200_TC_BATCHED=1: batches of tensor-core problems as one launch (unmeasured)
  bool panel_taper = false;   // env LASER_B200_PANEL_TAPER=1: cut the last row panel finer (shorter PCIe tail)
  bool cta_pair = true;   // env LASER_B200_CTA_PAIR=0 forces the single-CTA kernel
  int kc_faithful = 128;  // env LASER_B200_KC (K extent per TMEM accumulation block)
  bool profiling = false;
  std::vector<EventPair> prof;
  int dev = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;          // compute (and default) stream of the library
  cudaStream_t up = nullptr, down = nullptr;  // H2D / D2H streams of the pipelined host entry
  std::vector<cudaEvent_t> panel_ev;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  Buffer ws[8];      // per operand: hi, lo (fp32) and xb, lb (bf16) -- A then B
  Buffer stage[3];   // device staging of host A, B, C spans
  Buffer splitk;     // split-K partial-sum planes
  Buffer layer_ws;   // im2col workspace of the host-pointer convolution
  Buffer f16s;       // F16X3 mode: fp32 bits of max_k |a| per row of A (words [0, M)) and of max_k |b| per column of B (from
                     // f16_b_off on), written and read on the device
  cudaEvent_t ws_free = nullptr;  // recorded after the last kernel that reads ws[]
  std::mutex mu;       // workspace + tensor-map construction
  std::mutex host_mu;  // staging buffers of the host-pointer entry points
  std::atomic<bool> ready{false};
};

constexpr int kMaxDevices = 32;
Ctx g_ctx[kMaxDevices];
std::mutex g_ctx_mu;

int get_ctx(Ctx **out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(LASER_B200_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  }
  if (dev < 0 || dev >= kMaxDevices) return set_error(LASER_B200_ENODEVICE, "device index %d", dev);
  Ctx &c = g_ctx[dev];
  if (!c.ready) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!c.ready) {
      cudaDeviceProp prop;
      CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
      if (prop.major != 10)
        return set_error(LASER_B200_ENODEVICE,
                         "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)",
                         dev, prop.major, prop.minor);
      c.dev = dev;
      c.sm_count = prop.multiProcessorCount;
      CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaStreamCreateWithFlags(&c.up, cudaStreamNonBlocking));
      CUDA_TRY(cudaStreamCreateWithFlags(&c.down, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&c.ws_free, cudaEventDisableTiming));
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess)
        return set_error(LASER_B200_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
      c.encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
      if (const char *kc = getenv("LASER_B200_KC")) {
        const int v = atoi(kc);
        if (v >= 32) c.kc_faithful = v;
      }
      if (const char *cp = getenv("LASER_B200_CTA_PAIR")) c.cta_pair = atoi(cp) != 0;
      if (const char *rg = getenv("LASER_B200_RASTER")) c.raster_g = atoi(rg);
      if (const char *sk = getenv("LASER_B200_SPLITK")) c.splitk_enabled = atoi(sk) != 0;
      if (const char *pt = getenv("LASER_B200_PANEL_TAPER")) c.panel_taper = atoi(pt) != 0;
      if (const char *tb = getenv("LASER_B200_TC_BATCHED")) c.tc_batched = atoi(tb) != 0;
      if (const char *lh = getenv("LASER_B200_L2HINT")) c.l2_hint = atoi(lh);
      if (const char *pr = getenv("LASER_B200_PANEL_ROWS")) {
        const int64_t v = atoll(pr) / 256 * 256;   // whole CTA-pair tiles
        if (v >= 256) c.panel_rows = v;
      }
      const char *mode = getenv("LASER_B200_F32_MODE");
      if (g_f32_mode.load() < 0) {
        int m = LASER_B200_PATH_TF32_BF16C;
        if (mode) {
          if (!strcmp(mode, "tf32x3")) m = LASER_B200_PATH_TF32X3;
          if (!strcmp(mode, "tf32x1")) m = LASER_B200_PATH_TF32X1;
          else if (!strcmp(mode, "tf32_bf16c")) m = LASER_B200_PATH_TF32_BF16C;
          else if (!strcmp(mode, "bf16x3")) m = LASER_B200_PATH_BF16X3;
          else if (!strcmp(mode, "f16x3")) m = LASER_B200_PATH_F16X3;
          else if (!strcmp(mode, "simt")) m = LASER_B200_PATH_SIMT;
        }
        g_f32_mode.store(m);
      }
      c.ready = true;
    }
  }
  *out = &c;
  return LASER_B200_OK;
}

int ensure(Buffer &b, size_t bytes) {
  if (b.bytes >= bytes) return LASER_B200_OK;
  if (b.ptr) CUDA_TRY(cudaFree(b.ptr));
  b.ptr = nullptr;
  b.bytes = 0;
  size_t want = bytes + (bytes >> 3);  // slack: avoid re-allocation on slightly larger calls
  want = (want + 255) & ~static_cast<size_t>(255);
  CUDA_TRY(cudaMalloc(&b.ptr, want));
  b.bytes = want;
  return LASER_B200_OK;
}

int prof_open(Ctx &c, cudaStream_t s, EventPair *ep, int kind) {
  if (!c.profiling) return LASER_B200_OK;
  ep->kind = kind;
  ep->launches = 0;
  CUDA_TRY(cudaEventCreate(&ep->a));
  CUDA_TRY(cudaEventCreate(&ep->b));
  CUDA_TRY(cudaEventRecord(ep->a, s));
  return LASER_B200_OK;
}
int prof_close(Ctx &c, cudaStream_t s, EventPair *ep, int launches) {
  if (!c.profiling) return LASER_B200_OK;
  CUDA_TRY(cudaEventRecord(ep->b, s));
  ep->launches = launches;
  c.prof.push_back(*ep);
  return LASER_B200_OK;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int grid_for(const Ctx &c, int64_t work_items, int per_sm) {
  int64_t g = static_cast<int64_t>(c.sm_count) * per_sm;
  if (work_items < g) g = work_items > 0 ? work_items : 1;
  return static_cast<int>(g);
}
#define COUNT_LAUNCH() g_launches.fetch_add(1, std::memory_order_relaxed)
#define CHECK_LAUNCH() CUDA_TRY(cudaGetLastError())

// ---------------------------------------------------------------------------------------
//                                   exact SIMT path
// ---------------------------------------------------------------------------------------
template <typename T, int TM, int TN, int BK>
int launch_simt(Ctx &c, int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t rsA,
                int64_t csA, const T *B, int64_t rsB, int64_t csB, T beta, T *C, int64_t rsC,
                int64_t csC, cudaStream_t s, int64_t batch = 1, int64_t bsA = 0, int64_t bsB = 0,
                int64_t bsC = 0) {
  SimtParams<T> p;
  const int64_t tiles = simt_plan<T, TM, TN>(p, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
  if constexpr (std::is_same<T, float>::value) { p.bias = g_epi.bias; p.bias_per_row = g_epi.bias_per_row; p.act = g_epi.act; }
  if (tiles > 0x7fffffff) return set_error(LASER_B200_EINVAL, "too many tiles");
  p.batch = batch; p.bsA = bsA; p.bsB = bsB; p.bsC = bsC;
  const int grid = grid_for(c, tiles * batch, 2);
  if (batch > 1) gemm_simt_batched_kernel<T, TM, TN, BK><<<grid, 256, 0, s>>>(p);
  else gemm_simt_kernel<T, TM, TN, BK><<<grid, 256, 0, s>>>(p);
  COUNT_LAUNCH();
  CHECK_LAUNCH();
  return LASER_B200_OK;
}

template <typename T>
int gemm_simt(Ctx &c, int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t rsA,
              int64_t csA, const T *B, int64_t rsB, int64_t csB, T beta, T *C, int64_t rsC,
              int64_t csC, cudaStream_t s, int64_t batch = 1, int64_t bsA = 0, int64_t bsB = 0,
              int64_t bsC = 0) {
#ifndef LB200_SIMT_BK
#define LB200_SIMT_BK 16
#endif
  if constexpr (sizeof(T) == 4)
    return launch_simt<T, 8, 8, LB200_SIMT_BK>(c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, s, batch,
                                               bsA, bsB, bsC);
  else
    return launch_simt<T, 4, 4, 16>(c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, s, batch, bsA, bsB,
                                    bsC);
}

// ---------------------------------------------------------------------------------------
//                                  tensor-core path
// ---------------------------------------------------------------------------------------
// An operand of the contraction seen as [mn][k]: for A mn = M (s_mn = rowStrideA,
// s_k = colStrideA), for B mn = N (s_mn = colStrideB, s_k = rowStrideB).
struct Operand {
  const void *ptr;
  int64_t mn, k, s_mn, s_k;
};
enum Major { K_MAJOR = 0, MN_MAJOR = 1, GENERAL = 2 };

Major classify(const Operand &o, int esz) {
  const bool aligned = (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0;
  if (!aligned) return GENERAL;
  const int64_t lim = (static_cast<int64_t>(1) << 40) / esz;
  if (o.s_k == 1 && o.s_mn > 0 && (o.s_mn * esz) % 16 == 0 && o.s_mn < lim) return K_MAJOR;
  if (o.s_mn == 1 && o.s_k > 0 && (o.s_k * esz) % 16 == 0 && o.s_k < lim) return MN_MAJOR;
  return GENERAL;
}

int encode_map(Ctx &c, CUtensorMap *map, int esz, const void *base, int64_t inner, int64_t outer,
               int64_t outer_stride_elems, int box_inner, int box_outer, CUtensorMapSwizzle swz) {
  const MapKey key{base, inner, outer, outer_stride_elems, esz, box_inner, box_outer, static_cast<int>(swz)};
  for (int i = 0; i < kMapCacheSize; ++i)
    if (c.map_cache[i].valid && c.map_cache[i].key == key) {
      *map = c.map_cache[i].map;
      return LASER_B200_OK;
    }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(outer_stride_elems) * esz};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = c.encode(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LASER_B200_ECUDA,
                     "cuTensorMapEncodeTiled failed (%d): base=%p inner=%lld outer=%lld stride=%lld box=%dx%d",
                     static_cast<int>(r), base, (long long)inner, (long long)outer,
                     (long long)outer_stride_elems, box_inner, box_outer);
  MapCacheEntry &e = c.map_cache[c.map_cache_next];
  c.map_cache_next = (c.map_cache_next + 1) % kMapCacheSize;
  e.key = key;
  e.map = *map;
  e.valid = true;
  return LASER_B200_OK;
}

// tensor map for one operand given as compact/strided [mn][k] data with the stated major-ness
int operand_map(Ctx &c, CUtensorMap *map, int esz, const void *base, Major major, int64_t mn,
                int64_t k, int64_t ld, int block_mn) {
  const int block_k = TC_ROW_BYTES / esz;
  const int mn_atom = TC_ROW_BYTES / esz;
  if (major == K_MAJOR)
    return encode_map(c, map, esz, base, k, mn, ld, block_k, block_mn, CU_TENSOR_MAP_SWIZZLE_128B);
  // MN-major fp32/tf32 tiles need the 32-byte-atom flavour of the 128B swizzle (see ptx.cuh)
  return encode_map(c, map, esz, base, mn, k, ld, mn_atom, block_k,
                    esz == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

// Tensor maps (and workspace arrays) of one operand for the tensor-core kernel.
//   hi  : the operand itself (single pass) or tf32_rna(x) in fp32 containers
//   lo  : tf32_rna(x - hi)                       (fp32, 3-pass mode)
//   xb  : bf16(x),  lb : bf16(x - hi)            (bf16, mixed mode)
struct OperandMaps {
  CUtensorMap hi, lo, xb, lb;
  bool mn_major = false;
};
struct OperandWs {
  Buffer *hi, *lo, *xb, *lb;
  int64_t amax_off;   // F16X3: word offset of the operand's abs-max vector in Ctx::f16s (set by f16_scales)
};
enum SplitMode { SPLIT_NONE = 0, SPLIT_TF32 = 1, SPLIT_MIXED = 2, SPLIT_BF16X2 = 3 /* fp32 -> two bf16 arrays (xb, lb) */,
                 SPLIT_F16X2 = 4 /* fp32 -> abs-max word + two fp16 arrays of the scaled operand (xb, lb) */ };

template <int ESZ, typename OutT, bool PAIR>
int launch_tc(Ctx &c, const OperandMaps &A, const OperandMaps &B, const TcParams &p, cudaStream_t s) {
  const bool a_mn = A.mn_major, b_mn = B.mn_major;
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks * p.k_splits;  // work units
  // persistent: one CTA (or one CTA pair) per SM (pair of SMs), never more CTAs than units
  const int units = PAIR ? c.sm_count / 2 : c.sm_count;
  const int sched = static_cast<int>(tiles < units ? tiles : units);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(PAIR ? 2 * sched : sched);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TcCfg<PAIR>::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define LB200_LAUNCH(AMN, BMN)                                                                   \
  do {                                                                                           \
    auto kfn = gemm_tc_kernel<ESZ, AMN, BMN, OutT, PAIR>;                                        \
    static std::atomic<uint32_t> attr_set{0};   /* per device: function attributes live in the context */ \
    if (!(attr_set.load(std::memory_order_acquire) & (1u << c.dev))) {                           \
      CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                    TcCfg<PAIR>::SMEM_BYTES));                                   \
      attr_set.fetch_or(1u << c.dev, std::memory_order_release);                                 \
    }                                                                                            \
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kfn, A.hi, A.lo, B.hi, B.lo, A.xb, A.lb, B.xb, B.lb, p));  \
  } while (0)
#define LB200_LAUNCH_HINT(AMN, BMN)                                                              \
  do {                                                                                           \
    auto kfn = gemm_tc_hint_kernel<ESZ, AMN, BMN, OutT, PAIR>;                                   \
    static std::atomic<uint32_t> attr_set{0};   /* per device: function attributes live in the context */ \
    if (!(attr_set.load(std::memory_order_acquire) & (1u << c.dev))) {                           \
      CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                    TcCfg<PAIR>::SMEM_BYTES));                                   \
      attr_set.fetch_or(1u << c.dev, std::memory_order_release);                                 \
    }                                                                                            \
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kfn, A.hi, A.lo, B.hi, B.lo, A.xb, A.lb, B.xb, B.lb, ph)); \
  } while (0)
  bool hinted = false;
  if constexpr (ESZ == 4) hinted = (c.l2_hint == 1 || c.l2_hint == 2);   // the experiment covers the fp32 kernels only
  if constexpr (ESZ == 4) if (hinted) {
    TcHintParams ph;
    static_cast<TcParams &>(ph) = p;
    ph.hint_a = c.l2_hint == 1 ? ptx::kEvictLast : ptx::kEvictFirst;
    ph.hint_b = c.l2_hint == 1 ? ptx::kEvictFirst : ptx::kEvictLast;
    if (!a_mn && !b_mn) LB200_LAUNCH_HINT(false, false);
    else if (!a_mn && b_mn) LB200_LAUNCH_HINT(false, true);
    else if (a_mn && !b_mn) LB200_LAUNCH_HINT(true, false);
    else LB200_LAUNCH_HINT(true, true);
  }
  if (!hinted) {
    if (!a_mn && !b_mn) LB200_LAUNCH(false, false);
    else if (!a_mn && b_mn) LB200_LAUNCH(false, true);
    else if (a_mn && !b_mn) LB200_LAUNCH(true, false);
    else LB200_LAUNCH(true, true);
  }
#undef LB200_LAUNCH_HINT
#undef LB200_LAUNCH
  COUNT_LAUNCH();
  CHECK_LAUNCH();
  return LASER_B200_OK;
}

// gemm_tc_f16_kernel (LASER_B200_PATH_F16X3): same grid / cluster / shared-memory configuration as launch_tc
template <bool PAIR>
int launch_tc_f16(Ctx &c, const OperandMaps &A, const OperandMaps &B, const TcParams &p, const F16Scales &sc, cudaStream_t s) {
  const bool a_mn = A.mn_major, b_mn = B.mn_major;
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks * p.k_splits;
  const int units = PAIR ? c.sm_count / 2 : c.sm_count;
  const int sched = static_cast<int>(tiles < units ? tiles : units);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(PAIR ? 2 * sched : sched);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TcCfg<PAIR>::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TcF16Params pf;
  static_cast<TcParams &>(pf) = p;
  pf.amax_a = sc.a;
  pf.amax_b = sc.b;
#define LB200_LAUNCH_F16(AMN, BMN)                                                               \
  do {                                                                                           \
    auto kfn = gemm_tc_f16_kernel<2, AMN, BMN, float, PAIR>;                                     \
    static std::atomic<uint32_t> attr_set{0};   /* per device: function attributes live in the context */ \
    if (!(attr_set.load(std::memory_order_acquire) & (1u << c.dev))) {                           \
      CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                    TcCfg<PAIR>::SMEM_BYTES));                                   \
      attr_set.fetch_or(1u << c.dev, std::memory_order_release);                                 \
    }                                                                                            \
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kfn, A.hi, A.lo, B.hi, B.lo, A.xb, A.lb, B.xb, B.lb, pf)); \
  } while (0)
  if (!a_mn && !b_mn) LB200_LAUNCH_F16(false, false);
  else if (!a_mn && b_mn) LB200_LAUNCH_F16(false, true);
  else if (a_mn && !b_mn) LB200_LAUNCH_F16(true, false);
  else LB200_LAUNCH_F16(true, true);
#undef LB200_LAUNCH_F16
  COUNT_LAUNCH();
  CHECK_LAUNCH();
  return LASER_B200_OK;
}

// ---- batch of problems in one launch (gemm_tc_batched_kernel): 3-d tensor maps, box depth 1 ----
int encode_map3(Ctx &c, CUtensorMap *map, int esz, const void *base, int64_t inner, int64_t outer, int64_t nb,
                int64_t outer_stride_elems, int64_t batch_stride_elems, int box_inner, int box_outer,
                CUtensorMapSwizzle swz) {
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer), static_cast<cuuint64_t>(nb)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(outer_stride_elems) * esz,
                                 static_cast<cuuint64_t>(nb > 1 ? batch_stride_elems : outer_stride_elems * outer) * esz};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = c.encode(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LASER_B200_ECUDA, "cuTensorMapEncodeTiled (3-d) failed (%d): inner=%lld outer=%lld batch=%lld",
                     static_cast<int>(r), (long long)inner, (long long)outer, (long long)nb);
  return LASER_B200_OK;
}
// nb matrices seen as [mn][k] each, stacked densely: K-major [nb][mn][ld], MN-major [nb][k][ld]
int operand_map3(Ctx &c, CUtensorMap *map, int esz, const void *base, bool mn_major, int64_t mn, int64_t k, int64_t nb,
                 int64_t ld, int block_mn) {
  const int block_k = TC_ROW_BYTES / esz, mn_atom = TC_ROW_BYTES / esz;
  if (!mn_major) return encode_map3(c, map, esz, base, k, mn, nb, ld, mn * ld, block_k, block_mn, CU_TENSOR_MAP_SWIZZLE_128B);
  return encode_map3(c, map, esz, base, mn, k, nb, ld, k * ld, mn_atom, block_k,
                     esz == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

// F16X3 mode: abs-max per mn index of a row-contiguous fp32 operand [R][Cc] (mn along R, or along Cc when the operand is
// MN-major), then its two fp16 pieces (split.cuh, f16_scale.cuh); the caller made room with f16_scales()
int f16x2_prepare(Ctx &c, const float *src, int64_t R, int64_t Cc, int64_t src_ld, bool mn_along_cols, const OperandWs &w,
                  int64_t ld_b, int grid, cudaStream_t s) {
  uint32_t *words = static_cast<uint32_t *>(c.f16s.ptr) + w.amax_off;
  const int64_t n_mn = mn_along_cols ? Cc : R;
  if (c.f16s.bytes < static_cast<size_t>(w.amax_off + n_mn) * sizeof(uint32_t))
    return set_error(LASER_B200_ECUDA, "internal: F16X3 scale buffer not sized for this operand");
  CUDA_TRY(cudaMemsetAsync(words, 0, static_cast<size_t>(n_mn) * sizeof(uint32_t), s));
  uint16_t *xb = static_cast<uint16_t *>(w.xb->ptr), *lb = static_cast<uint16_t *>(w.lb->ptr);
  if (mn_along_cols) {
    const int64_t items = ((Cc + 3) / 4) * ((R + ABSMAX_COL_ROWS - 1) / ABSMAX_COL_ROWS);
    absmax_mn_kernel<true><<<grid_for(c, (items + 255) / 256, 8), 256, 0, s>>>(src, R, Cc, src_ld, words);
    COUNT_LAUNCH();
    CHECK_LAUNCH();
    split_rows_f16x2_kernel<true><<<grid, 256, 0, s>>>(src, R, Cc, src_ld, xb, lb, ld_b, words);
  } else {
    const int64_t warp_items = R * ((Cc + ABSMAX_ROW_CHUNK - 1) / ABSMAX_ROW_CHUNK);
    absmax_mn_kernel<false><<<grid_for(c, (warp_items + 7) / 8, 8), 256, 0, s>>>(src, R, Cc, src_ld, words);
    COUNT_LAUNCH();
    CHECK_LAUNCH();
    split_rows_f16x2_kernel<false><<<grid, 256, 0, s>>>(src, R, Cc, src_ld, xb, lb, ld_b, words);
  }
  return LASER_B200_OK;   // the caller counts and checks this last launch
}

template <int ESZ>
int prepare_operand(Ctx &c, const Operand &o, SplitMode mode, const OperandWs &w, int block_mn,
                    OperandMaps *m, bool *used_ws, cudaStream_t s) {
  using ET = typename std::conditional<ESZ == 4, float, uint16_t>::type;
  const Major mj = classify(o, ESZ);
  const int64_t vec = 16 / ESZ;
  int rc;
  if (mj != GENERAL && mode == SPLIT_NONE) {
    const int64_t ld = (mj == K_MAJOR) ? o.s_mn : o.s_k;
    m->mn_major = (mj == MN_MAJOR);
    if ((rc = operand_map(c, &m->hi, ESZ, o.ptr, mj, o.mn, o.k, ld, block_mn))) return rc;
    m->lo = m->xb = m->lb = m->hi;
    return LASER_B200_OK;
  }
  *used_ws = true;
  // layout of the prepared arrays: the operand's own major-ness when TMA can address it,
  // compact K-major [mn][k] after a gather otherwise
  const Major out_mj = (mj == GENERAL) ? K_MAJOR : mj;
  const int64_t R = (out_mj == K_MAJOR) ? o.mn : o.k;    // rows of the prepared arrays
  const int64_t Cc = (out_mj == K_MAJOR) ? o.k : o.mn;   // contiguous extent
  const int64_t ld = round_up(Cc, vec);
  const int64_t ld_b = round_up(Cc, 8);
  const size_t bytes = static_cast<size_t>(R) * ld * ESZ;
  const size_t bytes_b = static_cast<size_t>(R) * ld_b * 2;
  if (mode != SPLIT_BF16X2 && !(mode == SPLIT_F16X2 && mj != GENERAL) && (rc = ensure(*w.hi, bytes))) return rc;
  if (mode == SPLIT_TF32 && (rc = ensure(*w.lo, bytes))) return rc;
  if (mode == SPLIT_MIXED || mode == SPLIT_BF16X2 || mode == SPLIT_F16X2) {
    if ((rc = ensure(*w.xb, bytes_b))) return rc;
    if ((rc = ensure(*w.lb, bytes_b))) return rc;
  }
  if (mj != GENERAL) {
    // TMA-addressable: elementwise split that keeps the operand's major-ness (fp32 only)
    if constexpr (ESZ == 4) {
      const int64_t src_ld = (mj == K_MAJOR) ? o.s_mn : o.s_k;
      const int64_t items = R * ((Cc + 3) / 4);
      const int grid = grid_for(c, (items + 255) / 256, 8);
      if (mode == SPLIT_F16X2) {
        if ((rc = f16x2_prepare(c, static_cast<const float *>(o.ptr), R, Cc, src_ld, mj == MN_MAJOR, w, ld_b, grid, s))) return rc;
      } else if (mode == SPLIT_TF32)
        split_rows_tf32_kernel<<<grid, 256, 0, s>>>(static_cast<const float *>(o.ptr), R, Cc, src_ld,
                                                    static_cast<float *>(w.hi->ptr),
                                                    static_cast<float *>(w.lo->ptr), ld);
      else if (mode == SPLIT_BF16X2)
        split_rows_bf16x2_kernel<<<grid, 256, 0, s>>>(static_cast<const float *>(o.ptr), R, Cc, src_ld,
                                                      static_cast<uint16_t *>(w.xb->ptr),
                                                      static_cast<uint16_t *>(w.lb->ptr), ld_b);
      else
        split_rows_mixed_kernel<<<grid, 256, 0, s>>>(static_cast<const float *>(o.ptr), R, Cc, src_ld,
                                                     static_cast<float *>(w.hi->ptr), ld,
                                                     static_cast<uint16_t *>(w.xb->ptr),
                                                     static_cast<uint16_t *>(w.lb->ptr), ld_b);
    }
  } else {
    // general strides: one coalesced gather (the surviving descendant of pack_A / pack_B)
    const int64_t tiles = ((o.mn + 31) / 32) * ((o.k + 31) / 32);
    const int read_along_r = (llabs(o.s_mn) < llabs(o.s_k)) ? 1 : 0;
    const int grid = grid_for(c, tiles, 8);
    const ET *src = static_cast<const ET *>(o.ptr);
    ET *dhi = static_cast<ET *>(w.hi->ptr);
    if constexpr (ESZ == 4) {
      if (mode == SPLIT_F16X2) {
        // gather into the compact fp32 array first, then scale + split that (two more passes, general operands only)
        pack_general_kernel<float, 0><<<grid, 256, 0, s>>>(src, o.mn, o.k, o.s_mn, o.s_k, dhi, nullptr, ld,
                                                           read_along_r, nullptr, nullptr, 0);
        COUNT_LAUNCH();
        CHECK_LAUNCH();
        const int g2 = grid_for(c, (R * ((Cc + 3) / 4) + 255) / 256, 8);
        if ((rc = f16x2_prepare(c, dhi, R, Cc, ld, false, w, ld_b, g2, s))) return rc;
      } else if (mode == SPLIT_BF16X2)
        pack_general_kernel<float, 3><<<grid, 256, 0, s>>>(src, o.mn, o.k, o.s_mn, o.s_k, nullptr, nullptr, ld,
                                                           read_along_r, static_cast<uint16_t *>(w.xb->ptr),
                                                           static_cast<uint16_t *>(w.lb->ptr), ld_b);
      else if (mode == SPLIT_TF32)
        pack_general_kernel<float, 1><<<grid, 256, 0, s>>>(src, o.mn, o.k, o.s_mn, o.s_k, dhi,
                                                           static_cast<float *>(w.lo->ptr), ld,
                                                           read_along_r, nullptr, nullptr, 0);
      else if (mode == SPLIT_MIXED)
        pack_general_kernel<float, 2><<<grid, 256, 0, s>>>(src, o.mn, o.k, o.s_mn, o.s_k, dhi, nullptr, ld,
                                                           read_along_r, static_cast<uint16_t *>(w.xb->ptr),
                                                           static_cast<uint16_t *>(w.lb->ptr), ld_b);
      else
        pack_general_kernel<float, 0><<<grid, 256, 0, s>>>(src, o.mn, o.k, o.s_mn, o.s_k, dhi, nullptr, ld,
                                                           read_along_r, nullptr, nullptr, 0);
    } else {
      pack_general_kernel<ET, 0><<<grid, 256, 0, s>>>(src, o.mn, o.k, o.s_mn, o.s_k, dhi, nullptr, ld,
                                                      read_along_r, nullptr, nullptr, 0);
    }
  }
  COUNT_LAUNCH();
  CHECK_LAUNCH();
  m->mn_major = (out_mj == MN_MAJOR);
  if (mode == SPLIT_BF16X2 || mode == SPLIT_F16X2) {
    // the 16-bit kernel's three-pass order reads (hi, lo) = (xb, lb); TMA moves 16-bit words, whatever their format
    if ((rc = operand_map(c, &m->hi, 2, w.xb->ptr, out_mj, o.mn, o.k, ld_b, block_mn))) return rc;
    if ((rc = operand_map(c, &m->lo, 2, w.lb->ptr, out_mj, o.mn, o.k, ld_b, block_mn))) return rc;
    m->xb = m->lb = m->hi;
    return LASER_B200_OK;
  }
  if ((rc = operand_map(c, &m->hi, ESZ, w.hi->ptr, out_mj, o.mn, o.k, ld, block_mn))) return rc;
  m->lo = m->xb = m->lb = m->hi;
  if (mode == SPLIT_TF32) return operand_map(c, &m->lo, ESZ, w.lo->ptr, out_mj, o.mn, o.k, ld, block_mn);
  if (mode == SPLIT_MIXED) {
    if ((rc = operand_map(c, &m->xb, 2, w.xb->ptr, out_mj, o.mn, o.k, ld_b, block_mn))) return rc;
    return operand_map(c, &m->lb, 2, w.lb->ptr, out_mj, o.mn, o.k, ld_b, block_mn);
  }
  return LASER_B200_OK;
}

inline SplitMode split_mode(int npass) {
  return (npass == 3) ? SPLIT_TF32 : (npass == 2) ? SPLIT_MIXED : SPLIT_NONE;
}
inline OperandWs ws_of_A(Ctx &c) { return OperandWs{&c.ws[0], &c.ws[1], &c.ws[2], &c.ws[3], 0}; }
inline OperandWs ws_of_B(Ctx &c, int64_t amax_off = 0) { return OperandWs{&c.ws[4], &c.ws[5], &c.ws[6], &c.ws[7], amax_off}; }
// F16X3: room for M + N abs-max words; A's vector starts at word 0, B's at the returned offset
inline int f16_scales(Ctx &c, int64_t M, int64_t N, int64_t *b_off) {
  *b_off = round_up(M, 64);
  return ensure(c.f16s, static_cast<size_t>(*b_off + N) * sizeof(uint32_t));
}

// launch the tensor-core kernel on prepared operands (c.mu held by the caller)
template <int ESZ, typename OutT>
int tc_run(Ctx &c, int64_t M, int64_t N, int64_t K, float alpha, const OperandMaps &ma,
           const OperandMaps &mb, float beta, OutT *C, int64_t rsC, int64_t csC, int npass, bool pair,
           cudaStream_t s, const F16Scales *f16 = nullptr) {
  // f16 != nullptr: the operands are fp16 pieces of scaled fp32 matrices (F16X3; ESZ == 2, fp32 output only)
  auto launch = [&](const TcParams &q) -> int {
    if constexpr (ESZ == 2 && std::is_same<OutT, float>::value) {
      if (f16) return pair ? launch_tc_f16<true>(c, ma, mb, q, *f16, s) : launch_tc_f16<false>(c, ma, mb, q, *f16, s);
    }
    return pair ? launch_tc<ESZ, OutT, true>(c, ma, mb, q, s) : launch_tc<ESZ, OutT, false>(c, ma, mb, q, s);
  };
  TcParams p;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta;
  p.C = C; p.rsC = rsC; p.csC = csC; p.npass = npass; p.zero = 0; p.epi = g_epi;
  tc_plan<ESZ, std::is_same<OutT, float>::value>(p, npass, pair, TcPlanCfg{c.kc_faithful, c.raster_g, c.splitk_enabled, c.sm_count});
  EventPair ep;
  int rc = prof_open(c, s, &ep, 0);
  if (rc) return rc;
  if (p.k_splits > 1) {
    if constexpr (std::is_same<OutT, float>::value) {
      // partial sums of split s go to plane s of the workspace; a second kernel reduces
      const int64_t ld = round_up(N, 4);
      const int64_t plane = M * ld;
      if ((rc = ensure(c.splitk, static_cast<size_t>(p.k_splits) * plane * sizeof(float)))) return rc;
      TcParams q = p;
      q.C = c.splitk.ptr; q.rsC = ld; q.csC = 1; q.alpha = 1.0f; q.beta = 0.0f; q.epi = Epilogue();
      q.split_plane = plane;
      rc = launch(q);
      if (rc) return rc;
      const int64_t items = (M * N + 255) / 256;
      splitk_reduce_kernel<<<grid_for(c, items, 8), 256, 0, s>>>(
          static_cast<const float *>(c.splitk.ptr), p.k_splits, M, N, ld, plane, alpha, beta, C, rsC, csC,
          p.epi.bias, p.epi.bias_per_row, p.epi.act);
      COUNT_LAUNCH();
      CHECK_LAUNCH();
      CUDA_TRY(cudaEventRecord(c.ws_free, s));  // the planes are workspace too
      return prof_close(c, s, &ep, 2);
    }
  }
  rc = launch(p);
  if (rc) return rc;
  return prof_close(c, s, &ep, 1);
}

// SRC_ESZ: element size of the caller's operands; it differs from the kernel's ESZ only in the
// BF16X3 mode (fp32 operands split into two bf16 arrays each, multiplied by the bf16 kernel)
template <int ESZ, typename OutT, int SRC_ESZ = ESZ>
int gemm_tc(Ctx &c, int64_t M, int64_t N, int64_t K, float alpha, const void *A, int64_t rsA,
            int64_t csA, const void *B, int64_t rsB, int64_t csB, float beta, OutT *C, int64_t rsC,
            int64_t csC, int npass, cudaStream_t s, SplitMode forced = SPLIT_NONE) {
  if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL)
    return set_error(LASER_B200_EUNSUPPORTED, "tensor-core path: extents must fit in int32");
  std::lock_guard<std::mutex> lk(c.mu);  // workspace + descriptor construction are per context
  const SplitMode mode = (forced != SPLIT_NONE) ? forced : split_mode(npass);
  Operand oa{A, M, K, rsA, csA};
  Operand ob{B, N, K, csB, rsB};
  OperandMaps ma, mb;
  bool used_ws = false;
  // the previous call may still be reading the workspace on another stream
  CUDA_TRY(cudaStreamWaitEvent(s, c.ws_free, 0));
  EventPair ep;
  const int64_t launches_before = g_launches.load();
  int rc = prof_open(c, s, &ep, 1);
  if (rc) return rc;
  int64_t f16_b_off = 0;
  if (mode == SPLIT_F16X2 && (rc = f16_scales(c, M, N, &f16_b_off))) return rc;
  rc = prepare_operand<SRC_ESZ>(c, oa, mode, ws_of_A(c), TC_BLOCK_M, &ma, &used_ws, s);
  if (rc) return rc;
  // CTA pairs (cta_group::2, 256 x 256 tiles) whenever there are at least two 128-row blocks
  const bool pair = c.cta_pair && M > TC_BLOCK_M;
  rc = prepare_operand<SRC_ESZ>(c, ob, mode, ws_of_B(c, f16_b_off), pair ? TC_BLOCK_N / 2 : TC_BLOCK_N, &mb, &used_ws, s);
  if (rc) return rc;
  rc = prof_close(c, s, &ep, static_cast<int>(g_launches.load() - launches_before));
  if (rc) return rc;
  const F16Scales f16{static_cast<const uint32_t *>(c.f16s.ptr), static_cast<const uint32_t *>(c.f16s.ptr) + f16_b_off};
  rc = tc_run<ESZ, OutT>(c, M, N, K, alpha, ma, mb, beta, C, rsC, csC, npass, pair, s, mode == SPLIT_F16X2 ? &f16 : nullptr);
  if (rc) return rc;
  if (used_ws) CUDA_TRY(cudaEventRecord(c.ws_free, s));
  return LASER_B200_OK;
}

template <int ESZ, typename OutT, bool PAIR>
int launch_tc_batched(Ctx &c, const OperandMaps &A, const OperandMaps &B, const TcBatchedParams &p, cudaStream_t s) {
  const bool a_mn = A.mn_major, b_mn = B.mn_major;
  const int64_t units_total = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks * p.k_splits * p.batch;
  const int units = PAIR ? c.sm_count / 2 : c.sm_count;
  const int sched = static_cast<int>(units_total < units ? units_total : units);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(PAIR ? 2 * sched : sched);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TcCfg<PAIR>::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define LB200_LAUNCH(AMN, BMN)                                                                   \
  do {                                                                                           \
    auto kfn = gemm_tc_batched_kernel<ESZ, AMN, BMN, OutT, PAIR>;                                \
    static std::atomic<uint32_t> attr_set{0};                                                    \
    if (!(attr_set.load(std::memory_order_acquire) & (1u << c.dev))) {                           \
      CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                    TcCfg<PAIR>::SMEM_BYTES));                                   \
      attr_set.fetch_or(1u << c.dev, std::memory_order_release);                                 \
    }                                                                                            \
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kfn, A.hi, A.lo, B.hi, B.lo, A.xb, A.lb, B.xb, B.lb, p));  \
  } while (0)
  if (!a_mn && !b_mn) LB200_LAUNCH(false, false);
  else if (!a_mn && b_mn) LB200_LAUNCH(false, true);
  else if (a_mn && !b_mn) LB200_LAUNCH(true, false);
  else LB200_LAUNCH(true, true);
#undef LB200_LAUNCH
  COUNT_LAUNCH();
  CHECK_LAUNCH();
  return LASER_B200_OK;
}

// One operand of a batch: nb = 1 (shared) or `batch` matrices that must stack densely -- K-major rows one
// after the other (batch stride = mn * row pitch) or MN-major k-rows one after the other (batch stride =
// k * pitch); then the whole stack is ONE matrix for the preparation kernels.  Returns -2 when the layout
// does not stack (the caller falls back to one launch sequence per problem).
int prepare_operand_stack(Ctx &c, const void *ptr, int64_t mn, int64_t k, int64_t s_mn, int64_t s_k, int64_t bs, int64_t nb,
                          SplitMode mode, const OperandWs &w, int block_mn, OperandMaps *m, bool *used_ws, cudaStream_t s) {
  Operand one{ptr, mn, k, s_mn, s_k};
  const Major mj = classify(one, 4);
  if (mj == GENERAL && nb > 1) return -2;   // a single (shared) matrix of any strides is gathered as usual
  Operand stack = one;
  if (nb > 1) {
    if (mj == K_MAJOR) { if (bs != mn * s_mn) return -2; stack.mn = nb * mn; }
    else { if (bs != k * s_k) return -2; stack.k = nb * k; }
  }
  OperandMaps flat;
  bool used = false;
  int rc = prepare_operand<4>(c, stack, mode, w, block_mn, &flat, &used, s);   // the 2-d maps it builds are not used
  if (rc) return rc;
  *used_ws = *used_ws || used;
  const bool mn_major = (mj == MN_MAJOR);   // gathered operands come out K-major
  const int64_t Cc = mn_major ? mn : k;
  const int64_t ld = used ? round_up(Cc, 4) : (mn_major ? s_k : s_mn), ld_b = round_up(Cc, 8);
  const void *hi = used ? w.hi->ptr : ptr;
  m->mn_major = mn_major;
  if ((rc = operand_map3(c, &m->hi, 4, hi, mn_major, mn, k, nb, ld, block_mn))) return rc;
  m->lo = m->xb = m->lb = m->hi;
  if (mode == SPLIT_TF32) return operand_map3(c, &m->lo, 4, w.lo->ptr, mn_major, mn, k, nb, ld, block_mn);
  if (mode == SPLIT_MIXED) {
    if ((rc = operand_map3(c, &m->xb, 2, w.xb->ptr, mn_major, mn, k, nb, ld_b, block_mn))) return rc;
    return operand_map3(c, &m->lb, 2, w.lb->ptr, mn_major, mn, k, nb, ld_b, block_mn);
  }
  return LASER_B200_OK;
}

// `batch` float32 problems of one shape as ONE tensor-core launch (after at most one preparation launch
// per operand).  -2: the operands do not stack, nothing was launched.
int gemm_tc_batched(Ctx &c, int64_t batch, int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t rsA,
                    int64_t csA, int64_t bsA, const float *B, int64_t rsB, int64_t csB, int64_t bsB, float beta, float *C,
                    int64_t rsC, int64_t csC, int64_t bsC, int npass, cudaStream_t s) {
  if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL || batch > 0x7fffffffLL / 4 ||
      batch * M > 0x7fffffffLL || batch * K > 0x7fffffffLL || batch * N > 0x7fffffffLL)
    return -2;
  std::lock_guard<std::mutex> lk(c.mu);
  const SplitMode mode = split_mode(npass);
  const bool pair = c.cta_pair && M > TC_BLOCK_M;
  OperandMaps ma, mb;
  bool used_ws = false;
  // both operands must stack before anything is launched
  {
    Operand oa{A, M, K, rsA, csA}, ob{B, N, K, csB, rsB};
    const Major ja = classify(oa, 4), jb = classify(ob, 4);
    if ((ja == GENERAL && bsA != 0) || (jb == GENERAL && bsB != 0)) return -2;
    if (bsA != 0 && bsA != (ja == K_MAJOR ? M * rsA : K * csA)) return -2;
    if (bsB != 0 && bsB != (jb == K_MAJOR ? N * csB : K * rsB)) return -2;
  }
  CUDA_TRY(cudaStreamWaitEvent(s, c.ws_free, 0));
  int rc = prepare_operand_stack(c, A, M, K, rsA, csA, bsA, bsA == 0 ? 1 : batch, mode, ws_of_A(c), TC_BLOCK_M, &ma, &used_ws, s);
  if (rc) return rc;
  rc = prepare_operand_stack(c, B, N, K, csB, rsB, bsB, bsB == 0 ? 1 : batch, mode, ws_of_B(c), pair ? TC_BLOCK_N / 2 : TC_BLOCK_N,
                             &mb, &used_ws, s);
  if (rc) return rc;
  TcBatchedParams p;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta;
  p.C = C; p.rsC = rsC; p.csC = csC; p.npass = npass; p.zero = 0; p.epi = g_epi;
  tc_plan<4, true>(p, npass, pair, TcPlanCfg{c.kc_faithful, c.raster_g, false /* no split-K */, c.sm_count});
  p.batch = static_cast<int>(batch);
  p.a_shared = bsA == 0 ? 1 : 0;
  p.b_shared = bsB == 0 ? 1 : 0;
  p.bsC = bsC;
  if (pair) rc = launch_tc_batched<4, float, true>(c, ma, mb, p, s);
  else rc = launch_tc_batched<4, float, false>(c, ma, mb, p, s);
  if (rc) return rc;
  if (used_ws) CUDA_TRY(cudaEventRecord(c.ws_free, s));
  return LASER_B200_OK;
}

// ---------------------------------------------------------------------------------------
//                      pre-packed operands (gemm_prepacked.nim:63-292)
// ---------------------------------------------------------------------------------------
// Layout of a packed operand seen as [mn][k] (A: mn = M, B: mn = N), a pure function of (mn, k):
//   [hi : mn x ld  fp32][xb : mn x ld_b bf16][lb : mn x ld_b bf16], sections 256-byte aligned,
//   ld = round_up(k, 4), ld_b = round_up(k, 8): compact K-major arrays of the mixed mode.
int finish(Ctx &c, cudaStream_t user, cudaStream_t s);

struct PackedLayout {
  int64_t ld, ld_b;
  size_t off_hi, off_xb, off_lb, bytes;
};
inline PackedLayout packed_layout(int64_t mn, int64_t k) {
  PackedLayout L;
  L.ld = round_up(k, 4);
  L.ld_b = round_up(k, 8);
  auto al = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
  L.off_hi = 0;
  L.off_xb = al(static_cast<size_t>(mn) * L.ld * 4);
  L.off_lb = L.off_xb + al(static_cast<size_t>(mn) * L.ld_b * 2);
  L.bytes = L.off_lb + al(static_cast<size_t>(mn) * L.ld_b * 2);
  return L;
}

int prepack_dev(void *dst, int64_t mn, int64_t k, const float *src, int64_t s_mn, int64_t s_k,
                void *stream) {
  if (mn < 0 || k < 0) return set_error(LASER_B200_EINVAL, "negative extent");
  if (mn == 0 || k == 0) return LASER_B200_OK;
  if (!dst || !src) return set_error(LASER_B200_EINVAL, "null pointer");
  if (reinterpret_cast<uintptr_t>(dst) & 255) return set_error(LASER_B200_EINVAL, "packed buffer must be 256-byte aligned");
  Ctx *c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  const PackedLayout L = packed_layout(mn, k);
  uint8_t *base = static_cast<uint8_t *>(dst);
  const int64_t tiles = ((mn + 31) / 32) * ((k + 31) / 32);
  const int read_along_r = (llabs(s_mn) < llabs(s_k)) ? 1 : 0;
  pack_general_kernel<float, 2><<<grid_for(*c, tiles, 8), 256, 0, s>>>(
      src, mn, k, s_mn, s_k, reinterpret_cast<float *>(base + L.off_hi), nullptr, L.ld, read_along_r,
      reinterpret_cast<uint16_t *>(base + L.off_xb), reinterpret_cast<uint16_t *>(base + L.off_lb), L.ld_b);
  COUNT_LAUNCH();
  CHECK_LAUNCH();
  return finish(*c, static_cast<cudaStream_t>(stream), s);
}

int packed_maps(Ctx &c, const void *packed, int64_t mn, int64_t k, int block_mn, OperandMaps *m) {
  const PackedLayout L = packed_layout(mn, k);
  const uint8_t *base = static_cast<const uint8_t *>(packed);
  int rc;
  m->mn_major = false;
  if ((rc = operand_map(c, &m->hi, 4, base + L.off_hi, K_MAJOR, mn, k, L.ld, block_mn))) return rc;
  m->lo = m->hi;
  if ((rc = operand_map(c, &m->xb, 2, base + L.off_xb, K_MAJOR, mn, k, L.ld_b, block_mn))) return rc;
  return operand_map(c, &m->lb, 2, base + L.off_lb, K_MAJOR, mn, k, L.ld_b, block_mn);
}

// A: either raw (A != nullptr) or packed (packedA != nullptr); B always packed
int gemm_packed_dev(int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t rsA,
                    int64_t csA, const void *packedA, const void *packedB, float beta, float *C,
                    int64_t rsC, int64_t csC, void *stream) {
  if (M < 0 || N < 0 || K < 0) return set_error(LASER_B200_EINVAL, "negative extent");
  if (M == 0 || N == 0 || K == 0) return LASER_B200_OK;
  if ((!A && !packedA) || !packedB || !C) return set_error(LASER_B200_EINVAL, "null pointer");
  if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL)
    return set_error(LASER_B200_EUNSUPPORTED, "extents must fit in int32");
  Ctx *cp;
  int rc = get_ctx(&cp);
  if (rc) return rc;
  Ctx &c = *cp;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c.stream;
  {
    std::lock_guard<std::mutex> lk(c.mu);
    const bool pair = c.cta_pair && M > TC_BLOCK_M;
    OperandMaps ma, mb;
    bool used_ws = false;
    CUDA_TRY(cudaStreamWaitEvent(s, c.ws_free, 0));
    if (packedA) {
      if ((rc = packed_maps(c, packedA, M, K, TC_BLOCK_M, &ma))) return rc;
    } else {
      Operand oa{A, M, K, rsA, csA};
      EventPair ep;
      const int64_t before = g_launches.load();
      if ((rc = prof_open(c, s, &ep, 1))) return rc;
      if ((rc = prepare_operand<4>(c, oa, SPLIT_MIXED, ws_of_A(c), TC_BLOCK_M, &ma, &used_ws, s))) return rc;
      if ((rc = prof_close(c, s, &ep, static_cast<int>(g_launches.load() - before)))) return rc;
    }
    if ((rc = packed_maps(c, packedB, N, K, pair ? TC_BLOCK_N / 2 : TC_BLOCK_N, &mb))) return rc;
    if ((rc = tc_run<4, float>(c, M, N, K, alpha, ma, mb, beta, C, rsC, csC, 2, pair, s))) return rc;
    if (used_ws) CUDA_TRY(cudaEventRecord(c.ws_free, s));
  }
  g_last_path = LASER_B200_PATH_TF32_BF16C;
  return finish(c, static_cast<cudaStream_t>(stream), s);
}

// ---------------------------------------------------------------------------------------
//                                      dispatch
// ---------------------------------------------------------------------------------------
int check_args(int64_t M, int64_t N, int64_t K, const void *A, const void *B, const void *C) {
  if (M < 0 || N < 0 || K < 0) return set_error(LASER_B200_EINVAL, "negative extent M=%lld N=%lld K=%lld",
                                                (long long)M, (long long)N, (long long)K);
  if (M == 0 || N == 0 || K == 0) return -1;  // nothing to do (gemm.nim:150: C untouched)
  if (!A || !B || !C) return set_error(LASER_B200_EINVAL, "null matrix pointer");
  return LASER_B200_OK;
}

int finish(Ctx &, cudaStream_t user, cudaStream_t s) {
  if (!user) CUDA_TRY(cudaStreamSynchronize(s));
  return LASER_B200_OK;
}

int f32_dev(int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t rsA, int64_t csA,
            const float *B, int64_t rsB, int64_t csB, float beta, float *C, int64_t rsC, int64_t csC,
            int path, void *stream) {
  int rc = check_args(M, N, K, A, B, C);
  if (rc == -1) return LASER_B200_OK;
  if (rc) return rc;
  Ctx *c;
  rc = get_ctx(&c);
  if (rc) return rc;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  if (path == LASER_B200_PATH_AUTO) {
    // the reference switches behaviour at M*N*K > 128^3 (gemm.nim:140-141); below that
    // a 128x256 tensor-core tile is mostly padding and the exact kernel is used
    const double work = static_cast<double>(M) * N * K;
    if (work <= 128.0 * 128.0 * 128.0) path = LASER_B200_PATH_SIMT;
    else if (N <= 4 && M >= 1024 && !g_epi.bias && !g_epi.act) path = -1;  // skinny: warp-shuffle GEMV
    else path = g_f32_mode.load();
  }
  switch (path) {
    case -1: {
      const int grid = grid_for(*c, (M + 7) / 8, 8);
const bool vec = (csA == 1) && (rsA % 4 == 0) && (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
      const size_t bsmem = static_cast<size_t>(N) * K * sizeof(float);
      const bool use_smem = vec && bsmem <= 96 * 1024;
#define LB200_GEMV(NV)                                                                                         \
  do {                                                                                                         \
    if (use_smem) {                                                                                            \
      static std::atomic<uint32_t> attr_set{0};                                                                \
      if (!(attr_set.load(std::memory_order_acquire) & (1u << c->dev))) {                                      \
        CUDA_TRY(cudaFuncSetAttribute(gemv_warp_smem_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
        attr_set.fetch_or(1u << c->dev, std::memory_order_release);                                            \
      }                                                                                                        \
      gemv_warp_smem_kernel<NV><<<grid_for(*c, (M + 7) / 8, 2), 256, bsmem, s>>>(M, K, alpha, A, rsA, B, rsB, csB, beta, C, rsC, csC); \
    } else if (vec) gemv_warp_kernel<NV, true><<<grid, 256, 0, s>>>(M, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);  \
    else gemv_warp_kernel<NV, false><<<grid, 256, 0, s>>>(M, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);     \
  } while (0)
      if (N == 1) LB200_GEMV(1); else if (N == 2) LB200_GEMV(2); else if (N == 3) LB200_GEMV(3); else LB200_GEMV(4);
#undef LB200_GEMV
      COUNT_LAUNCH();
      CHECK_LAUNCH();
      g_last_path = LASER_B200_PATH_SIMT;
      break;
    }
    case LASER_B200_PATH_SIMT:
      rc = gemm_simt<float>(*c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, s);
      if (rc) return rc;
      g_last_path = LASER_B200_PATH_SIMT;
      break;
    case LASER_B200_PATH_TF32X1:
    case LASER_B200_PATH_TF32X3:
    case LASER_B200_PATH_TF32_BF16C:
      rc = gemm_tc<4, float>(*c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC,
                             path == LASER_B200_PATH_TF32X3 ? 3 : (path == LASER_B200_PATH_TF32_BF16C ? 2 : 1), s);
      if (rc) return rc;
      g_last_path = path;
      break;
    case LASER_B200_PATH_BF16X3:
      // two bf16 pieces per fp32 operand, three passes (h*l', l*h', h*h') of the bf16 kernel, fp32 output
      rc = gemm_tc<2, float, 4>(*c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, 3, s, SPLIT_BF16X2);
      if (rc) return rc;
      g_last_path = path;
      break;
    case LASER_B200_PATH_F16X3:
      // two fp16 pieces of each operand scaled by a power of two (device-side abs-max), three passes of the fp16 flavour
      // of the kernel, whose epilogue undoes the scales
      rc = gemm_tc<2, float, 4>(*c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, 3, s, SPLIT_F16X2);
      if (rc) return rc;
      g_last_path = path;
      break;
    default:
      return set_error(LASER_B200_EINVAL, "unknown path %d for float32", path);
  }
  return finish(*c, static_cast<cudaStream_t>(stream), s);
}

template <typename T>
int simt_dev(int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t rsA, int64_t csA,
             const T *B, int64_t rsB, int64_t csB, T beta, T *C, int64_t rsC, int64_t csC,
             void *stream) {
  int rc = check_args(M, N, K, A, B, C);
  if (rc == -1) return LASER_B200_OK;
  if (rc) return rc;
  Ctx *c;
  rc = get_ctx(&c);
  if (rc) return rc;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  rc = gemm_simt<T>(*c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, s);
  if (rc) return rc;
  g_last_path = LASER_B200_PATH_SIMT;
  return finish(*c, static_cast<cudaStream_t>(stream), s);
}

int bf16_dev(int64_t M, int64_t N, int64_t K, float alpha, const uint16_t *A, int64_t rsA,
             int64_t csA, const uint16_t *B, int64_t rsB, int64_t csB, float beta, uint16_t *C,
             int64_t rsC, int64_t csC, void *stream) {
  int rc = check_args(M, N, K, A, B, C);
  if (rc == -1) return LASER_B200_OK;
  if (rc) return rc;
  Ctx *c;
  rc = get_ctx(&c);
  if (rc) return rc;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  rc = gemm_tc<2, uint16_t>(*c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, 1, s);
  if (rc) return rc;
  g_last_path = LASER_B200_PATH_BF16;
  return finish(*c, static_cast<cudaStream_t>(stream), s);
}

// ---------------------------------------------------------------------------------------
//                         host-pointer (drop-in) variants
// ---------------------------------------------------------------------------------------
struct Span {
  int64_t lo, hi;   // element offsets relative to the base pointer, inclusive
  bool dense;       // every element of [lo, hi] belongs to the view
};
Span span_of(int64_t rows, int64_t cols, int64_t rs, int64_t cs) {
  Span sp;
  const int64_t r = (rows - 1) * rs, q = (cols - 1) * cs;
  sp.lo = (r < 0 ? r : 0) + (q < 0 ? q : 0);
  sp.hi = (r > 0 ? r : 0) + (q > 0 ? q : 0);
  const int64_t ars = llabs(rs), acs = llabs(cs);
  sp.dense = (acs == 1 && (ars == cols || rows == 1)) || (ars == 1 && (acs == rows || cols == 1)) ||
             (rows == 1 && cols == 1);
  return sp;
}

// Host-pointer fp32 GEMM, pipelined over row panels (the drop-in call's fast path).
// PCIe is the bound of a host-resident GEMM (805 MB cross the bus for 1.1 TFLOP at 8192^3), so
// the three phases run on three streams: B is uploaded and prepared once; then row panel p+1 of
// A is in flight H2D while panel p is split + multiplied and panel p-1 of C returns D2H.
// Preconditions (checked by the caller): tensor-core path, row panels of A and of C are
// (nearly) disjoint address ranges, C dense inside each panel span (with beta != 0 the old panel of
// C travels to the device next to its panel of A).
struct PanelPlan {
  int64_t rows;   // rows per panel
  int panels;
};
inline bool panel_separable(int64_t rows, int64_t cols, int64_t rs, int64_t cs) {
  // a panel of `rows` consecutive rows must cover an address span not much larger than its data
  const Span sp = span_of(rows, cols, rs, cs);
  const double span = static_cast<double>(sp.hi - sp.lo + 1);
  return llabs(rs) >= llabs(cs) && span <= 1.25 * static_cast<double>(rows) * static_cast<double>(cols);
}

int host_gemm_f32_pipelined(Ctx &c, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                            int64_t rsA, int64_t csA, const float *B, int64_t rsB, int64_t csB,
                            float beta, float *C, int64_t rsC, int64_t csC, int path) {
  const bool f16x3 = (path == LASER_B200_PATH_F16X3);
  const bool bf16x3 = (path == LASER_B200_PATH_BF16X3) || f16x3;   // two 16-bit pieces per operand: the 16-bit kernel, three passes
  const int npass = (path == LASER_B200_PATH_TF32X3 || bf16x3) ? 3 : (path == LASER_B200_PATH_TF32_BF16C ? 2 : 1);
  std::lock_guard<std::mutex> host_lk(c.host_mu);
  std::lock_guard<std::mutex> lk(c.mu);
  int rc;
  const int64_t panel_rows = c.panel_rows;
  // row panels (first row, rows).  The time after the last byte of A has crossed the bus is one
  // panel's split + GEMM + D2H: optionally the last panel is cut finer (halves down to 256 rows).
  std::vector<std::pair<int64_t, int64_t>> plist;
  for (int64_t m0 = 0; m0 < M; m0 += panel_rows) plist.emplace_back(m0, (M - m0 < panel_rows) ? (M - m0) : panel_rows);
  if (c.panel_taper && plist.size() >= 2) {
    int64_t m0 = plist.back().first, rows = plist.back().second;
    plist.pop_back();
    while (rows > 256) {
      const int64_t h = (rows / 2 + 255) / 256 * 256;
      if (h >= rows) break;
      plist.emplace_back(m0, h);
      m0 += h;
      rows -= h;
    }
    plist.emplace_back(m0, rows);
  }
  const int panels = static_cast<int>(plist.size());
  const Span sa = span_of(M, K, rsA, csA), sb = span_of(K, N, rsB, csB), sc = span_of(M, N, rsC, csC);
  const size_t na = static_cast<size_t>(sa.hi - sa.lo + 1) * 4, nb = static_cast<size_t>(sb.hi - sb.lo + 1) * 4;
  const size_t nc = static_cast<size_t>(sc.hi - sc.lo + 1) * 4;
  if ((rc = ensure(c.stage[0], na + 256))) return rc;
  if ((rc = ensure(c.stage[1], nb + 256))) return rc;
  if ((rc = ensure(c.stage[2], nc + 256))) return rc;
  float *dA = static_cast<float *>(c.stage[0].ptr) - sa.lo;   // device address of element A[0,0]
  float *dB = static_cast<float *>(c.stage[1].ptr) - sb.lo;
  float *dC = static_cast<float *>(c.stage[2].ptr) - sc.lo;
  if (static_cast<int>(c.panel_ev.size()) < 2 * panels + 1) {
    const size_t old = c.panel_ev.size();
    c.panel_ev.resize(2 * panels + 1);
    for (size_t i = old; i < c.panel_ev.size(); ++i)
      CUDA_TRY(cudaEventCreateWithFlags(&c.panel_ev[i], cudaEventDisableTiming));
  }
  cudaStream_t up = c.up, cmp = c.stream, down = c.down;
  // f16x3: B's abs-max words are written once, A's (one per row of the panel) are rewritten by every panel's preparation --
  // after the previous panel's GEMM, whose epilogue reads them, because everything of a panel runs on the compute stream
  const SplitMode mode = f16x3 ? SPLIT_F16X2 : (bf16x3 ? SPLIT_BF16X2 : split_mode(npass));
  const bool pair = c.cta_pair && panel_rows > TC_BLOCK_M && M > TC_BLOCK_M;
  // staging buffers / workspace may still be in use by an earlier call
  CUDA_TRY(cudaStreamWaitEvent(up, c.ws_free, 0));
  CUDA_TRY(cudaStreamWaitEvent(cmp, c.ws_free, 0));
  // ---- B: upload once, prepare once ----
  CUDA_TRY(cudaMemcpyAsync(dB + sb.lo, B + sb.lo, nb, cudaMemcpyHostToDevice, up));
  CUDA_TRY(cudaEventRecord(c.panel_ev[2 * panels], up));
  CUDA_TRY(cudaStreamWaitEvent(cmp, c.panel_ev[2 * panels], 0));
  OperandMaps mb;
  bool used_ws = false;
  Operand ob{dB, N, K, csB, rsB};
  int64_t f16_b_off = 0;       // f16x3: room for the longest panel's rows of A + the columns of B
  if (f16x3 && (rc = f16_scales(c, panel_rows < M ? panel_rows : M, N, &f16_b_off))) return rc;
  if ((rc = prepare_operand<4>(c, ob, mode, ws_of_B(c, f16_b_off), pair ? TC_BLOCK_N / 2 : TC_BLOCK_N, &mb, &used_ws, cmp)))
    return rc;
  const F16Scales f16{static_cast<const uint32_t *>(c.f16s.ptr), static_cast<const uint32_t *>(c.f16s.ptr) + f16_b_off};
  // ---- row panels ----
  for (int pnl = 0; pnl < panels; ++pnl) {
    const int64_t m0 = plist[pnl].first;
    const int64_t mp = plist[pnl].second;
    const float *Ap = A + m0 * rsA;
    float *Cp = C + m0 * rsC;
    const Span pa = span_of(mp, K, rsA, csA), pc = span_of(mp, N, rsC, csC);
    CUDA_TRY(cudaMemcpyAsync(dA + m0 * rsA + pa.lo, Ap + pa.lo, static_cast<size_t>(pa.hi - pa.lo + 1) * 4,
                             cudaMemcpyHostToDevice, up));
    if (beta != 0.0f)   // the old values of this panel of C are read by the epilogue
      CUDA_TRY(cudaMemcpyAsync(dC + m0 * rsC + pc.lo, Cp + pc.lo, static_cast<size_t>(pc.hi - pc.lo + 1) * 4,
                               cudaMemcpyHostToDevice, up));
    CUDA_TRY(cudaEventRecord(c.panel_ev[pnl], up));
    CUDA_TRY(cudaStreamWaitEvent(cmp, c.panel_ev[pnl], 0));
    OperandMaps ma;
    Operand oa{dA + m0 * rsA, mp, K, rsA, csA};
    if ((rc = prepare_operand<4>(c, oa, mode, ws_of_A(c), TC_BLOCK_M, &ma, &used_ws, cmp))) return rc;
    // B's tensor maps were built for `pair` (128- vs 256-column boxes): every panel, however
    // short, must run the same kernel variant
    if (bf16x3) rc = tc_run<2, float>(c, mp, N, K, alpha, ma, mb, beta, dC + m0 * rsC, rsC, csC, npass, pair, cmp,
                                      f16x3 ? &f16 : nullptr);
    else rc = tc_run<4, float>(c, mp, N, K, alpha, ma, mb, beta, dC + m0 * rsC, rsC, csC, npass, pair, cmp);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c.panel_ev[panels + pnl], cmp));
    CUDA_TRY(cudaStreamWaitEvent(down, c.panel_ev[panels + pnl], 0));
    CUDA_TRY(cudaMemcpyAsync(Cp + pc.lo, dC + m0 * rsC + pc.lo, static_cast<size_t>(pc.hi - pc.lo + 1) * 4,
                             cudaMemcpyDeviceToHost, down));
  }
  CUDA_TRY(cudaEventRecord(c.ws_free, cmp));
  CUDA_TRY(cudaStreamSynchronize(down));
  CUDA_TRY(cudaStreamSynchronize(cmp));
  g_last_path = path;
  return LASER_B200_OK;
}

template <typename T, typename Fn>
int host_gemm(int64_t M, int64_t N, int64_t K, const T *A, int64_t rsA, int64_t csA, const T *B,
              int64_t rsB, int64_t csB, bool beta_zero, T *C, int64_t rsC, int64_t csC, Fn run) {
  int rc = check_args(M, N, K, A, B, C);
  if (rc == -1) return LASER_B200_OK;
  if (rc) return rc;
  Ctx *c;
  rc = get_ctx(&c);
  if (rc) return rc;
  const Span sa = span_of(M, K, rsA, csA), sb = span_of(K, N, rsB, csB), sc = span_of(M, N, rsC, csC);
  const size_t na = static_cast<size_t>(sa.hi - sa.lo + 1) * sizeof(T);
  const size_t nb = static_cast<size_t>(sb.hi - sb.lo + 1) * sizeof(T);
  const size_t nc = static_cast<size_t>(sc.hi - sc.lo + 1) * sizeof(T);
  std::lock_guard<std::mutex> host_lk(c->host_mu);  // one host-pointer call at a time per device
  if ((rc = ensure(c->stage[0], na + 256))) return rc;
  if ((rc = ensure(c->stage[1], nb + 256))) return rc;
  if ((rc = ensure(c->stage[2], nc + 256))) return rc;
  T *dA = static_cast<T *>(c->stage[0].ptr);
  T *dB = static_cast<T *>(c->stage[1].ptr);
  T *dC = static_cast<T *>(c->stage[2].ptr);
  cudaStream_t s = c->stream;
  CUDA_TRY(cudaMemcpyAsync(dA, A + sa.lo, na, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(dB, B + sb.lo, nb, cudaMemcpyHostToDevice, s));
  // C travels to the device only if it is read (beta != 0) or if the span holds
  // elements outside the view that must survive the round trip
  if (!beta_zero || !sc.dense) CUDA_TRY(cudaMemcpyAsync(dC, C + sc.lo, nc, cudaMemcpyHostToDevice, s));
  rc = run(dA - sa.lo, dB - sb.lo, dC - sc.lo, static_cast<void *>(s));
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(C + sc.lo, dC, nc, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return LASER_B200_OK;
}

}  // namespace

// =======================================================================================
//                                     extern "C"
// =======================================================================================
extern "C" {

int laser_b200_init(void) {
  Ctx *c;
  return get_ctx(&c);
}

void laser_b200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < kMaxDevices; ++d) {
    Ctx &c = g_ctx[d];
    if (!c.ready) continue;
    cudaSetDevice(d);
    cudaStreamSynchronize(c.stream);
    for (auto &b : c.ws) { if (b.ptr) cudaFree(b.ptr); b = Buffer(); }
    for (auto &b : c.stage) { if (b.ptr) cudaFree(b.ptr); b = Buffer(); }
    if (c.splitk.ptr) { cudaFree(c.splitk.ptr); c.splitk = Buffer(); }
    if (c.layer_ws.ptr) { cudaFree(c.layer_ws.ptr); c.layer_ws = Buffer(); }
    if (c.f16s.ptr) { cudaFree(c.f16s.ptr); c.f16s = Buffer(); }
    cudaEventDestroy(c.ws_free);
    for (auto e : c.panel_ev) cudaEventDestroy(e);
    c.panel_ev.clear();
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.up);
    cudaStreamDestroy(c.down);
    c.ready = false;
  }
  cudaSetDevice(cur);
}

int laser_b200_profile_begin(void) {
  Ctx *c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  for (auto &e : c->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  c->prof.clear();
  c->profiling = true;
  return LASER_B200_OK;
}
int laser_b200_profile_end(double *gemm_ms, int64_t *gemm_launches, double *prep_ms,
                           int64_t *prep_launches) {
  Ctx *c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(c->mu);
  double ms[2] = {0.0, 0.0};
  int64_t n[2] = {0, 0};
  for (auto &e : c->prof) {
    float t = 0.0f;
    if (e.launches > 0) {
      CUDA_TRY(cudaEventElapsedTime(&t, e.a, e.b));
      ms[e.kind] += t;
      n[e.kind] += e.launches;
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  c->prof.clear();
  c->profiling = false;
  if (gemm_ms) *gemm_ms = ms[0];
  if (gemm_launches) *gemm_launches = n[0];
  if (prep_ms) *prep_ms = ms[1];
  if (prep_launches) *prep_launches = n[1];
  return LASER_B200_OK;
}

const char *laser_b200_last_error(void) { return g_last_error.c_str(); }
int laser_b200_version(void) { return 100; }
int64_t laser_b200_launch_count(void) { return g_launches.load(); }
int laser_b200_last_path(void) { return g_last_path; }
int laser_b200_set_f32_mode(int path) {
  if (path != LASER_B200_PATH_SIMT && path != LASER_B200_PATH_TF32X1 && path != LASER_B200_PATH_TF32X3 &&
      path != LASER_B200_PATH_TF32_BF16C && path != LASER_B200_PATH_BF16X3 && path != LASER_B200_PATH_F16X3)
    return set_error(LASER_B200_EINVAL, "f32 mode must be SIMT, TF32X1, TF32X3, TF32_BF16C, BF16X3 or F16X3");
  g_f32_mode.store(path);
  return LASER_B200_OK;
}
int laser_b200_get_f32_mode(void) {
  const int m = g_f32_mode.load();
  return m < 0 ? LASER_B200_PATH_TF32_BF16C : m;
}

// ---- device-resident -----------------------------------------------------------------
int laser_b200_gemm_strided_f32_dev(int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                                    int64_t rsA, int64_t csA, const float *B, int64_t rsB,
                                    int64_t csB, float beta, float *C, int64_t rsC, int64_t csC,
                                    int path, void *stream) {
  return f32_dev(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, path, stream);
}
int laser_b200_gemm_strided_f32_epi_dev(int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                                        int64_t rsA, int64_t csA, const float *B, int64_t rsB,
                                        int64_t csB, float beta, float *C, int64_t rsC, int64_t csC,
                                        const laser_b200_epilogue *epi, int path, void *stream) {
  if (epi) {
    if (epi->activation < 0 || epi->activation > 3) return set_error(LASER_B200_EINVAL, "unknown activation %d", epi->activation);
    g_epi.bias = epi->bias;
    g_epi.bias_per_row = epi->bias_per_row ? 1 : 0;
    g_epi.act = epi->activation;
  }
  const int rc = f32_dev(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, path, stream);
  g_epi = Epilogue();
  return rc;
}
int laser_b200_gemm_strided_f64_dev(int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                                    int64_t rsA, int64_t csA, const double *B, int64_t rsB,
                                    int64_t csB, double beta, double *C, int64_t rsC, int64_t csC,
                                    void *stream) {
  return simt_dev<double>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, stream);
}
int laser_b200_gemm_strided_i32_dev(int64_t M, int64_t N, int64_t K, int32_t alpha, const int32_t *A,
                                    int64_t rsA, int64_t csA, const int32_t *B, int64_t rsB,
                                    int64_t csB, int32_t beta, int32_t *C, int64_t rsC, int64_t csC,
                                    void *stream) {
  return simt_dev<int32_t>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, stream);
}
int laser_b200_gemm_strided_i64_dev(int64_t M, int64_t N, int64_t K, int64_t alpha, const int64_t *A,
                                    int64_t rsA, int64_t csA, const int64_t *B, int64_t rsB,
                                    int64_t csB, int64_t beta, int64_t *C, int64_t rsC, int64_t csC,
                                    void *stream) {
  return simt_dev<int64_t>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, stream);
}
int laser_b200_gemm_strided_bf16_dev(int64_t M, int64_t N, int64_t K, float alpha, const uint16_t *A,
                                     int64_t rsA, int64_t csA, const uint16_t *B, int64_t rsB,
                                     int64_t csB, float beta, uint16_t *C, int64_t rsC, int64_t csC,
                                     void *stream) {
  return bf16_dev(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, stream);
}

// ---- host pointers (the drop-in signature, gemm.nim:184-193) ---------------------------
int laser_b200_gemm_strided_f32(int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                                int64_t rsA, int64_t csA, const float *B, int64_t rsB, int64_t csB,
                                float beta, float *C, int64_t rsC, int64_t csC) {
  // large tensor-core problems whose row panels are separate address ranges: overlap the
  // PCIe transfers with the compute, panel by panel
  if (M >= 2048 && N > 4 && K > 0 && A && B && C &&
      M <= 0x7fffffffLL && N <= 0x7fffffffLL && K <= 0x7fffffffLL) {
    Ctx *c;
    int rc = get_ctx(&c);
    if (rc) return rc;
    const int mode = g_f32_mode.load();
    const bool tc_mode = mode == LASER_B200_PATH_TF32X3 || mode == LASER_B200_PATH_TF32_BF16C ||
                         mode == LASER_B200_PATH_TF32X1 || mode == LASER_B200_PATH_BF16X3 || mode == LASER_B200_PATH_F16X3;
    if (tc_mode && panel_separable(c->panel_rows, K, rsA, csA) && panel_separable(c->panel_rows, N, rsC, csC) &&
        span_of(c->panel_rows, N, rsC, csC).dense && rsA > 0 && rsC > 0)
      return host_gemm_f32_pipelined(*c, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, mode);
  }
  return host_gemm<float>(M, N, K, A, rsA, csA, B, rsB, csB, beta == 0.0f, C, rsC, csC,
                          [&](const float *a, const float *b, float *c, void *s) {
                            return f32_dev(M, N, K, alpha, a, rsA, csA, b, rsB, csB, beta, c, rsC, csC,
                                           LASER_B200_PATH_AUTO, s);
                          });
}
int laser_b200_gemm_strided_f64(int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                                int64_t rsA, int64_t csA, const double *B, int64_t rsB, int64_t csB,
                                double beta, double *C, int64_t rsC, int64_t csC) {
  return host_gemm<double>(M, N, K, A, rsA, csA, B, rsB, csB, beta == 0.0, C, rsC, csC,
                           [&](const double *a, const double *b, double *c, void *s) {
                             return simt_dev<double>(M, N, K, alpha, a, rsA, csA, b, rsB, csB, beta, c,
                                                     rsC, csC, s);
                           });
}
int laser_b200_gemm_strided_i32(int64_t M, int64_t N, int64_t K, int32_t alpha, const int32_t *A,
                                int64_t rsA, int64_t csA, const int32_t *B, int64_t rsB, int64_t csB,
                                int32_t beta, int32_t *C, int64_t rsC, int64_t csC) {
  return host_gemm<int32_t>(M, N, K, A, rsA, csA, B, rsB, csB, beta == 0, C, rsC, csC,
                            [&](const int32_t *a, const int32_t *b, int32_t *c, void *s) {
                              return simt_dev<int32_t>(M, N, K, alpha, a, rsA, csA, b, rsB, csB, beta, c,
                                                       rsC, csC, s);
                            });
}
int laser_b200_gemm_strided_i64(int64_t M, int64_t N, int64_t K, int64_t alpha, const int64_t *A,
                                int64_t rsA, int64_t csA, const int64_t *B, int64_t rsB, int64_t csB,
                                int64_t beta, int64_t *C, int64_t rsC, int64_t csC) {
  return host_gemm<int64_t>(M, N, K, A, rsA, csA, B, rsB, csB, beta == 0, C, rsC, csC,
                            [&](const int64_t *a, const int64_t *b, int64_t *c, void *s) {
                              return simt_dev<int64_t>(M, N, K, alpha, a, rsA, csA, b, rsB, csB, beta, c,
                                                       rsC, csC, s);
                            });
}
int laser_b200_gemm_strided_bf16(int64_t M, int64_t N, int64_t K, float alpha, const uint16_t *A,
                                 int64_t rsA, int64_t csA, const uint16_t *B, int64_t rsB,
                                 int64_t csB, float beta, uint16_t *C, int64_t rsC, int64_t csC) {
  return host_gemm<uint16_t>(M, N, K, A, rsA, csA, B, rsB, csB, beta == 0.0f, C, rsC, csC,
                             [&](const uint16_t *a, const uint16_t *b, uint16_t *c, void *s) {
                               return bf16_dev(M, N, K, alpha, a, rsA, csA, b, rsB, csB, beta, c, rsC,
                                               csC, s);
                             });
}

// ---- pre-packed operands (gemm_prepacked.nim:63-292) --------------------------------------
size_t laser_b200_gemm_prepackA_mem_required_f32(int64_t M, int64_t N, int64_t K) {
  (void)N;
  return (M > 0 && K > 0) ? packed_layout(M, K).bytes : 0;
}
size_t laser_b200_gemm_prepackB_mem_required_f32(int64_t M, int64_t N, int64_t K) {
  (void)M;
  return (N > 0 && K > 0) ? packed_layout(N, K).bytes : 0;
}
int laser_b200_gemm_prepackA_f32_dev(void *dst, int64_t M, int64_t N, int64_t K, const float *A,
                                     int64_t rsA, int64_t csA, void *stream) {
  (void)N;
  return prepack_dev(dst, M, K, A, rsA, csA, stream);
}
int laser_b200_gemm_prepackB_f32_dev(void *dst, int64_t M, int64_t N, int64_t K, const float *B,
                                     int64_t rsB, int64_t csB, void *stream) {
  (void)M;
  return prepack_dev(dst, N, K, B, csB, rsB, stream);   // B seen as [n][k]
}
int laser_b200_gemm_packed_f32_dev(int64_t M, int64_t N, int64_t K, float alpha, const void *packedA,
                                   const void *packedB, float beta, float *C, int64_t rsC, int64_t csC,
                                   void *stream) {
  return gemm_packed_dev(M, N, K, alpha, nullptr, 0, 0, packedA, packedB, beta, C, rsC, csC, stream);
}
int laser_b200_gemm_packedB_f32_dev(int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                                    int64_t rsA, int64_t csA, const void *packedB, float beta, float *C,
                                    int64_t rsC, int64_t csC, void *stream) {
  if (!A) return set_error(LASER_B200_EINVAL, "null pointer");
  return gemm_packed_dev(M, N, K, alpha, A, rsA, csA, nullptr, packedB, beta, C, rsC, csC, stream);
}

// ---- storage -----------------------------------------------------------------------------
int laser_b200_malloc(void **dev_ptr, size_t bytes) {
  if (!dev_ptr) return set_error(LASER_B200_EINVAL, "null out pointer");
  Ctx *c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  *dev_ptr = nullptr;
  CUDA_TRY(cudaMalloc(dev_ptr, bytes ? bytes : 1));
  return LASER_B200_OK;
}
int laser_b200_free(void *dev_ptr) {
  if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
  return LASER_B200_OK;
}
int laser_b200_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes) {
  CUDA_TRY(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
  return LASER_B200_OK;
}
int laser_b200_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes) {
  CUDA_TRY(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return LASER_B200_OK;
}
int laser_b200_memset_zero(void *dst_dev, size_t bytes) {
  CUDA_TRY(cudaMemset(dst_dev, 0, bytes));
  return LASER_B200_OK;
}
int laser_b200_synchronize(void) {
  CUDA_TRY(cudaDeviceSynchronize());
  return LASER_B200_OK;
}

int laser_b200_matmul_views(const laser_b200_tensor_view *A, const laser_b200_tensor_view *B,
                            laser_b200_tensor_view *C, double alpha, double beta, int path,
                            void *stream) {
  if (!A || !B || !C) return set_error(LASER_B200_EINVAL, "null view");
  if (A->rank != 2 || B->rank != 2 || C->rank != 2)
    return set_error(LASER_B200_EINVAL, "matmul needs rank-2 views (got %d, %d, %d)", A->rank, B->rank, C->rank);
  if (A->dtype != B->dtype || A->dtype != C->dtype) return set_error(LASER_B200_EINVAL, "dtype mismatch");
  const int64_t M = A->shape[0], K = A->shape[1], N = B->shape[1];
  if (B->shape[0] != K || C->shape[0] != M || C->shape[1] != N)
    return set_error(LASER_B200_EINVAL, "shape mismatch: A %lldx%lld B %lldx%lld C %lldx%lld", (long long)M,
                     (long long)K, (long long)B->shape[0], (long long)N, (long long)C->shape[0],
                     (long long)C->shape[1]);
#define LB200_RAW(T, v) (static_cast<T *>((v)->storage) + (v)->offset)
  switch (A->dtype) {
    case 0:
      return f32_dev(M, N, K, (float)alpha, LB200_RAW(float, A), A->strides[0], A->strides[1],
                     LB200_RAW(float, B), B->strides[0], B->strides[1], (float)beta, LB200_RAW(float, C),
                     C->strides[0], C->strides[1], path, stream);
    case 1:
      return simt_dev<double>(M, N, K, alpha, LB200_RAW(double, A), A->strides[0], A->strides[1],
                              LB200_RAW(double, B), B->strides[0], B->strides[1], beta,
                              LB200_RAW(double, C), C->strides[0], C->strides[1], stream);
    case 2:
      return simt_dev<int32_t>(M, N, K, (int32_t)alpha, LB200_RAW(int32_t, A), A->strides[0],
                               A->strides[1], LB200_RAW(int32_t, B), B->strides[0], B->strides[1],
                               (int32_t)beta, LB200_RAW(int32_t, C), C->strides[0], C->strides[1], stream);
    case 3:
      return simt_dev<int64_t>(M, N, K, (int64_t)alpha, LB200_RAW(int64_t, A), A->strides[0],
                               A->strides[1], LB200_RAW(int64_t, B), B->strides[0], B->strides[1],
                               (int64_t)beta, LB200_RAW(int64_t, C), C->strides[0], C->strides[1], stream);
    case 4:
      return bf16_dev(M, N, K, (float)alpha, LB200_RAW(uint16_t, A), A->strides[0], A->strides[1],
                      LB200_RAW(uint16_t, B), B->strides[0], B->strides[1], (float)beta,
                      LB200_RAW(uint16_t, C), C->strides[0], C->strides[1], stream);
    default:
      return set_error(LASER_B200_EINVAL, "unknown dtype %d", A->dtype);
  }
#undef LB200_RAW
}

int laser_b200_debug_classify(int elem_size, const void *base, int64_t s_mn, int64_t s_k) {
  if (elem_size != 2 && elem_size != 4) return -1;
  Operand o{base, 1, 1, s_mn, s_k};
  return static_cast<int>(classify(o, elem_size));
}
int laser_b200_debug_span(int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride, int64_t *lo,
                          int64_t *hi, int *dense) {
  if (rows <= 0 || cols <= 0 || !lo || !hi || !dense) return LASER_B200_EINVAL;
  const Span sp = span_of(rows, cols, row_stride, col_stride);
  *lo = sp.lo; *hi = sp.hi; *dense = sp.dense ? 1 : 0;
  return LASER_B200_OK;
}

int laser_b200_fill_uniform_f32_dev(float *dst_dev, int64_t n, uint64_t seed, float lo, float hi,
                                    void *stream) {
  if (n <= 0) return LASER_B200_OK;
  if (!dst_dev) return set_error(LASER_B200_EINVAL, "null pointer");
  Ctx *c;
  int rc = get_ctx(&c);
  if (rc) return rc;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  fill_uniform_f32_kernel<<<grid_for(*c, (n + 255) / 256, 8), 256, 0, s>>>(dst_dev, n, seed, lo, hi);
  COUNT_LAUNCH();
  CHECK_LAUNCH();
  return finish(*c, static_cast<cudaStream_t>(stream), s);
}

}  // extern "C"

#include "capi_layers.inc"
