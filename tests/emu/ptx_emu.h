// ptx_emu.h -- TEST INFRASTRUCTURE: a functional model of the PTX that laser_b200/csrc/ptx.cuh
// wraps (mbarrier, cp.async.bulk.tensor, tcgen05.alloc/mma/commit/ld, clusters of two CTAs), under
// the same names, so that gemm_tc.cuh -- the tcgen05 kernel -- can be compiled by g++ and run on host
// threads (cuda_emu.h).  What the model checks: the producer / MMA / epilogue protocol (barrier
// counts, phases, stage rings: a protocol error shows up as a deadlock or as a wrong sum), the tile
// scheduler, raster and split-K ranges, the k-block bookkeeping of every mode, operand addressing
// through the descriptors' start / LBO / SBO fields, out-of-bounds zero fill, the CTA-pair
// ownership of rows and columns, and every epilogue path.  What it cannot check: anything that is
// a property of the silicon -- the 128-byte swizzle patterns (tiles are kept unswizzled here, on
// both the TMA and the MMA side), instruction encodings, memory-proxy fences, register limits,
// the accumulator's rounding (modelled: tf32 operands truncated, exact products, one rounding per
// instruction).  Those are covered by the -m gpu tests only.
#pragma once

#include <cuda.h>
#include <stdint.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

// dynamic shared memory of the kernel: this CTA's buffer
#define LB200_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(emu::dyn_smem_ptr())

namespace emu {

// what the host side of the harness puts inside the 128 opaque bytes of a CUtensorMap: the
// arguments of cuTensorMapEncodeTiled for a 2-d tensor
struct TensorMap2D {
  uint64_t magic;
  const unsigned char *base;
  int64_t dim0, dim1;        // extents in elements, dim0 innermost
  int64_t stride1_bytes;     // pitch of dim1
  int32_t esz, box0, box1;   // element size, box extents in elements
  int64_t dim2 = 1, stride2_bytes = 0;   // rank 3: dim2 matrices, stride2 apart (box depth 1)
};
constexpr uint64_t kMapMagic = 0x4c42323030544d41ull;
static_assert(sizeof(TensorMap2D) <= sizeof(CUtensorMap), "fits in the opaque struct");

struct MBarrier {
  int init_count = 0, pending = 0;
  long tx = 0;
  unsigned phase = 0;
};
inline std::mutex mb_mu;
inline std::condition_variable mb_cv;
inline std::map<const void *, MBarrier> mbars;
inline float tmem[kMaxCluster][128][512];   // TMEM of each CTA: 128 lanes x 512 columns of 32 bits

inline void reset_state() {
  std::lock_guard<std::mutex> lk(mb_mu);
  mbars.clear();
}
inline void mb_check(MBarrier &b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1u;
    b.pending = b.init_count;
    mb_cv.notify_all();
  }
}
inline MBarrier &mb_get(const void *bar) {
  auto it = mbars.find(bar);
  if (it == mbars.end()) { std::fprintf(stderr, "emu: mbarrier %p used before init\n", bar); std::abort(); }
  return it->second;
}
inline void mb_arrive(const void *bar, long expect_tx) {
  std::lock_guard<std::mutex> lk(mb_mu);
  MBarrier &b = mb_get(bar);
  b.tx += expect_tx;
  if (--b.pending < 0) { std::fprintf(stderr, "emu: mbarrier %p over-arrived\n", bar); std::abort(); }
  mb_check(b);
}
inline void mb_complete_tx(const void *bar, long bytes) {
  std::lock_guard<std::mutex> lk(mb_mu);
  MBarrier &b = mb_get(bar);
  b.tx -= bytes;
  mb_check(b);
}
// the same offset in the shared memory of CTA `rank` of the cluster
template <typename T>
inline T *peer_ptr(T *p, unsigned rank) {
  const unsigned char *q = reinterpret_cast<const unsigned char *>(p);
  for (unsigned r = 0; r < kMaxCluster; ++r)
    if (q >= dyn_smem[r] && q < dyn_smem[r] + kDynSmemBytes)
      return reinterpret_cast<T *>(dyn_smem[rank] + (q - dyn_smem[r]));
  std::fprintf(stderr, "emu: pointer %p is not in shared memory\n", static_cast<const void *>(p));
  std::abort();
}
}  // namespace emu

namespace lb200 {
namespace ptx {

// shared-window address: offset inside the CTA's shared memory, CTA rank in bit 24
inline uint32_t smem_u32(const void *p) {
  const unsigned char *q = static_cast<const unsigned char *>(p);
  for (unsigned r = 0; r < emu::kMaxCluster; ++r)
    if (q >= emu::dyn_smem[r] && q < emu::dyn_smem[r] + emu::kDynSmemBytes)
      return static_cast<uint32_t>(q - emu::dyn_smem[r]) | (r << 24);
  std::fprintf(stderr, "emu: smem_u32 of a pointer outside shared memory\n");
  std::abort();
}
inline bool elect_one() { return (emu::t_idx.x & 31) == 0; }

// ------------------------------------------------------------------ mbarrier
inline void mbar_init(uint64_t *bar, uint32_t count) {
  std::lock_guard<std::mutex> lk(emu::mb_mu);
  emu::MBarrier b;
  b.init_count = b.pending = static_cast<int>(count);
  emu::mbars[bar] = b;
}
inline void fence_barrier_init() {}
inline void fence_proxy_async_smem() {}
inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { emu::mb_arrive(bar, bytes); }
inline void mbar_arrive(uint64_t *bar) { emu::mb_arrive(bar, 0); }
inline bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  std::lock_guard<std::mutex> lk(emu::mb_mu);
  return emu::mb_get(bar).phase != parity;
}
inline void mbar_wait(uint64_t *bar, uint32_t parity) {   // blocking (hundreds of waiters on a few cores)
  std::unique_lock<std::mutex> lk(emu::mb_mu);
  emu::MBarrier &b = emu::mb_get(bar);
  emu::mb_cv.wait(lk, [&]() { return b.phase != parity; });
}

// ----------------------------------------------------------------------- TMA
inline void prefetch_tensormap(const CUtensorMap *) {}
inline void prefetch_l2(const void *) {}
template <int N> inline void setmaxnreg_inc() {}
template <int N> inline void setmaxnreg_dec() {}
inline void tma_copy_box(void *smem_dst, const CUtensorMap *map, int32_t c0, int32_t c1, long *bytes, int32_t c2 = 0) {
  emu::TensorMap2D m;
  std::memcpy(&m, map, sizeof m);
  if (m.magic != emu::kMapMagic) { std::fprintf(stderr, "emu: not an emulated tensor map\n"); std::abort(); }
  unsigned char *dst = static_cast<unsigned char *>(smem_dst);
  {   // the destination must be this CTA's shared memory, whole box inside it, swizzle-atom (1024 B) aligned
    const unsigned char *lo = emu::dyn_smem[emu::cta_rank];
    const size_t box_bytes = static_cast<size_t>(m.box0) * m.box1 * m.esz;
    if (dst < lo || dst + box_bytes > lo + emu::kDynSmemBytes || ((dst - lo) & 1023)) {
      std::fprintf(stderr, "emu: TMA destination outside this CTA's shared memory or not 1024-byte aligned\n");
      std::abort();
    }
  }
  const bool z_inside = c2 >= 0 && c2 < m.dim2;
  m.base += static_cast<int64_t>(z_inside ? c2 : 0) * m.stride2_bytes;
  if (!z_inside) m.dim1 = 0;   // a matrix outside the batch reads as zeros
  for (int r = 0; r < m.box1; ++r)
    for (int e = 0; e < m.box0; ++e) {
      const int64_t i0 = static_cast<int64_t>(c0) + e, i1 = static_cast<int64_t>(c1) + r;
      unsigned char *d = dst + (static_cast<size_t>(r) * m.box0 + e) * m.esz;
      if (i0 >= 0 && i0 < m.dim0 && i1 >= 0 && i1 < m.dim1) std::memcpy(d, m.base + i1 * m.stride1_bytes + i0 * m.esz, m.esz);
      else std::memset(d, 0, m.esz);   // out-of-bounds elements are zero-filled (and still counted)
    }
  *bytes = static_cast<long>(m.box0) * m.box1 * m.esz;
}
inline void cp_async_8(void *smem_dst, const void *gsrc, bool valid) {   // completes at once: wait_group has nothing to wait for
  if (valid) std::memcpy(smem_dst, gsrc, 8);
  else std::memset(smem_dst, 0, 8);
}
inline void cp_async_16(void *smem_dst, const void *gsrc, uint32_t src_bytes) {
  if ((reinterpret_cast<uintptr_t>(smem_dst) | reinterpret_cast<uintptr_t>(gsrc)) & 15u) std::abort();   // the hardware faults
  std::memset(smem_dst, 0, 16);
  if (src_bytes) std::memcpy(smem_dst, gsrc, src_bytes);
}
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
inline int sw128_chunk(int, int j) { return j; }      // the model keeps tiles unswizzled on both sides
// cp.async.bulk.tensor store: the box leaves at once (commit / wait have nothing to track); elements outside the tensor are
// not written
inline void tma_store_2d(const CUtensorMap *map, const void *smem_src, int32_t c0, int32_t c1) {
  emu::TensorMap2D m;
  std::memcpy(&m, map, sizeof m);
  if (m.magic != emu::kMapMagic) { std::fprintf(stderr, "emu: not an emulated tensor map\n"); std::abort(); }
  const unsigned char *src = static_cast<const unsigned char *>(smem_src);
  const unsigned char *lo = emu::dyn_smem[emu::cta_rank];
  const size_t box_bytes = static_cast<size_t>(m.box0) * m.box1 * m.esz;
  if (src < lo || src + box_bytes > lo + emu::kDynSmemBytes || ((src - lo) & 1023)) {
    std::fprintf(stderr, "emu: TMA store source outside this CTA's shared memory or not 1024-byte aligned\n");
    std::abort();
  }
  for (int r = 0; r < m.box1; ++r)
    for (int e = 0; e < m.box0; ++e) {
      const int64_t i0 = static_cast<int64_t>(c0) + e, i1 = static_cast<int64_t>(c1) + r;
      if (i0 >= 0 && i0 < m.dim0 && i1 >= 0 && i1 < m.dim1)
        std::memcpy(const_cast<unsigned char *>(m.base) + i1 * m.stride1_bytes + i0 * m.esz, src + (static_cast<size_t>(r) * m.box0 + e) * m.esz, m.esz);
    }
}
inline void tma_store_commit() {}
template <int N> inline void tma_store_wait_read() {}
template <int N> inline void tma_store_wait() {}
inline void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  if ((reinterpret_cast<uintptr_t>(smem_dst) | reinterpret_cast<uintptr_t>(gsrc) | bytes) & 15u) std::abort();   // the hardware faults
  std::memcpy(smem_dst, gsrc, bytes);
  emu::mb_complete_tx(bar, bytes);
}
inline void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1) {
  long bytes;
  tma_copy_box(smem_dst, map, c0, c1, &bytes);
  emu::mb_complete_tx(bar, bytes);
}

inline void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1, int32_t c2) {
  long bytes;
  tma_copy_box(smem_dst, map, c0, c1, &bytes, c2);
  emu::mb_complete_tx(bar, bytes);
}

constexpr uint64_t kEvictNormal = 0x1000000000000000ull, kEvictFirst = 0x12F0000000000000ull, kEvictLast = 0x14F0000000000000ull;
inline void tma_load_2d_hint(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1, uint64_t) {
  tma_load_2d(smem_dst, map, bar, c0, c1);   // cache policies have no functional effect
}

// ------------------------------------------------------------------- tcgen05
inline void tc_fence_before_sync() {}
inline void tc_fence_after_sync() {}
template <uint32_t NCOLS> inline void tmem_alloc(uint32_t *smem_dst) { *smem_dst = 0; }   // whole warp, same value
template <uint32_t NCOLS> inline void tmem_dealloc(uint32_t) {}

constexpr uint32_t kLayoutSw128 = 2;
constexpr uint32_t kLayoutSw128Base32 = 1;
inline uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2;
constexpr uint32_t make_idesc(uint32_t fmt, uint32_t a_mn_major, uint32_t b_mn_major, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// one operand element as fp32: tile kept UNSWIZZLED in shared memory, addressed through the
// descriptor fields the way the canonical layouts define them:
//   K-major : row i at (i / 8) * SBO + (i % 8) * 128, k inside the 128-byte row
//   MN-major: 128-byte chunk c of the mn extent at c * LBO, k-row kk at (kk / R) * SBO + (kk % R) * 128,
//             R = 4 k-rows per atom for the 32-byte-atom layout, 8 otherwise
inline float operand_elem(const unsigned char *cta_smem, uint64_t desc, bool mn_major, int E, uint32_t fmt, int i, int kk) {
  const uint32_t start = static_cast<uint32_t>(desc & 0x3FFF) << 4;
  const uint32_t lbo = static_cast<uint32_t>((desc >> 16) & 0x3FFF) << 4;
  const uint32_t sbo = static_cast<uint32_t>((desc >> 32) & 0x3FFF) << 4;
  const uint32_t layout = static_cast<uint32_t>(desc >> 61);
  size_t off;
  if (!mn_major) {
    off = start + static_cast<size_t>(i / 8) * sbo + static_cast<size_t>(i % 8) * 128 + static_cast<size_t>(kk) * E;
  } else {
    const int per_chunk = 128 / E, R = (layout == kLayoutSw128Base32) ? 4 : 8;
    off = start + static_cast<size_t>(i / per_chunk) * lbo + static_cast<size_t>(kk / R) * sbo +
          static_cast<size_t>(kk % R) * 128 + static_cast<size_t>(i % per_chunk) * E;
  }
  if (off + E > emu::kDynSmemBytes) { std::fprintf(stderr, "emu: operand read outside shared memory\n"); std::abort(); }
  if (E == 4) {
    uint32_t u;
    std::memcpy(&u, cta_smem + off, 4);
    if (fmt == kFmtTF32) u &= 0xffffe000u;   // kind::tf32 ignores the low 13 mantissa bits
    return __uint_as_float(u);
  }
  uint16_t h;
  std::memcpy(&h, cta_smem + off, 2);
  if (fmt == kFmtF16) {   // IEEE binary16 (subnormals included), exactly representable in fp32
    const int e = (h >> 10) & 31, m = h & 1023;
    float v;
    if (e == 0) v = std::ldexp(static_cast<float>(m), -24);
    else if (e == 31) v = m ? std::nanf("") : HUGE_VALF;
    else v = std::ldexp(static_cast<float>(1024 + m), e - 25);
    return (h & 0x8000) ? -v : v;
  }
  if (fmt != kFmtBF16) { std::fprintf(stderr, "emu: unknown 16-bit operand format\n"); std::abort(); }
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}

// D[tmem] (+)= A * B for one instruction (K = 32 bytes per operand row).  ncta = 1: M x N from this
// CTA's shared memory into this CTA's TMEM.  ncta = 2: rows [0,128) of A and of D belong to CTA 0,
// rows [128,256) to CTA 1; columns [0,N/2) of B come from CTA 0, [N/2,N) from CTA 1.
inline void mma_model(int ncta, int E, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t fmt = (idesc >> 7) & 7;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  const int N = static_cast<int>((idesc >> 17) & 63) << 3, M = static_cast<int>((idesc >> 24) & 31) << 4;
  const int KI = 32 / E;
  if (M != 128 * ncta || N > 256 || N % 16) { std::fprintf(stderr, "emu: unsupported MMA shape %d x %d\n", M, N); std::abort(); }
  const uint32_t col0 = d_tmem & 0xFFFF, lane0 = d_tmem >> 16;
  if (lane0 != 0 || col0 + N > 512) { std::fprintf(stderr, "emu: accumulator outside TMEM\n"); std::abort(); }
  const unsigned self = emu::cta_rank;
  static thread_local float a[256][16], b[256][16];
  for (int i = 0; i < M; ++i) {
    const unsigned char *sm = emu::dyn_smem[ncta == 2 ? i / 128 : self];
    for (int kk = 0; kk < KI; ++kk) a[i][kk] = operand_elem(sm, a_desc, a_mn, E, fmt, i % 128, kk);
  }
  for (int j = 0; j < N; ++j) {
    const unsigned char *sm = emu::dyn_smem[ncta == 2 ? j / (N / 2) : self];
    const int jj = ncta == 2 ? j % (N / 2) : j;
    for (int kk = 0; kk < KI; ++kk) b[j][kk] = operand_elem(sm, b_desc, b_mn, E, fmt, jj, kk);
  }
  for (int i = 0; i < M; ++i) {
    float *drow = emu::tmem[ncta == 2 ? i / 128 : self][i % 128] + col0;
    for (int j = 0; j < N; ++j) {
      double s = accumulate ? static_cast<double>(drow[j]) : 0.0;
      for (int kk = 0; kk < KI; ++kk) s += static_cast<double>(a[i][kk]) * static_cast<double>(b[j][kk]);
      drow[j] = static_cast<float>(s);
    }
  }
}
inline void mma_tf32_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { mma_model(1, 4, d, ad, bd, idesc, acc); }
inline void mma_f16_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { mma_model(1, 2, d, ad, bd, idesc, acc); }
// the instruction completes before the call returns, so a commit is an immediate arrival
inline void mma_commit(uint64_t *bar) { emu::mb_arrive(bar, 0); }

inline void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  const uint32_t lane = (taddr >> 16) + (emu::t_idx.x & 31), col = taddr & 0xFFFF;
  if (lane >= 128 || col + 16 > 512) { std::fprintf(stderr, "emu: tcgen05.ld outside TMEM\n"); std::abort(); }
  if ((taddr >> 16) / 32 != ((emu::t_idx.x >> 5) & 3)) {   // a warp may only touch its own lane quarter
    std::fprintf(stderr, "emu: warp %u reads TMEM lanes of quarter %u\n", emu::t_idx.x >> 5, (taddr >> 16) / 32);
    std::abort();
  }
  for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(emu::tmem[emu::cta_rank][lane][col + j]);
}
inline void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  uint32_t lo[16], hi[16];
  tmem_ld_32x32b_x16(taddr, lo);
  tmem_ld_32x32b_x16(taddr + 16, hi);
  for (int j = 0; j < 16; ++j) { r[j] = lo[j]; r[16 + j] = hi[j]; }
}
inline void tmem_ld_wait(uint32_t (&)[16]) {}
inline void tmem_ld_wait(uint32_t (&)[32]) {}

// ------------------------------------------------------------ CTA pairs (cta_group::2)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
inline uint32_t cluster_ctarank() { return emu::cta_rank; }
inline void cluster_sync() { pthread_barrier_wait(&emu::cluster_barrier); }
inline void mbar_arrive_leader(uint64_t *bar) { emu::mb_arrive(emu::peer_ptr(bar, 0), 0); }
inline void mbar_arrive_cluster(uint64_t *bar, uint32_t rank) { emu::mb_arrive(emu::peer_ptr(bar, rank), 0); }
inline void mbar_arrive_cluster_relaxed(uint64_t *bar, uint32_t rank, uint32_t count) {
  for (uint32_t i = 0; i < count; ++i) emu::mb_arrive(emu::peer_ptr(bar, rank), 0);
}
inline void mbar_wait_cluster(uint64_t *bar, uint32_t parity) { mbar_wait(bar, parity); }
inline void st_shared_cluster_s32(int *p, uint32_t rank, int v) { *reinterpret_cast<volatile int *>(emu::peer_ptr(p, rank)) = v; }
inline void griddep_wait() {}
inline void griddep_launch_dependents() {}
inline void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1) {
  long bytes;
  tma_copy_box(smem_dst, map, c0, c1, &bytes);               // into THIS CTA's shared memory
  emu::mb_complete_tx(emu::peer_ptr(bar, 0), bytes);        // bytes credited to the LEADER's barrier
}
inline void tma_load_2d_pair_hint(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1, uint64_t) {
  tma_load_2d_pair(smem_dst, map, bar, c0, c1);
}
inline void tma_load_3d_pair(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1, int32_t c2) {
  long bytes;
  tma_copy_box(smem_dst, map, c0, c1, &bytes, c2);
  emu::mb_complete_tx(emu::peer_ptr(bar, 0), bytes);
}
template <uint32_t NCOLS> inline void tmem_alloc_pair(uint32_t *smem_dst) { *smem_dst = 0; }
template <uint32_t NCOLS> inline void tmem_dealloc_pair(uint32_t) {}
inline void mma_tf32_ss_pair(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { mma_model(2, 4, d, ad, bd, idesc, acc); }
inline void mma_f16_ss_pair(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { mma_model(2, 2, d, ad, bd, idesc, acc); }
inline void mma_commit_pair(uint64_t *bar) {   // one arrival on the barrier at this offset in BOTH CTAs
  emu::mb_arrive(emu::peer_ptr(bar, 0), 0);
  emu::mb_arrive(emu::peer_ptr(bar, 1), 0);
}

}  // namespace ptx
// mma.sync.aligned.m8n8k4.row.col.f64 (gemm_dmma.cuh): lane l holds A[l / 4][l % 4], B[l % 4][l / 4] and the two accumulators
// D[l / 4][2 * (l % 4) + {0, 1}]; per output element the four steps of the FMA chain in k order
inline void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  const unsigned t = emu::t_idx.x, w = t >> 5, r = emu::cta_rank, base = t & ~31u, lane = t & 31u;
  emu::warp_scratch_d[r][0][t] = a;
  emu::warp_scratch_d[r][1][t] = b;
  pthread_barrier_wait(&emu::warp_barrier[r][w]);
  const unsigned row = lane >> 2, c0 = 2 * (lane & 3);
  for (unsigned k = 0; k < 4; ++k) {
    const double ak = emu::warp_scratch_d[r][0][base + row * 4 + k];
    d0 = std::fma(ak, emu::warp_scratch_d[r][1][base + c0 * 4 + k], d0);
    d1 = std::fma(ak, emu::warp_scratch_d[r][1][base + (c0 + 1) * 4 + k], d1);
  }
  pthread_barrier_wait(&emu::warp_barrier[r][w]);
}
}  // namespace lb200
