// cuda_emu.h -- TEST INFRASTRUCTURE: runs a __global__ function of the product on host threads so
// that the CPU test suite can execute the kernels' index arithmetic, bounds handling and
// shared-memory choreography without a GPU (one std::thread per CUDA thread of a block, a pthread
// barrier for __syncthreads, blocks one after the other).  Only kernels written in plain CUDA C++
// (no inline PTX; of the warp intrinsics only full-mask __shfl_xor_sync) can be run this way: layers.cuh, gemm_simt_kernel, split.cuh (with a software tf32 rounding).
// It is a test of the product's source, not a fallback: nothing under laser_b200/ includes it.
#pragma once

#include <cuda_runtime.h>  // host-side definitions of float4 / make_float4
#include <pthread.h>
#include <stdint.h>

#include <thread>
#include <vector>

#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __forceinline__
#define __forceinline__ inline
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __shared__
#define __shared__ static  // one block at a time: a function-local static is the block's shared memory

namespace emu {
struct Idx {
  unsigned x, y, z;
};
inline thread_local Idx t_idx{0, 0, 0};
inline thread_local Idx b_idx{0, 0, 0};
inline Idx b_dim{1, 1, 1}, g_dim{1, 1, 1};
inline pthread_barrier_t barrier;
inline pthread_barrier_t warp_barrier[32];     // one per warp of the block (warp shuffles)
inline float warp_scratch[1024];
alignas(16) inline unsigned char dyn_smem[96 * 1024];   // dynamic shared memory of the block

// kernels whose threads all reach every __syncthreads (or that have none) only
template <typename Body>
void launch(unsigned grid, unsigned block, Body body) {
  g_dim = Idx{grid, 1, 1};
  b_dim = Idx{block, 1, 1};
  pthread_barrier_init(&barrier, nullptr, block);
  for (unsigned w = 0; w < block / 32; ++w) pthread_barrier_init(&warp_barrier[w], nullptr, 32);
  std::vector<std::thread> threads;
  threads.reserve(block);
  for (unsigned t = 0; t < block; ++t)
    threads.emplace_back([=]() {
      t_idx = Idx{t, 0, 0};
      for (unsigned b = 0; b < grid; ++b) {
        b_idx = Idx{b, 0, 0};
        body();
        pthread_barrier_wait(&barrier);  // next block only when this one is done (shared memory reuse)
      }
    });
  for (auto &th : threads) th.join();
  pthread_barrier_destroy(&barrier);
  for (unsigned w = 0; w < block / 32; ++w) pthread_barrier_destroy(&warp_barrier[w]);
}
}  // namespace emu

// correctly rounded single operations (compile with -ffp-contract=off so that a * b + c stays unfused)
#include <cmath>
inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
template <typename T>
inline T __ldg(const T *p) { return *p; }
#include <cstring>
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __fsub_rn(float a, float b) { return a - b; }

#define threadIdx (emu::t_idx)
#define blockIdx (emu::b_idx)
#define blockDim (emu::b_dim)
#define gridDim (emu::g_dim)
inline void __syncthreads() { pthread_barrier_wait(&emu::barrier); }
// full-mask butterfly shuffle: every lane of the warp must call it (true for the reductions it is used in)
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const unsigned t = emu::t_idx.x, w = t >> 5;
  emu::warp_scratch[t] = v;
  pthread_barrier_wait(&emu::warp_barrier[w]);
  const float r = emu::warp_scratch[t ^ static_cast<unsigned>(lane_mask)];
  pthread_barrier_wait(&emu::warp_barrier[w]);
  return r;
}
