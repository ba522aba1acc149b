// layers_emu.cpp -- TEST INFRASTRUCTURE: the kernels of laser_b200/csrc/layers.cuh compiled for
// the host (cuda_emu.h) behind a small C interface for ctypes.  The launch geometry comes from the
// same host functions the library uses (transpose_can_vec / im2col_plan / copy_plan).
#include "cuda_emu.h"

#include "../../laser_b200/csrc/layers.cuh"

using namespace lb200;

template <typename T>
static int run_transpose(T *dst, const T *src, int64_t N, int64_t NR, int64_t NC, int grid, int force_scalar) {
  const bool vec = !force_scalar && transpose_can_vec<T>(dst, src, NR, NC);
  const int64_t tiles = transpose_tiles(N, NR, NC);
  if (grid <= 0 || grid > tiles) grid = static_cast<int>(tiles);
  if (vec) emu::launch(grid, 256, [=]() { transpose_batched_kernel<T, 4>(dst, src, N, NR, NC); });
  else emu::launch(grid, 256, [=]() { transpose_batched_kernel<T, 1>(dst, src, N, NR, NC); });
  return vec ? 4 : 1;
}

extern "C" {

// returns the vector width used (4 or 1), -1 on a bad element size
int emu_transpose_batched(int elem_size, void *dst, const void *src, int64_t N, int64_t NR, int64_t NC, int grid,
                          int force_scalar) {
  switch (elem_size) {
    case 1: return run_transpose(static_cast<uint8_t *>(dst), static_cast<const uint8_t *>(src), N, NR, NC, grid, force_scalar);
    case 2: return run_transpose(static_cast<uint16_t *>(dst), static_cast<const uint16_t *>(src), N, NR, NC, grid, force_scalar);
    case 4: return run_transpose(static_cast<uint32_t *>(dst), static_cast<const uint32_t *>(src), N, NR, NC, grid, force_scalar);
    case 8: return run_transpose(static_cast<uint64_t *>(dst), static_cast<const uint64_t *>(src), N, NR, NC, grid, force_scalar);
  }
  return -1;
}

// geom = {C, H, W, kH, kW, pH, pW, sH, sW, outH, outW}; returns 4 (float4 stores) or 1
int emu_im2col(float *workspace, const float *input, int64_t images, const int64_t geom[11], int force_scalar) {
  Im2colParams p;
  bool vec;
  const int64_t blocks = im2col_plan(geom, images, workspace, &p, &vec);
  if (force_scalar) vec = false;
  if (vec) emu::launch(static_cast<unsigned>(blocks), 256, [=]() { im2col_kernel<true>(workspace, input, p); });
  else emu::launch(static_cast<unsigned>(blocks), 256, [=]() { im2col_kernel<false>(workspace, input, p); });
  return vec ? 4 : 1;
}

// returns the rank left after merging dimensions
int emu_copy_strided(int elem_size, void *dst, const void *src, int rank, const int64_t *shape,
                     const int64_t *dst_strides, const int64_t *src_strides, int grid) {
  CopyParams p;
  copy_plan(rank, shape, dst_strides, src_strides, &p);
  if (p.total == 0) return p.rank;
  if (elem_size == 4)
    emu::launch(grid, 256, [=]() { copy_strided_kernel<uint32_t>(static_cast<uint32_t *>(dst), static_cast<const uint32_t *>(src), p); });
  else if (elem_size == 8)
    emu::launch(grid, 256, [=]() { copy_strided_kernel<uint64_t>(static_cast<uint64_t *>(dst), static_cast<const uint64_t *>(src), p); });
  else if (elem_size == 2)
    emu::launch(grid, 256, [=]() { copy_strided_kernel<uint16_t>(static_cast<uint16_t *>(dst), static_cast<const uint16_t *>(src), p); });
  else return -1;
  return p.rank;
}

}  // extern "C"
