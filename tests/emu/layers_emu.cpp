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

// forEach opcodes: strides[t] may be NULL for operands the opcode does not read; returns the merged rank
template <typename T>
static int run_foreach(int op, T *o, const T *x, const T *y, const T *z, int rank, const int64_t *shape,
                       const int64_t *so, const int64_t *sx, const int64_t *sy, const int64_t *sz, double alpha, int grid) {
  const int64_t *strides[4] = {so, sx, sy, sz};
  ForeachParams p;
  foreach_plan(rank, shape, strides, &p);
  if (p.total == 0) return p.rank;
  const T a = static_cast<T>(alpha);
#define FE(OP) case OP: emu::launch(grid, 256, [=]() { foreach_strided_kernel<T, OP>(o, x, y, z, p, a); }); break
  switch (op) {
    FE(FE_COPY); FE(FE_FILL); FE(FE_SCALE); FE(FE_ADD); FE(FE_SUB); FE(FE_MUL); FE(FE_FMA); FE(FE_AXPY); FE(FE_BENCH);
    default: return -1;
  }
#undef FE
  return p.rank;
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

int emu_foreach(int elem_size, int op, void *o, const void *x, const void *y, const void *z, int rank, const int64_t *shape,
                const int64_t *so, const int64_t *sx, const int64_t *sy, const int64_t *sz, double alpha, int grid) {
  if (foreach_operands(op) < 0) return -1;
  if (elem_size == 4)
    return run_foreach(op, static_cast<float *>(o), static_cast<const float *>(x), static_cast<const float *>(y),
                       static_cast<const float *>(z), rank, shape, so, sx, sy, sz, alpha, grid);
  if (elem_size == 8)
    return run_foreach(op, static_cast<double *>(o), static_cast<const double *>(x), static_cast<const double *>(y),
                       static_cast<const double *>(z), rank, shape, so, sx, sy, sz, alpha, grid);
  return -1;
}

}  // extern "C"
