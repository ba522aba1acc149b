// laser_b200.hpp -- header-only C++ host mirror of the reference's interface for the hot path,
// on top of the C ABI (laser_b200.h).  The reference is Nim, which compiles to C/C++; no Nim
// toolchain exists in the build image, so this is the compiled-language host side: same names,
// argument order and meaning as the Nim procs, errors surfaced as exceptions (the Nim shim in
// nim/laser_b200.nim does the same with LaserB200Error).
//
//   laser::gemm_strided<T>(M, N, K, alpha, A, rowStrideA, colStrideA, B, rowStrideB, colStrideB,
//                          beta, C, rowStrideC, colStrideC)
//        == proc gemm_strided*[T: SomeNumber](...)   laser/primitives/matrix_multiplication/gemm.nim:184-193
//   laser::CudaTensor<T>  (shape / strides / offset / storage; rank, size, is_C_contiguous,
//                          unsafe_raw_data)           laser/tensor/datatypes.nim:12-88
//   laser::newTensor<T>, toTensor<T>, toHost          laser/tensor/initialization.nim:156-202
//   laser::gemm_prepackA/B_mem_required, gemm_prepackA/B, gemm_packed
//                                                     .../gemm_prepacked.nim:63-292
#pragma once

#include <cstdint>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "laser_b200.h"

namespace laser {

struct LaserB200Error : std::runtime_error {
  int code;
  LaserB200Error(int c, const char *msg) : std::runtime_error(std::string("laser_b200: ") + msg), code(c) {}
};
inline void check(int rc) {
  if (rc != LASER_B200_OK) throw LaserB200Error(rc, laser_b200_last_error());
}

// ---- gemm_strided: host pointers, synchronous (gemm.nim:184-193) -------------------------
inline void gemm_strided(int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t rowStrideA,
                         int64_t colStrideA, const float *B, int64_t rowStrideB, int64_t colStrideB, float beta,
                         float *C, int64_t rowStrideC, int64_t colStrideC) {
  check(laser_b200_gemm_strided_f32(M, N, K, alpha, A, rowStrideA, colStrideA, B, rowStrideB, colStrideB, beta, C,
                                    rowStrideC, colStrideC));
}
inline void gemm_strided(int64_t M, int64_t N, int64_t K, double alpha, const double *A, int64_t rowStrideA,
                         int64_t colStrideA, const double *B, int64_t rowStrideB, int64_t colStrideB, double beta,
                         double *C, int64_t rowStrideC, int64_t colStrideC) {
  check(laser_b200_gemm_strided_f64(M, N, K, alpha, A, rowStrideA, colStrideA, B, rowStrideB, colStrideB, beta, C,
                                    rowStrideC, colStrideC));
}
inline void gemm_strided(int64_t M, int64_t N, int64_t K, int32_t alpha, const int32_t *A, int64_t rowStrideA,
                         int64_t colStrideA, const int32_t *B, int64_t rowStrideB, int64_t colStrideB, int32_t beta,
                         int32_t *C, int64_t rowStrideC, int64_t colStrideC) {
  check(laser_b200_gemm_strided_i32(M, N, K, alpha, A, rowStrideA, colStrideA, B, rowStrideB, colStrideB, beta, C,
                                    rowStrideC, colStrideC));
}
inline void gemm_strided(int64_t M, int64_t N, int64_t K, int64_t alpha, const int64_t *A, int64_t rowStrideA,
                         int64_t colStrideA, const int64_t *B, int64_t rowStrideB, int64_t colStrideB, int64_t beta,
                         int64_t *C, int64_t rowStrideC, int64_t colStrideC) {
  check(laser_b200_gemm_strided_i64(M, N, K, alpha, A, rowStrideA, colStrideA, B, rowStrideB, colStrideB, beta, C,
                                    rowStrideC, colStrideC));
}

// ---- device tensor honouring the tensor contract (datatypes.nim:12-88) -------------------
constexpr int LASER_MAXRANK = LASER_B200_MAXRANK;  // laser/dynamic_stack_arrays.nim:6

template <typename T> struct dtype_code;
template <> struct dtype_code<float> { static constexpr int value = 0; };
template <> struct dtype_code<double> { static constexpr int value = 1; };
template <> struct dtype_code<int32_t> { static constexpr int value = 2; };
template <> struct dtype_code<int64_t> { static constexpr int value = 3; };

template <typename T>
struct CudaStorage {  // CpuStorage analogue: raw_buffer + ownership (datatypes.nim:24-30)
  T *raw_buffer = nullptr;
  bool memowner = false;
  explicit CudaStorage(size_t n) {
    void *p = nullptr;
    check(laser_b200_malloc(&p, n * sizeof(T)));
    check(laser_b200_memset_zero(p, n * sizeof(T)));
    raw_buffer = static_cast<T *>(p);
    memowner = true;
  }
  ~CudaStorage() {
    if (memowner && raw_buffer) laser_b200_free(raw_buffer);
  }
  CudaStorage(const CudaStorage &) = delete;
  CudaStorage &operator=(const CudaStorage &) = delete;
};

template <typename T>
struct CudaTensor {
  std::vector<int64_t> shape, strides;  // strides in elements
  int64_t offset = 0;
  std::shared_ptr<CudaStorage<T>> storage;  // reference semantics, like the `ref object` it mirrors

  int rank() const { return static_cast<int>(shape.size()); }
  int64_t size() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
  bool is_C_contiguous() const {  // datatypes.nim:37-47
    int64_t cur = 1;
    for (int i = rank() - 1; i >= 0; --i) {
      if (shape[i] != 1 && strides[i] != cur) return false;
      cur *= shape[i];
    }
    return true;
  }
  T *unsafe_raw_data() const { return storage->raw_buffer + offset; }  // datatypes.nim:64-88
  CudaTensor transpose() const {
    CudaTensor t = *this;
    t.shape = {shape[1], shape[0]};
    t.strides = {strides[1], strides[0]};
    return t;
  }
  laser_b200_tensor_view view() const {
    laser_b200_tensor_view v{};
    v.rank = rank();
    v.dtype = dtype_code<T>::value;
    for (int i = 0; i < rank(); ++i) { v.shape[i] = shape[i]; v.strides[i] = strides[i]; }
    v.offset = offset;
    v.storage = storage->raw_buffer;
    return v;
  }
};

template <typename T>
CudaTensor<T> newTensor(std::initializer_list<int64_t> shape) {  // zero-initialised, row-major
  if (shape.size() > static_cast<size_t>(LASER_MAXRANK)) throw std::invalid_argument("rank > LASER_MAXRANK");
  CudaTensor<T> t;
  t.shape.assign(shape.begin(), shape.end());
  t.strides.assign(t.shape.size(), 1);
  int64_t acc = 1;
  for (int i = t.rank() - 1; i >= 0; --i) { t.strides[i] = acc; acc *= t.shape[i]; }
  t.storage = std::make_shared<CudaStorage<T>>(static_cast<size_t>(acc));
  return t;
}
template <typename T>
CudaTensor<T> toTensor(const T *host, std::initializer_list<int64_t> shape) {
  CudaTensor<T> t = newTensor<T>(shape);
  check(laser_b200_memcpy_h2d(t.unsafe_raw_data(), host, static_cast<size_t>(t.size()) * sizeof(T)));
  return t;
}
template <typename T>
std::vector<T> toHost(const CudaTensor<T> &t) {  // C-contiguous tensors only
  if (!t.is_C_contiguous()) throw std::invalid_argument("toHost needs a C-contiguous tensor");
  std::vector<T> out(static_cast<size_t>(t.size()));
  check(laser_b200_memcpy_d2h(out.data(), t.unsafe_raw_data(), out.size() * sizeof(T)));
  return out;
}

// C <- alpha * A*B + beta * C on rank-2 device tensors of any strides
template <typename T>
void matmul(const CudaTensor<T> &A, const CudaTensor<T> &B, CudaTensor<T> &C, double alpha = 1.0, double beta = 0.0,
            int path = LASER_B200_PATH_AUTO) {
  const laser_b200_tensor_view va = A.view(), vb = B.view();
  laser_b200_tensor_view vc = C.view();
  check(laser_b200_matmul_views(&va, &vb, &vc, alpha, beta, path, nullptr));
}

// ---- pre-packed API (gemm_prepacked.nim:63-292); device pointers --------------------------
inline size_t gemm_prepackA_mem_required(int64_t M, int64_t N, int64_t K) { return laser_b200_gemm_prepackA_mem_required_f32(M, N, K); }
inline size_t gemm_prepackB_mem_required(int64_t M, int64_t N, int64_t K) { return laser_b200_gemm_prepackB_mem_required_f32(M, N, K); }
inline void gemm_prepackA(void *dst_packedA, int64_t M, int64_t N, int64_t K, const float *src_A, int64_t rowStrideA,
                          int64_t colStrideA) {
  check(laser_b200_gemm_prepackA_f32_dev(dst_packedA, M, N, K, src_A, rowStrideA, colStrideA, nullptr));
}
inline void gemm_prepackB(void *dst_packedB, int64_t M, int64_t N, int64_t K, const float *src_B, int64_t rowStrideB,
                          int64_t colStrideB) {
  check(laser_b200_gemm_prepackB_f32_dev(dst_packedB, M, N, K, src_B, rowStrideB, colStrideB, nullptr));
}
inline void gemm_packed(int64_t M, int64_t N, int64_t K, float alpha, const void *packedA, const void *packedB, float beta,
                        float *C, int64_t rowStrideC, int64_t colStrideC) {
  check(laser_b200_gemm_packed_f32_dev(M, N, K, alpha, packedA, packedB, beta, C, rowStrideC, colStrideC, nullptr));
}

}  // namespace laser
