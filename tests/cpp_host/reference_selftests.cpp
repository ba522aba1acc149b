// The reference's own GEMM self-tests (`when isMainModule` blocks of
// laser/primitives/matrix_multiplication/gemm.nim:255-507 and gemm_prepacked.nim:300-523),
// re-stated against the C++ host mirror (include/laser_b200.hpp): same call shape
//     gemm_strided(M, N, K, 1, a, K, 1, b, N, 1, 0, res_ab, N, 1);  doAssert res_ab == ab
// in the reference's element types (Nim `float` = double, `int` = int64), plus the pre-packed
// variants on device tensors.  `--link-only` returns before touching the GPU.
#include <cstdio>
#include <cstring>
#include <vector>

#include "laser_b200.hpp"

using laser::gemm_strided;

template <typename T>
static int run_case(const char *src, int M, int N, int K, std::vector<T> a, std::vector<T> b, std::vector<T> ab) {
  std::vector<T> res_ab(static_cast<size_t>(M) * N, T(-99));
  gemm_strided(M, N, K, T(1), a.data(), K, 1, b.data(), N, 1, T(0), res_ab.data(), N, 1);
  if (res_ab != ab) { std::fprintf(stderr, "FAILED %s\n", src); return 1; }
  std::printf("SUCCESS %s\n", src);
  return 0;
}

static int run_prepacked(const char *src, int M, int N, int K, std::vector<float> a, std::vector<float> b, std::vector<float> ab) {
  // gemm_prepacked.nim:300-347 `pack_and_test`: pack both operands, multiply, compare
  auto tA = laser::toTensor<float>(a.data(), {M, K}), tB = laser::toTensor<float>(b.data(), {K, N});
  auto tC = laser::newTensor<float>({M, N});
  auto pA = laser::newTensor<float>({static_cast<int64_t>(laser::gemm_prepackA_mem_required(M, N, K) / 4 + 64)});
  auto pB = laser::newTensor<float>({static_cast<int64_t>(laser::gemm_prepackB_mem_required(M, N, K) / 4 + 64)});
  laser::gemm_prepackA(pA.unsafe_raw_data(), M, N, K, tA.unsafe_raw_data(), K, 1);
  laser::gemm_prepackB(pB.unsafe_raw_data(), M, N, K, tB.unsafe_raw_data(), N, 1);
  laser::gemm_packed(M, N, K, 1.0f, pA.unsafe_raw_data(), pB.unsafe_raw_data(), 0.0f, tC.unsafe_raw_data(), N, 1);
  if (laser::toHost(tC) != ab) { std::fprintf(stderr, "FAILED prepacked %s\n", src); return 1; }
  std::printf("SUCCESS prepacked %s\n", src);
  return 0;
}

int main(int argc, char **argv) {
  if (argc > 1 && std::strcmp(argv[1], "--link-only") == 0) {
    std::printf("laser_b200 %d\n", laser_b200_version());
    return 0;
  }
  int bad = 0;
  try {
    bad += run_case<double>("gemm.nim:257-282", 3, 2, 3, {1, 2, 3, 1, 1, 1, 1, 1, 1}, {1, 1, 1, 1, 1, 1}, {6, 6, 3, 3, 3, 3});
    bad += run_case<double>("gemm.nim:284-309", 3, 2, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9}, {1, 1, 1, 1, 1, 1}, {6, 6, 15, 15, 24, 24});
    bad += run_case<double>("gemm.nim:311-334", 2, 2, 3, {1, 2, 3, 4, 5, 6}, {7, 8, 9, 10, 11, 12}, {58, 64, 139, 154});
    bad += run_case<int64_t>("gemm.nim:336-360", 2, 4, 3, {-2, -3, -1, 3, 0, 4}, {1, 5, 2, -1, -3, 0, 3, 4, 6, -2, 7, -4},
                             {1, -8, -20, -6, 27, 7, 34, -19});
    bad += run_case<int64_t>("gemm.nim:362-393", 5, 4, 4, {5, 6, 5, 8, 8, 2, 8, 8, 0, 5, 4, 0, 4, 0, 5, 6, 4, 5, 0, 3},
                             {5, 3, 6, 0, 5, 2, 3, 3, 8, 8, 2, 0, 7, 7, 0, 0},
                             {151, 123, 58, 18, 170, 148, 70, 6, 57, 42, 23, 15, 102, 94, 34, 0, 66, 43, 39, 15});
    bad += run_case<int64_t>("gemm.nim:395-424", 2, 2, 8, {2, 4, 3, 1, 3, 1, 3, 1, 4, 3, 2, 4, 1, 0, 0, 0},
                             {2, 2, 2, 1, 0, 3, 0, 1, 0, 2, 4, 3, 3, 3, 2, 1}, {27, 37, 14, 23});
    bad += run_case<int64_t>("gemm.nim:426-461", 8, 8, 2, {2, 1, 1, 3, 2, 1, 1, 0, 3, 4, 2, 4, 3, 1, 4, 0},
                             {2, 2, 0, 4, 0, 0, 4, 2, 2, 1, 2, 1, 2, 4, 4, 1},
                             {6, 5, 2, 9, 2, 4, 12, 5, 8, 5, 6, 7, 6, 12, 16, 5, 6, 5, 2, 9, 2, 4, 12, 5, 2, 2, 0, 4, 0, 0, 4, 2,
                              14, 10, 8, 16, 8, 16, 28, 10, 12, 8, 8, 12, 8, 16, 24, 8, 8, 7, 2, 13, 2, 4, 16, 7, 8, 8, 0, 16, 0, 0, 16, 8});
    bad += run_case<int64_t>("gemm.nim:463-507", 8, 8, 8,
                             {2, 4, 3, 1, 3, 1, 3, 1, 1, 2, 1, 1, 2, 0, 4, 3, 2, 0, 0, 3, 0, 4, 4, 1, 1, 1, 4, 0, 3, 1, 3, 0,
                              3, 4, 1, 1, 4, 2, 3, 4, 2, 4, 0, 2, 3, 3, 3, 4, 3, 0, 0, 3, 1, 4, 3, 1, 4, 3, 2, 4, 1, 0, 0, 0},
                             {2, 2, 0, 4, 0, 0, 4, 2, 2, 0, 0, 1, 1, 1, 3, 1, 0, 2, 2, 0, 2, 2, 3, 3, 0, 0, 1, 0, 4, 2, 4, 1,
                              0, 0, 1, 3, 4, 2, 4, 2, 4, 3, 4, 1, 4, 4, 0, 3, 3, 3, 0, 2, 1, 2, 3, 3, 2, 1, 2, 1, 2, 4, 4, 1},
                             {27, 23, 16, 29, 35, 32, 58, 37, 24, 19, 11, 23, 26, 30, 49, 27, 34, 29, 21, 21, 34, 34, 36, 32,
                              17, 22, 15, 21, 28, 25, 40, 33, 39, 27, 23, 40, 45, 46, 72, 41, 41, 26, 25, 34, 47, 48, 65, 38,
                              33, 28, 22, 26, 37, 34, 41, 33, 14, 12, 9, 22, 27, 17, 51, 23});
    // float32 flavours + the pre-packed path (gemm_prepacked.nim:354-367 and the shared vectors)
    bad += run_case<float>("gemm.nim:311-334 (f32)", 2, 2, 3, {1, 2, 3, 4, 5, 6}, {7, 8, 9, 10, 11, 12}, {58, 64, 139, 154});
    bad += run_prepacked("gemm_prepacked.nim:354-367", 3, 3, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9}, {1, 2, 3, 4, 5, 6, 7, 8, 9},
                         {30, 36, 42, 66, 81, 96, 102, 126, 150});
    bad += run_prepacked("gemm.nim:311-334 (prepacked)", 2, 2, 3, {1, 2, 3, 4, 5, 6}, {7, 8, 9, 10, 11, 12}, {58, 64, 139, 154});
    // the tensor contract on a device tensor: transposed view shares storage, matmul on views
    auto tA = laser::toTensor<float>(std::vector<float>{1, 2, 3, 4, 5, 6}.data(), {2, 3});
    auto tBt = laser::toTensor<float>(std::vector<float>{7, 9, 11, 8, 10, 12}.data(), {2, 3});  // B^T stored
    auto tB = tBt.transpose();
    auto tC = laser::newTensor<float>({2, 2});
    if (tA.rank() != 2 || tA.size() != 6 || !tA.is_C_contiguous() || tB.is_C_contiguous()) { std::fprintf(stderr, "tensor contract\n"); ++bad; }
    laser::matmul(tA, tB, tC);
    if (laser::toHost(tC) != std::vector<float>{58, 64, 139, 154}) { std::fprintf(stderr, "FAILED matmul on views\n"); ++bad; }
    else std::printf("SUCCESS matmul on strided device tensors\n");
  } catch (const laser::LaserB200Error &e) {
    std::fprintf(stderr, "exception: %s (code %d)\n", e.what(), e.code);
    return 2;
  }
  return bad ? 1 : 0;
}
