// capi_python_emu.cpp -- TEST INFRASTRUCTURE: capi_layers_emu.cpp (the host side of the layer entry
// points on the CPU) plus CPU stand-ins for the remaining symbols of include/laser_b200.h, so that the
// Python mirror (laser_b200/*.py) can be loaded against it (LASER_B200_LIB) and its argument
// marshalling, view handling and the expectations of the layer tests can be exercised without a GPU.
// GEMM entries run the emulated EXACT kernel whatever path is asked for; bf16 and the pre-packed API are
// not modelled (LASER_B200_EUNSUPPORTED).  Loaded only by tests/test_emulated_python_mirror.py.
#include "capi_layers_emu.cpp"

#include "../../laser_b200/csrc/split.cuh"

namespace {
int g_mode = LASER_B200_PATH_TF32_BF16C;

template <typename T>
int simt_entry(int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t rsA, int64_t csA, const T *B, int64_t rsB,
               int64_t csB, T beta, T *C, int64_t rsC, int64_t csC) {
  int rc = check_args(M, N, K, A, B, C);
  if (rc == -1) return LASER_B200_OK;
  if (rc) return rc;
  g_last_path = LASER_B200_PATH_SIMT;
  return gemm_simt<T>(g_ctx, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, nullptr);
}
}  // namespace

extern "C" {
int laser_b200_init(void) { return 0; }
void laser_b200_shutdown(void) {}
const char *laser_b200_last_error(void) { return g_last_error.c_str(); }
int laser_b200_version(void) { return 100; }
int64_t laser_b200_launch_count(void) { return g_launches.load(); }
int laser_b200_last_path(void) { return g_last_path; }
int laser_b200_profile_begin(void) { return 0; }
int laser_b200_profile_end(double *a, int64_t *b, double *c, int64_t *d) {
  if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; if (d) *d = 0;
  return 0;
}
int laser_b200_set_f32_mode(int path) { g_mode = path; return 0; }
int laser_b200_get_f32_mode(void) { return g_mode; }

#define GEMM_ENTRIES(SUF, T)                                                                                          \
  int laser_b200_gemm_strided_##SUF(int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t rsA, int64_t csA,    \
                                    const T *B, int64_t rsB, int64_t csB, T beta, T *C, int64_t rsC, int64_t csC) {   \
    return simt_entry<T>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);                                 \
  }
GEMM_ENTRIES(f32, float)
GEMM_ENTRIES(f64, double)
GEMM_ENTRIES(i32, int32_t)
GEMM_ENTRIES(i64, int64_t)
int laser_b200_gemm_strided_f32_dev(int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t rsA, int64_t csA,
                                    const float *B, int64_t rsB, int64_t csB, float beta, float *C, int64_t rsC,
                                    int64_t csC, int path, void *stream) {
  return f32_dev(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, path, stream);
}
int laser_b200_gemm_strided_f32_epi_dev(int64_t, int64_t, int64_t, float, const float *, int64_t, int64_t, const float *,
                                        int64_t, int64_t, float, float *, int64_t, int64_t, const laser_b200_epilogue *, int, void *) {
  return set_error(LASER_B200_EUNSUPPORTED, "not modelled");
}
int laser_b200_gemm_strided_f64_dev(int64_t M, int64_t N, int64_t K, double alpha, const double *A, int64_t rsA, int64_t csA,
                                    const double *B, int64_t rsB, int64_t csB, double beta, double *C, int64_t rsC,
                                    int64_t csC, void *) {
  return simt_entry<double>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
}
int laser_b200_gemm_strided_i32_dev(int64_t M, int64_t N, int64_t K, int32_t alpha, const int32_t *A, int64_t rsA,
                                    int64_t csA, const int32_t *B, int64_t rsB, int64_t csB, int32_t beta, int32_t *C,
                                    int64_t rsC, int64_t csC, void *) {
  return simt_entry<int32_t>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
}
int laser_b200_gemm_strided_i64_dev(int64_t M, int64_t N, int64_t K, int64_t alpha, const int64_t *A, int64_t rsA,
                                    int64_t csA, const int64_t *B, int64_t rsB, int64_t csB, int64_t beta, int64_t *C,
                                    int64_t rsC, int64_t csC, void *) {
  return simt_entry<int64_t>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
}
int laser_b200_gemm_strided_bf16(int64_t, int64_t, int64_t, float, const uint16_t *, int64_t, int64_t, const uint16_t *,
                                 int64_t, int64_t, float, uint16_t *, int64_t, int64_t) {
  return set_error(LASER_B200_EUNSUPPORTED, "not modelled");
}
int laser_b200_gemm_strided_bf16_dev(int64_t, int64_t, int64_t, float, const uint16_t *, int64_t, int64_t, const uint16_t *,
                                     int64_t, int64_t, float, uint16_t *, int64_t, int64_t, void *) {
  return set_error(LASER_B200_EUNSUPPORTED, "not modelled");
}
size_t laser_b200_gemm_prepackA_mem_required_f32(int64_t, int64_t, int64_t) { return 0; }
size_t laser_b200_gemm_prepackB_mem_required_f32(int64_t, int64_t, int64_t) { return 0; }
int laser_b200_gemm_prepackA_f32_dev(void *, int64_t, int64_t, int64_t, const float *, int64_t, int64_t, void *) {
  return set_error(LASER_B200_EUNSUPPORTED, "not modelled");
}
int laser_b200_gemm_prepackB_f32_dev(void *, int64_t, int64_t, int64_t, const float *, int64_t, int64_t, void *) {
  return set_error(LASER_B200_EUNSUPPORTED, "not modelled");
}
int laser_b200_gemm_packed_f32_dev(int64_t, int64_t, int64_t, float, const void *, const void *, float, float *, int64_t,
                                   int64_t, void *) {
  return set_error(LASER_B200_EUNSUPPORTED, "not modelled");
}
int laser_b200_gemm_packedB_f32_dev(int64_t, int64_t, int64_t, float, const float *, int64_t, int64_t, const void *, float,
                                    float *, int64_t, int64_t, void *) {
  return set_error(LASER_B200_EUNSUPPORTED, "not modelled");
}
int laser_b200_malloc(void **dev_ptr, size_t bytes) {
  if (!dev_ptr) return set_error(LASER_B200_EINVAL, "null out pointer");
  *dev_ptr = std::aligned_alloc(256, (bytes + 255) / 256 * 256 + 256);
  return *dev_ptr ? 0 : set_error(LASER_B200_ENOMEM, "out of memory");
}
int laser_b200_free(void *p) { std::free(p); return 0; }
int laser_b200_memcpy_h2d(void *d, const void *s, size_t n) { std::memcpy(d, s, n); return 0; }
int laser_b200_memcpy_d2h(void *d, const void *s, size_t n) { std::memcpy(d, s, n); return 0; }
int laser_b200_memset_zero(void *d, size_t n) { std::memset(d, 0, n); return 0; }
int laser_b200_synchronize(void) { return 0; }
int laser_b200_matmul_views(const laser_b200_tensor_view *A, const laser_b200_tensor_view *B, laser_b200_tensor_view *C,
                            double alpha, double beta, int, void *) {
  if (!A || !B || !C || A->rank != 2 || B->rank != 2 || C->rank != 2) return set_error(LASER_B200_EINVAL, "rank-2 views");
  const int64_t M = A->shape[0], K = A->shape[1], N = B->shape[1];
#define RAW(T, v) (static_cast<T *>((v)->storage) + (v)->offset)
  if (A->dtype == 0)
    return simt_entry<float>(M, N, K, (float)alpha, RAW(float, A), A->strides[0], A->strides[1], RAW(float, B), B->strides[0],
                             B->strides[1], (float)beta, RAW(float, C), C->strides[0], C->strides[1]);
  if (A->dtype == 1)
    return simt_entry<double>(M, N, K, alpha, RAW(double, A), A->strides[0], A->strides[1], RAW(double, B), B->strides[0],
                              B->strides[1], beta, RAW(double, C), C->strides[0], C->strides[1]);
#undef RAW
  return set_error(LASER_B200_EUNSUPPORTED, "not modelled");
}
int laser_b200_debug_classify(int, const void *, int64_t, int64_t) { return -1; }
int laser_b200_debug_span(int64_t, int64_t, int64_t, int64_t, int64_t *, int64_t *, int *) { return LASER_B200_EUNSUPPORTED; }
int laser_b200_fill_uniform_f32_dev(float *dst, int64_t n, uint64_t seed, float lo, float hi, void *) {
  emu::launch(4, 256, [=]() { fill_uniform_f32_kernel(dst, n, seed, lo, hi); });
  return 0;
}
}  // extern "C"
