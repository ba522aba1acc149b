// tc_tf32x1.cu -- the eight instantiations (operand major-ness x single CTA / CTA pair) of gemm_tc_kernel<4, ptx::kFmtBF16, 1, float, false>
#include "tc_launch_impl.cuh"

namespace lb200 {
int launch_tc_tf32x1(const TcLaunch &l) { return launch_tc_family<4, ptx::kFmtBF16, 1, float, false>(l); }
}  // namespace lb200
