// lip_gemm_tc.cu — tcgen05 3xTF32 GEMM (placeholder until the tensor-core kernel lands; SIMT path is used).
#include "lip_common.cuh"

namespace lip {
bool tc_available() { return false; }
int gemm_tc(const TcGemmProblem&, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return LIP_ERR_UNSUPPORTED;
}
int tf32_split(const float*, int64_t, float*, float*, int64_t, int64_t, int64_t, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return LIP_ERR_UNSUPPORTED;
}
}  // namespace lip

extern "C" int lip_selftest_tc_gemm(int32_t, int64_t, int64_t, int64_t, int64_t, float*, lip_stream_t) {
  lip::set_error("tcgen05 GEMM not built");
  return LIP_ERR_UNSUPPORTED;
}
