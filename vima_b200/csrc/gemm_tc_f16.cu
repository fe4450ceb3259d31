#include "gemm_tc_variants.cuh"
namespace vima {
cudaError_t launch_gemm_tc_f16(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream) {
  return launch_gemm_tc_impl<DT_F16>(p, l, grid, smem, max_smem, stream);
}
}  // namespace vima
