// Epilogue specialisations of gemm_tc_kernel that the VIMA path uses (ACT, GLU, MUL, RES, O32, O16); everything else
// runs the generic runtime-flag variant.  Instantiated once per 16-bit format in gemm_tc_f16.cu / gemm_tc_bf16.cu.
#pragma once
#include "gemm_tc.cuh"

namespace vima {

struct GemmLaunch {
  int act, glu, mul, res, o32, o16, dtype;
};

#define VIMA_GEMM_VARIANTS(X, DT)                                                   \
  X(ACT_NONE, false, false, false, false, true, DT)      /* q / kv / c_attn / T5 qkv */ \
  X(ACT_NONE, false, false, true, true, true, DT)        /* xattn out-proj, linear2 */  \
  X(ACT_NONE, false, false, false, true, false, DT)      /* gate, conv1, in_proj -> fp32 */ \
  X(ACT_GELU, false, true, false, false, true, DT)       /* linear1: gelu(.) * gate */   \
  X(ACT_NONE, false, false, true, true, false, DT)       /* c_proj / mlp c_proj / T5 o, wo + residual */ \
  X(ACT_GELU, true, false, false, false, true, DT)       /* c_fc || gated_layer GEGLU */ \
  X(ACT_RELU, false, false, false, false, true, DT)      /* MLP hidden layers, T5 wi */  \
  X(ACT_QUICKGELU, false, false, false, false, true, DT) /* ViT c_fc */

template <int DT>
cudaError_t launch_gemm_tc_dt(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream);

cudaError_t launch_gemm_tc_f16(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream);
cudaError_t launch_gemm_tc_bf16(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream);

template <class E, bool TWO_CTA>
inline cudaError_t launch_kernel(const GemmParams& p, int grid, size_t smem, int max_smem, cudaStream_t stream) {
  static bool attr_set = false;  // per instantiation, per process (one device per process)
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<E, TWO_CTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (p.mcast || TWO_CTA) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<E, TWO_CTA>, p);
  }
  gemm_tc_kernel<E, TWO_CTA><<<grid, GEMM_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

template <class E>
inline cudaError_t launch_one(const GemmParams& p, int grid, size_t smem, int max_smem, cudaStream_t stream) {
  return p.two_cta ? launch_kernel<E, true>(p, grid, smem, max_smem, stream) : launch_kernel<E, false>(p, grid, smem, max_smem, stream);
}

template <int DT>
inline cudaError_t launch_gemm_tc_impl(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream) {
#define VIMA_TRY(ACT, GLU, MUL, RES, O32, O16, DTT)                                                                   \
  if (l.act == ACT && l.glu == (int)GLU && l.mul == (int)MUL && l.res == (int)RES && l.o32 == (int)O32 && l.o16 == (int)O16) \
    return launch_one<EpiCfg<false, ACT, GLU, MUL, RES, O32, O16, DTT>>(p, grid, smem, max_smem, stream);
  VIMA_GEMM_VARIANTS(VIMA_TRY, DT)
#undef VIMA_TRY
  return launch_one<EpiCfg<true, 0, false, false, false, false, false, DT>>(p, grid, smem, max_smem, stream);
}

}  // namespace vima
