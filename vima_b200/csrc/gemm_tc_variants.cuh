// Epilogue specialisations of gemm_tc_kernel that the VIMA path uses (ACT, GLU, MUL, RES, O32, O16); everything else
// runs the generic runtime-flag variant.  Instantiated once per 16-bit format in gemm_tc_f16.cu / gemm_tc_bf16.cu.
#pragma once
#include "gemm_tc.cuh"

namespace vima {

struct GemmLaunch {
  int act, glu, mul, res, o32, o16, dtype;
  int lna, lnr, stats;  // folded A-side LayerNorm, LayerNorm'd residual, partial row statistics of the output
  int device;           // function attributes are per device
};

// (ACT, GLU, MUL, RES, O32, O16, DT, LNA, LNR, STATS)
#define VIMA_GEMM_VARIANTS(X, DT)                                                   \
  X(ACT_NONE, false, false, false, false, true, DT, false, false, false)      /* q / kv / c_attn / T5 qkv */ \
  X(ACT_NONE, false, false, true, true, true, DT, false, false, false)        /* linear2 (+ residual -> fp32 + operands) */  \
  X(ACT_NONE, false, false, true, true, true, DT, false, false, true)         /* xattn out-proj, c_proj: + row statistics for the folded LN */ \
  X(ACT_NONE, false, false, false, true, false, DT, false, false, false)      /* conv1, in_proj -> fp32 */ \
  X(ACT_GELU, false, true, false, false, true, DT, false, false, false)       /* linear1: gelu(.) * gate (unfused form) */   \
  X(ACT_NONE, false, false, true, true, false, DT, false, false, false)       /* T5 o, wo + residual */ \
  X(ACT_NONE, false, false, true, true, false, DT, false, true, false)        /* mlp c_proj + LayerNorm'd residual */ \
  X(ACT_GELU, true, false, false, false, true, DT, false, false, false)       /* GEGLU (pre-normalised operand) */ \
  X(ACT_GELU, true, false, false, false, true, DT, true, false, false)        /* GEGLU with the LayerNorm folded in (linear1||gate, c_fc||gate) */ \
  X(ACT_RELU, false, false, false, false, true, DT, false, false, false)      /* MLP hidden layers, T5 wi */  \
  X(ACT_QUICKGELU, false, false, false, false, true, DT, false, false, false) /* ViT c_fc (pre-normalised operand) */ \
  X(ACT_QUICKGELU, false, false, false, false, true, DT, true, false, false)  /* ViT c_fc with ln_2 folded in */ \
  X(ACT_NONE, false, false, false, true, false, DT, true, false, false)       /* ViT in_proj with ln_1 folded in -> fp32 */

template <int DT>
cudaError_t launch_gemm_tc_dt(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream);

cudaError_t launch_gemm_tc_f16(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream);
cudaError_t launch_gemm_tc_bf16(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream);

template <class E, bool TWO_CTA>
inline cudaError_t launch_kernel(const GemmParams& p, int grid, size_t smem, int max_smem, cudaStream_t stream, int device) {
  static bool attr_set[64] = {};  // per instantiation and device (function attributes are per device)
  const int di = device & 63;
  if (!attr_set[di]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<E, TWO_CTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return e;
    attr_set[di] = true;
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
inline cudaError_t launch_one(const GemmParams& p, int grid, size_t smem, int max_smem, cudaStream_t stream, int device) {
  return p.two_cta ? launch_kernel<E, true>(p, grid, smem, max_smem, stream, device) : launch_kernel<E, false>(p, grid, smem, max_smem, stream, device);
}

template <int DT>
inline cudaError_t launch_gemm_tc_impl(const GemmParams& p, const GemmLaunch& l, int grid, size_t smem, int max_smem, cudaStream_t stream) {
#define VIMA_TRY(ACT, GLU, MUL, RES, O32, O16, DTT, LNA, LNR, STATS)                                                     \
  if (l.act == ACT && l.glu == (int)GLU && l.mul == (int)MUL && l.res == (int)RES && l.o32 == (int)O32 && l.o16 == (int)O16 && \
      l.lna == (int)LNA && l.lnr == (int)LNR && l.stats == (int)STATS)                                                  \
    return launch_one<EpiCfg<false, ACT, GLU, MUL, RES, O32, O16, DTT, LNA, LNR, STATS>>(p, grid, smem, max_smem, stream, l.device);
  VIMA_GEMM_VARIANTS(VIMA_TRY, DT)
#undef VIMA_TRY
  return launch_one<EpiCfg<true, 0, false, false, false, false, false, DT>>(p, grid, smem, max_smem, stream, l.device);
}

}  // namespace vima
