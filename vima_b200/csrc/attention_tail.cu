// Stand-alone launch of the tail-row routine (attention_tail.cuh): a warp per (batch, head), 8 warps per CTA.  The default path runs
// the same routine INSIDE attention_tc_kernel (option attn_tail = "fused"); this kernel is the "kernel" setting, kept for A/B runs
// and as the reference the fused path is tested against.
#include "attention_tail.cuh"

namespace vima {

namespace {

constexpr int TAIL_WARPS = 8;

template <int DT>
__global__ void __launch_bounds__(TAIL_WARPS * 32) attention_tail_kernel(const AttnParams p, int row0, int nt, int lk_pad) {
  extern __shared__ __align__(16) float smt[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long unit = (long long)blockIdx.x * TAIL_WARPS + warp;
  if (unit >= (long long)p.B * p.H) return;  // whole warps leave; nothing below synchronises across warps
  const int b = (int)(unit / p.H), h = (int)(unit % p.H);
  attention_tail_rows<DT, 1>(p, b, h, row0, nt, lk_pad, smt + (size_t)warp * attention_tail_scratch_floats(lk_pad, 1), lane, 0);
}

}  // namespace

// rows [row0, row0 + nt) of every (batch, head) of p (nt <= ATTN_TAIL_MAX_ROWS); p.Lq / p.q_batch_rows give the row pitch
cudaError_t launch_attention_tail(const AttnParams& p, int row0, int nt, cudaStream_t stream) {
  if (p.B == 0 || nt <= 0) return cudaSuccess;
  if (nt > TAIL_NT || p.D != TAIL_D) return cudaErrorInvalidValue;
  const int lk_pad = (p.Lk + 31) & ~31;
  const size_t smem = (size_t)TAIL_WARPS * attention_tail_scratch_floats(lk_pad, 1) * sizeof(float);
  constexpr size_t smem_max = (size_t)TAIL_WARPS * attention_tail_scratch_floats(512, 1) * sizeof(float);  // Lk <= 512 (attention_tc's limit)
  if (smem > smem_max) return cudaErrorInvalidValue;
  static bool attr_set[2][64] = {};  // once per (format, device): not legal inside a CUDA-graph capture
  int dev = 0;
  cudaGetDevice(&dev);
  const int fi = p.dtype == DT_BF16 ? 1 : 0;
  if (!attr_set[fi][dev & 63]) {
    cudaError_t e = fi ? cudaFuncSetAttribute(attention_tail_kernel<DT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max)
                       : cudaFuncSetAttribute(attention_tail_kernel<DT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return e;
    attr_set[fi][dev & 63] = true;
  }
  const long long units = (long long)p.B * p.H;
  const unsigned grid = (unsigned)((units + TAIL_WARPS - 1) / TAIL_WARPS);
  if (fi) attention_tail_kernel<DT_BF16><<<grid, TAIL_WARPS * 32, smem, stream>>>(p, row0, nt, lk_pad);
  else attention_tail_kernel<DT_F16><<<grid, TAIL_WARPS * 32, smem, stream>>>(p, row0, nt, lk_pad);
  return cudaGetLastError();
}

}  // namespace vima
