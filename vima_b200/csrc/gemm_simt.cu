// Exact-fp32 grouped GEMM on CUDA cores for the path's tiny / odd-shaped layers (bbox MLP K=4, action embedding
// K=2|4, the 12 action-head MLPs with N=50|100, M = episodes).  y[M,n] = act(x[M,k] * w[n,k]^T + b), one group
// per blockIdx.z.  These layers are <0.1 % of the step's FLOPs; keeping them in fp32 FFMA means the action
// logits carry no tensor-core rounding of their own.
#include "kernels.h"

namespace vima {

constexpr int ST = 64;   // tile edge
constexpr int SK = 32;   // k step
constexpr int SIMT_LD = ST * SK / 256;  // elements of each operand tile a thread brings in per k step

struct SimtGroupsByValue {  // descriptors travel in the kernel's parameter space: no device-side array, CUDA-graph safe
  SimtGemmGroup g[SIMT_MAX_HOST_GROUPS];
};

// These launches are latency-bound (M = episodes: a few dozen CTAs, K up to 768 walked in dependent steps), so the operand tiles of
// step k+1 are requested into registers before the FMAs of step k: one global round trip per step is hidden behind the arithmetic.
// Every output still accumulates its products in ascending k with fmaf, so the result does not depend on SK or on the prefetch.
__device__ __forceinline__ void simt_gemm_tile(const SimtGemmGroup& g, int M, int act) {
  const int n0 = blockIdx.x * ST, m0 = blockIdx.y * ST;
  if (n0 >= g.n) return;
  __shared__ float xs[SK][ST + 4];
  __shared__ float ws[SK][ST + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4] = {};
  float xr[SIMT_LD], wr[SIMT_LD];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < SIMT_LD; ++e) {
      const int i = threadIdx.x + e * 256;
      const int r = i / SK, kk = i % SK;
      const int m = m0 + r, n = n0 + r, k = k0 + kk;
      xr[e] = (m < M && k < g.k) ? __ldg(g.x + (size_t)m * g.ldx + k) : 0.f;
      wr[e] = (n < g.n && k < g.k) ? __ldg(g.w + (size_t)n * g.ldw + k) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.k; k0 += SK) {
#pragma unroll
    for (int e = 0; e < SIMT_LD; ++e) {
      const int i = threadIdx.x + e * 256;
      xs[i % SK][i / SK] = xr[e];
      ws[i % SK][i / SK] = wr[e];
    }
    __syncthreads();
    if (k0 + SK < g.k) fetch(k0 + SK);
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = xs[kk][ty * 4 + i]; b[i] = ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.n) continue;
      float v = acc[i][j] + (g.b ? __ldg(g.b + n) : 0.f);
      g.y[(size_t)m * g.ldy + n] = apply_act(act, v);
    }
  }
}

__global__ void __launch_bounds__(256) simt_gemm_kernel(const SimtGemmGroup* __restrict__ groups, int M, int act) {
  const SimtGemmGroup g = groups[blockIdx.z];
  simt_gemm_tile(g, M, act);
}

__global__ void __launch_bounds__(256) simt_gemm_kernel_v(const __grid_constant__ SimtGroupsByValue groups, int M, int act) {
  simt_gemm_tile(groups.g[blockIdx.z], M, act);
}

cudaError_t launch_simt_gemm_grouped_host(const SimtGemmGroup* groups_host, int n_groups, int M, int max_n, int act, cudaStream_t stream) {
  if (M == 0 || n_groups == 0) return cudaSuccess;
  for (int g0 = 0; g0 < n_groups; g0 += SIMT_MAX_HOST_GROUPS) {
    const int n = n_groups - g0 < SIMT_MAX_HOST_GROUPS ? n_groups - g0 : SIMT_MAX_HOST_GROUPS;
    SimtGroupsByValue v;
    for (int i = 0; i < SIMT_MAX_HOST_GROUPS; ++i) v.g[i] = groups_host[g0 + (i < n ? i : 0)];
    dim3 grid((max_n + ST - 1) / ST, (M + ST - 1) / ST, n);
    simt_gemm_kernel_v<<<grid, 256, 0, stream>>>(v, M, act);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_simt_gemm_grouped(const SimtGemmGroup* groups_dev, int n_groups, int M, int max_n, int act, cudaStream_t stream) {
  if (M == 0 || n_groups == 0) return cudaSuccess;
  dim3 grid((max_n + ST - 1) / ST, (M + ST - 1) / ST, n_groups);
  simt_gemm_kernel<<<grid, 256, 0, stream>>>(groups_dev, M, act);
  return cudaGetLastError();
}

}  // namespace vima
