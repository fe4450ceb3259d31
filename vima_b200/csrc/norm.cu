// LayerNorm / T5 RMSNorm rows kernel: one warp per row, row kept in registers, 128-bit loads, warp-shuffle
// reductions.  Optionally adds a second tensor before normalising and chains a second LayerNorm, and emits the
// result as fp32 and/or as (hi, lo) 16-bit GEMM operands.  HBM-bound: reads 4 B, writes <= 12 B per element.
#include "kernels.h"

namespace vima {

constexpr int NORM_MAX_V4 = 8;  // up to 8 float4 per lane -> cols <= 1024

template <int DT>
__global__ void __launch_bounds__(256) norm_rows_kernel(const NormParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.rows) return;
  const size_t row = (size_t)warp;
  const int nv4 = p.cols >> 2;
  float4 v[NORM_MAX_V4];
  const float4* x4 = reinterpret_cast<const float4*>(p.x + row * p.ldx);
  const float4* a4 = p.add ? reinterpret_cast<const float4*>(p.add + row * p.ld_add) : nullptr;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_MAX_V4; ++i) {
    const int c = i * 32 + lane;
    if (c < nv4) {
      float4 t = __ldg(x4 + c);
      if (a4) {
        const float4 u = __ldg(a4 + c);
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      v[i] = t;
      sum += (t.x + t.y) + (t.z + t.w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float inv_n = 1.0f / (float)p.cols;

  auto normalise = [&](const float* w, const float* b, float eps, bool rms, float s) {
    float mean = 0.f;
    if (!rms) mean = warp_sum(s) * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_V4; ++i) {
      const int c = i * 32 + lane;
      if (c < nv4) {
        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
        sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float var = warp_sum(sq) * inv_n;
    const float rstd = rsqrtf(var + eps);
    float nsum = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_V4; ++i) {
      const int c = i * 32 + lane;
      if (c < nv4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x;
        o.y = (v[i].y - mean) * rstd * g.y;
        o.z = (v[i].z - mean) * rstd * g.z;
        o.w = (v[i].w - mean) * rstd * g.w;
        if (b) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + c);
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        v[i] = o;
        nsum += (o.x + o.y) + (o.z + o.w);
      }
    }
    return nsum;
  };

  if (p.w) sum = normalise(p.w, p.b, p.eps, p.rms != 0, sum);
  if (p.out_f32) {
    float4* o4 = reinterpret_cast<float4*>(p.out_f32 + row * p.ld_o32);
#pragma unroll
    for (int i = 0; i < NORM_MAX_V4; ++i) {
      const int c = i * 32 + lane;
      if (c < nv4) o4[c] = v[i];
    }
  }
  if (p.w2) {
    normalise(p.w2, p.b2, p.eps2, false, sum);
    if (p.out2_f32) {
      float4* o4 = reinterpret_cast<float4*>(p.out2_f32 + row * p.ld_o2);
#pragma unroll
      for (int i = 0; i < NORM_MAX_V4; ++i) {
        const int c = i * 32 + lane;
        if (c < nv4) o4[c] = v[i];
      }
    }
  }
  if (p.out_hi) {
    uint2* h2 = reinterpret_cast<uint2*>(p.out_hi + row * p.ld_o16);
    uint2* l2 = p.out_lo ? reinterpret_cast<uint2*>(p.out_lo + row * p.ld_o16) : nullptr;
#pragma unroll
    for (int i = 0; i < NORM_MAX_V4; ++i) {
      const int c = i * 32 + lane;
      if (c < nv4) {
        uint2 hv, lv;
        split4v<DT>(v[i], hv, lv);
        h2[c] = hv;
        if (l2) l2[c] = lv;
      }
    }
  }
  if (p.out_lo8) {
    uint32_t* l8 = reinterpret_cast<uint32_t*>(p.out_lo8 + row * p.ld_o8);
    uint32_t* h8 = reinterpret_cast<uint32_t*>(p.out_hi8 + row * p.ld_o8);
#pragma unroll
    for (int i = 0; i < NORM_MAX_V4; ++i) {
      const int c = i * 32 + lane;
      if (c < nv4) {
        uint2 h16;
        uint32_t lo8, hi8;
        split4_f8(v[i], F8_ACT_LO_SCALE, F8_ACT_HI_SCALE, h16, lo8, hi8);
        l8[c] = lo8;
        h8[c] = hi8;
      }
    }
  }
}

cudaError_t launch_norm(const NormParams& p, cudaStream_t stream) {
  if (p.rows == 0) return cudaSuccess;
  const int warps_per_block = 8;
  const long long blocks = (p.rows + warps_per_block - 1) / warps_per_block;
  if (p.dtype == DT_BF16)
    norm_rows_kernel<DT_BF16><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(p);
  else
    norm_rows_kernel<DT_F16><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace vima
