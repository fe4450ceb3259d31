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
  if (p.stats_out) {  // statistics of the rows just produced: the consumer GEMM applies the next LayerNorm in its epilogue
    const float mean = p.rms ? 0.f : warp_sum(sum) * inv_n;  // rms: statistics for a folded T5 RMSNorm (no centring)
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_V4; ++i) {
      const int c = i * 32 + lane;
      if (c < nv4) {
        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
        sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_n + p.stats_eps);
    if (lane == 0) reinterpret_cast<float2*>(p.stats_out)[row] = make_float2(mean, rstd);
  }
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

// partial [rows, parts, 2] = (sum, sum of squares) over disjoint column sets -> stats [rows, 2] = (mean, rstd).  Combined in fp64 in
// a fixed order (the partials themselves are fp32 sums of <= 128 values each, written by the GEMM epilogues).
__global__ void __launch_bounds__(256) row_stats_finalize_kernel(const float2* __restrict__ partial, long long rows, int parts, double inv_n, float eps,
                                                                 int rms, float2* __restrict__ stats) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < parts; ++i) {
    const float2 v = __ldg(partial + r * parts + i);
    s1 += (double)v.x;
    s2 += (double)v.y;
  }
  const double mean = rms ? 0.0 : s1 * inv_n;  // rms: T5 RMSNorm statistics (mean 0, rstd = 1/sqrt(mean(x^2) + eps))
  double var = s2 * inv_n - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[r] = make_float2((float)mean, rsqrtf((float)var + eps));
}

cudaError_t launch_row_stats_finalize(const float* partial, long long rows, int parts, int cols, float eps, int rms, float* stats,
                                      cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((rows + 255) / 256);
  row_stats_finalize_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float2*>(partial), rows, parts, 1.0 / (double)cols, eps, rms,
                                                        reinterpret_cast<float2*>(stats));
  return cudaGetLastError();
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
