// The step immediately before the policy path (scripts/example.py:243-473) on the GPU:
//   object_stats_kernel : per (image, object id) pixel count and bounding box of the segmentation mask
//   crop_resize_kernel  : crop, zero-pad to a square, cv2.INTER_AREA resize to 32x32 (bit-exact restatement of OpenCV's
//                         8-bit paths, see oracle/prepare_oracle.py), visible objects compacted to the front
// HBM-bound byte/integer work: the segmentation map is read once (1-8 B / pixel), each crop's source pixels once or twice.
#include "kernels.h"

namespace vima {

constexpr int OBJ_MAX = 64;  // object ids per image

__device__ __forceinline__ long long load_segm(const void* p, size_t i, int elem) {
  if (elem == 1) return (long long)reinterpret_cast<const unsigned char*>(p)[i];
  if (elem == 4) return (long long)reinterpret_cast<const int*>(p)[i];
  return reinterpret_cast<const long long*>(p)[i];
}

// stats[img, obj, 5] = {count, xmin, xmax, ymin, ymax}; ids: [n_obj] shared by all images, or [n_img, n_obj]
__global__ void __launch_bounds__(1024) object_stats_kernel(const void* __restrict__ segm, int elem, int H, int W,
                                                            const long long* __restrict__ ids, int n_obj, int ids_per_image,
                                                            int* __restrict__ stats) {
  __shared__ long long sid[OBJ_MAX];
  __shared__ int acc[OBJ_MAX][5];
  const int img = blockIdx.x;
  for (int o = threadIdx.x; o < n_obj; o += blockDim.x) {
    sid[o] = ids[(ids_per_image ? (size_t)img * n_obj : 0) + o];
    acc[o][0] = 0; acc[o][1] = INT_MAX; acc[o][2] = -1; acc[o][3] = INT_MAX; acc[o][4] = -1;
  }
  __syncthreads();
  const size_t base = (size_t)img * H * W;
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
    const long long v = load_segm(segm, base + p, elem);
    const int y = p / W, x = p - y * W;
    for (int o = 0; o < n_obj; ++o) {
      if (v == sid[o]) {  // ids may repeat: every matching slot is updated
        atomicAdd(&acc[o][0], 1);
        atomicMin(&acc[o][1], x); atomicMax(&acc[o][2], x);
        atomicMin(&acc[o][3], y); atomicMax(&acc[o][4], y);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_obj * 5; i += blockDim.x) stats[(size_t)img * n_obj * 5 + i] = acc[i / 5][i % 5];
}

// ---- cv2 INTER_AREA, 8-bit --------------------------------------------------------------------------------------------
struct Square {  // the zero-padded square crop (example.py:419-431), addressed in padded coordinates
  const unsigned char* rgb;  // this image's [3, H, W]
  int H, W, x0, y0, cw, ch, padx, pady;
  __device__ __forceinline__ int at(int c, int py, int px) const {
    const int y = py - pady, x = px - padx;
    if ((unsigned)y >= (unsigned)ch || (unsigned)x >= (unsigned)cw) return 0;
    return rgb[((size_t)c * H + (y0 + y)) * W + (x0 + x)];
  }
};

// computeResizeAreaTab for one destination index: source span [first partial | full cells | last partial]
struct AreaSpan {
  int s_first, sx1, sx2;            // partial-left source index (or -1), full cells [sx1, sx2), partial-right at sx2 (if has_last)
  float a_first, a_mid, a_last;
  bool has_last;
};
__device__ __forceinline__ AreaSpan area_span(int d, int n) {
  const double scale = 1.0 / (32.0 / (double)n);
  const double f1 = d * scale, f2 = f1 + scale;
  const double cell = fmin(scale, (double)n - f1);
  int sx1 = (int)ceil(f1), sx2 = (int)floor(f2);
  sx2 = min(sx2, n - 1);
  sx1 = min(sx1, sx2);
  AreaSpan s;
  s.sx1 = sx1; s.sx2 = sx2;
  s.s_first = -1; s.a_first = 0.f;
  if ((double)sx1 - f1 > 1e-3) { s.s_first = sx1 - 1; s.a_first = (float)(((double)sx1 - f1) / cell); }
  s.a_mid = (float)(1.0 / cell);
  s.has_last = (f2 - (double)sx2 > 1e-3);
  s.a_last = s.has_last ? (float)(fmin(fmin(f2 - (double)sx2, 1.0), cell) / cell) : 0.f;
  return s;
}

// "area" coefficients of the bilinear fallback used when enlarging (resize.cpp, area_mode): 11-bit fixed point
struct LinTap { int s; int c0, c1; bool edge; };
__device__ __forceinline__ LinTap lin_tap(int d, int n) {
  const double inv = 32.0 / (double)n, scale = 1.0 / inv;
  int s = (int)floor(d * scale);
  float f = (float)((double)(d + 1) - (double)(s + 1) * inv);
  f = (f <= 0.f) ? 0.f : f - floorf(f);
  LinTap t;
  t.edge = (s + 1 >= n);
  if (s >= n - 1) { f = 0.f; s = n - 1; }
  t.s = s;
  t.c0 = max(-32768, min(32767, __float2int_rn(__fmul_rn(1.f - f, 2048.f))));
  t.c1 = max(-32768, min(32767, __float2int_rn(__fmul_rn(f, 2048.f))));
  return t;
}

__device__ __forceinline__ unsigned char sat_u8(int v) { return (unsigned char)max(0, min(255, v)); }

// grid (n_obj, n_img), 32x32 threads: thread (dy, dx) produces the three channels of one output pixel.
__global__ void __launch_bounds__(1024) crop_resize_kernel(const unsigned char* __restrict__ rgb, int H, int W, const int* __restrict__ stats,
                                                          int n_obj, unsigned char* __restrict__ crops, long long* __restrict__ bbox,
                                                          unsigned char* __restrict__ mask, int* __restrict__ n_valid) {
  const int o = blockIdx.x, img = blockIdx.y;
  const int dx = threadIdx.x, dy = threadIdx.y;
  const int* st = stats + ((size_t)img * n_obj + o) * 5;
  const bool valid = st[0] >= 2;  // example.py:409-411 / 281-282: fewer than two pixels -> not visible
  // slot: visible objects keep their order at the front, the others fill the tail (example.py:441-456)
  int before_valid = 0, total_valid = 0;
  for (int j = 0; j < n_obj; ++j) {
    const bool vj = stats[((size_t)img * n_obj + j) * 5] >= 2;
    total_valid += vj;
    if (j < o) before_valid += vj;
  }
  const int slot = valid ? before_valid : total_valid + (o - before_valid);
  const size_t out_base = ((size_t)img * n_obj + slot) * 3 * 32 * 32;
  const int tid = dy * 32 + dx;
  if (tid == 0) {
    mask[(size_t)img * n_obj + slot] = valid ? 1 : 0;
    if (o == 0 && n_valid) n_valid[img] = total_valid;
  }
  if (!valid) {
    for (int c = 0; c < 3; ++c) crops[out_base + (size_t)c * 1024 + tid] = 0;
    if (tid < 4) bbox[((size_t)img * n_obj + slot) * 4 + tid] = 0;
    return;
  }
  const int xmin = st[1], xmax = st[2], ymin = st[3], ymax = st[4];
  if (tid == 0) {
    long long* b = bbox + ((size_t)img * n_obj + slot) * 4;
    b[0] = (xmin + xmax) / 2; b[1] = (ymin + ymax) / 2; b[2] = ymax - ymin; b[3] = xmax - xmin;
  }
  Square sq;
  sq.rgb = rgb + (size_t)img * 3 * H * W;
  sq.H = H; sq.W = W; sq.x0 = xmin; sq.y0 = ymin; sq.cw = xmax - xmin + 1; sq.ch = ymax - ymin + 1;
  const int n = max(sq.cw, sq.ch);
  sq.padx = (sq.ch > sq.cw) ? (sq.ch - sq.cw) / 2 : 0;
  sq.pady = (sq.cw > sq.ch) ? (sq.cw - sq.ch) / 2 : 0;

  int res[3];
  if (n == 32) {
    for (int c = 0; c < 3; ++c) res[c] = sq.at(c, dy, dx);
  } else if (n < 32) {  // bilinear with area coefficients, fixed point (HResizeLinear / VResizeLinear<uchar,int,short>)
    const LinTap tx = lin_tap(dx, n), ty = lin_tap(dy, n);
    const int r0 = min(max(ty.s, 0), n - 1), r1 = min(max(ty.s + 1, 0), n - 1);
    for (int c = 0; c < 3; ++c) {
      int h0, h1;
      if (!tx.edge) {
        h0 = sq.at(c, r0, tx.s) * tx.c0 + sq.at(c, r0, tx.s + 1) * tx.c1;
        h1 = sq.at(c, r1, tx.s) * tx.c0 + sq.at(c, r1, tx.s + 1) * tx.c1;
      } else {
        h0 = sq.at(c, r0, tx.s) * 2048;
        h1 = sq.at(c, r1, tx.s) * 2048;
      }
      res[c] = sat_u8((((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2);
    }
  } else if (n % 32 == 0) {  // ResizeAreaFast_: integer box sums
    const int k = n / 32;
    const float scale = 1.f / (float)(k * k);
    for (int c = 0; c < 3; ++c) {
      int s = 0;
      for (int yy = 0; yy < k; ++yy)
        for (int xx = 0; xx < k; ++xx) s += sq.at(c, dy * k + yy, dx * k + xx);
      res[c] = (k == 2) ? ((s + 2) >> 2) : sat_u8(__float2int_rn(__fmul_rn((float)s, scale)));
    }
  } else {  // ResizeArea_<uchar, float>: separately rounded float32 multiplies and adds, in OpenCV's order
    const AreaSpan ax = area_span(dx, n), ay = area_span(dy, n);
    float sum[3] = {0.f, 0.f, 0.f};
    bool first_row = true;
    auto row = [&](int sy, float beta) {
      float buf[3] = {0.f, 0.f, 0.f};
      if (ax.s_first >= 0)
        for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)sq.at(c, sy, ax.s_first), ax.a_first));
      for (int sx = ax.sx1; sx < ax.sx2; ++sx)
        for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)sq.at(c, sy, sx), ax.a_mid));
      if (ax.has_last)
        for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)sq.at(c, sy, ax.sx2), ax.a_last));
      for (int c = 0; c < 3; ++c) sum[c] = first_row ? __fmul_rn(beta, buf[c]) : __fadd_rn(sum[c], __fmul_rn(beta, buf[c]));
      first_row = false;
    };
    if (ay.s_first >= 0) row(ay.s_first, ay.a_first);
    for (int sy = ay.sx1; sy < ay.sx2; ++sy) row(sy, ay.a_mid);
    if (ay.has_last) row(ay.sx2, ay.a_last);
    for (int c = 0; c < 3; ++c) res[c] = sat_u8(__float2int_rn(sum[c]));
  }
  for (int c = 0; c < 3; ++c) crops[out_base + (size_t)c * 1024 + tid] = (unsigned char)res[c];
}

cudaError_t launch_object_stats(const void* segm, int elem, int n_img, int H, int W, const long long* ids, int n_obj, int ids_per_image,
                                int* stats, cudaStream_t s) {
  if (n_img == 0 || n_obj == 0) return cudaSuccess;
  object_stats_kernel<<<n_img, 1024, 0, s>>>(segm, elem, H, W, ids, n_obj, ids_per_image, stats);
  return cudaGetLastError();
}

cudaError_t launch_crop_resize(const unsigned char* rgb, int n_img, int H, int W, const int* stats, int n_obj, unsigned char* crops,
                               long long* bbox, unsigned char* mask, int* n_valid, cudaStream_t s) {
  if (n_img == 0 || n_obj == 0) return cudaSuccess;
  crop_resize_kernel<<<dim3(n_obj, n_img), dim3(32, 32), 0, s>>>(rgb, H, W, stats, n_obj, crops, bbox, mask, n_valid);
  return cudaGetLastError();
}

}  // namespace vima
