// Tail rows of the decoder attentions (head_dim 32): the <= 8 query rows of a (batch, head) past its last full 128-row tile.
//
// The tcgen05 kernel (attention_tc.cu) works on 128-query tiles.  The decoder's sequence lengths are T*(Q+1)-1 (263 for the 200M
// benchmark configuration, 392 = prompt | sep | history for VIMA-Gato): 7 / 8 rows spill into one more tile per (batch, head) that
// holds a CTA slot for its whole key range while one warp of four has work -- measured at 31 % of the self-attention and 29 % of
// the cross-attention kernel time (L = 263 vs 256, profiles/r2f_*).  `attention_tail_rows` computes those rows with plain fp32
// FMAs (packed f32x2) on the (hi + lo) operands, ~2 MFLOP per (batch, head), by NW cooperating warps:
//
//   NW = 4  inside attention_tc_kernel: the four softmax warps of the CTA that owns the LAST FULL tile of the (batch, head) run it
//           after their tile is stored -- no extra CTA slot, no extra launch, and K / V are L2-hot (the same CTA and its sibling
//           tiles streamed them through TMA moments earlier);
//   NW = 1  attention_tail_kernel (attention_tail.cu): a warp per (batch, head) in a launch of its own (option attn_tail = "kernel").
//
//   phase 1  thread = key:   y[j][i] = (q_i * scale*log2e) . k_j  + mask terms     -> shared [Lk][8], running row maxima
//   phase 2  thread = key:   p = exp2(y - max_i), row sums
//   phase 3  thread = (key group of 4*NW, 4-dim group of 8):  o_i += p[j][i] * v_j -> shuffle / shared-memory reduction
//
// Same mask semantics as attention.cu / attention_tc.cu (reference components.py:51-80, modeling_openai.py:86-115): causal
// replaces a hidden score by the soft -1e4, key padding adds finfo.min, keys beyond Lk are excluded, softmax over the full key
// range in fp32 (so a row whose visible keys are all padded gets the reference's degenerate weights without any special case).
#pragma once
#include "kernels.h"

namespace vima {

constexpr int TAIL_NT = ATTN_TAIL_MAX_ROWS;
constexpr int TAIL_D = 32;
constexpr float FP32_MIN_TL = -3.4028234663852886e38f;
constexpr float LOG2E_TL = 1.4426950408889634f;
constexpr float CAUSAL_L2_TL = -1e4f * LOG2E_TL;

// scratch floats the routine needs for `lk_pad` keys (lk_pad = Lk rounded up to 32)
__host__ __device__ constexpr int attention_tail_scratch_floats(int lk_pad, int nw) {
  return TAIL_NT * TAIL_D + lk_pad * TAIL_NT + 2 * nw * TAIL_NT + nw * TAIL_NT * TAIL_D;
}

template <int DT>
__device__ __forceinline__ float2 tl_unpack2(uint32_t w) {
  if constexpr (DT == DT_F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ unsigned long long tl_pack2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float2 tl_unpack64(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
// acc += a * b on two packed fp32 lanes
__device__ __forceinline__ void tl_ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

template <int DT>
__device__ __forceinline__ void tl_store4(const AttnParams& p, size_t row, int h, int dg, float4 o) {
  const size_t off = row * p.ldo + h * TAIL_D + dg * 4;
  if (p.o_lo8) {  // fp16 hi + e4m3 cross-term views (an "f16f8" consumer GEMM); the 16-bit lo part only if asked for
    uint2 h16;
    uint32_t l8, h8;
    split4_f8(o, F8_ACT_LO_SCALE, F8_ACT_HI_SCALE, h16, l8, h8);
    *reinterpret_cast<uint2*>(p.o_hi + off) = h16;
    const size_t off8 = row * p.ldo8 + h * TAIL_D + dg * 4;
    *reinterpret_cast<uint32_t*>(p.o_lo8 + off8) = l8;
    *reinterpret_cast<uint32_t*>(p.o_hi8 + off8) = h8;
    if (p.o_lo) {
      uint2 hi, lo;
      split4v<DT>(o, hi, lo);
      *reinterpret_cast<uint2*>(p.o_lo + off) = lo;
    }
  } else {
    uint2 hi, lo;
    split4v<DT>(o, hi, lo);
    *reinterpret_cast<uint2*>(p.o_hi + off) = hi;
    if (p.o_lo) *reinterpret_cast<uint2*>(p.o_lo + off) = lo;
  }
}

// Query rows [row0, row0 + nt) of (b, h), nt <= 8, by the 32*NW threads t = 0 .. 32*NW-1 (whole warps; every one of them must call).
// `scratch`: attention_tail_scratch_floats(lk_pad, NW) floats of shared memory, 16-byte aligned, private to the group;
// `bar_id`: a named barrier the group may use (NW > 1).
template <int DT, int NW>
__device__ __forceinline__ void attention_tail_rows(const AttnParams& p, int b, int h, int row0, int nt, int lk_pad, float* scratch, int t,
                                                    int bar_id) {
  constexpr int NTHR = 32 * NW;
  const int lane = t & 31, warp = t >> 5;
  const int Lk = p.Lk;
  const int kvb = p.kv_batch_rows ? p.kv_batch_rows : Lk;
  const int mld = p.mask_ld ? p.mask_ld : Lk;
  const int qbr = p.q_batch_rows ? p.q_batch_rows : p.Lq;
  float* qs = scratch;                               // [8][32] queries, pre-scaled by scale*log2e
  float* sc = qs + TAIL_NT * TAIL_D;                 // [lk_pad][8] scores -> weights
  float* red = sc + (size_t)lk_pad * TAIL_NT;        // [2][NW][8] per-warp row maxima / sums
  float* ored = red + 2 * NW * TAIL_NT;              // [NW][8][32] per-warp partial outputs
  auto group_sync = [&]() {
    if constexpr (NW == 1) __syncwarp(); else named_bar_sync(bar_id, NTHR);
  };

  // ---- queries ----
  const float c_l2 = p.scale * LOG2E_TL;
  for (int idx = t; idx < TAIL_NT * TAIL_D; idx += NTHR) {
    const int i = idx >> 5, d = idx & 31;
    float q = 0.f;
    if (i < nt) {
      const size_t off = ((size_t)b * qbr + row0 + i) * p.ldq + h * TAIL_D + d;
      q = Op16<DT>::back(p.q_hi[off]);
      if (p.q_lo) q += Op16<DT>::back(p.q_lo[off]);
    }
    qs[idx] = q * c_l2;
  }
  group_sync();

  // ---- phase 1: scores, thread = key ----
  float mx[TAIL_NT];
#pragma unroll
  for (int i = 0; i < TAIL_NT; ++i) mx[i] = -INFINITY;
  const int pos0 = row0 + p.q_pos0;  // key position of tail row 0 (causal)
  // raw (hi, lo) words of this thread's key row; the NEXT round's rows are requested as soon as the current ones are converted, so
  // the L2 round trip overlaps the 128 packed FMAs of the round in hand
  uint4 kh[4], kl[4];
  auto load_k = [&](int j) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { kh[c] = make_uint4(0u, 0u, 0u, 0u); kl[c] = kh[c]; }
    if (j < Lk) {
      const size_t rk = ((size_t)b * kvb + j) * p.ldk + h * TAIL_D;
      const uint4* ph = reinterpret_cast<const uint4*>(p.k_hi + rk);
#pragma unroll
      for (int c = 0; c < 4; ++c) kh[c] = __ldg(ph + c);
      if (p.k_lo) {
        const uint4* pl = reinterpret_cast<const uint4*>(p.k_lo + rk);
#pragma unroll
        for (int c = 0; c < 4; ++c) kl[c] = __ldg(pl + c);
      }
    }
  };
  load_k(t);
  for (int j0 = 0; j0 < lk_pad; j0 += NTHR) {
    const int j = j0 + t;
    const bool valid = j < Lk;
    unsigned long long kf[TAIL_D / 2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t aw[4] = {kh[c].x, kh[c].y, kh[c].z, kh[c].w};
      const uint32_t lw[4] = {kl[c].x, kl[c].y, kl[c].z, kl[c].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = tl_unpack2<DT>(aw[e]), y = tl_unpack2<DT>(lw[e]);
        kf[c * 4 + e] = tl_pack2(x.x + y.x, x.y + y.y);
      }
    }
    if (j0 + NTHR < lk_pad) load_k(j + NTHR);
    float madd = -INFINITY;  // beyond the sequence: excluded
    if (valid) madd = (p.key_mask == nullptr || p.key_mask[(size_t)b * mld + j]) ? 0.f : FP32_MIN_TL;
    float y[TAIL_NT];
#pragma unroll
    for (int i = 0; i < TAIL_NT; ++i) {
      const ulonglong2* q2 = reinterpret_cast<const ulonglong2*>(qs + i * TAIL_D);
      unsigned long long a0 = 0ull, a1 = 0ull;
#pragma unroll
      for (int c = 0; c < TAIL_D / 4; ++c) {
        const ulonglong2 q = q2[c];  // broadcast: every thread reads the same 16 bytes
        tl_ffma2(a0, q.x, kf[2 * c]);
        tl_ffma2(a1, q.y, kf[2 * c + 1]);
      }
      const float2 s0 = tl_unpack64(a0), s1 = tl_unpack64(a1);
      float v = ((s0.x + s0.y) + (s1.x + s1.y)) + madd;
      if (p.causal && j > pos0 + i) v = CAUSAL_L2_TL + madd;
      y[i] = v;
      mx[i] = fmaxf(mx[i], v);
    }
    if (j < lk_pad) {
      float4* dst = reinterpret_cast<float4*>(sc + (size_t)j * TAIL_NT);
      dst[0] = make_float4(y[0], y[1], y[2], y[3]);
      dst[1] = make_float4(y[4], y[5], y[6], y[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < TAIL_NT; ++i) mx[i] = warp_max(mx[i]);
  if constexpr (NW > 1) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < TAIL_NT; ++i) red[warp * TAIL_NT + i] = mx[i];
    }
    group_sync();
#pragma unroll
    for (int i = 0; i < TAIL_NT; ++i) {
      float m = red[i];
#pragma unroll
      for (int w = 1; w < NW; ++w) m = fmaxf(m, red[w * TAIL_NT + i]);
      mx[i] = m;
    }
  }

  // ---- phase 2: weights and row sums (every thread re-reads what it wrote itself) ----
  float ls[TAIL_NT];
#pragma unroll
  for (int i = 0; i < TAIL_NT; ++i) ls[i] = 0.f;
  for (int j0 = 0; j0 < lk_pad; j0 += NTHR) {
    if (j0 + t >= lk_pad) break;
    float4* cell = reinterpret_cast<float4*>(sc + (size_t)(j0 + t) * TAIL_NT);
    const float4 a = cell[0], c = cell[1];
    float w[TAIL_NT] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < TAIL_NT; ++i) {
      w[i] = ex2_approx(w[i] - mx[i]);
      ls[i] += w[i];
    }
    cell[0] = make_float4(w[0], w[1], w[2], w[3]);
    cell[1] = make_float4(w[4], w[5], w[6], w[7]);
  }
#pragma unroll
  for (int i = 0; i < TAIL_NT; ++i) ls[i] = warp_sum(ls[i]);
  if constexpr (NW > 1) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < TAIL_NT; ++i) red[(NW + warp) * TAIL_NT + i] = ls[i];
    }
  }
  group_sync();  // the weights of every key are visible to the whole group
  if constexpr (NW > 1) {
#pragma unroll
    for (int i = 0; i < TAIL_NT; ++i) {
      float s = red[NW * TAIL_NT + i];
#pragma unroll
      for (int w = 1; w < NW; ++w) s += red[(NW + w) * TAIL_NT + i];  // fixed order: deterministic
      ls[i] = s;
    }
  }

  // ---- phase 3: O = P V, thread = (key group kg of 4*NW, 4-dim group dg of 8) ----
  const int kg = t >> 3, dg = t & 7;
  constexpr int KG = 4 * NW;
  unsigned long long acc[TAIL_NT][2];
#pragma unroll
  for (int i = 0; i < TAIL_NT; ++i) { acc[i][0] = 0ull; acc[i][1] = 0ull; }
  for (int j = kg; j < Lk; j += 4 * KG) {  // 4 keys per thread per trip: their loads are all in flight before the first FMA
    uint2 vh[4], vl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = j + KG * u;
      vh[u] = make_uint2(0u, 0u); vl[u] = vh[u];
      if (jj < Lk) {
        const size_t rv = ((size_t)b * kvb + jj) * p.ldv + h * TAIL_D + dg * 4;
        vh[u] = __ldg(reinterpret_cast<const uint2*>(p.v_hi + rv));
        if (p.v_lo) vl[u] = __ldg(reinterpret_cast<const uint2*>(p.v_lo + rv));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = j + KG * u;
      if (jj >= Lk) break;
      const float2 h0 = tl_unpack2<DT>(vh[u].x), l0 = tl_unpack2<DT>(vl[u].x), h1 = tl_unpack2<DT>(vh[u].y), l1 = tl_unpack2<DT>(vl[u].y);
      const unsigned long long v01 = tl_pack2(h0.x + l0.x, h0.y + l0.y), v23 = tl_pack2(h1.x + l1.x, h1.y + l1.y);
      const float4* cell = reinterpret_cast<const float4*>(sc + (size_t)jj * TAIL_NT);
      const float4 a = cell[0], c = cell[1];
      const float w[TAIL_NT] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < TAIL_NT; ++i) {
        const unsigned long long ww = tl_pack2(w[i], w[i]);
        tl_ffma2(acc[i][0], ww, v01);
        tl_ffma2(acc[i][1], ww, v23);
      }
    }
  }
  // fold the 4 key groups of a warp (lanes differing in bits 3 and 4): afterwards every lane holds its warp's sums of its 4 dims
#pragma unroll
  for (int i = 0; i < TAIL_NT; ++i) {
    const float2 a01 = tl_unpack64(acc[i][0]), a23 = tl_unpack64(acc[i][1]);
    float4 o = make_float4(a01.x, a01.y, a23.x, a23.y);
    o.x += __shfl_xor_sync(0xffffffffu, o.x, 8); o.y += __shfl_xor_sync(0xffffffffu, o.y, 8);
    o.z += __shfl_xor_sync(0xffffffffu, o.z, 8); o.w += __shfl_xor_sync(0xffffffffu, o.w, 8);
    o.x += __shfl_xor_sync(0xffffffffu, o.x, 16); o.y += __shfl_xor_sync(0xffffffffu, o.y, 16);
    o.z += __shfl_xor_sync(0xffffffffu, o.z, 16); o.w += __shfl_xor_sync(0xffffffffu, o.w, 16);
    if constexpr (NW == 1) {
      if ((i & 3) != (lane >> 3) || i >= nt) continue;  // lane group g stores rows g and g + 4
      const float inv = 1.0f / ls[i];
      o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
      tl_store4<DT>(p, (size_t)b * qbr + row0 + i, h, dg, o);
    } else {
      if ((lane >> 3) == 0) *reinterpret_cast<float4*>(ored + ((size_t)warp * TAIL_NT + i) * TAIL_D + dg * 4) = o;
    }
  }
  if constexpr (NW > 1) {
    group_sync();
    if (t < TAIL_NT * 8) {  // thread = (row i, 4-dim group): fold the warps in a fixed order, normalise, store
      const int i = t >> 3, d4 = t & 7;
      if (i < nt) {
        float4 o = *reinterpret_cast<const float4*>(ored + (size_t)i * TAIL_D + d4 * 4);
#pragma unroll
        for (int w = 1; w < NW; ++w) {
          const float4 x = *reinterpret_cast<const float4*>(ored + ((size_t)w * TAIL_NT + i) * TAIL_D + d4 * 4);
          o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
        }
        float l = red[NW * TAIL_NT + i];
#pragma unroll
        for (int w = 1; w < NW; ++w) l += red[(NW + w) * TAIL_NT + i];
        const float inv = 1.0f / l;
        o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
        tl_store4<DT>(p, (size_t)b * qbr + row0 + i, h, d4, o);
      }
    }
  }
}

}  // namespace vima
