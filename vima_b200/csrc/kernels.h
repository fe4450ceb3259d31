// Internal launch interface between the C-ABI layer (api.cu) and the kernel translation units.
#pragma once
#include "common.cuh"

namespace vima {

struct NormParams {
  const float* x; long long rows; int cols; int ldx;
  const float* add; int ld_add;       // optional: normalise (x + add)
  const float* w; const float* b;     // first norm (w == null: no normalisation, passthrough/convert only)
  float eps; int rms;                 // rms=1: T5 RMSNorm (no mean, no bias)
  const float* w2; const float* b2; float eps2;  // optional chained LayerNorm on the first norm's output
  float* out_f32; int ld_o32;         // output of the first norm (fp32), optional
  float* out2_f32; int ld_o2;         // output of the second norm (fp32), optional
  unsigned short* out_hi; unsigned short* out_lo; int ld_o16;  // last norm's output as 16-bit operands, optional
  int dtype;
  unsigned char* out_lo8; unsigned char* out_hi8; int ld_o8;   // e4m3 cross-term views, optional
  float* stats_out; float stats_eps;  // optional [rows, 2] = (mean, rstd) of the FIRST norm's output rows (for a LayerNorm folded downstream)
};
cudaError_t launch_norm(const NormParams& p, cudaStream_t stream);
cudaError_t launch_row_stats_finalize(const float* partial, long long rows, int parts, int cols, float eps, int rms, float* stats,
                                      cudaStream_t stream);

struct AttnParams {
  const unsigned short *q_hi, *q_lo; int ldq;  // [B*Lq, ldq]; pointer already at head 0's first column
  const unsigned short *k_hi, *k_lo; int ldk;  // [B*Lk, ldk]
  const unsigned short *v_hi, *v_lo; int ldv;
  const unsigned char* key_mask;               // [B, Lk] 1 = attend, or null
  const float* rel_bias;                       // [H, 2*Lk-1] additive bias indexed by (j - i + Lk - 1), or null
  unsigned short *o_hi, *o_lo; int ldo;        // [B*Lq, ldo]
  int B, H, Lq, Lk, D;
  float scale; int causal; int split; int dtype;
  unsigned char *o_lo8, *o_hi8; int ldo8;      // e4m3 cross-term views of the output, optional
  int kv_batch_rows;                           // rows between consecutive batch elements in k/v (>= Lk; KV caches), 0 = Lk
  int mask_ld;                                 // row pitch of key_mask, 0 = Lk
  int q_pos0;                                  // causal: query row i sits at key position q_pos0 + i (incremental decode)
  int q_batch_rows;                            // rows between consecutive batch elements in q / o (>= Lq), 0 = Lq
};
constexpr int ATTN_TAIL_MAX_ROWS = 8;          // attention_tail.cu: query rows per (batch, head) the SIMT tail kernel takes
cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream);
size_t attention_smem_bytes(const AttnParams& p);             // dynamic shared memory the mma.sync kernel needs for p
int attention_max_lk(const AttnParams& p, size_t smem_limit);  // largest Lk (multiple of 64) that fits smem_limit at p's format
bool attention_tc_supported(const AttnParams& p);  // tcgen05 variant (attention_tc.cu): head_dim 32, split operands, no bias, Lk <= 512
cudaError_t launch_attention_tc(const AttnParams& p, void* encode_tiled_fn, cudaStream_t stream);  // encode_tiled_fn: cuTensorMapEncodeTiled
// query rows [row0, row0 + nt) of every (batch, head) (nt <= ATTN_TAIL_MAX_ROWS, head_dim 32, Lk <= 512): the rows that would
// otherwise occupy a nearly empty 128-row tile of the tcgen05 kernel
cudaError_t launch_attention_tail(const AttnParams& p, int row0, int nt, cudaStream_t stream);

struct SmallAttnParams {  // tiny-sequence fp32 attention (ViT: 5 tokens, 24 heads of 32)
  const float* qkv; int ld;        // [N*S, ld], q | k | v each W wide
  unsigned short *o_hi, *o_lo; int ldo; float* o_f32;  // [N*S, W]
  long long N; int S, H, W; float scale; int dtype;
};
cudaError_t launch_small_attention(const SmallAttnParams& p, cudaStream_t stream);

struct LatentAttnParams {  // fp32 attention of a few latent queries over <= 16 keys (Perceiver resampler)
  const float* q; int ldq; long long q_batch_stride;  // [N or 1][Lq, ldq]; stride 0 = the same queries for every image
  const float* k; int ldk;                            // [N*Lk, ldk]
  const float* v; int ldv;
  float* o; int ldo;                                  // [N*Lq, ldo]
  long long N; int Lq, Lk, H, d; float scale;
};
cudaError_t launch_latent_attention(const LatentAttnParams& p, cudaStream_t stream);

struct SimtGemmGroup {  // one fp32 problem: y[M, n] = act(x[M, k] * w[n, k]^T + b)
  const float* x; int ldx;
  const float* w; int ldw;
  const float* b;
  float* y; int ldy;
  int n, k;
};
cudaError_t launch_simt_gemm_grouped(const SimtGemmGroup* groups_dev, int n_groups, int M, int max_n, int act, cudaStream_t stream);
constexpr int SIMT_MAX_HOST_GROUPS = 16;  // descriptors per launch when they travel by value (kernel parameter space)
cudaError_t launch_simt_gemm_grouped_host(const SimtGemmGroup* groups_host, int n_groups, int M, int max_n, int act, cudaStream_t stream);

// element-wise / gather kernels (misc.cu)
cudaError_t launch_split(const float* x, long long rows, int cols, int ldx, unsigned short* hi, unsigned short* lo, int ld16,
                         int pad_cols, float scale, int dtype, cudaStream_t s);
cudaError_t launch_pack_weight(const float* w, int n, int k, int transposed, int ldw, unsigned short* hi, unsigned short* lo, int ld16,
                               float scale, int dtype, cudaStream_t s);
cudaError_t launch_pack_weight_f8(const float* w, int n, int k, int transposed, int ldw, unsigned char* hi8, unsigned char* lo8, int ld8,
                                  float scale, cudaStream_t s);
cudaError_t launch_split_f8(const float* x, long long rows, int cols, int ldx, unsigned char* lo8, unsigned char* hi8, int ld8, cudaStream_t s);
cudaError_t launch_assemble_history(const float* obs, const unsigned char* obs_mask, const float* act, int T, int B, int Q, int E,
                                    int La, float* tokens, unsigned char* masks_bl, long long* pos_bl, cudaStream_t s);
cudaError_t launch_mask_cumsum(const unsigned char* mask, int B, int L, long long* pos, cudaStream_t s);
cudaError_t launch_add_pos_embed(const float* tok, long long stride_b, long long stride_l, const long long* ids, const float* table,
                                 int n_pos, int B, int L, int E, float* out_f32, unsigned short* hi, unsigned short* lo, int ld16,
                                 int dtype, int* err_flag, cudaStream_t s);
cudaError_t launch_gather_prompt(const int* kind, const int* index, const long long* word_ids, const float* word_table,
                                 const float* img_emb, const unsigned char* img_mask, int B, int Lp, int D, float* out,
                                 unsigned char* mask_out, cudaStream_t s);
cudaError_t launch_patchify(const unsigned char* img, long long N, int H, int W, int P, unsigned short* hi, unsigned short* lo, int ld16,
                            int dtype, cudaStream_t s);
cudaError_t launch_vit_tokens(const float* patch_out, const float* cls, const float* pos, long long N, int S, int W, float* out,
                              cudaStream_t s);
cudaError_t launch_bbox_norm(const long long* bbox, long long n, float* out, cudaStream_t s);
cudaError_t launch_fill_ee(const long long* ee, const float* table, long long n_te, int Q, unsigned short* hi, unsigned short* lo,
                           int ld16, int col0, int n_pad, int dtype, cudaStream_t s);
cudaError_t launch_action_scale(const long long* idx, long long n, int width, const float* inv_bins, float* out, cudaStream_t s);
cudaError_t launch_object_stats(const void* segm, int elem, int n_img, int H, int W, const long long* ids, int n_obj, int ids_per_image,
                                int* stats, cudaStream_t s);
cudaError_t launch_crop_resize(const unsigned char* rgb, int n_img, int H, int W, const int* stats, int n_obj, unsigned char* crops,
                               long long* bbox, unsigned char* mask, int* n_valid, cudaStream_t s);
cudaError_t launch_action_post(const long long* idx, long long n, int width, const float* bins, const float* lo, const float* hi,
                               int bound_stride, float* out, cudaStream_t s);
cudaError_t launch_head_select(const float* logits, int B, int n_heads, const int* head_off, float* logits_norm, long long* modes,
                               cudaStream_t s);
cudaError_t launch_gato_positions(const unsigned char* prompt_mask, int B, int Lp, int L, unsigned char* mask_out, long long* pos_out,
                                  cudaStream_t s);
cudaError_t launch_max_u8(const unsigned char* x, long long n, int* out_max, cudaStream_t s);

}  // namespace vima
