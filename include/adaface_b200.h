/* adaface_b200.h -- C-ABI of the B200-native AdaFace hot path (libadaface_b200.so).
 *
 * Drop-in boundary for the reference's attention operator and SubjBasisGenerator transformer
 * (SURVEY.md section 8b).  Every entry point takes raw DEVICE pointers + int64 shapes/strides + a
 * cudaStream_t (passed as void*), never allocates or frees, and returns 0 on success; on failure it
 * returns non-zero and adaface_last_error() holds the message (the Python side raises RuntimeError --
 * never breakpoint()/abort as the reference does).
 * All activations are bf16 (uint16 storage) unless a flag says fp32; parameters of epilogues are fp32.
 * Strides are in ELEMENTS.  "dalc" = adaface/diffusers_attn_lora_capture.py of the reference.
 */
#ifndef ADAFACE_B200_H_
#define ADAFACE_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADAFACE_B200_ABI_VERSION 1

/* epilogue activation of adaface_proj_lora_fwd */
#define ADAFACE_ACT_NONE 0
#define ADAFACE_ACT_QUICK_GELU 1 /* x * sigmoid(1.702 x): HF CLIPMLP activation (arc2face_models.py path) */
#define ADAFACE_ACT_GEGLU 2      /* out[:, j] = a_j * gelu(g_j), weight rows packed [a(64) | g(64)] per 128-col tile
                                    (ldm/modules/attention.py:31-38) */

/* dtype flags */
#define ADAFACE_BF16 0
#define ADAFACE_F32 1

int adaface_version(void);
const char* adaface_last_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches evidence). */
int64_t adaface_launch_count(void);

/* ---- K1: projection GEMM with the LoRA/DoRA update folded in --------------------------------------
 * Replaces attn.to_q / to_k / to_v / to_out[0] (dalc:235, 283, 288, 331), their peft lora.Linear
 * substitutes (dalc:242, 281, 286, 329; formula SURVEY 8a row A4), ldm CrossAttention's Linear layers
 * (ldm/modules/attention.py:156-164, 172-178, 205), the GEGLU feed-forward (:31-58) and the CLIP-shaped
 * encoder's Linear layers (q/k/v/out_proj, fc1, fc2).
 *
 *   Y[M,N] = act( colscale[N] * ( X[M,K] W[N,K]^T + T[M,R] Bs[N,R]^T ) + bias[N] ) + residual[M,N]
 *
 * X, W, T, Bs bf16 row-major (tcgen05 K-major operands, loaded by TMA); T = X A^T is produced by a
 * first call with W = A; Bs = s*B.  Any of T/Bs (together), colscale, bias, residual may be NULL.
 * K and R must be multiples of 8 (16-byte TMA row pitch); ldx/ldt multiples of 8.
 */
int adaface_proj_lora_fwd(const void* x, int64_t ldx, const void* w, const void* t, int64_t ldt, const void* bs,
                          const float* colscale, const float* bias, const void* residual, int64_t ldr,
                          int residual_dtype, void* y, int64_t ldy, int y_dtype, int64_t M, int64_t N, int64_t K,
                          int64_t R, int act, void* stream);

/* Same GEMM, bf16 output scattered head-major: output column c = which*H*d + h*d + dd of row b*rows_per_batch + n is
 * stored at y[which][b][h][n][dd], rows padded to dpad (>= d) elements -- y is a [N/(H*d), M/rows_per_batch, H,
 * rows_per_batch, dpad] buffer whose pad columns the caller zeroed once.  This is the internal q/k/v layout between
 * the fused QKV projection and adaface_attn_headmajor_fwd: dense 128-byte rows per (batch, head), which TMA loads
 * ~3x faster than 80-byte head slices of the interleaved [B, L, H*d] layout (measured, profiles/). */
int adaface_proj_lora_heads_fwd(const void* x, int64_t ldx, const void* w, const void* t, int64_t ldt, const void* bs,
                                const float* colscale, const float* bias, void* y, int64_t M, int64_t N, int64_t K,
                                int64_t R, int64_t heads, int64_t d, int64_t dpad, int64_t rows_per_batch, void* stream);

/* ---- K2: flash attention forward (self-attention, fast cross-attention, CLIP causal multi-KV) ------
 * Replaces F.scaled_dot_product_attention at dalc:321, the einsum attention of
 * ldm/modules/attention.py:181-204 and CLIPAttentionMKV's bmm/softmax/bmm (arc2face_models.py:170-217).
 * q/k/v/o are [B, L, H*d] views (heads interleaved, head h at column h*d) with element strides
 * (batch, token); d in {40, 64, 80, 160}.  key_mask [B, Lk] (1 = attend) is the img_mask key mask of
 * dalc:254-273 / attention.py:185-194 or NULL.  causal_mult = 0: no causal mask; M >= 1: key j is
 * visible to query i iff j / M <= i (arc2face_models.py:192) and, because each token row then holds its M
 * keys back to back (arc2face_models.py:74-79), key j is read at token j / M, column (j % M) * H*d + h*d
 * (k_sn / v_sn are the TOKEN strides; Lk counts keys = tokens * M).  scale multiplies q.k before softmax.
 */
int adaface_attn_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                     const void* v, int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb, int64_t o_sn, int64_t B,
                     int64_t H, int64_t Lq, int64_t Lk, int64_t d, const uint8_t* key_mask, int causal_mult,
                     float scale, void* stream);

/* Same operator for q/k/v tensors with an explicit HEAD stride (element strides batch / head / token), e.g. the
 * head-major [which, B, H, L, d] buffer that adaface_proj_lora_fwd can scatter a fused QKV projection into: every
 * (batch, head) slice is then one dense [L, d] matrix, which is what the TMA loads of the tcgen05 kernel like best.
 * drow_q / drow_kv = elements that really exist in a q / k,v row (d, or the padded width with zeroed pad columns).
 * o keeps the reference layout [B, Lq, H*d].  Unmasked only; d in {40, 80, 160}. */
int adaface_attn_headmajor_fwd(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, const void* k, int64_t k_sb,
                               int64_t k_sh, int64_t k_sn, const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sn, void* o,
                               int64_t o_sb, int64_t o_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
                               int64_t drow_q, int64_t drow_kv, float scale, void* stream);

/* ---- K3: cross-attention with capture / normalize / mix (the slow SDPA of dalc:79-139) -------------
 * S = Lk <= 128 keys staged once in shared memory.  Optional outputs (NULL = not wanted):
 *   prob   [B,H,Lq,S] fp32  softmax probabilities          (cached_activations['attn'],      dalc:358)
 *   score  [B,H,Lq,S] fp32  score AFTER the edit           (cached_activations['attnscore'], dalc:359)
 *   prob_subj [B,H,Lq,n_subj] fp32: probabilities of the columns subj_cols[b, 0..n_subj) only
 *          (the optional subject-columns-only capture mode; entries < 0 are skipped).
 * normalize (dalc:119-133): for columns with col_flag[b, j] != 0:
 *          score <- (score - scale * qmean[b,h,:] . k[b,j,h,:]) * (*ca_scale)   (qmean = mean over queries of q;
 *          ca_scale = DEVICE pointer to cross_attn_scale_factor, so the nn.Parameter is never synced to the host)
 * mix (dalc:108-118): B even, instances [sc.., mc..]; score <- (score_b + score_{b +/- B/2}) / 2.
 * in_dtype: ADAFACE_BF16, or ADAFACE_F32 = q/k/v are fp32 views (the projection GEMM's fp32 output): q.k is then
 * evaluated with a bf16 hi/lo split (hi.hi + lo.hi + hi.lo) so that captured probabilities are good to 1e-3.
 */
int adaface_attn_cross_capture_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb,
                                   int64_t k_sn, const void* v, int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb,
                                   int64_t o_sn, int64_t B, int64_t H, int64_t Lq, int64_t S, int64_t d, float scale,
                                   float* prob, float* score, float* prob_subj, const int32_t* subj_cols,
                                   int64_t n_subj, const uint8_t* col_flag, const float* qmean,
                                   const float* ca_scale, int mix, int in_dtype, void* stream);

/* Mean over the Lq queries of q[b, :, c] -> qmean[B, C] fp32 (C = H*d); feeds `normalize` above
 * (mean_i(q_i . k_j) = (mean_i q_i) . k_j, dalc:126). */
int adaface_qmean(const void* q, int q_dtype, int64_t q_sb, int64_t q_sn, int64_t B, int64_t Lq, int64_t C,
                  float* qmean, void* stream);

/* Capture re-layout (dalc:349-362): dst[b, c, n] = factor * src[b, n, c], src bf16 or fp32 view with element
 * strides (batch, token), dst fp32 [B, C, L] contiguous  ('b h n d -> b (h d) n' times sqrt(scale)). */
int adaface_capture_chan_major(const void* src, int src_dtype, int64_t s_sb, int64_t s_sn, int64_t B, int64_t L,
                               int64_t C, float factor, float* dst, void* stream);

/* ---- K4: LayerNorm and the SubjBasisGenerator head -------------------------------------------------
 * y[M,C] = LayerNorm(x[M,C]) * w + b ; x bf16 or fp32 (x_dtype), y bf16; eps 1e-5
 * (ldm BasicTransformerBlock.norm1-3; CLIPEncoderLayer.layer_norm1/2). */
int adaface_layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const float* w, const float* b, void* y,
                          int y_dtype, int64_t ldy, int64_t M, int64_t C, float eps, void* stream);

/* out[r, :] = LayerNorm( sum_l wl[l] * h_l[r, :] ) * w + b for the n_layers (1..4) fp32 hidden states h0..h3
 * (arc2face_models.py:291-306; wl = HOST array of n_layers weights already divided by their sum; unused
 * h pointers may be NULL).  out fp32 [M, C], C <= 768. */
int adaface_sbg_head_fwd(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl,
                         int n_layers, int64_t ldh, const float* w, const float* b, float* out, int64_t ldo,
                         int64_t M, int64_t C, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADAFACE_B200_H_ */
