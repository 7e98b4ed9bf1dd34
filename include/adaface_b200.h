/* adaface_b200.h -- C-ABI of the B200-native AdaFace hot path (libadaface_b200.so).
 *
 * Drop-in boundary for the reference's attention operator, SubjBasisGenerator transformer and -- since ABI v3 -- the
 * convolutional blocks of the U-Net around them (SURVEY.md sections 8b, 8f).  Every entry point takes raw DEVICE pointers + int64 shapes/strides + a
 * cudaStream_t (passed as void*), never allocates or frees (one exception: adaface_conv3x3_fwd keeps a per-device fp32
 * split-K workspace, >= 32 MB, allocated on first use -- so warm a shape up before capturing it into a CUDA graph), and
 * returns 0 on success; on failure it
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

#define ADAFACE_B200_ABI_VERSION 4 /* v3 = v2 + the U-Net convolution set; v4 = v3 + sampler step, capture consumers (additive) */

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
/* Programmatic dependent launch for the tcgen05 kernels (launch attribute programmaticStreamSerialization; the kernels
 * run griddepcontrol.wait before their first global access).  Returns the previous mask.  Default: env ADAFACE_PDL,
 * else 1 (GEMMs only: the short projection kernels gain most from overlapping their launch ramp with the
 * predecessor's tail; +2.5 % on the whole step). */
int adaface_set_pdl(int mask);   /* bit 0: projection GEMMs, bit 1: attention kernels */

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
 * lse (optional, NULL = not wanted): fp32 [B, H, Lq], log2-domain log-sum-exp of every score row -- the only
 * thing the training forward leaves behind for adaface_attn_bwd.
 */
int adaface_attn_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                     const void* v, int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb, int64_t o_sn, int64_t B,
                     int64_t H, int64_t Lq, int64_t Lk, int64_t d, const uint8_t* key_mask, int causal_mult,
                     float scale, float* lse, void* stream);

/* Same operator for q/k/v tensors with an explicit HEAD stride (element strides batch / head / token), e.g. the
 * head-major [which, B, H, L, d] buffer that adaface_proj_lora_fwd can scatter a fused QKV projection into: every
 * (batch, head) slice is then one dense [L, d] matrix, which is what the TMA loads of the tcgen05 kernel like best.
 * drow_q / drow_kv = elements that really exist in a q / k,v row (d, or the padded width with zeroed pad columns).
 * o keeps the reference layout [B, Lq, H*d].  Unmasked only; d in {40, 80, 160}. */
int adaface_attn_headmajor_fwd(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, const void* k, int64_t k_sb,
                               int64_t k_sh, int64_t k_sn, const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sn, void* o,
                               int64_t o_sb, int64_t o_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
                               int64_t drow_q, int64_t drow_kv, float scale, float* lse, void* stream);

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

/* ---- SpatialTransformer entry / exit (ldm/modules/attention.py:287-304) -------------
 * y[b, hw, c] = GroupNorm(groups, eps)(x)[b, c, hw] * gamma[c] + beta[c]: the norm (attention.py:70-71, 291) fused with
 * 'b c h w -> b (h w) c' (:293).  x [B, C, HW] contiguous bf16 | fp32, y [B, HW, C] bf16 for the proj_in GEMM (the 1x1
 * convolution :268-272 is a Linear over channels).  a_ws, s_ws: fp32 [B, C] scratch (folded scale / shift). */
int adaface_groupnorm_tokens_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, int64_t B, int64_t C,
                                 int64_t HW, int64_t groups, float eps, float* a_ws, float* s_ws, void* y, void* stream);
/* out[b, c, hw] = t[b, hw, c] + x_in[b, c, hw]: 'b (h w) c -> b c h w' (:301) fused with the residual (:304);
 * t bf16 [B, HW, C] (output of the proj_out GEMM), x_in / out [B, C, HW] of x_dtype. */
int adaface_tokens_to_nchw_add(const void* t, const void* x_in, int x_dtype, void* out, int64_t B, int64_t C, int64_t HW,
                               void* stream);

/* ---- ResBlock / Downsample / Upsample of the U-Net (ldm/modules/diffusionmodules/openaimodel.py:92-277; SURVEY 8f
 * row 2) on NHWC activations: x [B, H, W, Cin] bf16 contiguous == the "tokens" layout [B, H*W, C] of the attention path,
 * so a ResBlock -> SpatialTransformer chain never leaves it.
 *
 * 3x3 convolution, padding 1, stride 1 | 2 (conv_nd(2, cin, cout, 3, padding=1 [, stride=2]) :110, :151, :206, :236) as an
 * implicit GEMM on tcgen05: M = B*Ho*Wo output pixels, N = Cout, K = 9 taps x Cin; the activation tile of every tap is a
 * shifted TMA box whose out-of-image part is zero-filled (= the padding), so no im2col buffer exists.
 *   w        bf16 [Cout, 9 * Kc] with Kc = Cin rounded up to 64: w[co][(ky*3 + kx) * Kc + ci] = weight[co][ci][ky][kx]
 *   bias     fp32 [Cout] or NULL;   rowbias fp32 [B, Cout] or NULL: added to every pixel of image b (h + emb_out :257)
 *   residual [B*Ho*Wo, ldr] bf16 | fp32 or NULL (skip_connection(x) + h :260);   y [B*Ho*Wo, ldy] bf16 | fp32
 *   t, bs, R, colscale: optional conv-LoRA / DoRA tail exactly as in adaface_proj_lora_fwd (t = lora_A conv output [M, R],
 *   bs = scaled lora_B 1x1 weight [Cout, R]; dalc:541-591).   Cin % 8 == 0, Wo <= 128, even H, W for stride 2. */
int adaface_conv3x3_fwd(const void* x, int64_t B, int64_t H, int64_t W, int64_t Cin, const void* w, const void* t, int64_t ldt,
                        const void* bs, int64_t R, const float* colscale, const float* bias, const float* rowbias,
                        const void* residual, int64_t ldr, int residual_dtype, void* y, int64_t ldy, int y_dtype, int64_t Cout,
                        int stride, int act, void* stream);
/* y = act(GroupNorm(groups, eps)(x) * gamma + beta) over tokens: x, y bf16 [B, HW, C]; act 0 = none, 1 = SiLU
 * (normalization() + nn.SiLU() in front of each convolution, openaimodel.py:203-205, 229-231; fp32 statistics like
 * GroupNorm32, util.py).  a_ws, s_ws: fp32 [B, C] scratch (folded scale / shift per image and channel); part_ws:
 * adaface_groupnorm_act_tokens_ws_floats(B, HW, C, groups) floats of scratch (may be 0 -> NULL allowed).  C / groups even. */
int adaface_groupnorm_act_tokens_fwd(const void* x, const float* gamma, const float* beta, int64_t B, int64_t HW, int64_t C,
                                     int64_t groups, float eps, int act, float* part_ws, float* a_ws, float* s_ws, void* y,
                                     void* stream);
int64_t adaface_groupnorm_act_tokens_ws_floats(int64_t B, int64_t HW, int64_t C, int64_t groups);
/* y = SiLU(x) as bf16, n elements (the nn.SiLU() in front of emb_layers' Linear, openaimodel.py:222-228). */
int adaface_silu_fwd(const void* x, int x_dtype, void* y, int64_t n, void* stream);
/* Nearest-neighbour 2x (F.interpolate(scale_factor=2, mode="nearest"), openaimodel.py:116): x bf16 [B, H, W, C] ->
 * y [B, 2H, 2W, C]. */
int adaface_upsample2x_tokens(const void* x, void* y, int64_t B, int64_t H, int64_t W, int64_t C, void* stream);
/* Sinusoidal timestep embedding (ldm/modules/diffusionmodules/util.py:154-174): t fp32 [B] -> out bf16 [B, dim] =
 * [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(max_period) i / (dim / 2)); an odd last column is zero. */
int adaface_timestep_embedding(const float* t, int64_t B, int64_t dim, float max_period, void* out, void* stream);

/* Backward of adaface_groupnorm_act_tokens_fwd w.r.t. x (gamma / beta are frozen U-Net weights, ddpm.py:637-638): x, dy, dx
 * bf16 [B, HW, C]; statistics are recomputed from x; coef_ws: fp32 [4, B, C] scratch.  Deterministic. */
int adaface_groupnorm_act_tokens_bwd(const void* x, const void* dy, const float* gamma, const float* beta, int64_t B, int64_t HW,
                                     int64_t C, int64_t groups, float eps, int act, float* coef_ws, void* dx, void* stream);
/* Re-sampling steps of the backward pass on NHWC bf16 (H, W = the LOW resolution):
 *   mode 0: y [B, H, W, C] = 2x2 sum-pool of x [B, 2H, 2W, C]          (backward of adaface_upsample2x_tokens)
 *   mode 1: y [B, 2H, 2W, C] = x [B, H, W, C] at even positions, else 0 (the backward of a stride-2 adaface_conv3x3_fwd is the
 *           stride-1 convolution of this tensor with the flipped, transposed weights) */
int adaface_resample2x_bwd(const void* x, void* y, int64_t B, int64_t H, int64_t W, int64_t C, int mode, void* stream);

/* ---- K5: backward kernels of the stage-2 training step (ddpm.py:1645-1707 back-propagates through the `sc`
 * instance of the U-Net into the LoRA / DoRA adapters, cross_attn_scale_factor and, via the context, SubjBasisGenerator).
 * The reference gets all of this from autograd over eager PyTorch ops; here every gradient is a kernel, recompute
 * form (nothing of size Lq x Lk is kept between forward and backward).
 *
 * Flash-attention backward of adaface_attn_fwd (same views, masks and causal multi-KV addressing):
 *   delta_i = dO_i . O_i;  P = exp2(scale log2e q.k - lse);  dV = P^T dO;  dS = P o (dO V^T - delta) scale;
 *   dQ = dS K;  dK = dS^T Q.   lse from the forward call; delta = fp32 [B, H, Lq] scratch.  dq/dk/dv bf16 views shaped
 * like q/k/v.  Deterministic (no atomics). */
int adaface_attn_bwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn, const void* v,
                     int64_t v_sb, int64_t v_sn, const void* o, int64_t o_sb, int64_t o_sn, const void* dout, int64_t do_sb,
                     int64_t do_sn, const float* lse, float* delta, void* dq, int64_t dq_sb, int64_t dq_sn, void* dk,
                     int64_t dk_sb, int64_t dk_sn, void* dv, int64_t dv_sb, int64_t dv_sn, int64_t B, int64_t H, int64_t Lq,
                     int64_t Lk, int64_t d, const uint8_t* key_mask, int causal_mult, float scale, void* stream);

/* Backward of adaface_attn_cross_capture_fwd (dalc:79-139; under mix only the sc half receives dq / dk, dalc:117).  Upstream gradients: dout [B,Lq,H*d]
 * bf16 and, optionally, dprob / dscore [B,H,Lq,S] fp32 (the losses on cached 'attn' / 'attnscore').  q/k/v, col_flag,
 * qmean, ca_scale, mix, in_dtype exactly as passed to the forward call.  Outputs: dq bf16 [B,Lq,H*d] view; dk, dv
 * [B,S,H*d] views of dkv_dtype; dca (optional) = dca_mul * d loss / d cross_attn_scale_factor (dca_mul carries the x10
 * GradientScaler of dalc:129).  Workspaces (fp32, caller-allocated): dk_part, dv_part of
 * B*H*chunks*S*d floats and dca_part of B*H*chunks floats, chunks = adaface_attn_cross_capture_bwd_chunks(B, H, Lq). */
int adaface_attn_cross_capture_bwd_chunks(int64_t B, int64_t H, int64_t Lq);
int adaface_attn_cross_capture_bwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                                   const void* v, int64_t v_sb, int64_t v_sn, const void* dout, int64_t do_sb,
                                   int64_t do_sn, const float* dprob, const float* dscore, int64_t B, int64_t H,
                                   int64_t Lq, int64_t S, int64_t d, float scale, const uint8_t* col_flag,
                                   const float* qmean, const float* ca_scale, int mix, int in_dtype, void* dq, int64_t dq_sb,
                                   int64_t dq_sn, void* dk, int64_t dk_sb, int64_t dk_sn, void* dv, int64_t dv_sb,
                                   int64_t dv_sn, int dkv_dtype, float* dca, float dca_mul, float* dk_part,
                                   float* dv_part, float* dca_part, void* stream);

/* dst[b, j, i] = alpha * colscale[j] * rowscale[i] * src[b, i, j]  (src [B, I, J], dst [B, J, I]; element strides
 * batch / row; colscale fp32 [J], rowscale fp32 [I], either may be NULL).  Operand re-layout of the weight-gradient GEMMs (dW = dY^T X, dA, dB of the LoRA pair,
 * with the DoRA column scale folded in) and the backward of adaface_capture_chan_major. */
int adaface_transpose(const void* src, int src_dtype, int64_t s_sb, int64_t s_ld, void* dst, int dst_dtype, int64_t d_sb,
                      int64_t d_ld, int64_t B, int64_t I, int64_t J, float alpha, const float* colscale,
                      const float* rowscale, void* stream);

/* out[j] += colmul[j] * sum_i a[i, j] * (b ? b[i, j] - bias[j] : 1)   (out fp32 [N], ACCUMULATED with atomics: zero
 * it first).  b = NULL: bias gradient.  b = Y, bias, colmul = 1/m: DoRA magnitude gradient. */
int adaface_colsum(const void* a, int a_dtype, int64_t lda, const void* b, int b_dtype, int64_t ldb, const float* bias,
                   const float* colmul, float* out, int64_t M, int64_t N, void* stream);

/* LayerNorm backward: dx[M,C] (dtype of x) from x, dy, w; dw / db fp32 [C] ACCUMULATED (zero first) or both NULL when
 * the affine parameters are frozen (U-Net blocks). */
int adaface_layernorm_bwd(const void* x, int x_dtype, int64_t ldx, const void* dy, int dy_dtype, int64_t lddy,
                          const float* w, void* dx, int64_t lddx, float* dw, float* db, int64_t M, int64_t C, float eps,
                          void* stream);

/* Stand-alone activation (training mode keeps the pre-activation u): h = act(u) and du = dh * act'(u), bf16.
 * ADAFACE_ACT_QUICK_GELU: u, h, dh, du all [M, n_out].  ADAFACE_ACT_GEGLU: u / du [M, 2 n_out] in packed
 * [a(64) | gate(64)] tiles, h / dh [M, n_out]. */
int adaface_act_fwd(const void* u, int64_t ldu, void* h, int64_t ldh, int64_t M, int64_t n_out, int act, void* stream);
int adaface_act_bwd(const void* u, int64_t ldu, const void* dh, int64_t lddh, void* du, int64_t lddu, int64_t M,
                    int64_t n_out, int act, void* stream);

/* Backward of adaface_sbg_head_fwd: dh_l = wl[l] * dmix (fp32 [M,C] each, written), dwl[l] += <dmix, h_l>,
 * dw / db of the final LayerNorm accumulated (zero dwl [4], dw, db first). */
int adaface_sbg_head_bwd(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl,
                         int n_layers, int64_t ldh, const float* w, const float* dout, int64_t lddo, float* dh0,
                         float* dh1, float* dh2, float* dh3, float* dwl, float* dw, float* db, int64_t M, int64_t C,
                         float eps, void* stream);

/* ---- K3c (ABI v4): capture with FUSED CONSUMERS -------------------------------------------------
 * The slow SDPA of adaface_attn_cross_capture_fwd (normalize supported, mix not) that REDUCES the probability map where it
 * sits in registers instead of writing [B,H,Lq,S] fp32 to HBM, for the two stage-2 losses that consume it:
 *   subj_sum[b,h,i] = sum_{j: sum_flag[b,j]} prob[b,h,i,j]   -- sel_emb_attns_by_indices(..., do_sum=True), the input of
 *       calc_subj_masked_bg_suppress_loss (ldm/util.py:1862-1868); sum_flag uint8 [B,S] marks each instance's subject tokens;
 *   sum_{h,i,j} (prob[b,h,i,j] - ref_prob[b,h,i,j])^2        -- the sc vs sc_rep probability MSE of
 *       calc_sc_rep_attn_distill_loss (ldm/util.py:2084-2089); ref_prob fp32 [B,H,Lq,S] (the detached sc_rep map); the
 *       kernel writes per-(CTA, warp) partials sq_part fp32 [B,H,sq_slots,4] (ZERO it first; sq_slots >= ceil(Lq/64) always
 *       suffices) which the caller adds up in a fixed order (deterministic).
 * prob (nullable) still receives the full map when a caller wants it (the sc_rep instance).  Any of the two consumers may be
 * absent (NULL pointers). */
int adaface_attn_cross_consume_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                                   const void* v, int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb, int64_t o_sn, int64_t B,
                                   int64_t H, int64_t Lq, int64_t S, int64_t d, float scale, float* prob,
                                   const uint8_t* col_flag, const float* qmean, const float* ca_scale, int in_dtype,
                                   const uint8_t* sum_flag, float* subj_sum, const float* ref_prob, float* sq_part,
                                   int64_t sq_slots, void* stream);
/* Backward of the above: like adaface_attn_cross_capture_bwd with the gradient of the map given IMPLICITLY --
 *   dprob[b,h,i,j] (nullable dense term) + g_subj[b,h,i] * sum_flag[b,j] + mse_coef[b] * (P[b,h,i,j] - ref_prob[b,h,i,j])
 * (P recomputed in the kernel; mse_coef DEVICE fp32 [B] = 2 * upstream gradient of instance b's squared-difference sum), so no [B,H,Lq,S]
 * gradient map exists either. */
int adaface_attn_cross_consume_bwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                                   const void* v, int64_t v_sb, int64_t v_sn, const void* dout, int64_t do_sb, int64_t do_sn,
                                   const float* dprob, int64_t B, int64_t H, int64_t Lq, int64_t S, int64_t d, float scale,
                                   const uint8_t* col_flag, const float* qmean, const float* ca_scale, int in_dtype, void* dq,
                                   int64_t dq_sb, int64_t dq_sn, void* dk, int64_t dk_sb, int64_t dk_sn, void* dv, int64_t dv_sb,
                                   int64_t dv_sn, int dkv_dtype, float* dca, float dca_mul, float* dk_part, float* dv_part,
                                   float* dca_part, const uint8_t* sum_flag, const float* g_subj, const float* ref_prob,
                                   const float* mse_coef, void* stream);

/* ---- K6 (ABI v4): sampler step around the U-Net ---------------------------------------------------------------------
 * One DDIM step for n_images latents of n_per_image fp32 elements each (ldm/models/diffusion/ddim.py:223-302, the
 * arithmetic after apply_model): classifier-free-guidance combine e = e_u + g (e_c - e_u) (:253-255) when has_uncond
 * (eps = [cond images.., uncond images..], the reference's CFG batch order :239-248), pred_x0 = (x - sqrt(1-a_t) e) / sqrt(a_t)
 * (:281), x_prev = sqrt(a_prev) pred_x0 + sqrt(1 - a_prev - sigma^2) e + sigma * noise * temperature (:285-301).
 * coef: DEVICE fp32[8] = {g, sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma, temperature, -} so that one
 * captured CUDA graph serves every step.  noise (fp32, same shape as x) may be NULL (eta = 0).  x_prev may alias x;
 * x_dup (nullable) receives a second copy of x_prev (the uncond half of the next step's CFG batch); pred_x0 nullable.
 * Separately rounded fp32 ops in the reference's order: bit-identical to torch fp32 for identical eps. */
int adaface_ddim_cfg_step(const float* eps, int64_t n_images, int64_t n_per_image, int has_uncond, const float* x,
                          const float* coef, const float* noise, float* x_prev, float* x_dup, float* pred_x0, void* stream);

/* adaface_sbg_head_fwd / _bwd with the (already sum-normalised) layer weights read from DEVICE memory (wl_dev: fp32
 * [n_layers]) instead of a host array: no host read of hidden_state_layer_weights per step, so the SubjBasisGenerator
 * training step can be captured into a CUDA graph. */
int adaface_sbg_head_fwd_dev(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl_dev, int n_layers,
                             int64_t ldh, const float* w, const float* b, float* out, int64_t ldo, int64_t M, int64_t C, float eps,
                             void* stream);
int adaface_sbg_head_bwd_dev(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl_dev, int n_layers,
                             int64_t ldh, const float* w, const float* dout, int64_t lddo, float* dh0, float* dh1, float* dh2,
                             float* dh3, float* dwl, float* dw, float* db, int64_t M, int64_t C, float eps, void* stream);

/* DoRA column scale of a LoRA adapter (peft DoraLinearLayer, eval form): out[n] = m[n] / ||W[n,:] + s BA[n,:]||_2.
 * W fp32 | bf16 [N, K] contiguous (a convolution weight flattened over cin, kh, kw); BA fp32 [N, K] with row pitch ldba = the
 * product B.A (from adaface_proj_lora_fwd), or NULL; m, out fp32 [N]. */
int adaface_dora_colscale(const void* W, int w_dtype, const float* BA, int64_t ldba, float s, const float* m, float* out, int64_t N,
                          int64_t K, void* stream);

/* im2col for the weight gradient of a 3x3 convolution adapter (conv-LoRA A matrix; adaface/diffusers_attn_lora_capture.py:541-591,
 * peft lora.Conv2d): x bf16 NHWC [B, H, W, C] -> col bf16 [B*H*W, 9*Kc], Kc = C rounded up to 64,
 * col[p, (ky*3+kx)*Kc + c] = x[b, y+ky-1, x+kx-1, c] (zero outside the image and for c >= C) -- the K order of the packed
 * weights of adaface_conv3x3_fwd, so that dA_packed = dT^T col is one call of adaface_proj_lora_fwd.  C multiple of 8. */
int adaface_im2col3x3_tokens(const void* x, void* col, int64_t B, int64_t H, int64_t W, int64_t C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADAFACE_B200_H_ */
