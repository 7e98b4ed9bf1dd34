// C-ABI entry points of libadaface_b200.so (declared in include/adaface_b200.h) + host-side helpers.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

long long g_launch_count = 0;
static thread_local char g_err[1024] = "";

static int g_pdl = -1;           // bit mask: 1 = projection GEMMs, 2 = attention kernels
bool pdl_enabled(int kind) {
  if (g_pdl < 0) {
    const char* e = getenv("ADAFACE_PDL");       // default 1 (GEMMs only): measured on the 117-kernel step graph
    g_pdl = e ? atoi(e) : 1;                     // 0: 3.221 ms, 1: 3.141 ms, 2: 3.192 ms, 3: 3.151 ms
  }
  return (g_pdl & kind) != 0;
}

int af_num_sms() {
  static std::atomic<int> sms[AF_MAX_DEV];
  const int dev = af_device();
  int n = sms[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// cuTensorMapEncodeTiled is resolved through the runtime so that the library has no link-time
// dependency on libcuda.so (it must load on a machine without a driver for the symbol-export test).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box_rows=%u", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return 1;
  }
  return 0;
}

// Store-side map of a bf16 row-major [rows, cols] output (row pitch `ld` elements): {32 columns x 128 rows} boxes with the 64-byte
// swizzle -- the layout of the GEMM epilogue's output slabs (16-byte piece j of row r at j ^ ((r >> 1) & 3)).
int make_tmap_bf16_store32(CUtensorMap* out, void* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(store) failed (%d): base=%p rows=%llu cols=%llu ld=%llu", (int)r, base, (unsigned long long)rows,
              (unsigned long long)cols, (unsigned long long)ld);
    return 1;
  }
  return 0;
}

int make_tmap_bf16_heads(CUtensorMap* out, const void* base, uint64_t d, uint64_t heads, uint64_t L, uint64_t B,
                         uint64_t sh, uint64_t sn, uint64_t sb, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t gdim[4] = {d, heads, L, B};
  cuuint64_t gstride[3] = {sh * 2, sn * 2, sb * 2};
  cuuint32_t box[4] = {64, 1, box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4D) failed (%d): base=%p d=%llu heads=%llu L=%llu B=%llu sn=%llu sb=%llu", (int)r,
              base, (unsigned long long)d, (unsigned long long)heads, (unsigned long long)L, (unsigned long long)B,
              (unsigned long long)sn, (unsigned long long)sb);
    return 1;
  }
  return 0;
}

int make_tmap_bf16_nhwc(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint64_t sw,
                        uint64_t sh, uint64_t sb, uint32_t box_w, uint32_t box_h, uint32_t box_b) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t gdim[4] = {C, W, H, B};
  cuuint64_t gstride[3] = {sw * 2, sh * 2, sb * 2};
  cuuint32_t box[4] = {64, box_w, box_h, box_b};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(NHWC) failed (%d): base=%p C=%llu W=%llu H=%llu B=%llu box=%ux%ux%u", (int)r, base,
              (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B, box_w, box_h, box_b);
    return 1;
  }
  return 0;
}

int proj_lora_fwd(const void*, int64_t, const void*, const void*, int64_t, const void*, const float*, const float*,
                  const void*, int64_t, int, void*, int64_t, int, int64_t, int64_t, int64_t, int64_t, int, int64_t, int64_t,
                  int64_t, int64_t, cudaStream_t);
int attn_fwd(const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*, int64_t, int64_t, void*, int64_t,
             int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, const uint8_t*, int, float, float*, cudaStream_t);
int attn_fwd_tcgen05(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, const void*, int64_t,
                     int64_t, int64_t, void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                     float, float*, const uint8_t*, cudaStream_t);
int attn_bwd(const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*, int64_t,
             int64_t, const void*, int64_t, int64_t, const float*, float*, void*, int64_t, int64_t, void*, int64_t, int64_t,
             void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, const uint8_t*, int, float, cudaStream_t);
int attn_cross_capture_bwd_chunks(int64_t, int64_t, int64_t);
int attn_cross_capture_bwd(const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*, int64_t, int64_t,
                           const void*, int64_t, int64_t, const float*, const float*, int64_t, int64_t, int64_t, int64_t,
                           int64_t, float, const uint8_t*, const float*, const float*, int, int, void*, int64_t, int64_t, void*,
                           int64_t, int64_t, void*, int64_t, int64_t, int, float*, float, float*, float*, float*, cudaStream_t,
                           const uint8_t*, const float*, const float*, const float*);
int transpose(const void*, int, int64_t, int64_t, void*, int, int64_t, int64_t, int64_t, int64_t, int64_t, float, const float*,
              const float*, cudaStream_t);
int colsum(const void*, int, int64_t, const void*, int, int64_t, const float*, const float*, float*, int64_t, int64_t,
           cudaStream_t);
int layernorm_bwd(const void*, int, int64_t, const void*, int, int64_t, const float*, void*, int64_t, float*, float*, int64_t,
                  int64_t, float, cudaStream_t);
int act_fwd(const void*, int64_t, void*, int64_t, int64_t, int64_t, int, cudaStream_t);
int act_bwd(const void*, int64_t, const void*, int64_t, void*, int64_t, int64_t, int64_t, int, cudaStream_t);
int sbg_head_bwd(const float*, const float*, const float*, const float*, const float*, int, int64_t, const float*, const float*,
                 int64_t, float*, float*, float*, float*, float*, float*, float*, int64_t, int64_t, float, cudaStream_t, const float*);
int attn_cross_capture_fwd(const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*, int64_t, int64_t,
                           void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, float, float*, float*,
                           float*, const int32_t*, int64_t, const uint8_t*, const float*, const float*, int, int, cudaStream_t,
                           const uint8_t*, float*, const float*, float*, int64_t);
int qmean(const void*, int, int64_t, int64_t, int64_t, int64_t, int64_t, float*, cudaStream_t);
int capture_chan_major(const void*, int, int64_t, int64_t, int64_t, int64_t, int64_t, float, float*, cudaStream_t);
int layernorm_fwd(const void*, int, int64_t, const float*, const float*, void*, int, int64_t, int64_t, int64_t, float,
                  cudaStream_t);
int sbg_head_fwd(const float*, const float*, const float*, const float*, const float*, int, int64_t, const float*,
                 const float*, float*, int64_t, int64_t, int64_t, float, cudaStream_t, const float*);
int groupnorm_tokens_fwd(const void*, int, const float*, const float*, int64_t, int64_t, int64_t, int64_t, float, float*, float*,
                         void*, cudaStream_t);
int tokens_to_nchw_add(const void*, const void*, int, void*, int64_t, int64_t, int64_t, cudaStream_t);
int conv3x3_fwd(const void*, int64_t, int64_t, int64_t, int64_t, const void*, const void*, int64_t, const void*, int64_t, const float*,
                const float*, const float*, const void*, int64_t, int, void*, int64_t, int, int64_t, int, int, cudaStream_t);
int groupnorm_act_tokens_fwd(const void*, const float*, const float*, int64_t, int64_t, int64_t, int64_t, float, int, float*, float*,
                             float*, void*, cudaStream_t);
int64_t groupnorm_act_tokens_ws_floats(int64_t, int64_t, int64_t, int64_t);
int silu_fwd(const void*, int, void*, int64_t, cudaStream_t);
int im2col3x3_tokens(const void*, void*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
int dora_colscale(const void*, int, const float*, int64_t, float, const float*, float*, int64_t, int64_t, cudaStream_t);
int ddim_cfg_step(const float*, int64_t, int64_t, int, const float*, const float*, const float*, float*, float*, float*, cudaStream_t);
int upsample2x_tokens(const void*, void*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
int timestep_embedding(const float*, int64_t, int64_t, float, void*, cudaStream_t);
int groupnorm_act_tokens_bwd(const void*, const void*, const float*, const float*, int64_t, int64_t, int64_t, int64_t, float, int, float*,
                             void*, cudaStream_t);
int resample2x_bwd(const void*, void*, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);

}  // namespace adaface

using namespace adaface;

extern "C" {

int adaface_version(void) { return ADAFACE_B200_ABI_VERSION; }
const char* adaface_last_error(void) { return g_err; }
int64_t adaface_launch_count(void) { return g_launch_count; }
int adaface_set_pdl(int mask) {
  pdl_enabled(1);
  const int prev = g_pdl;
  g_pdl = mask;
  return prev;
}

int adaface_proj_lora_fwd(const void* x, int64_t ldx, const void* w, const void* t, int64_t ldt, const void* bs,
                          const float* colscale, const float* bias, const void* residual, int64_t ldr,
                          int residual_dtype, void* y, int64_t ldy, int y_dtype, int64_t M, int64_t N, int64_t K,
                          int64_t R, int act, void* stream) {
  return proj_lora_fwd(x, ldx, w, t, ldt, bs, colscale, bias, residual, ldr, residual_dtype, y, ldy, y_dtype, M, N, K, R,
                       act, 0, 0, 0, 0, (cudaStream_t)stream);
}

int adaface_proj_lora_heads_fwd(const void* x, int64_t ldx, const void* w, const void* t, int64_t ldt, const void* bs,
                                const float* colscale, const float* bias, void* y, int64_t M, int64_t N, int64_t K,
                                int64_t R, int64_t heads, int64_t d, int64_t dpad, int64_t rows_per_batch, void* stream) {
  return proj_lora_fwd(x, ldx, w, t, ldt, bs, colscale, bias, nullptr, 0, ADAFACE_BF16, y, 0, ADAFACE_BF16, M, N, K, R,
                       ADAFACE_ACT_NONE, heads, d, dpad, rows_per_batch, (cudaStream_t)stream);
}

int adaface_attn_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                     const void* v, int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb, int64_t o_sn, int64_t B,
                     int64_t H, int64_t Lq, int64_t Lk, int64_t d, const uint8_t* key_mask, int causal_mult,
                     float scale, float* lse, void* stream) {
  return attn_fwd(q, q_sb, q_sn, k, k_sb, k_sn, v, v_sb, v_sn, o, o_sb, o_sn, B, H, Lq, Lk, d, key_mask, causal_mult,
                  scale, lse, (cudaStream_t)stream);
}

int adaface_attn_headmajor_fwd(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, const void* k, int64_t k_sb,
                               int64_t k_sh, int64_t k_sn, const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sn, void* o,
                               int64_t o_sb, int64_t o_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
                               int64_t drow_q, int64_t drow_kv, float scale, float* lse, void* stream) {
  const int rc = attn_fwd_tcgen05(q, q_sb, q_sh, q_sn, k, k_sb, k_sh, k_sn, v, v_sb, v_sh, v_sn, o, o_sb, o_sn, B, H, Lq, Lk,
                                  d, drow_q, drow_kv, scale, lse, nullptr, (cudaStream_t)stream);
  if (rc < 0) set_error("adaface_attn_headmajor_fwd: unsupported head dim %lld (40, 80, 160)", (long long)d);
  return rc < 0 ? 1 : rc;
}

int adaface_attn_cross_capture_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb,
                                   int64_t k_sn, const void* v, int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb,
                                   int64_t o_sn, int64_t B, int64_t H, int64_t Lq, int64_t S, int64_t d, float scale,
                                   float* prob, float* score, float* prob_subj, const int32_t* subj_cols,
                                   int64_t n_subj, const uint8_t* col_flag, const float* qmean_,
                                   const float* ca_scale, int mix, int in_dtype, void* stream) {
  return attn_cross_capture_fwd(q, q_sb, q_sn, k, k_sb, k_sn, v, v_sb, v_sn, o, o_sb, o_sn, B, H, Lq, S, d, scale, prob,
                                score, prob_subj, subj_cols, n_subj, col_flag, qmean_, ca_scale, mix, in_dtype,
                                (cudaStream_t)stream, nullptr, nullptr, nullptr, nullptr, 0);
}

int adaface_attn_cross_consume_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                                   const void* v, int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb, int64_t o_sn, int64_t B,
                                   int64_t H, int64_t Lq, int64_t S, int64_t d, float scale, float* prob,
                                   const uint8_t* col_flag, const float* qmean_, const float* ca_scale, int in_dtype,
                                   const uint8_t* sum_flag, float* subj_sum, const float* ref_prob, float* sq_part,
                                   int64_t sq_slots, void* stream) {
  return attn_cross_capture_fwd(q, q_sb, q_sn, k, k_sb, k_sn, v, v_sb, v_sn, o, o_sb, o_sn, B, H, Lq, S, d, scale, prob, nullptr,
                                nullptr, nullptr, 0, col_flag, qmean_, ca_scale, 0, in_dtype, (cudaStream_t)stream, sum_flag,
                                subj_sum, ref_prob, sq_part, sq_slots);
}

int adaface_qmean(const void* q, int q_dtype, int64_t q_sb, int64_t q_sn, int64_t B, int64_t Lq, int64_t C, float* out,
                  void* stream) {
  return qmean(q, q_dtype, q_sb, q_sn, B, Lq, C, out, (cudaStream_t)stream);
}

int adaface_capture_chan_major(const void* src, int src_dtype, int64_t s_sb, int64_t s_sn, int64_t B, int64_t L,
                               int64_t C, float factor, float* dst, void* stream) {
  return capture_chan_major(src, src_dtype, s_sb, s_sn, B, L, C, factor, dst, (cudaStream_t)stream);
}

int adaface_layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const float* w, const float* b, void* y,
                          int y_dtype, int64_t ldy, int64_t M, int64_t C, float eps, void* stream) {
  return layernorm_fwd(x, x_dtype, ldx, w, b, y, y_dtype, ldy, M, C, eps, (cudaStream_t)stream);
}

int adaface_sbg_head_fwd(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl,
                         int n_layers, int64_t ldh, const float* w, const float* b, float* out, int64_t ldo,
                         int64_t M, int64_t C, float eps, void* stream) {
  return sbg_head_fwd(h0, h1, h2, h3, wl, n_layers, ldh, w, b, out, ldo, M, C, eps, (cudaStream_t)stream, nullptr);
}
int adaface_sbg_head_fwd_dev(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl_dev, int n_layers,
                             int64_t ldh, const float* w, const float* b, float* out, int64_t ldo, int64_t M, int64_t C, float eps,
                             void* stream) {
  return sbg_head_fwd(h0, h1, h2, h3, nullptr, n_layers, ldh, w, b, out, ldo, M, C, eps, (cudaStream_t)stream, wl_dev);
}

int adaface_groupnorm_tokens_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, int64_t B, int64_t C,
                                 int64_t HW, int64_t groups, float eps, float* a_ws, float* s_ws, void* y, void* stream) {
  return groupnorm_tokens_fwd(x, x_dtype, gamma, beta, B, C, HW, groups, eps, a_ws, s_ws, y, (cudaStream_t)stream);
}

int adaface_conv3x3_fwd(const void* x, int64_t B, int64_t H, int64_t W, int64_t Cin, const void* w, const void* t, int64_t ldt,
                        const void* bs, int64_t R, const float* colscale, const float* bias, const float* rowbias,
                        const void* residual, int64_t ldr, int residual_dtype, void* y, int64_t ldy, int y_dtype, int64_t Cout,
                        int stride, int act, void* stream) {
  return conv3x3_fwd(x, B, H, W, Cin, w, t, ldt, bs, R, colscale, bias, rowbias, residual, ldr, residual_dtype, y, ldy, y_dtype, Cout,
                     stride, act, (cudaStream_t)stream);
}

int adaface_groupnorm_act_tokens_fwd(const void* x, const float* gamma, const float* beta, int64_t B, int64_t HW, int64_t C,
                                     int64_t groups, float eps, int act, float* part_ws, float* a_ws, float* s_ws, void* y,
                                     void* stream) {
  return groupnorm_act_tokens_fwd(x, gamma, beta, B, HW, C, groups, eps, act, part_ws, a_ws, s_ws, y, (cudaStream_t)stream);
}
int64_t adaface_groupnorm_act_tokens_ws_floats(int64_t B, int64_t HW, int64_t C, int64_t groups) {
  return groupnorm_act_tokens_ws_floats(B, HW, C, groups);
}
int adaface_ddim_cfg_step(const float* eps, int64_t n_images, int64_t n_per_image, int has_uncond, const float* x, const float* coef,
                          const float* noise, float* x_prev, float* x_dup, float* pred_x0, void* stream) {
  return ddim_cfg_step(eps, n_images, n_per_image, has_uncond, x, coef, noise, x_prev, x_dup, pred_x0, (cudaStream_t)stream);
}
int adaface_dora_colscale(const void* W, int w_dtype, const float* BA, int64_t ldba, float s, const float* m, float* out, int64_t N,
                          int64_t K, void* stream) {
  return dora_colscale(W, w_dtype, BA, ldba, s, m, out, N, K, (cudaStream_t)stream);
}
int adaface_im2col3x3_tokens(const void* x, void* col, int64_t B, int64_t H, int64_t W, int64_t C, void* stream) {
  return im2col3x3_tokens(x, col, B, H, W, C, (cudaStream_t)stream);
}
int adaface_silu_fwd(const void* x, int x_dtype, void* y, int64_t n, void* stream) { return silu_fwd(x, x_dtype, y, n, (cudaStream_t)stream); }
int adaface_upsample2x_tokens(const void* x, void* y, int64_t B, int64_t H, int64_t W, int64_t C, void* stream) {
  return upsample2x_tokens(x, y, B, H, W, C, (cudaStream_t)stream);
}

int adaface_groupnorm_act_tokens_bwd(const void* x, const void* dy, const float* gamma, const float* beta, int64_t B, int64_t HW,
                                     int64_t C, int64_t groups, float eps, int act, float* coef_ws, void* dx, void* stream) {
  return groupnorm_act_tokens_bwd(x, dy, gamma, beta, B, HW, C, groups, eps, act, coef_ws, dx, (cudaStream_t)stream);
}
int adaface_resample2x_bwd(const void* x, void* y, int64_t B, int64_t H, int64_t W, int64_t C, int mode, void* stream) {
  return resample2x_bwd(x, y, B, H, W, C, mode, (cudaStream_t)stream);
}

int adaface_timestep_embedding(const float* t, int64_t B, int64_t dim, float max_period, void* out, void* stream) {
  return timestep_embedding(t, B, dim, max_period, out, (cudaStream_t)stream);
}

int adaface_tokens_to_nchw_add(const void* t, const void* x_in, int x_dtype, void* out, int64_t B, int64_t C, int64_t HW,
                               void* stream) {
  return tokens_to_nchw_add(t, x_in, x_dtype, out, B, C, HW, (cudaStream_t)stream);
}

int adaface_attn_bwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn, const void* v,
                     int64_t v_sb, int64_t v_sn, const void* o, int64_t o_sb, int64_t o_sn, const void* dout, int64_t do_sb,
                     int64_t do_sn, const float* lse, float* delta, void* dq, int64_t dq_sb, int64_t dq_sn, void* dk,
                     int64_t dk_sb, int64_t dk_sn, void* dv, int64_t dv_sb, int64_t dv_sn, int64_t B, int64_t H, int64_t Lq,
                     int64_t Lk, int64_t d, const uint8_t* key_mask, int causal_mult, float scale, void* stream) {
  return attn_bwd(q, q_sb, q_sn, k, k_sb, k_sn, v, v_sb, v_sn, o, o_sb, o_sn, dout, do_sb, do_sn, lse, delta, dq, dq_sb, dq_sn,
                  dk, dk_sb, dk_sn, dv, dv_sb, dv_sn, B, H, Lq, Lk, d, key_mask, causal_mult, scale, (cudaStream_t)stream);
}

int adaface_attn_cross_capture_bwd_chunks(int64_t B, int64_t H, int64_t Lq) {
  return attn_cross_capture_bwd_chunks(B, H, Lq);
}

int adaface_attn_cross_capture_bwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                                   const void* v, int64_t v_sb, int64_t v_sn, const void* dout, int64_t do_sb,
                                   int64_t do_sn, const float* dprob, const float* dscore, int64_t B, int64_t H,
                                   int64_t Lq, int64_t S, int64_t d, float scale, const uint8_t* col_flag,
                                   const float* qmean_, const float* ca_scale, int mix, int in_dtype, void* dq, int64_t dq_sb,
                                   int64_t dq_sn, void* dk, int64_t dk_sb, int64_t dk_sn, void* dv, int64_t dv_sb,
                                   int64_t dv_sn, int dkv_dtype, float* dca, float dca_mul, float* dk_part,
                                   float* dv_part, float* dca_part, void* stream) {
  return attn_cross_capture_bwd(q, q_sb, q_sn, k, k_sb, k_sn, v, v_sb, v_sn, dout, do_sb, do_sn, dprob, dscore, B, H, Lq, S, d,
                                scale, col_flag, qmean_, ca_scale, mix, in_dtype, dq, dq_sb, dq_sn, dk, dk_sb, dk_sn, dv, dv_sb,
                                dv_sn, dkv_dtype, dca, dca_mul, dk_part, dv_part, dca_part, (cudaStream_t)stream, nullptr, nullptr,
                                nullptr, nullptr);
}

int adaface_attn_cross_consume_bwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                                   const void* v, int64_t v_sb, int64_t v_sn, const void* dout, int64_t do_sb, int64_t do_sn,
                                   const float* dprob, int64_t B, int64_t H, int64_t Lq, int64_t S, int64_t d, float scale,
                                   const uint8_t* col_flag, const float* qmean_, const float* ca_scale, int in_dtype, void* dq,
                                   int64_t dq_sb, int64_t dq_sn, void* dk, int64_t dk_sb, int64_t dk_sn, void* dv, int64_t dv_sb,
                                   int64_t dv_sn, int dkv_dtype, float* dca, float dca_mul, float* dk_part, float* dv_part,
                                   float* dca_part, const uint8_t* sum_flag, const float* g_subj, const float* ref_prob,
                                   const float* mse_coef, void* stream) {
  return attn_cross_capture_bwd(q, q_sb, q_sn, k, k_sb, k_sn, v, v_sb, v_sn, dout, do_sb, do_sn, dprob, nullptr, B, H, Lq, S, d,
                                scale, col_flag, qmean_, ca_scale, 0, in_dtype, dq, dq_sb, dq_sn, dk, dk_sb, dk_sn, dv, dv_sb,
                                dv_sn, dkv_dtype, dca, dca_mul, dk_part, dv_part, dca_part, (cudaStream_t)stream, sum_flag, g_subj,
                                ref_prob, mse_coef);
}

int adaface_transpose(const void* src, int src_dtype, int64_t s_sb, int64_t s_ld, void* dst, int dst_dtype, int64_t d_sb,
                      int64_t d_ld, int64_t B, int64_t I, int64_t J, float alpha, const float* colscale,
                      const float* rowscale, void* stream) {
  return transpose(src, src_dtype, s_sb, s_ld, dst, dst_dtype, d_sb, d_ld, B, I, J, alpha, colscale, rowscale,
                   (cudaStream_t)stream);
}

int adaface_colsum(const void* a, int a_dtype, int64_t lda, const void* b, int b_dtype, int64_t ldb, const float* bias,
                   const float* colmul, float* out, int64_t M, int64_t N, void* stream) {
  return colsum(a, a_dtype, lda, b, b_dtype, ldb, bias, colmul, out, M, N, (cudaStream_t)stream);
}

int adaface_layernorm_bwd(const void* x, int x_dtype, int64_t ldx, const void* dy, int dy_dtype, int64_t lddy,
                          const float* w, void* dx, int64_t lddx, float* dw, float* db, int64_t M, int64_t C, float eps,
                          void* stream) {
  return layernorm_bwd(x, x_dtype, ldx, dy, dy_dtype, lddy, w, dx, lddx, dw, db, M, C, eps, (cudaStream_t)stream);
}

int adaface_act_fwd(const void* u, int64_t ldu, void* h, int64_t ldh, int64_t M, int64_t n_out, int act, void* stream) {
  return act_fwd(u, ldu, h, ldh, M, n_out, act, (cudaStream_t)stream);
}

int adaface_act_bwd(const void* u, int64_t ldu, const void* dh, int64_t lddh, void* du, int64_t lddu, int64_t M,
                    int64_t n_out, int act, void* stream) {
  return act_bwd(u, ldu, dh, lddh, du, lddu, M, n_out, act, (cudaStream_t)stream);
}

int adaface_sbg_head_bwd(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl,
                         int n_layers, int64_t ldh, const float* w, const float* dout, int64_t lddo, float* dh0,
                         float* dh1, float* dh2, float* dh3, float* dwl, float* dw, float* db, int64_t M, int64_t C,
                         float eps, void* stream) {
  return sbg_head_bwd(h0, h1, h2, h3, wl, n_layers, ldh, w, dout, lddo, dh0, dh1, dh2, dh3, dwl, dw, db, M, C, eps,
                      (cudaStream_t)stream, nullptr);
}
int adaface_sbg_head_bwd_dev(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl_dev, int n_layers,
                             int64_t ldh, const float* w, const float* dout, int64_t lddo, float* dh0, float* dh1, float* dh2,
                             float* dh3, float* dwl, float* dw, float* db, int64_t M, int64_t C, float eps, void* stream) {
  return sbg_head_bwd(h0, h1, h2, h3, nullptr, n_layers, ldh, w, dout, lddo, dh0, dh1, dh2, dh3, dwl, dw, db, M, C, eps,
                      (cudaStream_t)stream, wl_dev);
}

}  // extern "C"
