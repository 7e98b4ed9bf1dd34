// Sampler-side elementwise kernels around the U-Net (BASELINE config 4: 50-step DDIM with classifier-free guidance).
//   ddim_cfg_step_kernel   ldm/models/diffusion/ddim.py:253-255 (CFG combine) + :280-301 (x0 prediction, x_{t-1}) in ONE pass.
// HBM-bound and tiny (32 KB per image); what matters is that the whole step stays on the device and inside the step's CUDA
// graph: the per-step scalars come from a DEVICE coefficient row, so one captured graph serves all 50 steps.
// Every operation is a separately rounded IEEE fp32 op in the reference's order (no FMA contraction, no fast-math division),
// so for identical eps the result is bit-identical to the reference's torch fp32 arithmetic.
#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

// coef (device, fp32[8]): 0 guidance scale, 1 sqrt(1 - a_t), 2 sqrt(a_t), 3 sqrt(a_prev), 4 sqrt(1 - a_prev - sigma^2), 5 sigma_t,
// 6 temperature, 7 unused
__device__ __forceinline__ float ddim_one(float ec, float eu, float x, float nz, bool cfg, const float (&c)[8], float& pred) {
  const float e = cfg ? __fadd_rn(eu, __fmul_rn(c[0], __fsub_rn(ec, eu))) : ec;            // ddim.py:255
  pred = __fdiv_rn(__fsub_rn(x, __fmul_rn(c[1], e)), c[2]);                                // :281
  const float dir = __fmul_rn(c[4], e);                                                    // :285
  const float noise = __fmul_rn(__fmul_rn(c[5], nz), c[6]);                                // :289
  return __fadd_rn(__fadd_rn(__fmul_rn(c[3], pred), dir), noise);                          // :301
}

__global__ void __launch_bounds__(256) ddim_cfg_step_kernel(const float4* __restrict__ eps, const float4* x, const float4* __restrict__ noise,
                                                            const float* __restrict__ coef, float4* x_prev, float4* x_dup,
                                                            float4* __restrict__ pred_x0, long long n4, int cfg) {
  float c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = __ldg(coef + i);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 ec = eps[i];
    const float4 eu = cfg ? eps[n4 + i] : ec;
    const float4 xv = x[i];
    const float4 nz = noise ? noise[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 p, o;
    o.x = ddim_one(ec.x, eu.x, xv.x, nz.x, cfg, c, p.x);
    o.y = ddim_one(ec.y, eu.y, xv.y, nz.y, cfg, c, p.y);
    o.z = ddim_one(ec.z, eu.z, xv.z, nz.z, cfg, c, p.z);
    o.w = ddim_one(ec.w, eu.w, xv.w, nz.w, cfg, c, p.w);
    x_prev[i] = o;
    if (x_dup) x_dup[i] = o;
    if (pred_x0) pred_x0[i] = p;
  }
}

int ddim_cfg_step(const float* eps, int64_t n_images, int64_t n_per_image, int has_uncond, const float* x, const float* coef,
                  const float* noise, float* x_prev, float* x_dup, float* pred_x0, cudaStream_t stream) {
  AF_CHECK(eps && x && coef && x_prev, "ddim_cfg_step: null pointer");
  AF_CHECK(n_images > 0 && n_per_image > 0 && n_per_image % 4 == 0, "ddim_cfg_step: n_per_image must be a positive multiple of 4 (got %lld)",
           (long long)n_per_image);
  for (const void* p : {(const void*)eps, (const void*)x, (const void*)noise, (const void*)x_prev, (const void*)x_dup, (const void*)pred_x0})
    AF_CHECK(((uintptr_t)p & 15) == 0, "ddim_cfg_step: pointers must be 16-byte aligned");
  const long long n4 = n_images * n_per_image / 4;
  long long grid = (n4 + 255) / 256;
  const long long cap = (long long)af_num_sms() * 8;
  if (grid > cap) grid = cap;
  ddim_cfg_step_kernel<<<(unsigned)grid, 256, 0, stream>>>((const float4*)eps, (const float4*)x, (const float4*)noise, coef, (float4*)x_prev,
                                                          (float4*)x_dup, (float4*)pred_x0, n4, has_uncond ? 1 : 0);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// DoRA column scale (peft DoraLinearLayer / DoraConv2dLayer in eval form): colscale[n] = m[n] / ||W[n,:] + s BA[n,:]||_2,
// the norm detached as in peft.  One warp per output row; W fp32 | bf16 [N, K] (a convolution weight flattened over cin, kh, kw),
// BA fp32 [N, K] = B.A from the projection GEMM (NULL: plain ||W||).  Runs once per parameter update -- with the B.A product on
// the tcgen05 GEMM this removes the last library arithmetic (cuBLAS sgemm + torch reduce) from the training step.
template <typename TW>
__global__ void __launch_bounds__(256) dora_colscale_kernel(const TW* __restrict__ W, const float* __restrict__ BA, long long ldba, float s,
                                                            const float* __restrict__ m, float* __restrict__ out, int N, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float w;
    if constexpr (sizeof(TW) == 2) w = __bfloat162float(W[(long long)row * K + k]);
    else w = W[(long long)row * K + k];
    if (BA) w += s * BA[(long long)row * ldba + k];
    acc += w * w;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = m[row] / sqrtf(acc);
}

int dora_colscale(const void* W, int w_dtype, const float* BA, int64_t ldba, float s, const float* m, float* out, int64_t N, int64_t K,
                  cudaStream_t stream) {
  AF_CHECK(W && m && out && N > 0 && K > 0, "dora_colscale: null pointer / empty");
  const unsigned grid = (unsigned)((N + 7) / 8);
  if (w_dtype == ADAFACE_F32) dora_colscale_kernel<float><<<grid, 256, 0, stream>>>((const float*)W, BA, ldba, s, m, out, (int)N, (int)K);
  else if (w_dtype == ADAFACE_BF16) dora_colscale_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)W, BA, ldba, s, m, out, (int)N, (int)K);
  else {
    set_error("dora_colscale: bad dtype %d", w_dtype);
    return 1;
  }
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// im2col of an NHWC activation for the WEIGHT gradient of a 3x3 convolution adapter (conv-LoRA A matrix, dalc:541-591):
//   col[(b, y, x), tap * Kc + c] = X[b, y + ky - 1, x + kx - 1, c]   (zero outside the image / for c >= C),  tap = ky * 3 + kx
// -- the K index of ops.pack_conv3x3_weight, so dA_packed [r, 9 Kc] = dT^T col is one K-major GEMM on the projection kernel.
// Training only, three adapters per U-Net pass at B = 1 (47 MB of col at level A): a plain 16-byte-vector copy kernel.
__global__ void __launch_bounds__(256) im2col3x3_tokens_kernel(const uint4* __restrict__ x, uint4* __restrict__ col, int H, int W, int vecC,
                                                               int vecK, long long n_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int v = (int)(i % vecK);
  long long r = i / vecK;
  const int tap = (int)(r % 9);
  r /= 9;
  const int px = (int)(r % W);
  r /= W;
  const int py = (int)(r % H);
  const long long b = r / H;
  const int sy = py + tap / 3 - 1, sx = px + tap % 3 - 1;
  uint4 val = make_uint4(0, 0, 0, 0);
  if (v < vecC && sy >= 0 && sy < H && sx >= 0 && sx < W) val = x[((b * H + sy) * W + sx) * vecC + v];
  col[i] = val;
}

int im2col3x3_tokens(const void* x, void* col, int64_t B, int64_t H, int64_t W, int64_t C, cudaStream_t stream) {
  AF_CHECK(x && col && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "im2col3x3_tokens: null pointer / empty / C not a multiple of 8");
  const int64_t kc = (C + 63) / 64 * 64;
  const long long n_out = (long long)B * H * W * 9 * (kc / 8);
  im2col3x3_tokens_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, stream>>>((const uint4*)x, (uint4*)col, (int)H, (int)W, (int)(C / 8),
                                                                              (int)(kc / 8), n_out);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

}  // namespace adaface
