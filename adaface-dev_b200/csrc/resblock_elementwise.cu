// HBM-bound helpers of the U-Net ResBlock / Upsample path on NHWC ("tokens", [B, H*W, C] bf16) activations
// (ldm/modules/diffusionmodules/openaimodel.py:92-277).  The convolutions themselves are the CONV
// instantiation of the tcgen05 GEMM (gemm_tcgen05.cu); here:
//   gn_tokens_stats | gn_tokens_partial + gn_tokens_finalize, then gn_tokens_apply
//                     GroupNorm32 (fp32 statistics, util.py normalization() = GroupNorm(32, C)) + SiLU in front of each
//                     3x3 convolution (openaimodel.py:203-207, 229-236), tokens in -> tokens out
//   silu              SiLU of the time embedding in front of emb_layers' Linear (openaimodel.py:222-228)
//   upsample2x        nearest-neighbour 2x of Upsample.forward (openaimodel.py:109-119) in NHWC
#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

constexpr int GN_ROWS = 64;        // most rows of one image summed by one CTA of the partial kernel (fewer for small maps)
constexpr int GN_MAX_PAIRS = 1280; // C <= 2560

// part[((b * chunks) + chunk) * groups + g] = (sum, sum of squares) over rows [chunk * GN_ROWS, +GN_ROWS) of group g.
// Threads walk channel PAIRS (a pair never straddles a group: C / groups is even), so every load instruction of a warp
// reads 128 contiguous bytes of one row.
__global__ void __launch_bounds__(256) gn_tokens_partial_kernel(const bf16* __restrict__ x, float2* __restrict__ part, int C, int HW,
                                                                 int groups, int rows) {
  __shared__ float ps[GN_MAX_PAIRS], pq[GN_MAX_PAIRS];
  const int chunk = blockIdx.x, b = blockIdx.y, chunks = gridDim.x;
  const int r0 = chunk * rows, r1 = min(HW, r0 + rows);
  const int pairs = C >> 1;
  const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(x + ((long long)b * HW + r0) * C);
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    float s = 0.f, q = 0.f;
#pragma unroll 8
    for (int r = 0; r < r1 - r0; ++r) {
      const float2 v = __bfloat1622float2(xb[(long long)r * pairs + p]);
      s += v.x + v.y;
      q += v.x * v.x + v.y * v.y;
    }
    ps[p] = s;
    pq[p] = q;
  }
  __syncthreads();
  const int ppg = pairs / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int i = 0; i < ppg; ++i) {
      s += ps[g * ppg + i];
      q += pq[g * ppg + i];
    }
    part[((long long)b * chunks + chunk) * groups + g] = make_float2(s, q);
  }
}

// One CTA per image: group statistics from the partials (fixed summation order: deterministic), folded with the affine
// parameters into a[b, c] = rstd * gamma[c], s[b, c] = beta[c] - mean * rstd * gamma[c].
__global__ void __launch_bounds__(256) gn_tokens_finalize_kernel(const float2* __restrict__ part, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float* __restrict__ a,
                                                                  float* __restrict__ s, int C, int HW, int groups, int chunks, float eps) {
  __shared__ float mean_s[256], rstd_s[256];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpg = C / groups;
  for (int g = warp; g < groups; g += 8) {
    float sum = 0.f, sq = 0.f;
    for (int c = lane; c < chunks; c += 32) {
      const float2 v = part[((long long)b * chunks + c) * groups + g];
      sum += v.x;
      sq += v.y;
    }
    sum = warp_sum(sum);
    sq = warp_sum(sq);
    if (lane == 0) {
      const float n = (float)cpg * (float)HW;
      const float mean = sum / n;
      mean_s[g] = mean;
      rstd_s[g] = rsqrtf(fmaxf(sq / n - mean * mean, 0.f) + eps);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float ga = gamma[c] * rstd_s[g];
    a[(long long)b * C + c] = ga;
    s[(long long)b * C + c] = beta[c] - mean_s[g] * ga;
  }
}

constexpr int GN_STATS_THREADS = 1024;

// One CTA per (group, image): mean / rstd of the group's [HW, C / groups] slab of the tokens, folded with the affine
// parameters into a[b, c] = rstd * gamma[c], s[b, c] = beta[c] - mean * rstd * gamma[c].  Threads form a TR x TP grid
// (TP = pairs per group rounded up to a power of two): a row of the slab is ppg adjacent bf16 pairs, rows are C apart.
// 1024 threads so that each walks only HW / TR rows (the loop is latency-bound: one 4-byte load per row).
// Fixed summation order: deterministic.
template <int TP>
__global__ void __launch_bounds__(GN_STATS_THREADS) gn_tokens_stats_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                                            const float* __restrict__ beta, float* __restrict__ a,
                                                                            float* __restrict__ s, int C, int HW, int groups, float eps) {
  constexpr int TR = GN_STATS_THREADS / TP;
  __shared__ float red[2][32];
  const int g = blockIdx.x, b = blockIdx.y, cpg = C / groups, ppg = cpg >> 1, pairs = C >> 1;
  const int p = threadIdx.x % TP, r0 = threadIdx.x / TP;
  const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(x + (long long)b * HW * C + g * cpg);
  float sum = 0.f, sq = 0.f;
  if (p < ppg) {
#pragma unroll 4
    for (int r = r0; r < HW; r += TR) {
      const float2 v = __bfloat1622float2(xb[(long long)r * pairs + p]);
      sum += v.x + v.y;
      sq += v.x * v.x + v.y * v.y;
    }
  }
  sum = warp_sum(sum);
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sum;
    red[1][threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  float ts = 0.f, tq = 0.f;
#pragma unroll
  for (int i = 0; i < GN_STATS_THREADS / 32; ++i) {
    ts += red[0][i];
    tq += red[1][i];
  }
  const float n = (float)cpg * (float)HW;
  const float mean = ts / n;
  const float rstd = rsqrtf(fmaxf(tq / n - mean * mean, 0.f) + eps);
  for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
    const int ch = g * cpg + c;
    const float ga = gamma[ch] * rstd;
    a[(long long)b * C + ch] = ga;
    s[(long long)b * C + ch] = beta[ch] - mean * ga;
  }
}

// SiLU.  The GroupNorm + SiLU apply pass is MUFU-bound, not HBM-bound: v / (1 + exp(-v)) costs an ex2 AND a reciprocal on the
// 16-lane special-function unit (B = 8, 64 x 64, 320 channels: 84 M elements x 2 = 36 us of MUFU time against 6.5 us of HBM time).
// x sigmoid(x) = 0.5 x (1 + tanh(x / 2)) needs ONE MUFU op (tanh.approx.f32, relative error 2^-11: below the bf16 rounding of y).
__device__ __forceinline__ float silu_f(float v) {
  float t;
  const float h = 0.5f * v;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// y[b, r, c] = act(x[b, r, c] * a[b, c] + s[b, c]); 8 channels (16 bytes) per thread, a / s of the image in shared memory.
template <int ACT>
__global__ void __launch_bounds__(256) gn_tokens_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ a,
                                                               const float* __restrict__ s, bf16* __restrict__ y, int C, int HW,
                                                               int rows_per_cta) {
  extern __shared__ float as_s[];    // [2][C]
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    as_s[c] = a[(long long)b * C + c];
    as_s[C + c] = s[(long long)b * C + c];
  }
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
  const int vec = C >> 3;
  const long long base = ((long long)b * HW + r0) * vec;
  const uint4* xv = reinterpret_cast<const uint4*>(x) + base;
  uint4* yv = reinterpret_cast<uint4*>(y) + base;
  const int n = (r1 - r0) * vec;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c0 = (i % vec) << 3;
    const uint4 u = xv[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    const float4 a0 = *reinterpret_cast<const float4*>(as_s + c0), a1 = *reinterpret_cast<const float4*>(as_s + c0 + 4);
    const float4 s0 = *reinterpret_cast<const float4*>(as_s + C + c0), s1 = *reinterpret_cast<const float4*>(as_s + C + c0 + 4);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
      lo = lo * av[2 * j] + sv[2 * j];
      hi = hi * av[2 * j + 1] + sv[2 * j + 1];
      if (ACT == 1) {
        lo = silu_f(lo);
        hi = silu_f(hi);
      }
      o[j] = pack_bf16(lo, hi);
    }
    yv[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// narrow groups take the row-split statistics path (see groupnorm_act_tokens_fwd)
static bool gn_row_split(int64_t C, int64_t groups) { return C / groups <= 16 && groups <= 256 && C <= 2 * GN_MAX_PAIRS; }

// rows per CTA of the partial kernel: ~2 CTAs per SM when the map is small, never more than GN_ROWS
static int gn_partial_rows(int64_t B, int64_t HW) {
  long long rows = (B * HW + 295) / 296;
  if (rows < 4) rows = 4;
  if (rows > GN_ROWS) rows = GN_ROWS;
  return (int)rows;
}

int groupnorm_act_tokens_fwd(const void* x, const float* gamma, const float* beta, int64_t B, int64_t HW, int64_t C, int64_t groups,
                             float eps, int act, float* part_ws, float* a_ws, float* s_ws, void* y, cudaStream_t stream) {
  AF_CHECK(x && gamma && beta && a_ws && s_ws && y, "groupnorm_act_tokens_fwd: null pointer");
  AF_CHECK(B > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 65535 && C % groups == 0 && (C / groups) % 2 == 0 && C % 8 == 0 &&
               C / groups <= 256 && C <= 6144 && B <= 65535,
           "groupnorm_act_tokens_fwd: bad shape B=%lld HW=%lld C=%lld groups=%lld", (long long)B, (long long)HW, (long long)C,
           (long long)groups);
  AF_CHECK(act == 0 || act == 1, "groupnorm_act_tokens_fwd: act %d (0 none | 1 SiLU)", act);
  AF_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "groupnorm_act_tokens_fwd: x / y must be 16-byte aligned");
  const int ppg = (int)(C / groups / 2);
  int n_launch = 2;
  if (gn_row_split(C, groups)) {
    // narrow groups (level A: 10 channels = 20 bytes per row): a CTA per group would gather 20-byte pieces; sum whole
    // rows per CTA instead and reduce the per-chunk partials in a second small kernel
    AF_CHECK(part_ws != nullptr && groups <= 256 && C <= 2 * GN_MAX_PAIRS, "groupnorm_act_tokens_fwd: row-split path needs part_ws, groups <= 256, C <= 2560");
    const int prow = gn_partial_rows(B, HW);
    const int chunks = (int)((HW + prow - 1) / prow);
    gn_tokens_partial_kernel<<<dim3((unsigned)chunks, (unsigned)B), 256, 0, stream>>>((const bf16*)x, (float2*)part_ws, (int)C, (int)HW, (int)groups, prow);
    gn_tokens_finalize_kernel<<<(unsigned)B, 256, 0, stream>>>((const float2*)part_ws, gamma, beta, a_ws, s_ws, (int)C, (int)HW, (int)groups, chunks, eps);
    n_launch = 3;
  } else {
    const dim3 gs((unsigned)groups, (unsigned)B);
#define AF_GN_STATS(TP) \
  gn_tokens_stats_kernel<TP><<<gs, GN_STATS_THREADS, 0, stream>>>((const bf16*)x, gamma, beta, a_ws, s_ws, (int)C, (int)HW, (int)groups, eps)
    if (ppg <= 16) AF_GN_STATS(16);
    else if (ppg <= 32) AF_GN_STATS(32);
    else if (ppg <= 64) AF_GN_STATS(64);
    else AF_GN_STATS(128);
#undef AF_GN_STATS
  }
  long long rows = (B * HW + 295) / 296;            // ~2 CTAs per SM
  if (rows < 4) rows = 4;
  if (rows > 64) rows = 64;
  const dim3 grid((unsigned)((HW + rows - 1) / rows), (unsigned)B);
  const size_t smem = (size_t)(2 * C * sizeof(float));
  if (act == 1) gn_tokens_apply_kernel<1><<<grid, 256, smem, stream>>>((const bf16*)x, a_ws, s_ws, (bf16*)y, (int)C, (int)HW, (int)rows);
  else gn_tokens_apply_kernel<0><<<grid, 256, smem, stream>>>((const bf16*)x, a_ws, s_ws, (bf16*)y, (int)C, (int)HW, (int)rows);
  AF_CUDA(cudaGetLastError());
  g_launch_count += n_launch;
  return 0;
}

int64_t groupnorm_act_tokens_ws_floats(int64_t B, int64_t HW, int64_t C, int64_t groups) {
  if (!gn_row_split(C, groups)) return 0;
  const int prow = gn_partial_rows(B, HW);
  return 2 * B * ((HW + prow - 1) / prow) * groups;
}

// ---------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) silu_kernel(const TIn* __restrict__ x, bf16* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v;
    if constexpr (sizeof(TIn) == 2) v = __bfloat162float(x[i]);
    else v = x[i];
    y[i] = __float2bfloat16(silu_f(v));
  }
}

int silu_fwd(const void* x, int x_dtype, void* y, int64_t n, cudaStream_t stream) {
  AF_CHECK(x && y && n > 0, "silu_fwd: null pointer / empty");
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (x_dtype == ADAFACE_F32) silu_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (bf16*)y, (long long)n);
  else if (x_dtype == ADAFACE_BF16) silu_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)x, (bf16*)y, (long long)n);
  else {
    set_error("silu_fwd: bad dtype %d", x_dtype);
    return 1;
  }
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// y[b, 2h + i, 2w + j, :] = x[b, h, w, :]  (i, j in {0, 1}); one thread per 16 bytes of OUTPUT
__global__ void __launch_bounds__(256) upsample2x_tokens_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int vec,
                                                                 long long n_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int v = (int)(i % vec);
  long long p = i / vec;
  const int ox = (int)(p % (2 * W));
  p /= 2 * W;
  const int oy = (int)(p % (2 * H));
  const long long b = p / (2 * H);
  y[i] = x[((b * H + (oy >> 1)) * W + (ox >> 1)) * vec + v];
}

int upsample2x_tokens(const void* x, void* y, int64_t B, int64_t H, int64_t W, int64_t C, cudaStream_t stream) {
  AF_CHECK(x && y, "upsample2x_tokens: null pointer");
  AF_CHECK(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "upsample2x_tokens: bad shape");
  AF_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "upsample2x_tokens: x / y must be 16-byte aligned");
  const long long n_out = B * 4 * H * W * (C / 8);
  upsample2x_tokens_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, stream>>>((const uint4*)x, (uint4*)y, (int)H, (int)W, (int)(C / 8), n_out);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Sinusoidal timestep embedding (ldm/modules/diffusionmodules/util.py:154-174): out[b] = [cos(t_b f_i) | sin(t_b f_i)],
// f_i = exp(-ln(max_period) i / half).  B x dim values: evaluated in double so that --use_fast_math cannot touch them.
__global__ void __launch_bounds__(256) timestep_embedding_kernel(const float* __restrict__ t, bf16* __restrict__ out, int B, int dim,
                                                                  float max_period) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const int b = idx / dim, j = idx - b * dim, half = dim / 2;
  float v = 0.f;
  if (j < 2 * half) {
    const int i = j < half ? j : j - half;
    const float freq = (float)exp(-log((double)max_period) * (double)i / (double)half);
    const float arg = t[b] * freq;
    v = (float)(j < half ? cos((double)arg) : sin((double)arg));
  }
  out[idx] = __float2bfloat16(v);
}

int timestep_embedding(const float* t, int64_t B, int64_t dim, float max_period, void* out, cudaStream_t stream) {
  AF_CHECK(t && out && B > 0 && dim > 1 && B * dim < (1ll << 31), "timestep_embedding: bad arguments");
  timestep_embedding_kernel<<<(unsigned)((B * dim + 255) / 256), 256, 0, stream>>>(t, (bf16*)out, (int)B, (int)dim, max_period);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Backward of y = act(GroupNorm(x) gamma + beta) with frozen gamma / beta (the U-Net is frozen in stage 2, ddpm.py:637-638):
//   y0 = xhat gamma + beta, xhat = (x - mu) r;   g = dy act'(y0);   dyh = g gamma;
//   dx = r (dyh - mean_group(dyh) - xhat mean_group(dyh xhat))  =  a_c g + P_g x + Q_g
// with a_c = r gamma_c, P_g = -r^2 s2 / n, Q_g = -r s1 / n + r^2 mu s2 / n, s1 = sum dyh, s2 = sum dyh xhat over the group.
// gn_bwd_stats: one CTA per (group, image), two passes over the group's slab (statistics recomputed exactly as in the
// forward kernel, then s1 / s2); writes a, s (for y0 = a x + s), P, Q per (image, channel).  gn_bwd_apply: one vector pass.
__device__ __forceinline__ float silu_grad_f(float y0) {
  const float sg = 1.f / (1.f + __expf(-y0));
  return sg * (1.f + y0 * (1.f - sg));
}

template <int TP, int ACT>
__global__ void __launch_bounds__(GN_STATS_THREADS) gn_bwd_stats_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                         float* __restrict__ coef, int B, int C, int HW, int groups, float eps) {
  constexpr int TR = GN_STATS_THREADS / TP;
  __shared__ float red[2][32];
  __shared__ float stat[2];
  const int g = blockIdx.x, b = blockIdx.y, cpg = C / groups, ppg = cpg >> 1, pairs = C >> 1;
  const int p = threadIdx.x % TP, r0 = threadIdx.x / TP;
  const long long off = (long long)b * HW * C + g * cpg;
  const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(x + off);
  const __nv_bfloat162* db = reinterpret_cast<const __nv_bfloat162*>(dy + off);
  const float n = (float)cpg * (float)HW;
  auto block_sum2 = [&](float& u, float& v) {
    u = warp_sum(u);
    v = warp_sum(v);
    __syncthreads();                      // red[] may still be read from the previous reduction
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = u;
      red[1][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    float tu = 0.f, tv = 0.f;
#pragma unroll
    for (int i = 0; i < GN_STATS_THREADS / 32; ++i) {
      tu += red[0][i];
      tv += red[1][i];
    }
    u = tu;
    v = tv;
  };
  float sum = 0.f, sq = 0.f;
  if (p < ppg) {
#pragma unroll 4
    for (int r = r0; r < HW; r += TR) {
      const float2 v = __bfloat1622float2(xb[(long long)r * pairs + p]);
      sum += v.x + v.y;
      sq += v.x * v.x + v.y * v.y;
    }
  }
  block_sum2(sum, sq);
  const float mean = sum / n;
  const float rstd = rsqrtf(fmaxf(sq / n - mean * mean, 0.f) + eps);
  float s1 = 0.f, s2 = 0.f;
  if (p < ppg) {
    const int c0 = g * cpg + 2 * p;
    const float ga0 = gamma[c0], ga1 = gamma[c0 + 1];
    const float a0 = rstd * ga0, a1 = rstd * ga1;
    const float sh0 = beta[c0] - mean * a0, sh1 = beta[c0 + 1] - mean * a1;
#pragma unroll 4
    for (int r = r0; r < HW; r += TR) {
      const float2 v = __bfloat1622float2(xb[(long long)r * pairs + p]);
      float2 d = __bfloat1622float2(db[(long long)r * pairs + p]);
      if (ACT == 1) {
        d.x *= silu_grad_f(v.x * a0 + sh0);
        d.y *= silu_grad_f(v.y * a1 + sh1);
      }
      const float h0 = d.x * ga0, h1 = d.y * ga1;
      s1 += h0 + h1;
      s2 += h0 * (v.x - mean) * rstd + h1 * (v.y - mean) * rstd;
    }
  }
  block_sum2(s1, s2);
  const float P = -rstd * rstd * s2 / n;
  const float Q = -rstd * s1 / n - P * mean;
  const long long plane = (long long)B * C;
  for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
    const int ch = g * cpg + c;
    const float a = gamma[ch] * rstd;
    const long long i = (long long)b * C + ch;
    coef[i] = a;
    coef[plane + i] = beta[ch] - mean * a;
    coef[2 * plane + i] = P;
    coef[3 * plane + i] = Q;
  }
}

template <int ACT>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                            const float* __restrict__ coef, bf16* __restrict__ dx, int B, int C, int HW,
                                                            int rows_per_cta) {
  extern __shared__ float cs[];      // [4][C]: a, s, P, Q of this image
  const int b = blockIdx.y;
  const long long plane = (long long)B * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 4; ++k) cs[k * C + c] = coef[k * plane + (long long)b * C + c];
  }
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
  const int vec = C >> 3;
  const long long base = ((long long)b * HW + r0) * vec;
  const uint4* xv = reinterpret_cast<const uint4*>(x) + base;
  const uint4* dv = reinterpret_cast<const uint4*>(dy) + base;
  uint4* ov = reinterpret_cast<uint4*>(dx) + base;
  const int n = (r1 - r0) * vec;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c0 = (i % vec) << 3;
    const uint4 u = xv[i], w = dv[i];
    const uint32_t xw[4] = {u.x, u.y, u.z, u.w}, dw[4] = {w.x, w.y, w.z, w.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float xe[2] = {__uint_as_float(xw[j] << 16), __uint_as_float(xw[j] & 0xffff0000u)};
      float de[2] = {__uint_as_float(dw[j] << 16), __uint_as_float(dw[j] & 0xffff0000u)};
      float r[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c0 + 2 * j + e;
        float gq = de[e];
        if (ACT == 1) gq *= silu_grad_f(xe[e] * cs[c] + cs[C + c]);
        r[e] = cs[c] * gq + cs[2 * C + c] * xe[e] + cs[3 * C + c];
      }
      o[j] = pack_bf16(r[0], r[1]);
    }
    ov[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

int groupnorm_act_tokens_bwd(const void* x, const void* dy, const float* gamma, const float* beta, int64_t B, int64_t HW, int64_t C,
                             int64_t groups, float eps, int act, float* coef_ws, void* dx, cudaStream_t stream) {
  AF_CHECK(x && dy && gamma && beta && coef_ws && dx, "groupnorm_act_tokens_bwd: null pointer");
  AF_CHECK(B > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 65535 && C % groups == 0 && (C / groups) % 2 == 0 && C % 8 == 0 &&
               C / groups <= 256 && C <= 3072 && B <= 65535,
           "groupnorm_act_tokens_bwd: bad shape B=%lld HW=%lld C=%lld groups=%lld", (long long)B, (long long)HW, (long long)C, (long long)groups);
  AF_CHECK(act == 0 || act == 1, "groupnorm_act_tokens_bwd: act %d (0 none | 1 SiLU)", act);
  AF_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
           "groupnorm_act_tokens_bwd: x / dy / dx must be 16-byte aligned");
  const dim3 gs((unsigned)groups, (unsigned)B);
  const int ppg = (int)(C / groups / 2);
#define AF_GNB(TP)                                                                                                                   \
  do {                                                                                                                               \
    if (act == 1)                                                                                                                    \
      gn_bwd_stats_kernel<TP, 1><<<gs, GN_STATS_THREADS, 0, stream>>>((const bf16*)x, (const bf16*)dy, gamma, beta, coef_ws, (int)B, \
                                                                      (int)C, (int)HW, (int)groups, eps);                            \
    else                                                                                                                             \
      gn_bwd_stats_kernel<TP, 0><<<gs, GN_STATS_THREADS, 0, stream>>>((const bf16*)x, (const bf16*)dy, gamma, beta, coef_ws, (int)B, \
                                                                      (int)C, (int)HW, (int)groups, eps);                            \
  } while (0)
  if (ppg <= 8) AF_GNB(8);
  else if (ppg <= 16) AF_GNB(16);
  else if (ppg <= 32) AF_GNB(32);
  else if (ppg <= 64) AF_GNB(64);
  else AF_GNB(128);
#undef AF_GNB
  long long rows = (B * HW + 295) / 296;
  if (rows < 4) rows = 4;
  if (rows > 64) rows = 64;
  const dim3 grid((unsigned)((HW + rows - 1) / rows), (unsigned)B);
  const size_t smem = (size_t)(4 * C * sizeof(float));       // <= 48 KB for C <= 3072
  if (act == 1) gn_bwd_apply_kernel<1><<<grid, 256, smem, stream>>>((const bf16*)x, (const bf16*)dy, coef_ws, (bf16*)dx, (int)B, (int)C, (int)HW, (int)rows);
  else gn_bwd_apply_kernel<0><<<grid, 256, smem, stream>>>((const bf16*)x, (const bf16*)dy, coef_ws, (bf16*)dx, (int)B, (int)C, (int)HW, (int)rows);
  AF_CUDA(cudaGetLastError());
  g_launch_count += 2;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Re-sampling helpers of the backward pass (NHWC, one thread per 16 bytes of OUTPUT):
//   MODE 0  sum-pool 2x2: y[b, h, w] = sum of x[b, 2h + i, 2w + j]  -- backward of the nearest 2x of Upsample (:116)
//   MODE 1  zero-insert:  y[b, 2h, 2w] = x[b, h, w], other positions 0 -- turns the backward of the stride-2 convolution of
//           Downsample (:151) into a stride-1 convolution with flipped weights over the full-resolution grid
template <int MODE>
__global__ void __launch_bounds__(256) resample2x_bwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int vec,
                                                              long long n_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int v = (int)(i % vec);
  long long p = i / vec;
  if (MODE == 0) {                   // output [B, H, W], input [B, 2H, 2W]
    const int ox = (int)(p % W);
    p /= W;
    const int oy = (int)(p % H);
    const long long b = p / H;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dyi = 0; dyi < 2; ++dyi)
#pragma unroll
      for (int dxi = 0; dxi < 2; ++dxi) {
        const uint4 u = x[((b * 2 * H + 2 * oy + dyi) * (2 * W) + 2 * ox + dxi) * vec + v];
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(w4[j] << 16);
          acc[2 * j + 1] += __uint_as_float(w4[j] & 0xffff0000u);
        }
      }
    y[i] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  } else {                           // output [B, 2H, 2W], input [B, H, W]
    const int ox = (int)(p % (2 * W));
    p /= 2 * W;
    const int oy = (int)(p % (2 * H));
    const long long b = p / (2 * H);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (((ox | oy) & 1) == 0) o = x[((b * H + (oy >> 1)) * W + (ox >> 1)) * vec + v];
    y[i] = o;
  }
}

int resample2x_bwd(const void* x, void* y, int64_t B, int64_t H, int64_t W, int64_t C, int mode, cudaStream_t stream) {
  AF_CHECK(x && y, "resample2x_bwd: null pointer");
  AF_CHECK(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (mode == 0 || mode == 1), "resample2x_bwd: bad arguments");
  AF_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "resample2x_bwd: x / y must be 16-byte aligned");
  const int vec = (int)(C / 8);
  const long long n_out = (mode == 0 ? B * H * W : B * 4 * H * W) * (long long)vec;
  const unsigned grid = (unsigned)((n_out + 255) / 256);
  if (mode == 0) resample2x_bwd_kernel<0><<<grid, 256, 0, stream>>>((const uint4*)x, (uint4*)y, (int)H, (int)W, vec, n_out);
  else resample2x_bwd_kernel<1><<<grid, 256, 0, stream>>>((const uint4*)x, (uint4*)y, (int)H, (int)W, vec, n_out);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

}  // namespace adaface
