// K4 and the small HBM-bound helpers around the attention kernels:
//   layernorm_fwd        LayerNorm of the transformer blocks (ldm/modules/attention.py:232-234, 244-250) and of the
//                        CLIP-shaped encoder (layer_norm1/2), fp32 statistics, bf16 output for the following GEMM
//   sbg_head_fwd         sum-normalised mix of the last hidden states + final LayerNorm (arc2face_models.py:291-306)
//   qmean                column mean of Q over the queries (normalize_cross_attn, dalc:123-126)
//   capture_chan_major   'b h n d -> b (h d) n' * sqrt(scale) re-layout of cached q/q2/k/v/attn_out (dalc:349-362)
#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<bf16>(const bf16* p) { return __bfloat162float(*p); }

// One warp per row; the row lives in registers between the statistics and the normalisation pass.
template <typename TIn, typename TOut, int MAXC>
__global__ void __launch_bounds__(256) layernorm_kernel(const TIn* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                         const float* __restrict__ b, TOut* __restrict__ y, long long ldy,
                                                         int M, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const TIn* xr = x + (long long)row * ldx;
  constexpr int PER = MAXC / 32;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    v[i] = c < C ? ld_as_float(xr + c) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    const float d = c < C ? v[i] - mean : 0.f;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) / C + eps);
  TOut* yr = y + (long long)row * ldy;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    if (c < C) {
      const float o = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
      if constexpr (sizeof(TOut) == 2) yr[c] = __float2bfloat16(o);
      else yr[c] = o;
    }
  }
}

template <typename TIn, typename TOut>
static int launch_ln(const void* x, long long ldx, const float* w, const float* b, void* y, long long ldy, int M, int C,
                     float eps, cudaStream_t stream) {
  const int rows_per_block = 8;
  dim3 grid((M + rows_per_block - 1) / rows_per_block);
  if (C <= 320)
    layernorm_kernel<TIn, TOut, 320><<<grid, 256, 0, stream>>>((const TIn*)x, ldx, w, b, (TOut*)y, ldy, M, C, eps);
  else if (C <= 768)
    layernorm_kernel<TIn, TOut, 768><<<grid, 256, 0, stream>>>((const TIn*)x, ldx, w, b, (TOut*)y, ldy, M, C, eps);
  else
    layernorm_kernel<TIn, TOut, 1280><<<grid, 256, 0, stream>>>((const TIn*)x, ldx, w, b, (TOut*)y, ldy, M, C, eps);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const float* w, const float* b, void* y, int y_dtype,
                  int64_t ldy, int64_t M, int64_t C, float eps, cudaStream_t stream) {
  AF_CHECK(x && w && b && y, "layernorm_fwd: null pointer");
  AF_CHECK(M > 0 && C > 0 && C <= 1280, "layernorm_fwd: unsupported shape M=%lld C=%lld (C <= 1280)", (long long)M,
           (long long)C);
  if (x_dtype == ADAFACE_BF16 && y_dtype == ADAFACE_BF16) return launch_ln<bf16, bf16>(x, ldx, w, b, y, ldy, (int)M, (int)C, eps, stream);
  if (x_dtype == ADAFACE_F32 && y_dtype == ADAFACE_BF16) return launch_ln<float, bf16>(x, ldx, w, b, y, ldy, (int)M, (int)C, eps, stream);
  if (x_dtype == ADAFACE_F32 && y_dtype == ADAFACE_F32) return launch_ln<float, float>(x, ldx, w, b, y, ldy, (int)M, (int)C, eps, stream);
  set_error("layernorm_fwd: unsupported dtype combination x=%d y=%d", x_dtype, y_dtype);
  return 1;
}

// ---------------------------------------------------------------------------------------------
struct HeadPtrs {
  const float* h[4];
  float wl[4];
  const float* wl_dev;     // when set, the layer weights are read from DEVICE memory (CUDA-graph-capturable training step)
  int n;
};

template <int MAXC>
__global__ void __launch_bounds__(256) sbg_head_kernel(const HeadPtrs hp, long long ldh, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ out,
                                                        long long ldo, int M, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int PER = MAXC / 32;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    float a = 0.f;
    if (c < C)
      for (int l = 0; l < hp.n; ++l) a += (hp.wl_dev ? __ldg(hp.wl_dev + l) : hp.wl[l]) * hp.h[l][(long long)row * ldh + c];
    v[i] = a;
    s += a;
  }
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    const float d = c < C ? v[i] - mean : 0.f;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) / C + eps);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    if (c < C) out[(long long)row * ldo + c] = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
  }
}

int sbg_head_fwd(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl, int n_layers,
                 int64_t ldh, const float* w, const float* b, float* out, int64_t ldo, int64_t M, int64_t C, float eps,
                 cudaStream_t stream, const float* wl_dev) {
  AF_CHECK(n_layers >= 1 && n_layers <= 4 && (wl || wl_dev), "sbg_head_fwd: n_layers must be 1..4 (got %d)", n_layers);
  AF_CHECK(M > 0 && C > 0 && C <= 768, "sbg_head_fwd: unsupported shape M=%lld C=%lld", (long long)M, (long long)C);
  HeadPtrs hp;
  const float* hs[4] = {h0, h1, h2, h3};
  for (int i = 0; i < 4; ++i) {
    hp.h[i] = hs[i];
    hp.wl[i] = (i < n_layers && wl) ? wl[i] : 0.f;   // wl is a HOST array of n_layers floats
    AF_CHECK(i >= n_layers || hs[i], "sbg_head_fwd: hidden state %d is null", i);
  }
  hp.n = n_layers;
  hp.wl_dev = wl_dev;
  dim3 grid(((int)M + 7) / 8);
  sbg_head_kernel<768><<<grid, 256, 0, stream>>>(hp, ldh, w, b, out, ldo, (int)M, (int)C, eps);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// qmean[b, c] = mean_n q[b, n, c].  grid (C/64, B, splits); 256 threads = 32 channel pairs x 8 row lanes.
template <typename T>
__global__ void __launch_bounds__(256) qmean_kernel(const T* __restrict__ q, long long sb, long long sn, int L, int C,
                                                     float* __restrict__ out, float inv_L) {
  __shared__ float red[8][64];
  const int cp = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + cp * 2;
  const int b = blockIdx.y;
  const int rows_per = (L + gridDim.z - 1) / gridDim.z;
  const int r_begin = blockIdx.z * rows_per, r_end = min(L, r_begin + rows_per);
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    const T* base = q + (long long)b * sb + c;
    for (int r = r_begin + ry; r < r_end; r += 8) {
      a0 += ld_as_float(base + (long long)r * sn);
      a1 += ld_as_float(base + (long long)r * sn + 1);
    }
  }
  red[ry][cp * 2] = a0;
  red[ry][cp * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < C) atomicAdd(out + (long long)b * C + cc, s * inv_L);
  }
}

int qmean(const void* q, int q_dtype, int64_t q_sb, int64_t q_sn, int64_t B, int64_t Lq, int64_t C, float* out,
          cudaStream_t stream) {
  AF_CHECK(q && out && B > 0 && Lq > 0 && C > 0 && C % 2 == 0, "qmean: bad arguments");
  AF_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * B * C, stream));
  const int splits = (int)((Lq + 511) / 512);
  dim3 grid((unsigned)((C + 63) / 64), (unsigned)B, (unsigned)splits);
  if (q_dtype == ADAFACE_F32)
    qmean_kernel<float><<<grid, 256, 0, stream>>>((const float*)q, q_sb, q_sn, (int)Lq, (int)C, out, 1.f / (float)Lq);
  else
    qmean_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)q, q_sb, q_sn, (int)Lq, (int)C, out, 1.f / (float)Lq);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// dst[b, c, n] = factor * src[b, n, c]   (32 x 32 smem transpose)
template <typename T>
__global__ void __launch_bounds__(256) chan_major_kernel(const T* __restrict__ src, long long sb, long long sn, int L, int C,
                                                          float factor, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, c = c0 + tx;
    tile[i][tx] = (n < L && c < C) ? ld_as_float(src + (long long)b * sb + (long long)n * sn + c) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, n = n0 + tx;
    if (c < C && n < L) dst[((long long)b * C + c) * L + n] = tile[tx][i] * factor;
  }
}

int capture_chan_major(const void* src, int src_dtype, int64_t s_sb, int64_t s_sn, int64_t B, int64_t L, int64_t C,
                       float factor, float* dst, cudaStream_t stream) {
  AF_CHECK(src && dst && B > 0 && L > 0 && C > 0, "capture_chan_major: bad arguments");
  dim3 grid((unsigned)((L + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
  if (src_dtype == ADAFACE_BF16)
    chan_major_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)src, s_sb, s_sn, (int)L, (int)C, factor, dst);
  else
    chan_major_kernel<float><<<grid, 256, 0, stream>>>((const float*)src, s_sb, s_sn, (int)L, (int)C, factor, dst);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}


// ---------------------------------------------------------------------------------------------
// SpatialTransformer entry / exit (ldm/modules/attention.py:287-304):
//   groupnorm_stats      per (batch, group) mean / rstd of x [B, C, HW], folded with the affine parameters into
//                        a[b, c] = rstd * gamma[c],  s[b, c] = beta[c] - mean * rstd * gamma[c]
//   groupnorm_tokens     y[b, hw, c] = x[b, c, hw] * a[b, c] + s[b, c]   ('b c h w -> b (h w) c' fused with the norm; bf16)
//   tokens_to_nchw_add   out[b, c, hw] = t[b, hw, c] + x_in[b, c, hw]    ('b (h w) c -> b c h w' fused with the residual)
template <typename T>
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ a,
                                                               float* __restrict__ s, int C, int HW, int groups, float eps) {
  const int g = blockIdx.x, b = blockIdx.y, cpg = C / groups;
  const T* xg = x + ((long long)b * C + (long long)g * cpg) * HW;
  const long long n = (long long)cpg * HW;
  float sum = 0.f, sq = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = ld_as_float(xg + i);
    sum += v;
    sq += v * v;
  }
  __shared__ float red[2][8];
  sum = warp_sum(sum);
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sum;
    red[1][threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  float ts = 0.f, tq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ts += red[0][i];
    tq += red[1][i];
  }
  const float mean = ts / n;
  const float rstd = rsqrtf(fmaxf(tq / n - mean * mean, 0.f) + eps);
  for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
    const int ch = g * cpg + c;
    const float ga = gamma[ch] * rstd;
    a[(long long)b * C + ch] = ga;
    s[(long long)b * C + ch] = beta[ch] - mean * ga;
  }
}

// MODE 0: y[b, hw, c] = x[b, c, hw] * a[b, c] + s[b, c] (TIn -> bf16);  MODE 1: out[b, c, hw] = t[b, hw, c] + res[b, c, hw]
template <typename TIn, typename TOut, int MODE>
__global__ void __launch_bounds__(256) nchw_tokens_kernel(const TIn* __restrict__ src, TOut* __restrict__ dst, const float* __restrict__ a,
                                                           const float* __restrict__ s, const TOut* __restrict__ res, int I, int J) {
  // src [B, I, J] -> dst [B, J, I]; MODE 0: I = C, J = HW; MODE 1: I = HW, J = C
  __shared__ float tile[64][65];
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64, b = blockIdx.z;
  const TIn* sb = src + (long long)b * I * J;
  TOut* db = dst + (long long)b * I * J;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll 4
  for (int r = ty; r < 64; r += 4) {
    const int i = i0 + r, j = j0 + tx;
    float v = 0.f;
    if (i < I && j < J) {
      v = ld_as_float(sb + (long long)i * J + j);
      if (MODE == 0) v = v * a[(long long)b * I + i] + s[(long long)b * I + i];
    }
    tile[r][tx] = v;
  }
  __syncthreads();
#pragma unroll 4
  for (int r = ty; r < 64; r += 4) {
    const int j = j0 + r, i = i0 + tx;
    if (j < J && i < I) {
      float v = tile[tx][r];
      const long long o = (long long)j * I + i;
      if (MODE == 1) v += ld_as_float(res + (long long)b * I * J + o);
      if constexpr (sizeof(TOut) == 2) db[o] = __float2bfloat16(v);
      else db[o] = v;
    }
  }
}

int groupnorm_tokens_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, int64_t B, int64_t C, int64_t HW,
                         int64_t groups, float eps, float* a_ws, float* s_ws, void* y, cudaStream_t stream) {
  AF_CHECK(x && gamma && beta && a_ws && s_ws && y, "groupnorm_tokens_fwd: null pointer");
  AF_CHECK(B > 0 && C > 0 && HW > 0 && groups > 0 && C % groups == 0 && B <= 65535, "groupnorm_tokens_fwd: bad shape");
  const dim3 gs((unsigned)groups, (unsigned)B), gt((unsigned)((HW + 63) / 64), (unsigned)((C + 63) / 64), (unsigned)B);
  if (x_dtype == ADAFACE_F32) {
    groupnorm_stats_kernel<float><<<gs, 256, 0, stream>>>((const float*)x, gamma, beta, a_ws, s_ws, (int)C, (int)HW, (int)groups, eps);
    nchw_tokens_kernel<float, bf16, 0><<<gt, 256, 0, stream>>>((const float*)x, (bf16*)y, a_ws, s_ws, nullptr, (int)C, (int)HW);
  } else if (x_dtype == ADAFACE_BF16) {
    groupnorm_stats_kernel<bf16><<<gs, 256, 0, stream>>>((const bf16*)x, gamma, beta, a_ws, s_ws, (int)C, (int)HW, (int)groups, eps);
    nchw_tokens_kernel<bf16, bf16, 0><<<gt, 256, 0, stream>>>((const bf16*)x, (bf16*)y, a_ws, s_ws, nullptr, (int)C, (int)HW);
  } else {
    set_error("groupnorm_tokens_fwd: bad dtype %d", x_dtype);
    return 1;
  }
  AF_CUDA(cudaGetLastError());
  g_launch_count += 2;
  return 0;
}

int tokens_to_nchw_add(const void* t, const void* x_in, int x_dtype, void* out, int64_t B, int64_t C, int64_t HW,
                       cudaStream_t stream) {
  AF_CHECK(t && x_in && out, "tokens_to_nchw_add: null pointer");
  AF_CHECK(B > 0 && C > 0 && HW > 0 && B <= 65535, "tokens_to_nchw_add: bad shape");
  const dim3 grid((unsigned)((C + 63) / 64), (unsigned)((HW + 63) / 64), (unsigned)B);
  if (x_dtype == ADAFACE_F32)
    nchw_tokens_kernel<bf16, float, 1><<<grid, 256, 0, stream>>>((const bf16*)t, (float*)out, nullptr, nullptr, (const float*)x_in, (int)HW, (int)C);
  else if (x_dtype == ADAFACE_BF16)
    nchw_tokens_kernel<bf16, bf16, 1><<<grid, 256, 0, stream>>>((const bf16*)t, (bf16*)out, nullptr, nullptr, (const bf16*)x_in, (int)HW, (int)C);
  else {
    set_error("tokens_to_nchw_add: bad dtype %d", x_dtype);
    return 1;
  }
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

}  // namespace adaface
