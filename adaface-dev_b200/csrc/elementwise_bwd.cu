// K5c: the HBM-bound helpers of the stage-2 backward pass (north-star item 5) -- one read + one write each:
//   transpose          out[b, j, i] = alpha * colscale[j] * rowscale[i] * in[b, i, j]: operand re-layout for the weight-gradient
//                      GEMMs (dW = dY^T X runs on the K-major tcgen05 GEMM over transposed operands), the DoRA column
//                      scale folded in, and the backward of the 'b n c -> b c n' capture re-layout (dalc:349-362)
//   colsum             out[j] += colmul[j] * sum_i a[i, j] * (b[i, j] - bias[j]): bias gradients and the DoRA magnitude
//                      gradient  dm_j = (1 / m_j) sum_i dY_ij (Y_ij - bias_j)   (norm term detached)
//   layernorm_bwd      dx (+ dw, db) of LayerNorm, statistics recomputed from x
//   act_fwd / act_bwd  quick-GELU (CLIPMLP) and packed GEGLU (ldm FeedForward) kept out of the GEMM epilogue in
//                      training mode so that the pre-activation survives for the backward pass
//   sbg_head_bwd       backward of the weighted last-layers mix + final LayerNorm (arc2face_models.py:291-306)
#include <math.h>

#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

// ---------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) transpose_kernel(const TIn* __restrict__ src, long long s_sb, long long s_ld,
                                                         TOut* __restrict__ dst, long long d_sb, long long d_ld, int I, int J,
                                                         float alpha, const float* __restrict__ colscale,
                                                         const float* __restrict__ rowscale) {
  __shared__ float tile[64][65];
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const TIn* s = src + (long long)blockIdx.z * s_sb;
  TOut* d = dst + (long long)blockIdx.z * d_sb;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
#pragma unroll 4
  for (int r = ty; r < 64; r += 4) {
    const int i = i0 + r, j = j0 + tx;
    tile[r][tx] = (i < I && j < J) ? ldf(s + (long long)i * s_ld + j) : 0.f;
  }
  __syncthreads();
#pragma unroll 4
  for (int r = ty; r < 64; r += 4) {
    const int j = j0 + r, i = i0 + tx;
    if (j < J && i < I) {
      float cs = colscale ? __ldg(colscale + j) * alpha : alpha;
      if (rowscale) cs *= __ldg(rowscale + i);
      stf(d + (long long)j * d_ld + i, tile[tx][r] * cs);
    }
  }
}

int transpose(const void* src, int src_dtype, int64_t s_sb, int64_t s_ld, void* dst, int dst_dtype, int64_t d_sb, int64_t d_ld,
              int64_t B, int64_t I, int64_t J, float alpha, const float* colscale, const float* rowscale, cudaStream_t stream) {
  AF_CHECK(src && dst, "transpose: null pointer");
  AF_CHECK(B > 0 && I > 0 && J > 0 && B <= 65535 && (I + 63) / 64 <= 65535, "transpose: bad shape B=%lld I=%lld J=%lld",
           (long long)B, (long long)I, (long long)J);
  const dim3 grid((unsigned)((J + 63) / 64), (unsigned)((I + 63) / 64), (unsigned)B);
#define AF_TR(TI, TO)                                                                                                  \
  transpose_kernel<TI, TO><<<grid, 256, 0, stream>>>((const TI*)src, s_sb, s_ld, (TO*)dst, d_sb, d_ld, (int)I, (int)J, \
                                                      alpha, colscale, rowscale)
  if (src_dtype == ADAFACE_BF16 && dst_dtype == ADAFACE_BF16) AF_TR(bf16, bf16);
  else if (src_dtype == ADAFACE_F32 && dst_dtype == ADAFACE_BF16) AF_TR(float, bf16);
  else if (src_dtype == ADAFACE_F32 && dst_dtype == ADAFACE_F32) AF_TR(float, float);
  else if (src_dtype == ADAFACE_BF16 && dst_dtype == ADAFACE_F32) AF_TR(bf16, float);
  else {
    set_error("transpose: bad dtypes %d -> %d", src_dtype, dst_dtype);
    return 1;
  }
#undef AF_TR
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// grid (ceil(N / 64), row_chunks); a warp covers 64 columns (2 per lane), the 8 warps stride over the chunk's rows.
template <typename TA, typename TB>
__global__ void __launch_bounds__(256) colsum_kernel(const TA* __restrict__ a, long long lda, const TB* __restrict__ b,
                                                      long long ldb, const float* __restrict__ bias,
                                                      const float* __restrict__ colmul, float* __restrict__ out, int M, int N,
                                                      int rows_per_cta) {
  __shared__ float red[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 64 + lane * 2;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float s0 = 0.f, s1 = 0.f;
  const bool ok0 = j < N, ok1 = j + 1 < N;
  const float bi0 = (bias && ok0) ? bias[j] : 0.f, bi1 = (bias && ok1) ? bias[j + 1] : 0.f;
  for (int r = r0 + warp; r < r1; r += 8) {
    float a0 = ok0 ? ldf(a + (long long)r * lda + j) : 0.f, a1 = ok1 ? ldf(a + (long long)r * lda + j + 1) : 0.f;
    if (b) {
      a0 *= ok0 ? ldf(b + (long long)r * ldb + j) - bi0 : 0.f;
      a1 *= ok1 ? ldf(b + (long long)r * ldb + j + 1) - bi1 : 0.f;
    }
    s0 += a0;
    s1 += a1;
  }
  red[warp][lane * 2] = s0;
  red[warp][lane * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int jj = blockIdx.x * 64 + threadIdx.x;
    if (jj < N) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
      if (colmul) s *= colmul[jj];
      atomicAdd(out + jj, s);
    }
  }
}

int colsum(const void* a, int a_dtype, int64_t lda, const void* b, int b_dtype, int64_t ldb, const float* bias,
           const float* colmul, float* out, int64_t M, int64_t N, cudaStream_t stream) {
  AF_CHECK(a && out && M > 0 && N > 0, "colsum: bad arguments");
  const int gx = (int)((N + 63) / 64);
  int gy = (int)((M + 255) / 256);
  const int want = (148 * 4 + gx - 1) / gx;
  if (gy > want) gy = want;
  if (gy < 1) gy = 1;
  const int rows = (int)((M + gy - 1) / gy);
  gy = (int)((M + rows - 1) / rows);
  const dim3 grid(gx, gy);
#define AF_CS(TA, TB) \
  colsum_kernel<TA, TB><<<grid, 256, 0, stream>>>((const TA*)a, lda, (const TB*)b, ldb, bias, colmul, out, (int)M, (int)N, rows)
  const int bd = b ? b_dtype : ADAFACE_BF16;
  if (a_dtype == ADAFACE_BF16 && bd == ADAFACE_BF16) AF_CS(bf16, bf16);
  else if (a_dtype == ADAFACE_BF16 && bd == ADAFACE_F32) AF_CS(bf16, float);
  else if (a_dtype == ADAFACE_F32 && bd == ADAFACE_F32) AF_CS(float, float);
  else if (a_dtype == ADAFACE_F32 && bd == ADAFACE_BF16) AF_CS(float, bf16);
  else {
    set_error("colsum: bad dtypes");
    return 1;
  }
#undef AF_CS
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// One warp per row (grid-stride): xhat = (x - mean) rstd, g = dy w,
//   dx = rstd (g - mean(g) - xhat mean(g xhat));  dw += dy xhat;  db += dy.
template <typename TX, typename TDY, int MAXC, bool WG>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const TX* __restrict__ x, long long ldx, const TDY* __restrict__ dy,
                                                             long long lddy, const float* __restrict__ w, TX* __restrict__ dx,
                                                             long long lddx, float* __restrict__ dw, float* __restrict__ db,
                                                             int M, int C, float eps) {
  constexpr int PER = MAXC / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dwa[WG ? PER : 1], dba[WG ? PER : 1];
  if constexpr (WG) {
#pragma unroll
    for (int i = 0; i < PER; ++i) dwa[i] = dba[i] = 0.f;
  }
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const TX* xr = x + (long long)row * ldx;
    const TDY* gr = dy + (long long)row * lddy;
    float v[PER], gy[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      v[i] = c < C ? ldf(xr + c) : 0.f;
      gy[i] = c < C ? ldf(gr + c) : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      v[i] = c < C ? v[i] - mean : 0.f;
      ss += v[i] * v[i];
    }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      v[i] *= rstd;                                   // xhat
      if constexpr (WG) {
        dwa[i] += gy[i] * v[i];
        dba[i] += gy[i];
      }
      gy[i] *= c < C ? __ldg(w + c) : 0.f;            // g
      sg += gy[i];
      sgx += gy[i] * v[i];
    }
    const float mg = warp_sum(sg) / C, mgx = warp_sum(sgx) / C;
    TX* dr = dx + (long long)row * lddx;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      if (c < C) stf(dr + c, rstd * (gy[i] - mg - v[i] * mgx));
    }
  }
  if constexpr (WG) {
    __shared__ float red[8][MAXC + 1];
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < PER; ++i) red[warp][i * 32 + lane] = pass ? dba[i] : dwa[i];
      __syncthreads();
      float* out = pass ? db : dw;
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) t += red[ww][c];
        atomicAdd(out + c, t);
      }
    }
  }
}

template <typename TX, typename TDY>
static int launch_ln_bwd(const void* x, long long ldx, const void* dy, long long lddy, const float* w, void* dx, long long lddx,
                         float* dw, float* db, int M, int C, float eps, cudaStream_t stream) {
  int grid = (M + 7) / 8;
  if (grid > 148 * 2) grid = 148 * 2;
#define AF_LNB(MAXC, WG)                                                                                             \
  layernorm_bwd_kernel<TX, TDY, MAXC, WG><<<grid, 256, 0, stream>>>((const TX*)x, ldx, (const TDY*)dy, lddy, w, (TX*)dx, \
                                                                     lddx, dw, db, M, C, eps)
  const bool wg = dw != nullptr;
  if (C <= 320) { if (wg) AF_LNB(320, true); else AF_LNB(320, false); }
  else if (C <= 768) { if (wg) AF_LNB(768, true); else AF_LNB(768, false); }
  else { if (wg) AF_LNB(1280, true); else AF_LNB(1280, false); }
#undef AF_LNB
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int layernorm_bwd(const void* x, int x_dtype, int64_t ldx, const void* dy, int dy_dtype, int64_t lddy, const float* w, void* dx,
                  int64_t lddx, float* dw, float* db, int64_t M, int64_t C, float eps, cudaStream_t stream) {
  AF_CHECK(x && dy && w && dx, "layernorm_bwd: null pointer");
  AF_CHECK((dw == nullptr) == (db == nullptr), "layernorm_bwd: dw and db go together");
  AF_CHECK(M > 0 && C > 0 && C <= 1280, "layernorm_bwd: unsupported shape M=%lld C=%lld (C <= 1280)", (long long)M, (long long)C);
  if (x_dtype == ADAFACE_BF16 && dy_dtype == ADAFACE_BF16) return launch_ln_bwd<bf16, bf16>(x, ldx, dy, lddy, w, dx, lddx, dw, db, (int)M, (int)C, eps, stream);
  if (x_dtype == ADAFACE_F32 && dy_dtype == ADAFACE_BF16) return launch_ln_bwd<float, bf16>(x, ldx, dy, lddy, w, dx, lddx, dw, db, (int)M, (int)C, eps, stream);
  if (x_dtype == ADAFACE_F32 && dy_dtype == ADAFACE_F32) return launch_ln_bwd<float, float>(x, ldx, dy, lddy, w, dx, lddx, dw, db, (int)M, (int)C, eps, stream);
  if (x_dtype == ADAFACE_BF16 && dy_dtype == ADAFACE_F32) return launch_ln_bwd<bf16, float>(x, ldx, dy, lddy, w, dx, lddx, dw, db, (int)M, (int)C, eps, stream);
  set_error("layernorm_bwd: bad dtypes");
  return 1;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// BWD = false: h = act(u).  BWD = true: du = dh * act'(u).  One thread per output element pair.
// GEGLU: u is [M, 2 n_out] in packed tiles [a(64) | gate(64)] (the layout the packed fc weight produces), h is [M, n_out].
template <int ACT, bool BWD>
__global__ void __launch_bounds__(256) act_kernel(const bf16* __restrict__ u, long long ldu, const bf16* __restrict__ dh,
                                                   long long lddh, bf16* __restrict__ out, long long ldo, int M, int n_out) {
  const long long total = (long long)M * (n_out / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / (n_out / 2)), c = (int)(i % (n_out / 2)) * 2;
    if (ACT == ADAFACE_ACT_QUICK_GELU) {
      const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(u + (long long)r * ldu + c));
      const float s0 = 1.f / (1.f + __expf(-1.702f * x.x)), s1 = 1.f / (1.f + __expf(-1.702f * x.y));
      float2 o;
      if (!BWD) {
        o = make_float2(x.x * s0, x.y * s1);
      } else {
        const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dh + (long long)r * lddh + c));
        o = make_float2(g.x * s0 * (1.f + 1.702f * x.x * (1.f - s0)), g.y * s1 * (1.f + 1.702f * x.y * (1.f - s1)));
      }
      *reinterpret_cast<__nv_bfloat162*>(out + (long long)r * ldo + c) = __floats2bfloat162_rn(o.x, o.y);
    } else {
      const int tile = c >> 6, cc = c & 63;
      const bf16* ur = u + (long long)r * ldu + tile * 128 + cc;
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ur));
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ur + 64));
      if (!BWD) {
        *reinterpret_cast<__nv_bfloat162*>(out + (long long)r * ldo + c) = __floats2bfloat162_rn(a.x * gelu_erf_f(g.x), a.y * gelu_erf_f(g.y));
      } else {
        const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dh + (long long)r * lddh + c));
        bf16* orow = out + (long long)r * ldo + tile * 128 + cc;
        *reinterpret_cast<__nv_bfloat162*>(orow) = __floats2bfloat162_rn(d.x * gelu_erf_f(g.x), d.y * gelu_erf_f(g.y));
        *reinterpret_cast<__nv_bfloat162*>(orow + 64) = __floats2bfloat162_rn(d.x * a.x * gelu_erf_grad(g.x), d.y * a.y * gelu_erf_grad(g.y));
      }
    }
  }
}

static int launch_act(const void* u, int64_t ldu, const void* dh, int64_t lddh, void* out, int64_t ldo, int64_t M, int64_t n_out,
                      int act, bool bwd, cudaStream_t stream) {
  AF_CHECK(u && out && (!bwd || dh), "act: null pointer");
  AF_CHECK(M > 0 && n_out > 0 && n_out % 2 == 0 && ldu % 2 == 0 && ldo % 2 == 0 && lddh % 2 == 0, "act: bad shape / strides");
  AF_CHECK(act != ADAFACE_ACT_GEGLU || n_out % 64 == 0, "act: GEGLU width must be a multiple of 64");
  const long long total = M * (n_out / 2);
  const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  const bf16 *pu = (const bf16*)u, *pd = (const bf16*)dh;
  bf16* po = (bf16*)out;
  if (act == ADAFACE_ACT_QUICK_GELU) {
    if (bwd) act_kernel<ADAFACE_ACT_QUICK_GELU, true><<<grid, 256, 0, stream>>>(pu, ldu, pd, lddh, po, ldo, (int)M, (int)n_out);
    else act_kernel<ADAFACE_ACT_QUICK_GELU, false><<<grid, 256, 0, stream>>>(pu, ldu, pd, lddh, po, ldo, (int)M, (int)n_out);
  } else if (act == ADAFACE_ACT_GEGLU) {
    if (bwd) act_kernel<ADAFACE_ACT_GEGLU, true><<<grid, 256, 0, stream>>>(pu, ldu, pd, lddh, po, ldo, (int)M, (int)n_out);
    else act_kernel<ADAFACE_ACT_GEGLU, false><<<grid, 256, 0, stream>>>(pu, ldu, pd, lddh, po, ldo, (int)M, (int)n_out);
  } else {
    set_error("act: unsupported activation %d", act);
    return 1;
  }
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}
int act_fwd(const void* u, int64_t ldu, void* h, int64_t ldh, int64_t M, int64_t n_out, int act, cudaStream_t stream) {
  return launch_act(u, ldu, nullptr, 0, h, ldh, M, n_out, act, false, stream);
}
int act_bwd(const void* u, int64_t ldu, const void* dh, int64_t lddh, void* du, int64_t lddu, int64_t M, int64_t n_out, int act,
            cudaStream_t stream) {
  return launch_act(u, ldu, dh, lddh, du, lddu, M, n_out, act, true, stream);
}

// ---------------------------------------------------------------------------------------------
struct HeadBwdPtrs {
  const float* h[4];
  float* dh[4];
  float wl[4];
  const float* wl_dev;     // device copy of the layer weights (used when set)
  int n;
};

// mix = sum_l wl[l] h_l; out = LN(mix) w + b.  Given dout: dmix (LayerNorm backward), dh_l = wl[l] dmix,
// dwl[l] += <dmix, h_l>, dw += dout xhat, db += dout.
template <int MAXC>
__global__ void __launch_bounds__(256) sbg_head_bwd_kernel(const HeadBwdPtrs hp, long long ldh, const float* __restrict__ w,
                                                            const float* __restrict__ dout, long long lddo, float* __restrict__ dwl,
                                                            float* __restrict__ dw, float* __restrict__ db, int M, int C, float eps) {
  constexpr int PER = MAXC / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dwa[PER], dba[PER], dl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < PER; ++i) dwa[i] = dba[i] = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    float v[PER], gy[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      float m = 0.f;
      if (c < C) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (l < hp.n) m += (hp.wl_dev ? __ldg(hp.wl_dev + l) : hp.wl[l]) * hp.h[l][(long long)row * ldh + c];
      }
      v[i] = m;
      gy[i] = c < C ? dout[(long long)row * lddo + c] : 0.f;
      s += m;
    }
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      v[i] = c < C ? v[i] - mean : 0.f;
      ss += v[i] * v[i];
    }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      v[i] *= rstd;
      dwa[i] += gy[i] * v[i];
      dba[i] += gy[i];
      gy[i] *= c < C ? __ldg(w + c) : 0.f;
      sg += gy[i];
      sgx += gy[i] * v[i];
    }
    const float mg = warp_sum(sg) / C, mgx = warp_sum(sgx) / C;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      if (c < C) {
        const float dm = rstd * (gy[i] - mg - v[i] * mgx);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          if (l < hp.n) {
            const long long o = (long long)row * ldh + c;
            dl[l] += dm * hp.h[l][o];
            hp.dh[l][o] = (hp.wl_dev ? __ldg(hp.wl_dev + l) : hp.wl[l]) * dm;
          }
        }
      }
    }
  }
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    if (l < hp.n) {
      const float t = warp_sum(dl[l]);
      if (lane == 0) atomicAdd(dwl + l, t);
    }
  }
  __shared__ float red[8][MAXC + 1];
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; ++i) red[warp][i * 32 + lane] = pass ? dba[i] : dwa[i];
    __syncthreads();
    float* out = pass ? db : dw;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) t += red[ww][c];
      atomicAdd(out + c, t);
    }
  }
}

int sbg_head_bwd(const float* h0, const float* h1, const float* h2, const float* h3, const float* wl, int n_layers, int64_t ldh,
                 const float* w, const float* dout, int64_t lddo, float* dh0, float* dh1, float* dh2, float* dh3, float* dwl,
                 float* dw, float* db, int64_t M, int64_t C, float eps, cudaStream_t stream, const float* wl_dev) {
  AF_CHECK(n_layers >= 1 && n_layers <= 4 && (wl || wl_dev) && w && dout && dwl && dw && db, "sbg_head_bwd: bad arguments");
  AF_CHECK(M > 0 && C > 0 && C <= 768, "sbg_head_bwd: unsupported shape M=%lld C=%lld (C <= 768)", (long long)M, (long long)C);
  HeadBwdPtrs hp;
  const float* hs[4] = {h0, h1, h2, h3};
  float* dhs[4] = {dh0, dh1, dh2, dh3};
  for (int i = 0; i < 4; ++i) {
    hp.h[i] = i < n_layers ? hs[i] : nullptr;
    hp.dh[i] = i < n_layers ? dhs[i] : nullptr;
    hp.wl[i] = (i < n_layers && wl) ? wl[i] : 0.f;
    AF_CHECK(i >= n_layers || (hs[i] != nullptr && dhs[i] != nullptr), "sbg_head_bwd: null hidden state / gradient %d", i);
  }
  hp.n = n_layers;
  hp.wl_dev = wl_dev;
  int grid = (int)((M + 7) / 8);
  if (grid > 148 * 2) grid = 148 * 2;
  sbg_head_bwd_kernel<768><<<grid, 256, 0, stream>>>(hp, ldh, w, dout, lddo, dwl, dw, db, (int)M, (int)C, eps);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

}  // namespace adaface
