// K5: flash-attention backward on warp-level tensor-core MMA (mma.sync m16n8k16, bf16 -> fp32), recompute form:
// nothing of size L x L is ever stored; the forward pass only leaves lse[b,h,i] = log2(sum_j exp2(s_ij)).
//
//   delta_i = sum_d dO_id O_id
//   P_ij    = exp2(scale_log2 * q_i.k_j - lse_i)
//   dV_j    = sum_i P_ij dO_i                     dP_ij = dO_i . v_j
//   dS_ij   = P_ij (dP_ij - delta_i) * scale      dK_j  = sum_i dS_ij q_i        dQ_i = sum_j dS_ij k_j
//
// Two kernels, no atomics, deterministic:
//   attn_bwd_dkdv_kernel  one CTA per 64 keys (a warp owns 16 keys), loops over the query blocks; it computes the
//                         TRANSPOSED tiles S^T = K Q^T and dP^T = V dO^T so that P^T and dS^T come out of the MMA in
//                         exactly the register layout the next MMA wants as its A operand (no shared-memory transpose).
//   attn_bwd_dq_kernel    one CTA per 64 queries (a warp owns 16 queries), loops over the key blocks.
// This is the backward of F.scaled_dot_product_attention at dalc:321 / ldm attention.py:181-204 (optional key mask,
// dalc:254-273), used by the stage-2 training step through the `sc` instance (ddpm.py:1645-1707).
#include <math.h>

#include "attn_common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

struct BwdParams {
  const bf16 *q, *k, *v, *o, *dout;
  long long q_sb, q_sn, k_sb, k_sn, v_sb, v_sn, o_sb, o_sn, do_sb, do_sn;
  const float* lse;     // [B, H, Lq] log2 domain
  float* delta;         // [B, H, Lq]
  bf16 *dq, *dk, *dv;
  long long dq_sb, dq_sn, dk_sb, dk_sn, dv_sb, dv_sn;
  int B, H, Lq, Lk;
  const uint8_t* key_mask;
  int causal_mult;      // 0: none; M >= 1: key j visible to query i iff j / M <= i, keys of a token stored back to back
  float scale, scale_log2;
};

// ---------------------------------------------------------------------------------------------
// delta[b,h,i] = sum_d dO * O : one warp per (b, h, i)
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const BwdParams p) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31, h = blockIdx.y, b = blockIdx.z;
  if (row >= p.Lq) return;
  const bf16* o = p.o + (long long)b * p.o_sb + (long long)row * p.o_sn + h * D;
  const bf16* d = p.dout + (long long)b * p.do_sb + (long long)row * p.do_sn + h * D;
  float s = 0.f;
  for (int c = lane * 2; c < D; c += 64) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(o + c);
    const __nv_bfloat162 g = *reinterpret_cast<const __nv_bfloat162*>(d + c);
    s += __bfloat162float(a.x) * __bfloat162float(g.x) + __bfloat162float(a.y) * __bfloat162float(g.y);
  }
  s = warp_sum(s);
  if (lane == 0) p.delta[((long long)b * p.H + h) * p.Lq + row] = s;
}

// A fragments (16 rows x DP) of this warp's 16 rows of a [rows][LD] smem tile
template <int D>
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[AttDims<D>::KT][4], const bf16* tile, int row0, int lane) {
  constexpr int LD = AttDims<D>::LD;
#pragma unroll
  for (int kk = 0; kk < AttDims<D>::KT; ++kk)
    ldsm_x4(smem_u32(tile + (row0 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8), f[kk][0], f[kk][1], f[kk][2], f[kk][3]);
}
// acc[16 x 64] += A(16 x DP, frags) * B^T where B = smem tile [64 rows (n)][DP (k)]
template <int D>
__device__ __forceinline__ void mma_a_bt(float (&acc)[8][4], const uint32_t (&af)[AttDims<D>::KT][4], const bf16* tileB, int lane) {
  constexpr int LD = AttDims<D>::LD;
#pragma unroll
  for (int kk = 0; kk < AttDims<D>::KT; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(smem_u32(tileB + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * LD + kk * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
      mma_bf16_16816(acc[2 * np], af[kk], b0, b1);
      mma_bf16_16816(acc[2 * np + 1], af[kk], b2, b3);
    }
  }
}
// out[16 x D] += A(16 x 64, accumulator-layout values `pv`) * B where B = smem tile [64 rows (k)][D (n)]
template <int D>
__device__ __forceinline__ void mma_p_b(float (&out)[AttDims<D>::NT_O][4], const float (&pv)[8][4], const bf16* tileB, int lane) {
  constexpr int LD = AttDims<D>::LD, NT_O = AttDims<D>::NT_O;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(pv[2 * kk][0], pv[2 * kk][1]);
    a[1] = pack_bf16(pv[2 * kk][2], pv[2 * kk][3]);
    a[2] = pack_bf16(pv[2 * kk + 1][0], pv[2 * kk + 1][1]);
    a[3] = pack_bf16(pv[2 * kk + 1][2], pv[2 * kk + 1][3]);
    const bf16* brow = tileB + (kk * 16 + (lane & 15)) * LD;
#pragma unroll
    for (int nt = 0; nt + 1 < NT_O; nt += 2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_trans(smem_u32(brow + nt * 8 + (lane >> 4) * 8), b0, b1, b2, b3);
      mma_bf16_16816(out[nt], a, b0, b1);
      mma_bf16_16816(out[nt + 1], a, b2, b3);
    }
    if (NT_O & 1) {
      uint32_t b0, b1;
      ldsm_x2_trans(smem_u32(brow + (NT_O - 1) * 8), b0, b1);
      mma_bf16_16816(out[NT_O - 1], a, b0, b1);
    }
  }
}
// this warp's 16 x D fp32 accumulator -> bf16 -> smem staging (its own 16 rows) -> coalesced 16-byte stores
template <int D>
__device__ __forceinline__ void store_rows(const float (&acc)[AttDims<D>::NT_O][4], float mul, bf16* stage, bf16* gbase,
                                           long long stride_n, int grow0, int L, int lane, int mult = 1, int sub = 0) {
  constexpr int LD = AttDims<D>::LD, CH = AttDims<D>::CH;
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < AttDims<D>::NT_O; ++nt) {
    *reinterpret_cast<uint32_t*>(stage + g * LD + nt * 8 + 2 * t) = pack_bf16(acc[nt][0] * mul, acc[nt][1] * mul);
    *reinterpret_cast<uint32_t*>(stage + (g + 8) * LD + nt * 8 + 2 * t) = pack_bf16(acc[nt][2] * mul, acc[nt][3] * mul);
  }
  __syncwarp();
  for (int c = lane; c < 16 * CH; c += 32) {
    const int r = c / CH, ch = c - r * CH;
    const int j = grow0 + r;
    if (j < L) {
      const long long off = mult > 1 ? (long long)(j / mult) * stride_n + (long long)(j % mult) * sub : (long long)j * stride_n;
      *reinterpret_cast<uint4*>(gbase + off + ch * 8) = *reinterpret_cast<const uint4*>(stage + r * LD + ch * 8);
    }
  }
}

// MODE 0: dK and dV in one pass; 1: dV only; 2: dK only (d = 160 runs two passes to fit the register file)
template <int D, int MODE>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkdv_kernel(const BwdParams p) {
  using A = AttDims<D>;
  constexpr int LD = A::LD, KT = A::KT, NT_O = A::NT_O;
  extern __shared__ __align__(16) uint8_t smem_bw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_bw);     // [64][LD]
  bf16* sV = sK + 64 * LD;                         // [64][LD]
  bf16* sQ = sV + 64 * LD;                         // [2][64][LD]
  bf16* sdO = sQ + 2 * 64 * LD;                    // [2][64][LD]
  float* sLse = reinterpret_cast<float*>(sdO + 2 * 64 * LD);   // [2][64]
  float* sDelta = sLse + 2 * 64;                   // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const bf16* gq = p.q + (long long)b * p.q_sb + h * D;
  const bf16* gdo = p.dout + (long long)b * p.do_sb + h * D;
  const long long stat0 = ((long long)b * p.H + h) * p.Lq;
  const int n_qblk = (p.Lq + 63) / 64;
  const int M = p.causal_mult, sub = p.H * D;
  const int blk0 = M > 0 ? min((n0 / M) / 64, n_qblk) : 0;   // causal: earlier queries see none of this CTA's keys

  auto load_q = [&](int blk, int buf) {
    load_rows<D>(sQ + buf * 64 * LD, gq, p.q_sn, blk * 64, p.Lq, 64);
    load_rows<D>(sdO + buf * 64 * LD, gdo, p.do_sn, blk * 64, p.Lq, 64);
    if (threadIdx.x < 64) {
      const int i = blk * 64 + threadIdx.x;
      sLse[buf * 64 + threadIdx.x] = i < p.Lq ? p.lse[stat0 + i] : INFINITY;   // +inf => P = 0 for missing queries
      sDelta[buf * 64 + threadIdx.x] = i < p.Lq ? p.delta[stat0 + i] : 0.f;
    }
  };
  zero_pad_cols<D>(sK, 2 * 64);
  zero_pad_cols<D>(sQ, 4 * 64);
  load_rows<D>(sK, p.k + (long long)b * p.k_sb + h * D, p.k_sn, n0, p.Lk, 64, M, sub);
  load_rows<D>(sV, p.v + (long long)b * p.v_sb + h * D, p.v_sn, n0, p.Lk, 64, M, sub);
  if (blk0 < n_qblk) load_q(blk0, 0);
  cp_async_commit();

  // validity of this thread's two key rows
  bool kok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int key = n0 + warp * 16 + g + i * 8;
    kok[i] = key < p.Lk && (!p.key_mask || p.key_mask[(long long)b * p.Lk + key] != 0);
  }
  uint32_t kf[KT][4], vf[KT][4];
  float dk_acc[NT_O][4], dv_acc[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk_acc[i][e] = dv_acc[i][e] = 0.f;

  for (int blk = blk0; blk < n_qblk; ++blk) {
    const int buf = (blk - blk0) & 1;
    if (blk + 1 < n_qblk) {
      load_q(blk + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (blk == blk0) {
      load_a_frags<D>(kf, sK, warp * 16, lane);
      load_a_frags<D>(vf, sV, warp * 16, lane);
    }
    const bf16* tQ = sQ + buf * 64 * LD;
    const bf16* tdO = sdO + buf * 64 * LD;
    const float* tL = sLse + buf * 64;
    const float* tD = sDelta + buf * 64;

    float st[8][4];                       // S^T tile: rows = this warp's keys, columns = the 64 queries
#pragma unroll
    for (int i = 0; i < 8; ++i) st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
    mma_a_bt<D>(st, kf, tQ, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qc = nt * 8 + 2 * t + (e & 1);
        bool ok = kok[e >> 1];
        if (M > 0) ok = ok && ((long long)(n0 + warp * 16 + g + (e >> 1) * 8) < (long long)(blk * 64 + qc + 1) * M);
        st[nt][e] = ok ? fast_exp2(st[nt][e] * p.scale_log2 - tL[qc]) : 0.f;             // P^T
      }
    }
    if (MODE != 2) mma_p_b<D>(dv_acc, st, tdO, lane);                                     // dV += P^T dO
    if (MODE != 1) {
      float dpt[8][4];                    // dP^T = V dO^T
#pragma unroll
      for (int i = 0; i < 8; ++i) dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      mma_a_bt<D>(dpt, vf, tdO, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qc = nt * 8 + 2 * t + (e & 1);
          dpt[nt][e] = st[nt][e] * (dpt[nt][e] - tD[qc]);                                 // dS^T (without the scale)
        }
      }
      mma_p_b<D>(dk_acc, dpt, tQ, lane);                                                  // dK += dS^T Q
    }
    __syncthreads();
  }
  bf16* stage = sQ + warp * 16 * LD;      // the Q buffers are free now; each warp stages through its own 16 rows
  if (MODE != 2)
    store_rows<D>(dv_acc, 1.f, stage, p.dv + (long long)b * p.dv_sb + h * D, p.dv_sn, n0 + warp * 16, p.Lk, lane, M, sub);
  if (MODE != 1)
    store_rows<D>(dk_acc, p.scale, stage, p.dk + (long long)b * p.dk_sb + h * D, p.dk_sn, n0 + warp * 16, p.Lk, lane, M, sub);
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_kernel(const BwdParams p) {
  using A = AttDims<D>;
  constexpr int LD = A::LD, KT = A::KT, NT_O = A::NT_O;
  extern __shared__ __align__(16) uint8_t smem_bq[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_bq);     // [64][LD]
  bf16* sdO = sQ + 64 * LD;                        // [64][LD]
  bf16* sK = sdO + 64 * LD;                        // [2][64][LD]
  bf16* sV = sK + 2 * 64 * LD;                     // [2][64][LD]
  uint8_t* sValid = reinterpret_cast<uint8_t*>(sV + 2 * 64 * LD);   // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const bf16* gk = p.k + (long long)b * p.k_sb + h * D;
  const bf16* gv = p.v + (long long)b * p.v_sb + h * D;
  const long long stat0 = ((long long)b * p.H + h) * p.Lq;
  const int M = p.causal_mult, sub = p.H * D;
  int n_kblk = (p.Lk + 63) / 64;
  if (M > 0) n_kblk = min(n_kblk, (int)(((long long)min(p.Lq, m0 + 64) * M + 63) / 64));

  auto load_kv = [&](int blk, int buf) {
    load_rows<D>(sK + buf * 64 * LD, gk, p.k_sn, blk * 64, p.Lk, 64, M, sub);
    load_rows<D>(sV + buf * 64 * LD, gv, p.v_sn, blk * 64, p.Lk, 64, M, sub);
    if (threadIdx.x < 64) {
      const int j = blk * 64 + threadIdx.x;
      sValid[buf * 64 + threadIdx.x] = j < p.Lk && (!p.key_mask || p.key_mask[(long long)b * p.Lk + j] != 0);
    }
  };
  zero_pad_cols<D>(sQ, 2 * 64);
  zero_pad_cols<D>(sK, 4 * 64);
  load_rows<D>(sQ, p.q + (long long)b * p.q_sb + h * D, p.q_sn, m0, p.Lq, 64);
  load_rows<D>(sdO, p.dout + (long long)b * p.do_sb + h * D, p.do_sn, m0, p.Lq, 64);
  load_kv(0, 0);
  cp_async_commit();

  float lse[2], dl[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = m0 + warp * 16 + g + i * 8;
    lse[i] = r < p.Lq ? p.lse[stat0 + r] : INFINITY;
    dl[i] = r < p.Lq ? p.delta[stat0 + r] : 0.f;
  }
  float dq_acc[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f;

  for (int blk = 0; blk < n_kblk; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < n_kblk) {
      load_kv(blk + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const bf16* tK = sK + buf * 64 * LD;
    const bf16* tV = sV + buf * 64 * LD;
    const uint8_t* tOk = sValid + buf * 64;

    // Q / dO fragments are re-read from shared memory every block: holding them would not fit the register file at d = 160
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    {
      uint32_t f[KT][4];
      load_a_frags<D>(f, sQ, warp * 16, lane);
      mma_a_bt<D>(s, f, tK, lane);                 // S = Q K^T
      load_a_frags<D>(f, sdO, warp * 16, lane);
      mma_a_bt<D>(dp, f, tV, lane);                // dP = dO V^T
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kc = nt * 8 + 2 * t + (e & 1);
        bool ok = tOk[kc] != 0;
        if (M > 0) ok = ok && ((long long)(blk * 64 + kc) < (long long)(m0 + warp * 16 + g + (e >> 1) * 8 + 1) * M);
        const float pr = ok ? fast_exp2(s[nt][e] * p.scale_log2 - lse[e >> 1]) : 0.f;
        s[nt][e] = pr * (dp[nt][e] - dl[e >> 1]);  // dS (without the scale)
      }
    }
    mma_p_b<D>(dq_acc, s, tK, lane);               // dQ += dS K
    __syncthreads();
  }
  store_rows<D>(dq_acc, p.scale, sQ + warp * 16 * LD, p.dq + (long long)b * p.dq_sb + h * D, p.dq_sn, m0 + warp * 16, p.Lq, lane);
}

template <int D>
static int launch_bwd(const BwdParams& p, cudaStream_t stream) {
  using A = AttDims<D>;
  constexpr int smem_kv = 6 * 64 * A::LD * 2 + 4 * 64 * 4;
  constexpr int smem_q = 6 * 64 * A::LD * 2 + 2 * 64;
  static DevOnce configured;
  const int cfg_dev = af_device();
  if (!configured.done(cfg_dev)) {
    AF_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    AF_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    AF_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    AF_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
    configured.set(cfg_dev);
  }
  const dim3 gkv((p.Lk + 63) / 64, p.H, p.B), gq((p.Lq + 63) / 64, p.H, p.B);
  if (D <= 80) {
    attn_bwd_dkdv_kernel<D, 0><<<gkv, ATT_THREADS, smem_kv, stream>>>(p);
    g_launch_count += 1;
  } else {
    attn_bwd_dkdv_kernel<D, 1><<<gkv, ATT_THREADS, smem_kv, stream>>>(p);
    attn_bwd_dkdv_kernel<D, 2><<<gkv, ATT_THREADS, smem_kv, stream>>>(p);
    g_launch_count += 2;
  }
  attn_bwd_dq_kernel<D><<<gq, ATT_THREADS, smem_q, stream>>>(p);
  AF_CUDA(cudaGetLastError());
  g_launch_count += 1;
  return 0;
}

template <int D>
static int launch_delta(const BwdParams& p, cudaStream_t stream) {
  attn_bwd_delta_kernel<D><<<dim3((p.Lq + 7) / 8, p.H, p.B), 256, 0, stream>>>(p);
  AF_CUDA(cudaGetLastError());
  g_launch_count += 1;
  return 0;
}

int attn_bwd_tcgen05(const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*,
                     int64_t, int64_t, const float*, const float*, void*, int64_t, int64_t, void*, int64_t, int64_t, void*, int64_t,
                     int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, float, const uint8_t*, cudaStream_t);

int attn_bwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn, const void* v, int64_t v_sb,
             int64_t v_sn, const void* o, int64_t o_sb, int64_t o_sn, const void* dout, int64_t do_sb, int64_t do_sn,
             const float* lse, float* delta, void* dq, int64_t dq_sb, int64_t dq_sn, void* dk, int64_t dk_sb, int64_t dk_sn,
             void* dv, int64_t dv_sb, int64_t dv_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
             const uint8_t* key_mask, int causal_mult, float scale, cudaStream_t stream) {
  AF_CHECK(q && k && v && o && dout && lse && delta && dq && dk && dv, "attn_bwd: null pointer");
  AF_CHECK(causal_mult >= 0, "attn_bwd: causal_mult must be >= 0");
  AF_CHECK(B > 0 && H > 0 && Lq > 0 && Lk > 0 && B <= 65535 && H <= 65535, "attn_bwd: bad problem size");
  const int64_t strides[] = {q_sb, q_sn, k_sb, k_sn, v_sb, v_sn, o_sb, o_sn, do_sb, do_sn, dq_sb, dq_sn, dk_sb, dk_sn, dv_sb, dv_sn};
  for (int64_t s : strides) AF_CHECK(s % 8 == 0, "attn_bwd: strides must be multiples of 8 elements");
  BwdParams p;
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (const bf16*)o; p.dout = (const bf16*)dout;
  p.q_sb = q_sb; p.q_sn = q_sn; p.k_sb = k_sb; p.k_sn = k_sn; p.v_sb = v_sb; p.v_sn = v_sn; p.o_sb = o_sb; p.o_sn = o_sn;
  p.do_sb = do_sb; p.do_sn = do_sn;
  p.lse = lse; p.delta = delta;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.dq_sb = dq_sb; p.dq_sn = dq_sn; p.dk_sb = dk_sb; p.dk_sn = dk_sn; p.dv_sb = dv_sb; p.dv_sn = dv_sn;
  p.B = (int)B; p.H = (int)H; p.Lq = (int)Lq; p.Lk = (int)Lk;
  p.key_mask = key_mask;
  p.causal_mult = causal_mult;
  p.scale = scale;
  p.scale_log2 = scale * LOG2E;
  int rc;
  switch (d) {
    case 40: rc = launch_delta<40>(p, stream); break;
    case 64: rc = launch_delta<64>(p, stream); break;
    case 80: rc = launch_delta<80>(p, stream); break;
    case 160: rc = launch_delta<160>(p, stream); break;
    default:
      set_error("attn_bwd: unsupported head dim %lld (supported: 40, 64, 80, 160)", (long long)d);
      return 1;
  }
  if (rc) return rc;
  if (causal_mult == 0) {
    // non-causal attention (key mask or none): tcgen05 / TMEM kernels (attn_bwd_tcgen05.cu)
    rc = attn_bwd_tcgen05(q, q_sb, q_sn, k, k_sb, k_sn, v, v_sb, v_sn, dout, do_sb, do_sn, lse, delta, dq, dq_sb, dq_sn, dk, dk_sb,
                          dk_sn, dv, dv_sb, dv_sn, B, H, Lq, Lk, d, scale, key_mask, stream);
    if (rc >= 0) return rc;
  }
  switch (d) {
    case 40: return launch_bwd<40>(p, stream);
    case 64: return launch_bwd<64>(p, stream);
    case 80: return launch_bwd<80>(p, stream);
  }
  return launch_bwd<160>(p, stream);
}

}  // namespace adaface
