// K5b: backward of the capture / normalize cross-attention (the slow SDPA of dalc:79-139) in ONE streaming pass.
//
// The stage-2 losses back-propagate through the captured probabilities (`attn`), the edited scores (`attnscore`) and
// the attention output at once (ldm/util.py:1822-1918, 2047-2121), so the kernel takes up to three upstream
// gradients -- dO [B,Lq,C] bf16, dprob and dscore [B,H,Lq,S] fp32 (either may be NULL) -- and recomputes S / P from
// q, k exactly as the forward kernel does (bf16 hi/lo split when q, k, v are fp32).  Nothing of size Lq x S is
// written.  With s' the edited score (dalc:119-133) and P = softmax(s'):
//
//   dP    = dO V^T + dprob                           ds' = P o (dP - rowsum(P o dP)) + dscore
//   normalize: s' = (scale q.k - mean_i) * c on flagged columns, mean detached  =>  ds = ds' * c * scale,
//              dc = sum_{flagged} ds' * (scale q.k - mean_i)        (c = cross_attn_scale_factor)
//   dQ = ds K         dK = ds^T Q         dV = P^T dO
//
// HBM-bound (S <= 128 keys, K/V staged once per CTA): a CTA owns a contiguous chunk of 64-query tiles of one
// (batch, head), streams q / dO / dprob / dscore once, writes dQ once, and keeps dK / dV in fp32 shared-memory
// accumulators that are flushed as per-chunk partials; cross_bwd_reduce_kernel then sums the few partials per
// (batch, head) deterministically (no atomics) and emits dk / dv in the layout of k / v plus d(cross_attn_scale_factor).
// mix_attn_mats_in_batch (dalc:108-118): the score tile of (sc, mc) is shared, so their dP add up; only the sc half gets dq / dk.
#include <math.h>

#include "attn_common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

constexpr int CB_BN = 128;          // max context keys
constexpr int CB_LDP = CB_BN + 8;   // row pitch of the bf16 P / dS tiles (conflict-free ldmatrix)

struct CapBwdParams {
  const void *q, *k, *v;            // bf16, or fp32 when instantiated with F32IN
  const bf16* dout;
  long long q_sb, q_sn, k_sb, k_sn, v_sb, v_sn, do_sb, do_sn;
  const float* dprob;
  const float* dscore;
  const uint8_t* col_flag;
  const float* qmean;
  const float* ca_scale;
  bf16* dq;
  long long dq_sb, dq_sn;
  float* dk_part;                   // [B, H, chunks, S, D]
  float* dv_part;
  float* dca_part;                  // [B * H * chunks]
  int B, H, Lq, S, tiles_per_cta;
  float scale;
  // ---- implicit dprob terms of the fused capture consumers (the dense [B,H,Lq,S] gradient map is never materialised)
  const uint8_t* sum_flag;          // [B, S]
  const float* g_subj;              // [B, H, Lq]: dprob[b,h,r,c] += g_subj[b,h,r] * sum_flag[b,c]      (d subj_sum)
  const float* ref_prob;            // [B, H, Lq, S]
  const float* mse_coef;            // device [B]: dprob[b] += mse_coef[b] * (P - ref_prob)              (d sum (P - ref)^2 = 2 (P - ref))
};

// out[16 keys x D] = sum over the 64 queries of A^T B: A = [64 queries][CB_LDP] tile (P or dS, columns = keys),
// B = [64 queries][LD] tile (dO or Q).  A^T fragments come out of ldmatrix.trans, no shared-memory transpose.
template <int D>
__device__ __forceinline__ void mma_at_b(float (&out)[AttDims<D>::NT_O][4], const bf16* tileA, int key0, const bf16* tileB,
                                         int lane) {
  constexpr int LD = AttDims<D>::LD, NT_O = AttDims<D>::NT_O;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    const int mi = lane >> 3, r = lane & 7;
    ldsm_x4_trans(smem_u32(tileA + (kk * 16 + (mi >> 1) * 8 + r) * CB_LDP + key0 + (mi & 1) * 8), a[0], a[1], a[2], a[3]);
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

template <int D, bool MIX, bool F32IN>
__global__ void __launch_bounds__(ATT_THREADS) attn_cross_capture_bwd_kernel(const CapBwdParams p) {
  using A = AttDims<D>;
  constexpr int LD = A::LD, KT = A::KT, NT_O = A::NT_O, NT_S = CB_BN / 8, NP = F32IN ? 2 : 1, NI = MIX ? 2 : 1;
  extern __shared__ __align__(16) uint8_t smem_cb[];
  bf16* sK = reinterpret_cast<bf16*>(smem_cb);           // [NP][NI][128][LD]   (NP: hi / lo parts)
  bf16* sKlo = sK + NI * CB_BN * LD;
  bf16* sV = sK + NP * NI * CB_BN * LD;                  // [NI][128][LD]
  bf16* sQ = sV + NI * CB_BN * LD;                       // [NP][NI][64][LD]
  bf16* sQlo = sQ + NI * ATT_BM * LD;
  bf16* sdO = sQ + NP * NI * ATT_BM * LD;                // [NI][64][LD]
  bf16* sP = sdO + NI * ATT_BM * LD;                     // [64][CB_LDP]
  bf16* sdS = sP + ATT_BM * CB_LDP;                      // [64][CB_LDP]
  float* sAccK = reinterpret_cast<float*>(sdS + ATT_BM * CB_LDP);   // [128][D]      (instance 0 only under MIX)
  float* sAccV = sAccK + CB_BN * D;                      // [NI][128][D]
  float* sColMean = sAccV + NI * CB_BN * D;              // [128]
  float* sRed = sColMean + CB_BN;                        // [4]
  uint8_t* sFlag = reinterpret_cast<uint8_t*>(sRed + 4); // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int chunk = blockIdx.x, h = blockIdx.y, b0 = blockIdx.z;
  const int half = p.B / 2;                              // MIX: instances b0 (sc) and b0 + half (mc)
  const int S = p.S, nts = (S + 7) / 8;
  const int n_tiles_all = (p.Lq + ATT_BM - 1) / ATT_BM;
  const int tile_begin = chunk * p.tiles_per_cta, tile_end = min(n_tiles_all, tile_begin + p.tiles_per_cta);

  zero_pad_cols<D>(sK, (NP + 1) * NI * CB_BN);           // sK (hi, lo) and sV are contiguous
  zero_pad_cols<D>(sQ, (NP + 1) * NI * ATT_BM);          // sQ (hi, lo) and sdO
  for (int i = threadIdx.x; i < (1 + NI) * CB_BN * D; i += blockDim.x) sAccK[i] = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int b = b0 + i * half;
    if constexpr (F32IN) {
      load_rows_f32_split<D>(sK + i * CB_BN * LD, sKlo + i * CB_BN * LD, (const float*)p.k + (long long)b * p.k_sb + h * D,
                             p.k_sn, 0, S, CB_BN);
      load_rows_f32_split<D>(sV + i * CB_BN * LD, nullptr, (const float*)p.v + (long long)b * p.v_sb + h * D, p.v_sn, 0, S,
                             CB_BN);
    } else {
      load_rows<D>(sK + i * CB_BN * LD, (const bf16*)p.k + (long long)b * p.k_sb + h * D, p.k_sn, 0, S, CB_BN);
      load_rows<D>(sV + i * CB_BN * LD, (const bf16*)p.v + (long long)b * p.v_sb + h * D, p.v_sn, 0, S, CB_BN);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  bool my_flag = false;
  if (threadIdx.x < CB_BN) {
    const int j = threadIdx.x;
    float cm = 0.f;
    if (!MIX && p.col_flag && j < S && p.col_flag[(long long)b0 * S + j]) {
      my_flag = true;
      const float* qm = p.qmean + ((long long)b0 * p.H + h) * D;
      const bf16* kr = sK + j * LD;
      const bf16* kl = sKlo + j * LD;
#pragma unroll 8
      for (int dd = 0; dd < D; ++dd) {
        float kv = __bfloat162float(kr[dd]);
        if constexpr (F32IN) kv += __bfloat162float(kl[dd]);
        cm += qm[dd] * kv;
      }
      cm *= p.scale;
    }
    sColMean[j] = cm;
    sFlag[j] = my_flag ? 1 : 0;
  }
  const bool any_flag = __syncthreads_or(my_flag ? 1 : 0) != 0;
  const float ca_scale = p.ca_scale ? __ldg(p.ca_scale) : 1.f;
  const float sc = MIX ? 0.5f * p.scale : p.scale;       // (score_sc + score_mc) / 2, dalc:117
  float dc_local = 0.f;
  static_assert(NT_S <= 16, "sum_bits holds two columns for each of at most 16 key tiles");
  uint32_t sum_bits = 0;       // this thread's score columns that belong to the subject-column sum (fused consumers)
  if (!MIX && p.g_subj) {
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = nt * 8 + 2 * t + e;
        if (c < S && p.sum_flag[(long long)b0 * S + c]) sum_bits |= 1u << (2 * nt + e);
      }
    }
  }

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int m0 = tile * ATT_BM;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int b = b0 + i * half;
      if constexpr (F32IN)
        load_rows_f32_split<D>(sQ + i * ATT_BM * LD, sQlo + i * ATT_BM * LD, (const float*)p.q + (long long)b * p.q_sb + h * D,
                               p.q_sn, m0, p.Lq, ATT_BM);
      else
        load_rows<D>(sQ + i * ATT_BM * LD, (const bf16*)p.q + (long long)b * p.q_sb + h * D, p.q_sn, m0, p.Lq, ATT_BM);
      load_rows<D>(sdO + i * ATT_BM * LD, p.dout + (long long)b * p.do_sb + h * D, p.do_sn, m0, p.Lq, ATT_BM);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- S = Q K^T (hi.hi + lo.hi + hi.lo for fp32 inputs, exactly as the forward kernel; MIX: both instances summed)
    float acc_s[NT_S][4];
#pragma unroll
    for (int i = 0; i < NT_S; ++i) acc_s[i][0] = acc_s[i][1] = acc_s[i][2] = acc_s[i][3] = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        const int q_off = i * ATT_BM * LD + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8;
        uint32_t qf[4], ql[4];
        ldsm_x4(smem_u32(sQ + q_off), qf[0], qf[1], qf[2], qf[3]);
        if constexpr (F32IN) ldsm_x4(smem_u32(sQlo + q_off), ql[0], ql[1], ql[2], ql[3]);
#pragma unroll
        for (int np = 0; np < NT_S / 2; ++np) {
          if (2 * np < nts) {
            const int k_off = i * CB_BN * LD + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * LD + kk * 16 + ((lane >> 3) & 1) * 8;
            uint32_t r0, r1, r2, r3;
            ldsm_x4(smem_u32(sK + k_off), r0, r1, r2, r3);
            mma_bf16_16816(acc_s[2 * np], qf, r0, r1);
            mma_bf16_16816(acc_s[2 * np + 1], qf, r2, r3);
            if constexpr (F32IN) {
              mma_bf16_16816(acc_s[2 * np], ql, r0, r1);
              mma_bf16_16816(acc_s[2 * np + 1], ql, r2, r3);
              ldsm_x4(smem_u32(sKlo + k_off), r0, r1, r2, r3);
              mma_bf16_16816(acc_s[2 * np], qf, r0, r1);
              mma_bf16_16816(acc_s[2 * np + 1], qf, r2, r3);
            }
          }
        }
      }
    }
    // ---- edit + exact softmax
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = nt * 8 + 2 * t + (e & 1);
        float s = -INFINITY;
        if (c < S) {
          s = acc_s[nt][e] * sc;
          if (sFlag[c]) s = (s - sColMean[c]) * ca_scale;
        }
        acc_s[nt][e] = s;
        mx[e >> 1] = fmaxf(mx[e >> 1], s);
      }
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f((acc_s[nt][e] - mx[e >> 1]) * LOG2E);
        acc_s[nt][e] = pv;
        sum[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
      sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
      sum[i] = 1.f / sum[i];
    }
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc_s[nt][e] *= sum[e >> 1];
    }

    // ---- dP = dO V^T (+ dprob); MIX: the score tile is shared, so the two instances' dP simply add up
    float acc_dp[NT_S][4];
#pragma unroll
    for (int i = 0; i < NT_S; ++i) acc_dp[i][0] = acc_dp[i][1] = acc_dp[i][2] = acc_dp[i][3] = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        uint32_t df[4];
        ldsm_x4(smem_u32(sdO + i * ATT_BM * LD + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8), df[0], df[1], df[2], df[3]);
#pragma unroll
        for (int np = 0; np < NT_S / 2; ++np) {
          if (2 * np < nts) {
            uint32_t r0, r1, r2, r3;
            ldsm_x4(smem_u32(sV + i * CB_BN * LD + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * LD + kk * 16 + ((lane >> 3) & 1) * 8),
                    r0, r1, r2, r3);
            mma_bf16_16816(acc_dp[2 * np], df, r0, r1);
            mma_bf16_16816(acc_dp[2 * np + 1], df, r2, r3);
          }
        }
      }
    }
    const int row_lo = m0 + warp * 16 + g;
    if constexpr (!MIX) {
      if (p.g_subj) {
        const float* gs = p.g_subj + ((long long)b0 * p.H + h) * p.Lq;
        const float g0 = row_lo < p.Lq ? __ldg(gs + row_lo) : 0.f, g1 = row_lo + 8 < p.Lq ? __ldg(gs + row_lo + 8) : 0.f;
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (sum_bits & (1u << (2 * nt + (e & 1)))) acc_dp[nt][e] += (e >> 1) ? g1 : g0;
        }
      }
      if (p.ref_prob) {
        const float coef = __ldg(p.mse_coef + b0);
        const float* ref = p.ref_prob + (((long long)b0 * p.H + h) * p.Lq) * S;
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = nt * 8 + 2 * t + (e & 1), r = row_lo + (e >> 1) * 8;
            if (c < S && r < p.Lq) acc_dp[nt][e] += coef * (acc_s[nt][e] - __ldg(ref + (long long)r * S + c));
          }
        }
      }
    }
    if (p.dprob) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const long long map0 = (((long long)(b0 + i * half) * p.H + h) * p.Lq) * S;
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = nt * 8 + 2 * t + (e & 1), r = row_lo + (e >> 1) * 8;
            if (c < S && r < p.Lq) acc_dp[nt][e] += __ldg(p.dprob + map0 + (long long)r * S + c);
          }
        }
      }
    }
    float rd[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) rd[e >> 1] += acc_s[nt][e] * acc_dp[nt][e];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      rd[i] += __shfl_xor_sync(0xffffffffu, rd[i], 1);
      rd[i] += __shfl_xor_sync(0xffffffffu, rd[i], 2);
    }
    // ds' = P (dP - rowdot) + dscore ; P -> shared memory (bf16) for dV = P^T dO
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = nt * 8 + 2 * t + (e & 1), r = row_lo + (e >> 1) * 8;
        float ds = acc_s[nt][e] * (acc_dp[nt][e] - rd[e >> 1]);
        if (p.dscore && c < S && r < p.Lq) {
#pragma unroll
          for (int i = 0; i < NI; ++i)
            ds += __ldg(p.dscore + (((long long)(b0 + i * half) * p.H + h) * p.Lq + r) * S + c);
        }
        if (r >= p.Lq) ds = 0.f;
        acc_dp[nt][e] = ds;
      }
      const bool lo_ok = row_lo < p.Lq, hi_ok = row_lo + 8 < p.Lq;
      *reinterpret_cast<uint32_t*>(sP + (warp * 16 + g) * CB_LDP + nt * 8 + 2 * t) =
          lo_ok ? pack_bf16(acc_s[nt][0], acc_s[nt][1]) : 0u;
      *reinterpret_cast<uint32_t*>(sP + (warp * 16 + g + 8) * CB_LDP + nt * 8 + 2 * t) =
          hi_ok ? pack_bf16(acc_s[nt][2], acc_s[nt][3]) : 0u;
    }
    // ---- normalize: d(cross_attn_scale_factor) needs u = scale q.k - mean on the flagged columns: recompute q.k
    if (!MIX && any_flag) {
#pragma unroll
      for (int i = 0; i < NT_S; ++i) acc_s[i][0] = acc_s[i][1] = acc_s[i][2] = acc_s[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        uint32_t qf[4];
        ldsm_x4(smem_u32(sQ + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8), qf[0], qf[1], qf[2], qf[3]);
#pragma unroll
        for (int np = 0; np < NT_S / 2; ++np) {
          if (2 * np < nts) {
            uint32_t r0, r1, r2, r3;
            ldsm_x4(smem_u32(sK + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * LD + kk * 16 + ((lane >> 3) & 1) * 8), r0, r1, r2, r3);
            mma_bf16_16816(acc_s[2 * np], qf, r0, r1);
            mma_bf16_16816(acc_s[2 * np + 1], qf, r2, r3);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = nt * 8 + 2 * t + (e & 1);
          if (c < S && sFlag[c]) {
            dc_local += acc_dp[nt][e] * (acc_s[nt][e] * p.scale - sColMean[c]);
            acc_dp[nt][e] *= ca_scale;
          }
        }
      }
    }
    // ds = ds' * scale (* c on flagged columns, applied above; * 1/2 under MIX) -> bf16 tile + A operand of dQ
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc_dp[nt][e] *= sc;
      *reinterpret_cast<uint32_t*>(sdS + (warp * 16 + g) * CB_LDP + nt * 8 + 2 * t) = pack_bf16(acc_dp[nt][0], acc_dp[nt][1]);
      *reinterpret_cast<uint32_t*>(sdS + (warp * 16 + g + 8) * CB_LDP + nt * 8 + 2 * t) = pack_bf16(acc_dp[nt][2], acc_dp[nt][3]);
    }
    // ---- dQ = ds K  (instance 0 = sc; the mc half is detached, dalc:117: its dq rows stay zero)
    {
      float dq_acc[NT_O][4];
#pragma unroll
      for (int i = 0; i < NT_O; ++i) dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < CB_BN / 16; ++kk) {
        if (2 * kk < nts) {
          uint32_t a[4];
          a[0] = pack_bf16(acc_dp[2 * kk][0], acc_dp[2 * kk][1]);
          a[1] = pack_bf16(acc_dp[2 * kk][2], acc_dp[2 * kk][3]);
          a[2] = pack_bf16(acc_dp[2 * kk + 1][0], acc_dp[2 * kk + 1][1]);
          a[3] = pack_bf16(acc_dp[2 * kk + 1][2], acc_dp[2 * kk + 1][3]);
          const bf16* krow = sK + (kk * 16 + (lane & 15)) * LD;
#pragma unroll
          for (int nt = 0; nt + 1 < NT_O; nt += 2) {
            uint32_t r0, r1, r2, r3;
            ldsm_x4_trans(smem_u32(krow + nt * 8 + (lane >> 4) * 8), r0, r1, r2, r3);
            mma_bf16_16816(dq_acc[nt], a, r0, r1);
            mma_bf16_16816(dq_acc[nt + 1], a, r2, r3);
          }
          if (NT_O & 1) {
            uint32_t r0, r1;
            ldsm_x2_trans(smem_u32(krow + (NT_O - 1) * 8), r0, r1);
            mma_bf16_16816(dq_acc[NT_O - 1], a, r0, r1);
          }
        }
      }
      bf16* gdq = p.dq + (long long)b0 * p.dq_sb + h * D;
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        if (row_lo < p.Lq)
          *reinterpret_cast<uint32_t*>(gdq + (long long)row_lo * p.dq_sn + nt * 8 + 2 * t) = pack_bf16(dq_acc[nt][0], dq_acc[nt][1]);
        if (row_lo + 8 < p.Lq)
          *reinterpret_cast<uint32_t*>(gdq + (long long)(row_lo + 8) * p.dq_sn + nt * 8 + 2 * t) = pack_bf16(dq_acc[nt][2], dq_acc[nt][3]);
      }
    }
    __syncthreads();      // every warp's P / dS rows are in shared memory

    // ---- dV_i += P^T dO_i, dK += ds^T Q_0 : warp w owns keys [32 w, 32 w + 32)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int key0 = warp * 32 + hf * 16;
      if (key0 < S) {
#pragma unroll
        for (int which = 0; which < 1 + NI; ++which) {   // 0: dK, 1..NI: dV of instance which-1
          float acc[NT_O][4];
#pragma unroll
          for (int i = 0; i < NT_O; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
          mma_at_b<D>(acc, which ? sP : sdS, key0, which ? sdO + (which - 1) * ATT_BM * LD : sQ, lane);
          float* dst = sAccK + which * CB_BN * D;
#pragma unroll
          for (int nt = 0; nt < NT_O; ++nt) {
            float2* lo = reinterpret_cast<float2*>(dst + (key0 + g) * D + nt * 8 + 2 * t);
            float2* hi = reinterpret_cast<float2*>(dst + (key0 + g + 8) * D + nt * 8 + 2 * t);
            float2 x = *lo, y = *hi;
            x.x += acc[nt][0]; x.y += acc[nt][1];
            y.x += acc[nt][2]; y.y += acc[nt][3];
            *lo = x; *hi = y;
          }
        }
      }
    }
    __syncthreads();      // before the next tile overwrites sQ / sdO / sP / sdS
  }

  // ---- flush the per-chunk partials (MIX: the mc instance gets dK = 0)
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const long long part0 = ((((long long)(b0 + i * half) * p.H + h) * gridDim.x) + chunk) * (long long)S * D;
    for (int x = threadIdx.x; x < S * D; x += blockDim.x) {
      p.dk_part[part0 + x] = i == 0 ? sAccK[x] : 0.f;
      p.dv_part[part0 + x] = sAccV[i * CB_BN * D + x];
    }
  }
  dc_local = warp_sum(dc_local);
  if (lane == 0) sRed[warp] = dc_local;
  __syncthreads();
  if (threadIdx.x == 0) {
    p.dca_part[((long long)b0 * p.H + h) * gridDim.x + chunk] = sRed[0] + sRed[1] + sRed[2] + sRed[3];
    if (MIX) p.dca_part[((long long)(b0 + half) * p.H + h) * gridDim.x + chunk] = 0.f;
  }
}

// dk[b, s, h*D + dd] = sum over chunks of the partials (same for dv); block 0 also reduces d(cross_attn_scale_factor).
template <typename TOut>
__global__ void __launch_bounds__(256) cross_bwd_reduce_kernel(const float* __restrict__ dk_part, const float* __restrict__ dv_part,
                                                               const float* __restrict__ dca_part, TOut* __restrict__ dk,
                                                               long long dk_sb, long long dk_sn, TOut* __restrict__ dv,
                                                               long long dv_sb, long long dv_sn, float* __restrict__ dca,
                                                               float dca_mul, int B, int H, int S, int D, int chunks) {
  const long long total = (long long)B * H * S * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dd = (int)(i % D);
    const int s = (int)((i / D) % S);
    const int h = (int)((i / ((long long)D * S)) % H);
    const int b = (int)(i / ((long long)D * S * H));
    const long long base = (((long long)b * H + h) * chunks) * (long long)S * D + (long long)s * D + dd;
    float a = 0.f, c = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {
      a += dk_part[base + (long long)ch * S * D];
      c += dv_part[base + (long long)ch * S * D];
    }
    const long long ok = (long long)b * dk_sb + (long long)s * dk_sn + h * D + dd;
    const long long ov = (long long)b * dv_sb + (long long)s * dv_sn + h * D + dd;
    if constexpr (sizeof(TOut) == 2) {
      dk[ok] = __float2bfloat16(a);
      dv[ov] = __float2bfloat16(c);
    } else {
      dk[ok] = a;
      dv[ov] = c;
    }
  }
  if (blockIdx.x == 0 && dca) {
    __shared__ float red[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < B * H * chunks; i += blockDim.x) s += dca_part[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < 8; ++i) tot += red[i];
      *dca = tot * dca_mul;
    }
  }
}

template <int D, bool MIX, bool F32IN>
static int launch_cap_bwd(const CapBwdParams& p, int chunks, cudaStream_t stream) {
  using A = AttDims<D>;
  constexpr int NP = F32IN ? 2 : 1, NI = MIX ? 2 : 1;
  constexpr int smem = NI * ((NP + 1) * CB_BN + (NP + 1) * ATT_BM) * A::LD * 2 + 2 * ATT_BM * CB_LDP * 2 +
                       (1 + NI) * CB_BN * D * 4 + CB_BN * 4 + 16 + CB_BN;
  static_assert(smem <= 227 * 1024, "capture backward kernel shared memory exceeds the SM");
  static DevOnce configured;
  const int cfg_dev = af_device();
  if (!configured.done(cfg_dev)) {
    AF_CUDA(cudaFuncSetAttribute(attn_cross_capture_bwd_kernel<D, MIX, F32IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.set(cfg_dev);
  }
  attn_cross_capture_bwd_kernel<D, MIX, F32IN><<<dim3(chunks, p.H, MIX ? p.B / 2 : p.B), ATT_THREADS, smem, stream>>>(p);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// Number of query chunks (= partials per (batch, head)) the caller must size the workspaces for.
int attn_cross_capture_bwd_chunks(int64_t B, int64_t H, int64_t Lq) {
  const int64_t tiles = (Lq + ATT_BM - 1) / ATT_BM;
  int64_t chunks = (2 * 148 + B * H - 1) / (B * H);     // ~2 CTAs per SM in flight
  if (chunks > tiles) chunks = tiles;
  if (chunks > 32) chunks = 32;
  if (chunks < 1) chunks = 1;
  const int64_t per = (tiles + chunks - 1) / chunks;
  return (int)((tiles + per - 1) / per);
}

int attn_cross_capture_bwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                           const void* v, int64_t v_sb, int64_t v_sn, const void* dout, int64_t do_sb, int64_t do_sn,
                           const float* dprob, const float* dscore, int64_t B, int64_t H, int64_t Lq, int64_t S, int64_t d,
                           float scale, const uint8_t* col_flag, const float* qmean, const float* ca_scale, int mix, int in_dtype,
                           void* dq, int64_t dq_sb, int64_t dq_sn, void* dk, int64_t dk_sb, int64_t dk_sn, void* dv,
                           int64_t dv_sb, int64_t dv_sn, int dkv_dtype, float* dca, float dca_mul, float* dk_part,
                           float* dv_part, float* dca_part, cudaStream_t stream, const uint8_t* sum_flag, const float* g_subj,
                           const float* ref_prob, const float* mse_coef) {
  AF_CHECK(q && k && v && dout && dq && dk && dv && dk_part && dv_part && dca_part, "attn_cross_capture_bwd: null pointer");
  AF_CHECK(!(mix && (g_subj || ref_prob)), "attn_cross_capture_bwd: the fused consumers are not defined for mix_attn_mats_in_batch");
  AF_CHECK(!g_subj == !sum_flag || !g_subj, "attn_cross_capture_bwd: g_subj needs sum_flag");
  AF_CHECK(!ref_prob == !mse_coef, "attn_cross_capture_bwd: ref_prob and mse_coef go together");
  AF_CHECK(in_dtype == ADAFACE_BF16 || in_dtype == ADAFACE_F32, "attn_cross_capture_bwd: bad in_dtype %d", in_dtype);
  AF_CHECK(dkv_dtype == ADAFACE_BF16 || dkv_dtype == ADAFACE_F32, "attn_cross_capture_bwd: bad dkv_dtype %d", dkv_dtype);
  AF_CHECK(B > 0 && H > 0 && Lq > 0 && S > 0 && B <= 65535 && H <= 65535, "attn_cross_capture_bwd: bad problem size");
  AF_CHECK(S <= CB_BN, "attn_cross_capture_bwd: context length %lld exceeds %d keys", (long long)S, CB_BN);
  AF_CHECK(!(col_flag && !qmean), "attn_cross_capture_bwd: normalize needs qmean");
  AF_CHECK(!(mix && (B % 2)), "attn_cross_capture_bwd: mix needs an even batch [sc.., mc..] (dalc:113)");
  AF_CHECK(!(mix && col_flag), "attn_cross_capture_bwd: normalize and mix are mutually exclusive (dalc:108-119)");
  const int64_t al = in_dtype == ADAFACE_F32 ? 4 : 8;
  AF_CHECK(q_sb % al == 0 && q_sn % al == 0 && k_sb % al == 0 && k_sn % al == 0 && v_sb % al == 0 && v_sn % al == 0 &&
               do_sb % 8 == 0 && do_sn % 8 == 0 && dq_sb % 2 == 0 && dq_sn % 2 == 0,
           "attn_cross_capture_bwd: strides must keep rows 16-byte aligned");
  CapBwdParams p;
  p.q = q; p.k = k; p.v = v; p.dout = (const bf16*)dout;
  p.q_sb = q_sb; p.q_sn = q_sn; p.k_sb = k_sb; p.k_sn = k_sn; p.v_sb = v_sb; p.v_sn = v_sn; p.do_sb = do_sb; p.do_sn = do_sn;
  p.dprob = dprob; p.dscore = dscore; p.col_flag = col_flag; p.qmean = qmean; p.ca_scale = ca_scale;
  p.dq = (bf16*)dq; p.dq_sb = dq_sb; p.dq_sn = dq_sn;
  p.dk_part = dk_part; p.dv_part = dv_part; p.dca_part = dca_part;
  p.B = (int)B; p.H = (int)H; p.Lq = (int)Lq; p.S = (int)S;
  p.scale = scale;
  p.sum_flag = sum_flag; p.g_subj = g_subj; p.ref_prob = ref_prob; p.mse_coef = mse_coef;
  const int chunks = attn_cross_capture_bwd_chunks(B, H, Lq);
  const int tiles = (int)((Lq + ATT_BM - 1) / ATT_BM);
  p.tiles_per_cta = (tiles + chunks - 1) / chunks;
  const bool f32 = in_dtype == ADAFACE_F32;
  int rc;
  if (mix) {
    // the mc half of dq is detached (dalc:117): the kernel never writes it
    for (int64_t b = B / 2; b < B; ++b)
      AF_CUDA(cudaMemset2DAsync((bf16*)dq + b * dq_sb, dq_sn * 2, 0, H * d * 2, Lq, stream));
    if (d != 40) {
      set_error("attn_cross_capture_bwd: mix is built for head dim 40 (the captured SD-1.5 layers), got %lld", (long long)d);
      return 1;
    }
    rc = f32 ? launch_cap_bwd<40, true, true>(p, chunks, stream) : launch_cap_bwd<40, true, false>(p, chunks, stream);
  } else switch (d) {
    case 40: rc = f32 ? launch_cap_bwd<40, false, true>(p, chunks, stream) : launch_cap_bwd<40, false, false>(p, chunks, stream); break;
    case 80: rc = f32 ? launch_cap_bwd<80, false, true>(p, chunks, stream) : launch_cap_bwd<80, false, false>(p, chunks, stream); break;
    default:
      set_error("attn_cross_capture_bwd: unsupported head dim %lld (capture layers of SD-1.5 have d = 40; 80 also built)",
                (long long)d);
      return 1;
  }
  if (rc) return rc;
  const long long total = (long long)B * H * S * d;
  const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  if (dkv_dtype == ADAFACE_BF16)
    cross_bwd_reduce_kernel<bf16><<<grid, 256, 0, stream>>>(dk_part, dv_part, dca_part, (bf16*)dk, dk_sb, dk_sn, (bf16*)dv, dv_sb,
                                                            dv_sn, dca, dca_mul, (int)B, (int)H, (int)S, (int)d, chunks);
  else
    cross_bwd_reduce_kernel<float><<<grid, 256, 0, stream>>>(dk_part, dv_part, dca_part, (float*)dk, dk_sb, dk_sn, (float*)dv,
                                                             dv_sb, dv_sn, dca, dca_mul, (int)B, (int)H, (int)S, (int)d, chunks);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

}  // namespace adaface
